/* lq_widx.cu -- minimizer keys wider than the direct-address table (k > 15; index.c:69-86 and :150-201 use one khash per bucket there).
 *
 * Everything mapping observes of the reference's hash index is, per minimizer key, (count, positions ascending in y).  For k <= 15 the
 * key itself addresses `counts`/`offs` (lq_index.cu).  For 16 <= k <= 28 the 2k-bit keys of a part go through an open-address table in
 * HBM (linear probing, 64-bit compare-and-swap) that hands every DISTINCT key a dense id in [0, n_ids); the ids then play the role of
 * the keys for everything that follows -- count table, offsets, stable sort, occurrence threshold, look-up -- so those kernels are the
 * same for every k.  Which key gets which id depends on the order the inserts win their slots; nothing downstream can see it (ids
 * only address; a key's positions stay in y order because the sort by id is stable; mid_occ is a statistic of the multiset of counts).
 * The query minimizers are translated per part with look-ups only: a key the part does not hold maps to id n_ids, whose count is 0. */
#include "lq_widx.h"
#include "lq_prof.h"

#define LQW_EMPTY 0xffffffffffffffffULL   /* never a key: keys have at most 56 bits */

__device__ __forceinline__ uint64_t lqw_mix(uint64_t x)
{
    x ^= x >> 31; x *= 0x7fb5d329728ea185ULL; x ^= x >> 27; x *= 0x81dadef4bc2dd44dULL; x ^= x >> 33;
    return x;
}

__global__ void lqw_insert_k(const uint64_t *__restrict__ key, uint64_t n, unsigned long long *__restrict__ slots, uint32_t *__restrict__ vals, uint64_t mask,
                             uint32_t *__restrict__ n_ids)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const unsigned long long k = key[i];
        uint64_t h = lqw_mix(k) & mask;
        for (;;) {
            unsigned long long cur = slots[h];
            if (cur == LQW_EMPTY) cur = atomicCAS(&slots[h], LQW_EMPTY, k);
            if (cur == LQW_EMPTY) { vals[h] = atomicAdd(n_ids, 1u); break; }   /* this thread created the entry */
            if (cur == k) break;
            h = (h + 1) & mask;
        }
    }
}

__global__ void lqw_translate_k(const uint64_t *__restrict__ key, uint64_t n, const unsigned long long *__restrict__ slots, const uint32_t *__restrict__ vals, uint64_t mask,
                                const uint32_t *__restrict__ n_ids, uint32_t *__restrict__ out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t miss = *n_ids;
    for (; i < n; i += stride) {
        const unsigned long long k = key[i];
        uint64_t h = lqw_mix(k) & mask;
        uint32_t id = miss;
        for (;;) {
            const unsigned long long cur = slots[h];
            if (cur == k) { id = vals[h]; break; }
            if (cur == LQW_EMPTY) break;
            h = (h + 1) & mask;
        }
        out[i] = id;
    }
}

int lq_wide_build(LqWideTable *t, const uint64_t *d_key64, uint64_t n, uint32_t *d_ids, cudaStream_t st)
{
    uint64_t cap = 1024;
    while (cap < 2 * n + 16) cap <<= 1;
    if (n >= 0xfffffff0ULL) { fprintf(stderr, "[lqcov] ERROR: more than 2^32 minimizers in one index part\n"); return -1; }
    LQ_TRY(t->slots.ensure((size_t)cap * 8)); LQ_TRY(t->vals.ensure((size_t)cap * 4)); LQ_TRY(t->ctr.ensure(64));
    t->mask = cap - 1; t->n_ids = 0;
    LQ_CUDA_OK(cudaMemsetAsync(t->slots.p, 0xff, (size_t)cap * 8, st));
    LQ_CUDA_OK(cudaMemsetAsync(t->ctr.p, 0, 64, st));
    if (n) {
        unsigned grid = lq_grid(n, 256 * 4); if (grid > 148 * 32) grid = 148 * 32;
        { LqProfScope ps("widx_insert", st, 1, n * 8 + n * 16);
          lqw_insert_k<<<grid, 256, 0, st>>>(d_key64, n, t->slots.as<unsigned long long>(), t->vals.as<uint32_t>(), t->mask, t->ctr.as<uint32_t>()); }
        LQ_CUDA_OK(cudaGetLastError());
        LQ_TRY(lq_wide_translate(t, d_key64, n, d_ids, st));
    }
    uint32_t n_ids = 0;
    LQ_CUDA_OK(cudaMemcpyAsync(&n_ids, t->ctr.p, 4, cudaMemcpyDeviceToHost, st));
    LQ_CUDA_OK(cudaStreamSynchronize(st));
    t->n_ids = n_ids;
    return 0;
}

int lq_wide_translate(const LqWideTable *t, const uint64_t *d_key64, uint64_t n, uint32_t *d_ids, cudaStream_t st)
{
    if (n == 0) return 0;
    unsigned grid = lq_grid(n, 256 * 4); if (grid > 148 * 32) grid = 148 * 32;
    LqProfScope ps("widx_translate", st, 1, n * 12 + n * 12);
    lqw_translate_k<<<grid, 256, 0, st>>>(d_key64, n, t->slots.as<unsigned long long>(), t->vals.as<uint32_t>(), t->mask, t->ctr.as<uint32_t>(), d_ids);
    LQ_CUDA_OK(cudaGetLastError());
    return 0;
}
