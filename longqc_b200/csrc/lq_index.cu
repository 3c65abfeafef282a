/* lq_index.cu -- K2/K3: minimizer index of one target part.
 *
 * Replaces mm_idx_add + worker_post + mm_idx_get + mm_idx_cal_max_occ (reference index.c:229-236,
 * 150-201, 69-86, 123-144).  The reference scatters minimizers to 2^14 buckets, sorts each bucket and
 * builds one khash per bucket; what mapping can observe of that is only, per minimizer key,
 * (count, positions ascending in y).  On the B200 the key space of LongQC's k (12, 15) is small enough
 * for a DIRECT-ADDRESS table in HBM -- the degenerate, collision-free open-address hash:
 *     counts[4^k] (u32)  -- also the table that is all-reduced across GPUs
 *     offs  [4^k+1] (u64) = exclusive scan of counts
 *     pos   [n]      (u64) = every minimizer's y, stable-sorted by key
 * Records arrive ordered by (read, position) == ascending y, so a STABLE sort by key leaves every key's
 * positions ascending in y, which is what index.c:188 (radix_sort_64 per key) produces.  The sort is a
 * hand-written LSD radix sort, 8-bit digits: per-CTA digit histograms -> device-wide scan -> stable
 * scatter (warp match_any ranking), ceil(2k/8) passes.
 */
#include "lq_cuda.cuh"
#include "lq_device.h"
#include "lq_index.h"
#include <utility>
#include <vector>

#define RS_WARPS 8
#define RS_THREADS (RS_WARPS * 32)
#define RS_ROWS 8
#define RS_CHUNK (RS_THREADS * RS_ROWS) /* records per CTA */

/* ------------------------------------------------------------------ counts */

__global__ void lq_count_k(const uint32_t *__restrict__ key, uint64_t n, uint32_t *__restrict__ counts)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) atomicAdd(&counts[key[i]], 1u);
}

/* ------------------------------------------------------------------ stable LSD radix sort by key */

__global__ void __launch_bounds__(RS_THREADS) lq_rs_hist_k(const uint32_t *__restrict__ key, uint64_t n, int shift, uint32_t nblk, uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_CHUNK;
    #pragma unroll 4
    for (int r = 0; r < RS_ROWS; ++r) {
        uint64_t i = base + (uint64_t)r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(key[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    ghist[(uint64_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

/* Stable scatter.  Warp w of the CTA owns records [base + w*512, base + (w+1)*512) and walks them in rows of 32, so
 * "CTA order" == "warp, row, lane" order == input order.  The CTA's 4096 records are first ordered by digit in shared
 * memory (stable), then written out run by run, so that consecutive threads store to consecutive addresses. */
#define RS_SMEM_BYTES (RS_CHUNK * (4 + 8 + 1) + RS_WARPS * 256 * 4 + 256 * 4 + 256 * 8 + 33 * 4 + 64)
__global__ void __launch_bounds__(RS_THREADS, 5) lq_rs_scatter_k(const uint32_t *__restrict__ key_in, const uint64_t *__restrict__ y_in, const uint8_t *__restrict__ sp_in,
                                                               uint64_t n, int shift, uint32_t nblk, const uint64_t *__restrict__ gbase,
                                                               uint32_t *__restrict__ key_out, uint64_t *__restrict__ y_out, uint8_t *__restrict__ sp_out)
{
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t *s_y = (uint64_t*)rs_smem;                                   /* RS_CHUNK */
    uint64_t *s_gb = s_y + RS_CHUNK;                                      /* 256: global base of the digit minus its start inside the CTA */
    uint32_t *s_key = (uint32_t*)(s_gb + 256);                            /* RS_CHUNK */
    uint32_t (*wbase)[256] = (uint32_t (*)[256])(s_key + RS_CHUNK);       /* RS_WARPS x 256 */
    uint32_t *s_dstart = (uint32_t*)(wbase + RS_WARPS);                   /* 256 */
    uint32_t *s_scan = s_dstart + 256;                                    /* 33 */
    uint8_t *s_sp = (uint8_t*)(s_scan + 36);                              /* RS_CHUNK */
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t blk0 = (uint64_t)blockIdx.x * RS_CHUNK;
    const uint64_t base = blk0 + (uint64_t)wid * (RS_ROWS * 32);
    const uint32_t nblkrec = (uint32_t)(n - blk0 < RS_CHUNK ? n - blk0 : RS_CHUNK);
    const uint32_t lt = (1u << lane) - 1;
    for (int d = lane; d < 256; d += 32) wbase[wid][d] = 0;
    __syncwarp();
    /* 1. per-warp digit counts, and for every record its rank among the warp's records of the same digit (rows in order, lanes in
     *    order inside a row): the only sequential chain of the kernel -- one match + one counter update per row */
    uint32_t kv[RS_ROWS]; uint64_t yv[RS_ROWS];   /* the payload is fetched up front too: the chain below is a chain of warp syncs, no load may wait inside it */
    uint16_t rk[RS_ROWS];
    #pragma unroll
    for (int r = 0; r < RS_ROWS; ++r) { const uint64_t i = base + (uint64_t)r * 32 + lane; kv[r] = i < n ? key_in[i] : 0; yv[r] = i < n ? y_in[i] : 0; }
    #pragma unroll
    for (int r = 0; r < RS_ROWS; ++r) {
        const uint64_t i = base + (uint64_t)r * 32 + lane;
        const bool ok = i < n;
        const uint32_t act = __ballot_sync(0xffffffffu, ok);
        uint32_t d = 0, peers = 0;
        if (ok) {
            d = (kv[r] >> shift) & 255u;
            peers = __match_any_sync(act, d);
            rk[r] = (uint16_t)(wbase[wid][d] + __popc(peers & lt));
        }
        __syncwarp();
        if (ok && (peers & lt) == 0) wbase[wid][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    /* 2. digit d: exclusive prefix over the warps, CTA-wide start of the digit's run, global base */
    {
        const uint32_t d = threadIdx.x;
        uint32_t run = 0;
        #pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { const uint32_t t = wbase[w][d]; wbase[w][d] = run; run += t; }
        uint32_t tot;
        const uint32_t ds = lq_block_excl_scan(run, s_scan, &tot);
        s_dstart[d] = ds;
        s_gb[d] = gbase[(uint64_t)d * nblk + blockIdx.x] - ds;
    }
    __syncthreads();
    /* 3. order the CTA's records by digit in shared memory: every destination is known, nothing waits on anything */
    #pragma unroll
    for (int r = 0; r < RS_ROWS; ++r) {
        const uint64_t i = base + (uint64_t)r * 32 + lane;
        if (i < n) {
            const uint32_t d = (kv[r] >> shift) & 255u;
            const uint32_t dst = s_dstart[d] + wbase[wid][d] + rk[r];
            s_key[dst] = kv[r]; s_y[dst] = yv[r];
            if (sp_in) s_sp[dst] = sp_in[i];
        }
    }
    __syncthreads();
    /* 4. write out: record j of the CTA goes to (global base of its digit) + j */
    for (uint32_t j = threadIdx.x; j < nblkrec; j += RS_THREADS) {
        const uint32_t k = s_key[j];
        const uint64_t dst = s_gb[(k >> shift) & 255u] + j;
        if (key_out) key_out[dst] = k;
        y_out[dst] = s_y[j];
        if (sp_in) sp_out[dst] = s_sp[j];
    }
}

int lq_sort_by_key(LqMinimizers *m, int key_bits, LqDevBuf &tmp_key, LqDevBuf &tmp_y, LqDevBuf &tmp_sp, LqDevBuf &hist, LqDevBuf &ws, cudaStream_t st, int keep_keys)
{
    /* on return m->y (and m->span) are stable-sorted by key; m->key is sorted too when keep_keys, else unspecified */
    const uint64_t n = m->n;
    if (n == 0) return 0;
    const int npass = (key_bits + 7) / 8;
    const uint32_t nblk = (uint32_t)((n + RS_CHUNK - 1) / RS_CHUNK);
    LQ_TRY(tmp_key.ensure((size_t)n * 4)); LQ_TRY(tmp_y.ensure((size_t)n * 8));
    if (m->has_span) LQ_TRY(tmp_sp.ensure((size_t)n));
    LQ_TRY(hist.ensure((size_t)256 * nblk * 4 + (size_t)(256 * (size_t)nblk + 2) * 8 + 64));
    uint32_t *gh = hist.as<uint32_t>();
    uint64_t *gb = (uint64_t*)((char*)hist.p + (((size_t)256 * nblk * 4 + 63) & ~(size_t)63));
    uint32_t *kin = m->key.as<uint32_t>(), *kout = tmp_key.as<uint32_t>();
    uint64_t *yin = m->y.as<uint64_t>(), *yout = tmp_y.as<uint64_t>();
    uint8_t *sin = m->has_span ? m->span.as<uint8_t>() : 0, *sout = m->has_span ? tmp_sp.as<uint8_t>() : 0;
    LQ_CUDA_OK(cudaFuncSetAttribute(lq_rs_scatter_k, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM_BYTES));
    for (int p = 0; p < npass; ++p) {
        const int shift = 8 * p;
        { LqProfScope ps("radix_hist", st, 1, n * 4);
          lq_rs_hist_k<<<nblk, RS_THREADS, 0, st>>>(kin, n, shift, nblk, gh); }
        LQ_CUDA_OK(cudaGetLastError());
        LQ_TRY((lq_exclusive_scan<uint32_t, uint64_t>(gh, gb, (size_t)256 * nblk, 0, ws, st)));
        { LqProfScope ps("radix_scatter", st, 1, n * (12 + ((p == npass - 1 && !keep_keys) ? 8 : 12)));
          lq_rs_scatter_k<<<nblk, RS_THREADS, RS_SMEM_BYTES, st>>>(kin, yin, sin, n, shift, nblk, gb, (p == npass - 1 && !keep_keys) ? (uint32_t*)0 : kout, yout, sout); }
        LQ_CUDA_OK(cudaGetLastError());
        { uint32_t *t = kin; kin = kout; kout = t; } { uint64_t *t = yin; yin = yout; yout = t; } { uint8_t *t = sin; sin = sout; sout = t; }
    }
    if (npass & 1) { /* result sits in the tmp buffers: swap ownership */
        std::swap(m->y, tmp_y); std::swap(m->key, tmp_key);
        if (m->has_span) std::swap(m->span, tmp_sp);
    }
    return 0;
}

/* ------------------------------------------------------------------ mid_occ (index.c:123-144) */

#define OCC_BINS 65536
#define OCC_SMEM 2048
__global__ void lq_occ_hist_k(const uint32_t *__restrict__ counts, uint64_t nkeys, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t h[OCC_SMEM];
    for (int j = threadIdx.x; j < OCC_SMEM; j += blockDim.x) h[j] = 0;
    __syncthreads();
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < nkeys; i += stride) {
        const uint32_t c = counts[i];
        if (c == 0) continue;
        if (c < OCC_SMEM) atomicAdd(&h[c], 1u);
        else atomicAdd(&hist[c < OCC_BINS - 1 ? c : OCC_BINS - 1], 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < OCC_SMEM; j += blockDim.x) if (h[j]) atomicAdd(&hist[j], h[j]);
}

__global__ void lq_occ_big_k(const uint32_t *__restrict__ counts, uint64_t nkeys, uint32_t *__restrict__ list, uint32_t *__restrict__ n_list, uint32_t cap)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < nkeys; i += stride) {
        const uint32_t c = counts[i];
        if (c >= OCC_BINS - 1) { uint32_t at = atomicAdd(n_list, 1u); if (at < cap) list[at] = c; }
    }
}

static int cmp_u32(const void *a, const void *b) { uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return x < y ? -1 : x > y; }

int lq_index_mid_occ(const LqIndexDev *ix, float frac, int32_t *mid_occ, uint64_t *n_distinct, LqDevBuf &ws, cudaStream_t st)
{
    /* thres = (kk-th smallest occurrence count over distinct minimizers) + 1, kk = (uint32)((1. - f) * n)  -- index.c:141 */
    std::vector<uint32_t> h(OCC_BINS);
    LQ_TRY(ws.ensure((size_t)OCC_BINS * 4 + 16));
    LQ_CUDA_OK(cudaMemsetAsync(ws.p, 0, (size_t)OCC_BINS * 4 + 16, st));
    const uint64_t nkeys = ix->n_keyspace;
    unsigned grid = lq_grid(nkeys, 256 * 16); if (grid > 148 * 16) grid = 148 * 16;
    lq_prof_count_launch(1); lq_prof_d2h((uint64_t)OCC_BINS * 4);
    lq_occ_hist_k<<<grid, 256, 0, st>>>(ix->counts.as<uint32_t>(), nkeys, ws.as<uint32_t>());
    LQ_CUDA_OK(cudaGetLastError());
    LQ_CUDA_OK(cudaMemcpyAsync(h.data(), ws.p, (size_t)OCC_BINS * 4, cudaMemcpyDeviceToHost, st));
    LQ_CUDA_OK(cudaStreamSynchronize(st));
    uint64_t n = 0;
    for (int c = 1; c < OCC_BINS; ++c) n += h[c];
    *n_distinct = n;
    if (frac <= 0.f) { *mid_occ = INT32_MAX; return 0; }
    if (n == 0) { *mid_occ = 1; return 0; } /* the reference reads uninitialised memory here; no seeds exist either way */
    const uint64_t kk = (uint32_t)((1. - frac) * n);
    uint64_t seen = 0;
    for (int c = 1; c < OCC_BINS - 1; ++c) {
        seen += h[c];
        if (kk < seen) { *mid_occ = c + 1; return 0; }
    }
    { /* the quantile falls among counts >= 65535: fetch and sort those few */
        const uint32_t nbig = h[OCC_BINS - 1];
        LqDevBuf lst; LQ_TRY(lst.ensure((size_t)(nbig + 2) * 4));
        uint32_t *d_n = lst.as<uint32_t>() + nbig + 1;
        LQ_CUDA_OK(cudaMemsetAsync(d_n, 0, 4, st));
        lq_occ_big_k<<<grid, 256, 0, st>>>(ix->counts.as<uint32_t>(), nkeys, lst.as<uint32_t>(), d_n, nbig);
        LQ_CUDA_OK(cudaGetLastError());
        std::vector<uint32_t> big(nbig);
        LQ_CUDA_OK(cudaMemcpyAsync(big.data(), lst.p, (size_t)nbig * 4, cudaMemcpyDeviceToHost, st));
        LQ_CUDA_OK(cudaStreamSynchronize(st));
        lst.release();
        qsort(big.data(), nbig, 4, cmp_u32);
        uint64_t idx = kk - seen;
        *mid_occ = (int32_t)(big[idx < nbig ? idx : nbig - 1] + 1);
    }
    return 0;
}

/* ------------------------------------------------------------------ part build */

int lq_index_alloc(LqIndexDev *ix, int k, uint64_t n_ids, cudaStream_t st)
{
    if (k < 1 || k > LQ_MAX_K || (k > LQ_MAX_K_DIRECT) != (n_ids != 0)) {
        fprintf(stderr, "[lqcov] k=%d: unsupported by the index of this build (1..%d)\n", k, LQ_MAX_K);
        return -1;
    }
    ix->k = k;
    if (n_ids) { ix->n_keyspace = n_ids; ix->key_bits = 1; while (ix->key_bits < 32 && (n_ids >> ix->key_bits)) ++ix->key_bits; }
    else { ix->n_keyspace = 1ULL << (2 * k); ix->key_bits = 2 * k; }
    LQ_TRY(ix->counts.ensure((size_t)(ix->n_keyspace + 1) * 4));
    LQ_TRY(ix->offs.ensure((size_t)(ix->n_keyspace + 2) * 8));
    LQ_CUDA_OK(cudaMemsetAsync(ix->counts.p, 0, (size_t)(ix->n_keyspace + 1) * 4, st));
    return 0;
}

int lq_index_count(LqIndexDev *ix, const LqMinimizers *m, cudaStream_t st)
{
    if (m->n == 0) return 0;
    unsigned grid = lq_grid(m->n, 256 * 8); if (grid > 148 * 32) grid = 148 * 32;
    LqProfScope ps("idx_count", st, 1, m->n * 8);
    lq_count_k<<<grid, 256, 0, st>>>(m->key.as<uint32_t>(), m->n, ix->counts.as<uint32_t>());
    LQ_CUDA_OK(cudaGetLastError());
    return 0;
}

int lq_index_finish(LqIndexDev *ix, LqMinimizers *m, LqDevBuf &ws, cudaStream_t st)
{
    /* counts are final (all-reduced when several GPUs share the part); m holds ALL records of the part in y order */
    { LqProfScope ps("offs_scan", st, 0, ix->n_keyspace * 16);
      LQ_TRY((lq_exclusive_scan<uint32_t, uint64_t>(ix->counts.as<uint32_t>(), ix->offs.as<uint64_t>(), (size_t)ix->n_keyspace, 1, ws, st))); }
    LQ_TRY(lq_sort_by_key(m, ix->key_bits, ix->tmp_key, ix->tmp_y, ix->tmp_sp, ix->hist, ws, st, 0));
    ix->n_rec = m->n;
    return 0;
}
