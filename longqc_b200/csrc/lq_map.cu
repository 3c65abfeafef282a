/* lq_map.cu -- K4..K7: seed lookup, exact seed sort, chaining, coverage accounting for a batch of queries.
 *
 * Replaces, per query and index part, lq_map_frag_mod() of the reference (lqmap.c:207-326):
 *   collect_seed_hits   lqmap.c:140-205   -> lq_lookup_k, lq_qstat_k, lq_fill_k
 *   radix_sort_128x     ksort.h:84-134    -> lq_af_* (exact permutation, see lq_afsort_core.h)
 *   mm_chain_dp         chain.c:22-157    -> lq_chain_k (one warp per (query, strand, target) group)
 *   mm_gen_regs         hit.c:23-88       -> region coordinates inside lq_chain_k
 *   lq_cnt_match        esterr.c:72-140   -> overlap filter, lambda/lambda2, per-minimizer counters
 * Queries are processed in batches bounded by a seed budget; all per-seed arrays live in two arenas.
 */
#include <vector>
#include <algorithm>
#include "lq_cuda.cuh"
#include "lq_map.h"
#include "lq_afsort_core.h"
#include "lq_chain_core.h"

void LqQueryDev::release()
{
    reads.release(); mins.release(); first.release(); qtied.release(); dup.release(); dup_tmp.release(); dup_tk.release(); dup_ty.release(); dup_ts.release(); dup_hist.release(); nmatch_buf.release(); lambda.release(); lambda2.release(); mcnt.release();
    keep.release(); neff.release(); krank.release(); soff.release(); qstat.release(); fmask.release();
    self_off.release(); self_list.release(); qrank.release(); trank.release();
}

struct MapTables {
    int no_self, ava;
    const uint32_t *self_off, *self_list, *qrank, *trank;
};

/* lqmap.c:180-189: the exact self-diagonal / dual-mapping skips */
__device__ __forceinline__ bool lq_seed_skipped(const MapTables &t, uint64_t r, uint32_t qpos, uint32_t q)
{
    const uint32_t rid = (uint32_t)(r >> 32), rpos = (uint32_t)r >> 1;
    if (t.no_self && rpos == qpos)
        for (uint32_t s = t.self_off[q]; s < t.self_off[q + 1]; ++s) if (t.self_list[s] == rid) return true;
    if (t.ava && t.qrank[q] > t.trank[rid]) return true;
    return false;
}

/* ------------------------------------------------------------------ K4a: lookup (one thread per query minimizer) */

__global__ void lq_lookup_k(const uint32_t *__restrict__ qkey, const uint64_t *__restrict__ qy, uint64_t n_min,
                            const uint32_t *__restrict__ counts, const uint64_t *__restrict__ offs, const uint64_t *__restrict__ pos,
                            int mid_occ, MapTables t, const uint64_t *__restrict__ lambda, const uint32_t *__restrict__ qlen, int covt,
                            uint32_t *__restrict__ keep, uint32_t *__restrict__ neff)
{
    const uint64_t mi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (mi >= n_min) return;
    const uint64_t y = qy[mi];
    const uint32_t q = (uint32_t)(y >> 32), key = qkey[mi];
    const bool gate = lambda[q] / (uint64_t)qlen[q] > (uint64_t)covt; /* esterr.c:87: nothing of this query is counted in this part */
    const uint32_t c = counts[key];
    const bool kept = !gate && (int64_t)c < (int64_t)mid_occ;          /* lqmap.c:159,166 */
    uint32_t n = kept ? c : 0;
    if (n && t.ava) {
        const uint64_t o = offs[key];
        const uint32_t qpos = (uint32_t)y >> 1;
        uint32_t skipped = 0;
        for (uint32_t j = 0; j < c; ++j) skipped += lq_seed_skipped(t, pos[o + j], qpos, q);
        n -= skipped;
    } else if (n && t.no_self) {
        /* only the self-diagonal skip: occurrences (rid, rpos == qpos) of a target named like the query, either strand.  The
         * occurrence list ascends in y = rid<<32 | rpos<<1 | strand, so they are found by binary search. */
        const uint64_t *pp = pos + offs[key];
        const uint32_t qpos = (uint32_t)y >> 1;
        for (uint32_t s = t.self_off[q]; s < t.self_off[q + 1]; ++s) {
            const uint64_t lo_y = (uint64_t)t.self_list[s] << 32 | (uint64_t)qpos << 1;
            uint32_t lo = 0, hi = c;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (pp[mid] < lo_y) lo = mid + 1; else hi = mid; }
            while (lo < c && pp[lo] <= (lo_y | 1ULL)) { --n; ++lo; }
        }
    }
    keep[mi] = kept; neff[mi] = n;
}

/* ------------------------------------------------------------------ K4b: per-query statistics (one warp per query) */

__global__ void lq_qstat_k(uint32_t nq, const uint64_t *__restrict__ first, const uint32_t *__restrict__ keep, const uint32_t *__restrict__ neff,
                           const uint8_t *__restrict__ span, int k, const uint64_t *__restrict__ lambda, const uint32_t *__restrict__ qlen, int covt,
                           LqQStat *__restrict__ out)
{
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    uint32_t nk = 0, ssk = 0; uint64_t ns = 0, sss = 0;
    for (uint64_t mi = first[q] + lane; mi < first[q + 1]; mi += 32) {
        const uint32_t sp = span ? span[mi] : (uint32_t)k;
        if (keep[mi]) { ++nk; ssk += sp; }
        ns += neff[mi]; sss += (uint64_t)neff[mi] * sp;
    }
    nk = lq_warp_sum(nk); ssk = lq_warp_sum(ssk); ns = lq_warp_sum(ns); sss = lq_warp_sum(sss);
    if (lane == 0) {
        LqQStat s;
        s.n_kept = nk; s.sum_span_kept = ssk; s.n_seeds = ns; s.sum_span_seeds = sss;
        s.gate_closed = qlen[q] ? lambda[q] / (uint64_t)qlen[q] > (uint64_t)covt : 0;
        s.avg_span = ns ? __fdiv_rn(__ull2float_rn(sss), __ll2float_rn((long long)ns)) : 0.f; /* chain.c:38 */
        out[q] = s;
    }
}

/* ------------------------------------------------------------------ K4c: seed fill (one warp per kept query minimizer of the batch) */

struct SeedArrays { uint64_t *sx; uint32_t *sq, *sm, *idx; };   /* idx[at] = at | tie mark: the sort's (key, idx) pairs start out here */

__global__ void lq_fill_k(uint64_t mi0, uint64_t mi1, const uint32_t *__restrict__ qkey, const uint64_t *__restrict__ qy, const uint8_t *__restrict__ qspan, int k,
                          const uint64_t *__restrict__ first, const uint32_t *__restrict__ keep, const uint32_t *__restrict__ neff,
                          const uint32_t *__restrict__ krank, const uint64_t *__restrict__ soff, uint64_t seed_base,
                          const uint32_t *__restrict__ counts, const uint64_t *__restrict__ offs, const uint64_t *__restrict__ pos,
                          MapTables t, const uint32_t *__restrict__ qlen, const uint8_t *__restrict__ dup, const uint8_t *__restrict__ qtied /* null: every query */, SeedArrays s)
{
    const uint64_t mi = mi0 + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (mi >= mi1 || !keep[mi] || neff[mi] == 0) return;
    const uint64_t y = qy[mi];
    const uint32_t q = (uint32_t)(y >> 32), key = qkey[mi], c = counts[key];
    if (qtied && !qtied[q]) return;                       /* queries without tied keys are written by lq_fill_filtered_k */
    const uint32_t qpos = (uint32_t)y >> 1, qstrand = (uint32_t)y & 1, span = qspan ? qspan[mi] : (uint32_t)k;
    const uint32_t rank = krank[mi] - krank[first[q]];       /* index into the query's mini_pos (lqmap.c:174) */
    const bool filt = t.ava || (t.no_self && t.self_off[q + 1] > t.self_off[q]);
    const uint64_t o = offs[key];
    uint64_t out = soff[mi] - seed_base;
    const int32_t ql = (int32_t)qlen[q];
    const uint32_t tie = dup[mi] ? 0x80000000u : 0u;       /* carried through the sort only (bit 31 of sq) */
    for (uint32_t j0 = 0; j0 < c; j0 += 32) {
        const uint32_t j = j0 + lane;
        uint64_t r = 0; bool ok = j < c;
        if (ok) { r = pos[o + j]; if (filt && lq_seed_skipped(t, r, qpos, q)) ok = false; }
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const uint64_t at = out + __popc(m & ((1u << lane) - 1));
            const uint32_t rpos = (uint32_t)r >> 1;
            if (((uint32_t)r & 1) == qstrand) { /* lqmap.c:191-193 */
                s.sx[at] = (r & 0xffffffff00000000ULL) | rpos;
                s.sq[at] = qpos | tie;
            } else {                             /* lqmap.c:194-197 */
                s.sx[at] = 1ULL << 63 | (r & 0xffffffff00000000ULL) | rpos;
                s.sq[at] = (uint32_t)(ql - ((int32_t)qpos + 1 - (int32_t)span) - 1) | tie;
            }
            s.sm[at] = span << 24 | rank;
            s.idx[at] = (uint32_t)at | tie;
        }
        out += __popc(m);
    }
}


/* ------------------------------------------------------------------ seed pre-filter for queries without tied keys
 *
 * A (strand, target) run of fewer than T = max(min_cnt, ceil(min_sc / max_span)) anchors cannot hold a chain (chain.c:116-119, and a
 * chain's score is at most the sum of its anchors' spans), so its seeds influence nothing that is printed -- EXCEPT through the
 * reference's unstable sort, whose permutation of tied keys depends on every element.  For a query none of whose minimizers
 * repeats (no tied keys, lq_map_flag_dups) the sorted order is unique, so the seeds of short runs can be dropped before they
 * are ever written.  The run lengths are bounded from above with 4-bit saturating counters over a hash of (strand, target) in
 * shared memory (collisions only make a run look longer); ~93 % of the seeds of a typical query are chance hits and go away.
 * Queries with tied keys keep every seed. */
#define FL_THREADS 1024
#define FL_ILP 8
#define FL_BINS_LOG 18
#define FL_WORDS (1u << (FL_BINS_LOG - 3))      /* 4-bit counters, 8 per word: 128 KiB */

__device__ __forceinline__ uint32_t fl_hash(uint64_t r, uint32_t qstrand)
{
    const uint32_t v = (uint32_t)(r >> 32) << 1 | ((((uint32_t)r & 1u) != qstrand) ? 1u : 0u);
    return (v * 0x9E3779B1u) >> (32 - FL_BINS_LOG);
}
__device__ __forceinline__ void fl_inc(uint32_t *bins, uint32_t h)
{
    const uint32_t w = h >> 3, sh = (h & 7u) * 4;
    uint32_t old = bins[w];
    while (((old >> sh) & 15u) < 15u) { const uint32_t as = old; old = atomicCAS(&bins[w], as, as + (1u << sh)); if (old == as) break; }
}
__device__ __forceinline__ uint32_t fl_get(const uint32_t *bins, uint32_t h) { return (bins[h >> 3] >> ((h & 7u) * 4)) & 15u; }

struct FlArgs {
    const uint32_t *qkey; const uint64_t *qy; const uint8_t *qspan; int k;
    const uint64_t *first; const uint32_t *keep; uint32_t *neff; const uint8_t *qtied;
    const uint32_t *counts; const uint64_t *offs, *pos;
    MapTables t; const uint32_t *qlen;
    uint32_t thr;                 /* T */
    uint32_t *mask; uint32_t mstride;   /* survivor bits per occurrence: mstride words per minimizer (lq_filter_count_k -> lq_fill_masked_k) */
    /* fill only */
    const uint32_t *krank; const uint64_t *soff; uint64_t seed_base; SeedArrays s;
    uint32_t q0, q1;
};

/* Both kernels walk the query's occurrence lists one THREAD per minimizer (a list is ~70 consecutive 8-byte entries: each thread
 * streams its own list, and the 16..32 resident warps hide the latency); the fill then writes each minimizer's survivors to its own
 * contiguous output range, which preserves the reference's order by construction.  Measured alternatives on B200 (same workload,
 * 11.3 ms for this form): a warp per list 14.2 ms, eight lanes per list 16.3 ms -- coalescing the 64..256-byte lists buys less than
 * the loads in flight it costs. */

/* pass 1 of both kernels: run-length bounds of query q into bins[] */
__device__ __forceinline__ void fl_count_runs(const FlArgs &a, uint32_t q, uint32_t *bins)
{
    for (uint32_t j = threadIdx.x; j < FL_WORDS; j += blockDim.x) bins[j] = 0;
    __syncthreads();
    const bool filt = a.t.ava || (a.t.no_self && a.t.self_off[q + 1] > a.t.self_off[q]);
    for (uint64_t mi = a.first[q] + threadIdx.x; mi < a.first[q + 1]; mi += blockDim.x) {
        if (!a.keep[mi] || a.neff[mi] == 0) continue;
        const uint64_t y = a.qy[mi];
        const uint32_t key = a.qkey[mi], c = a.counts[key], qpos = (uint32_t)y >> 1, qstrand = (uint32_t)y & 1;
        const uint64_t *pp = a.pos + a.offs[key];
        for (uint32_t j0 = 0; j0 < c; j0 += FL_ILP) {   /* FL_ILP loads in flight per thread */
            uint64_t rr[FL_ILP];
            #pragma unroll
            for (int u = 0; u < FL_ILP; ++u) rr[u] = j0 + u < c ? pp[j0 + u] : 0;
            #pragma unroll
            for (int u = 0; u < FL_ILP; ++u) {
                if (j0 + u >= c) break;
                if (filt && lq_seed_skipped(a.t, rr[u], qpos, q)) continue;
                fl_inc(bins, fl_hash(rr[u], qstrand));
            }
        }
    }
    __syncthreads();
}

extern __shared__ uint32_t fl_bins[];

/* survivors per minimizer -> neff (queries without tied keys only) */
__global__ void __launch_bounds__(FL_THREADS) lq_filter_count_k(FlArgs a)
{
    for (uint32_t q = a.q0 + blockIdx.x; q < a.q1; q += gridDim.x) {
        if (a.qtied[q]) continue;
        __syncthreads();
        fl_count_runs(a, q, fl_bins);
        const bool filt = a.t.ava || (a.t.no_self && a.t.self_off[q + 1] > a.t.self_off[q]);
        for (uint64_t mi = a.first[q] + threadIdx.x; mi < a.first[q + 1]; mi += blockDim.x) {
            if (!a.keep[mi] || a.neff[mi] == 0) continue;
            const uint64_t y = a.qy[mi];
            const uint32_t key = a.qkey[mi], c = a.counts[key], qpos = (uint32_t)y >> 1, qstrand = (uint32_t)y & 1;
            const uint64_t *pp = a.pos + a.offs[key];
            uint32_t n = 0, mword = 0;
            uint32_t *mw = a.mask + mi * a.mstride;
            for (uint32_t j0 = 0; j0 < c; j0 += FL_ILP) {
                uint64_t rr[FL_ILP];
                #pragma unroll
                for (int u = 0; u < FL_ILP; ++u) rr[u] = j0 + u < c ? pp[j0 + u] : 0;
                #pragma unroll
                for (int u = 0; u < FL_ILP; ++u) {
                    if (j0 + u >= c) break;
                    if (filt && lq_seed_skipped(a.t, rr[u], qpos, q)) continue;
                    const uint32_t live = fl_get(fl_bins, fl_hash(rr[u], qstrand)) >= a.thr;
                    n += live; mword |= live << ((j0 + u) & 31);
                }
                if (((j0 + FL_ILP) & 31) == 0 || j0 + FL_ILP >= c) { mw[j0 >> 5] = mword; mword = 0; }   /* FL_ILP divides 32 */
            }
            a.neff[mi] = n;
        }
    }
}

/* the seeds that survive, in the reference's order (minimizer order x ascending target position): one thread per minimizer walks
 * the survivor bits lq_filter_count_k left and writes its own contiguous output range */
__global__ void lq_fill_masked_k(FlArgs a, uint64_t mi0, uint64_t mi1)
{
    const uint64_t mi = mi0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (mi >= mi1 || !a.keep[mi] || a.neff[mi] == 0) return;
    const uint64_t y = a.qy[mi];
    const uint32_t q = (uint32_t)(y >> 32);
    if (a.qtied[q]) return;                                /* written by lq_fill_k */
    const uint32_t key = a.qkey[mi], c = a.counts[key], qpos = (uint32_t)y >> 1, qstrand = (uint32_t)y & 1;
    const uint32_t span = a.qspan ? a.qspan[mi] : (uint32_t)a.k;
    const uint32_t sm = span << 24 | (a.krank[mi] - a.krank[a.first[q]]);
    const uint32_t sq_rev = (uint32_t)((int32_t)a.qlen[q] - ((int32_t)qpos + 1 - (int32_t)span) - 1);
    const uint64_t *pp = a.pos + a.offs[key];
    const uint32_t *mw = a.mask + mi * a.mstride;
    uint64_t at = a.soff[mi] - a.seed_base;
    for (uint32_t w0 = 0; w0 * 32 < c; ++w0) {
        uint32_t m = mw[w0];
        while (m) {
            const uint32_t j = w0 * 32 + (uint32_t)(__ffs(m) - 1); m &= m - 1;
            const uint64_t r = pp[j];
            const uint32_t rpos = (uint32_t)r >> 1;
            if (((uint32_t)r & 1) == qstrand) { a.s.sx[at] = (r & 0xffffffff00000000ULL) | rpos; a.s.sq[at] = qpos; }
            else { a.s.sx[at] = 1ULL << 63 | (r & 0xffffffff00000000ULL) | rpos; a.s.sq[at] = sq_rev; }
            a.s.sm[at] = sm;
            a.s.idx[at] = (uint32_t)at;
            ++at;
        }
    }
}

/* per query: does any of its minimizers repeat (=> tied sort keys)? */
__global__ void lq_qtied_k(uint32_t nq, const uint64_t *__restrict__ first, const uint8_t *__restrict__ dup, uint8_t *__restrict__ qtied)
{
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    uint32_t t = 0;
    for (uint64_t mi = first[q] + lane; mi < first[q + 1]; mi += 32) t |= dup[mi];
    t = __any_sync(0xffffffffu, t);
    if (lane == 0) qtied[q] = (uint8_t)t;
}
/* seeds of each query after filtering */
__global__ void lq_qseeds_k(uint32_t nq, const uint64_t *__restrict__ first, const uint64_t *__restrict__ soff, LqQStat *__restrict__ qs)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) qs[q].n_sorted = soff[first[q + 1]] - soff[first[q]];
}

/* ------------------------------------------------------------------ K5: exact seed sort */

struct AfBkt { uint32_t beg, end; };
/* The sort moves (key, idx) pairs: kx[p] = sort key of the element now at p, idx[p] = its seed number (bit 31: the seed belongs to a
 * repeated query minimizer, so its key may be tied).  Nothing is gathered through idx until the very end (lq_gather_k).
 * Two buffers, no copy-back: a level reads its buckets from (kx, idx) and leaves every element it handles in (kx2, idx2); the roles
 * swap from level to level.  An element is FINISHED when its sub-bucket has <= 64 elements (sorted on the spot) -- it must then sit
 * in the primary buffer (kxp, idxp), which is this level's destination on every other level and costs one extra copy otherwise. */
struct AfArgs {
    uint64_t *kx, *kx2, *kxp; uint32_t *idx, *idx2, *idxp, *dest, *ord; uint8_t *dig;   /* (ord, dest): pick-up order and slots of a walked bucket */
    int dst_primary;              /* (kx2, idx2) == (kxp, idxp) */
    const AfBkt *cur; const uint32_t *n_cur; AfBkt *nxt; uint32_t *n_nxt; uint32_t *cursor; uint32_t *n_walk;
    AfBkt *wlist; uint32_t *n_wlist; uint32_t *wcursor, *wcursor_f;   /* tied buckets with > 2 digits: walked by lq_af_walk_k (>= AFW_SMALL elements) ... */
    AfBkt *wlist_s; uint32_t *n_wlist_s; uint32_t *wcursor_s;   /* ... or by lq_af_walk_small_k (a warp per bucket) */
    AfBkt *wlist_w; uint32_t *n_wlist_w; uint32_t *wcursor_w;   /* ... or, all digits < 16 (few regions), by lq_af_walkf_k */
    lq_afq_phase *wph;            /* 256 phase entries per bucket of wlist, then of wlist_w (capacity wph_cap each) */
    uint32_t wph_cap;
    unsigned long long *n_elem;   /* elements this launch handled (profiling: algorithmic bytes of the launch) */
    const AfBkt *curb; const uint32_t *n_curb; AfBkt *nxtb; uint32_t *n_nxtb; uint32_t *cursorb;   /* buckets of >= AFB_MIN elements: a CTA each (lq_af_big_k) */
    int shift;
    int two;                      /* this level's digit takes at most two values in every bucket (strand; rid >> 16 of a part of <= 131072 reads) */
    uint32_t bitw;                /* 32-bit words of lq_af_big_k's shared-memory digit bitmap */
};

#define AFB_MIN 4096               /* buckets at least this long are sorted by a whole CTA (lq_af_big_k), shorter ones by a warp (lq_af_level_k) */
#define AFW_SMALL 2048             /* walks shorter than this: a warp per bucket (lq_af_walk_small_k); longer: a lane per bucket (lq_af_walk_k) */

__global__ void lq_af_init_k(uint32_t nqb, const uint64_t *__restrict__ qoff /* nqb+1 seed offsets in the batch */, uint64_t *__restrict__ kx,
                             uint32_t *__restrict__ idx, AfBkt *__restrict__ bkt, uint32_t *__restrict__ n_bkt, AfBkt *__restrict__ bktb, uint32_t *__restrict__ n_bktb)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nqb) return;
    const uint32_t beg = (uint32_t)qoff[q], end = (uint32_t)qoff[q + 1], n = end - beg;
    if (n >= AFB_MIN) { const uint32_t at = atomicAdd(n_bktb, 1u); bktb[at].beg = beg; bktb[at].end = end; }
    else if (n > LQ_RS_MIN) { const uint32_t at = atomicAdd(n_bkt, 1u); bkt[at].beg = beg; bkt[at].end = end; }
    else if (n > 1) lq_af_insertion_kv(kx + beg, idx + beg, n); /* ksort.h:132 */
}



#define AF_WARPS 4
#define AF_U 8

/* append a sub-bucket to the next level's list: one atomic per warp (the lanes that append are counted with a ballot) */
__device__ __forceinline__ void af_append(const AfArgs &a, uint32_t beg, uint32_t c, bool doit)
{
    const uint32_t lane = threadIdx.x & 31;
    const bool big = doit && c >= AFB_MIN, small = doit && !big;
    const uint32_t ms = __ballot_sync(0xffffffffu, small), mb = __ballot_sync(0xffffffffu, big);
    if ((ms | mb) == 0) return;
    uint32_t bs = 0, bb = 0;
    if (lane == 0) { if (ms) bs = atomicAdd(a.n_nxt, (uint32_t)__popc(ms)); if (mb) bb = atomicAdd(a.n_nxtb, (uint32_t)__popc(mb)); }
    bs = __shfl_sync(0xffffffffu, bs, 0); bb = __shfl_sync(0xffffffffu, bb, 0);
    if (small) { const uint32_t at = bs + __popc(ms & ((1u << lane) - 1)); a.nxt[at].beg = beg; a.nxt[at].end = beg + c; }
    if (big) { const uint32_t at = bb + __popc(mb & ((1u << lane) - 1)); a.nxtb[at].beg = beg; a.nxtb[at].end = beg + c; }
}

/* a tied bucket with more than two digits goes to one of the two walk kernels (called by one thread) */
__device__ __forceinline__ void af_push_walk(const AfArgs &a, uint32_t beg, uint32_t n, bool few /* no digit above 15 */)
{
    atomicAdd(a.n_elem, 0ULL - (unsigned long long)n);    /* profiling: the level only looked at these elements, the walk's place kernel moves them */
    if (n < AFW_SMALL) { const uint32_t at = atomicAdd(a.n_wlist_s, 1u); a.wlist_s[at].beg = beg; a.wlist_s[at].end = beg + n; }
    else if (few) { const uint32_t at = atomicAdd(a.n_wlist_w, 1u); a.wlist_w[at].beg = beg; a.wlist_w[at].end = beg + n; }
    else { const uint32_t at = atomicAdd(a.n_wlist, 1u); a.wlist[at].beg = beg; a.wlist[at].end = beg + n; }
}

/* stable sort of a sub-bucket of 9..64 elements by key, one warp: each lane holds two elements, rank = #smaller + #equal-before
 * (== the order ksort.h's insertion sort leaves) */
__device__ __forceinline__ void af_warp_ranksort(uint64_t *key, uint32_t *idx, uint32_t n, uint32_t lane, uint64_t *key2 = 0, uint32_t *idx2 = 0)
{
    const uint32_t e0 = lane < n ? idx[lane] : 0, e1 = lane + 32 < n ? idx[lane + 32] : 0;
    const uint64_t k0 = lane < n ? key[lane] : ~0ULL, k1 = lane + 32 < n ? key[lane + 32] : ~0ULL;
    uint32_t r0 = 0, r1 = 0;
    for (uint32_t j = 0; j < n; ++j) {
        const uint64_t kj = __shfl_sync(0xffffffffu, j < 32 ? k0 : k1, j & 31);
        r0 += kj < k0 || (kj == k0 && j < lane);
        r1 += kj < k1 || (kj == k1 && j < lane + 32);
    }
    __syncwarp();
    if (lane < n) { idx[r0] = e0; key[r0] = k0; if (key2) { idx2[r0] = e0; key2[r0] = k0; } }
    if (lane + 32 < n) { idx[r1] = e1; key[r1] = k1; if (key2) { idx2[r1] = e1; key2[r1] = k1; } }
    __syncwarp();
}

/* after dest[] is known: permute the payload, then hand the sub-buckets on (ksort.h:124-133) */
#define AF_SN 512
template <int SN>
__device__ __forceinline__ void af_finish_bucket(const AfArgs &a, uint32_t beg, uint32_t n, uint32_t nb, const uint32_t *cnt, const uint32_t *start,
                                                 const uint32_t *dest, uint32_t lane, uint64_t *s_k, uint32_t *s_i)
{
    const uint64_t *ks = a.kx + beg; const uint32_t *is = a.idx + beg;        /* where the bucket is */
    uint64_t *kd = a.kx2 + beg; uint32_t *id = a.idx2 + beg;                  /* where this level leaves it */
    uint64_t *kp = a.kxp + beg; uint32_t *ip = a.idxp + beg;                  /* where finished elements belong */
    if (a.shift > 0 && n <= SN) {
        /* small bucket: permuted straight into shared memory, so that the insertion sorts of its sub-buckets run at shared-memory
         * latency, and written out once */
        if (nb > 1) for (uint32_t p = lane; p < n; p += 32) { const uint32_t d = dest[p]; s_i[d] = is[p]; s_k[d] = ks[p]; }
        else for (uint32_t p = lane; p < n; p += 32) { s_i[p] = is[p]; s_k[p] = ks[p]; }
        __syncwarp();
        for (uint32_t d = lane; d < 256; d += 32) {
            const uint32_t c = cnt[d];
            const bool big_ = c > LQ_RS_MIN;
            af_append(a, beg + start[d], c, big_);
            if (!big_ && c > 1) lq_af_insertion_kv(s_k + start[d], s_i + start[d], c); /* ksort.h:88-98 on the staged copy */
        }
        __syncwarp();
        for (uint32_t p = lane; p < n; p += 32) { id[p] = s_i[p]; kd[p] = s_k[p]; }
        if (!a.dst_primary) for (uint32_t p = lane; p < n; p += 32) { ip[p] = s_i[p]; kp[p] = s_k[p]; }
        __syncwarp();
        return;
    }
    if (nb > 1) {
        for (uint32_t p0 = lane; p0 < n; p0 += 32 * AF_U) {
            uint32_t dd[AF_U], ee[AF_U]; uint64_t kk[AF_U];
            #pragma unroll
            for (int u = 0; u < AF_U; ++u) { const uint32_t p = p0 + u * 32; if (p < n) { dd[u] = dest[p]; ee[u] = is[p]; kk[u] = ks[p]; } }
            #pragma unroll
            for (int u = 0; u < AF_U; ++u) { const uint32_t p = p0 + u * 32; if (p < n) { id[dd[u]] = ee[u]; kd[dd[u]] = kk[u]; } }
        }
    } else {
        for (uint32_t p = lane; p < n; p += 32) { id[p] = is[p]; kd[p] = ks[p]; }
    }
    __syncwarp();
    if (a.shift > 0) {
        uint32_t mid = 0;   /* digits (bit per owned digit) whose sub-bucket has 9..64 elements: sorted by the whole warp afterwards */
        for (uint32_t d = lane; d < 256; d += 32) {
            const uint32_t c = cnt[d];
            const bool big_ = c > LQ_RS_MIN;
            af_append(a, beg + start[d], c, big_);
            if (big_ || c == 0) {}
            else if (c > 8) mid |= 1u << (d >> 5);
            else {
                if (c > 1) lq_af_insertion_kv(kd + start[d], id + start[d], c);
                if (!a.dst_primary) for (uint32_t x = 0; x < c; ++x) { kp[start[d] + x] = kd[start[d] + x]; ip[start[d] + x] = id[start[d] + x]; }
            }
        }
        __syncwarp();
        for (uint32_t l = 0; l < 32; ++l) {
            uint32_t m = __shfl_sync(0xffffffffu, mid, l);
            while (m) {
                const uint32_t d = (uint32_t)(__ffs(m) - 1) * 32 + l; m &= m - 1;
                af_warp_ranksort(kd + start[d], id + start[d], cnt[d], lane, a.dst_primary ? (uint64_t*)0 : kp + start[d], a.dst_primary ? (uint32_t*)0 : ip + start[d]);
            }
        }
    } else if (!a.dst_primary) {
        for (uint32_t p = lane; p < n; p += 32) { ip[p] = id[p]; kp[p] = kd[p]; }
    }
    __syncwarp();
}
/* A bucket of at most AF_SN elements none of which can tie: its keys are all distinct, so the order the reference leaves is simply
 * "sorted by key" whatever its permutations did, and any partition by digit will do.  One warp, one pass over the bucket: histogram
 * and scatter with shared-memory atomics (the arrival order inside a digit does not matter), then every digit group is sorted by
 * key in shared memory (<= 8: a lane each; 9..64: the warp's rank sort) and the bucket is written back once.  Groups of more than
 * 64 go on to the next level as usual.  Returns false (nothing written) when the bucket holds a tie mark. */
__device__ __forceinline__ bool af_small_untied(const AfArgs &a, uint32_t beg, uint32_t n, uint32_t lane, uint64_t *s_k, uint32_t *s_i,
                                                uint32_t *cnt, uint32_t *start, uint32_t *head)
{
    const uint64_t *kx = a.kx + beg; const uint32_t *idx = a.idx + beg;
    for (uint32_t d = lane; d < 256; d += 32) cnt[d] = 0;
    __syncwarp();
    uint32_t tied = 0;
    for (uint32_t p = lane; p < n; p += 32) { tied |= idx[p] >> 31; atomicAdd(&cnt[(uint32_t)(kx[p] >> a.shift) & 255u], 1u); }
    if (__any_sync(0xffffffffu, tied)) return false;
    __syncwarp();
    uint32_t loc = 0, cc[8];
    #pragma unroll
    for (int j = 0; j < 8; ++j) { cc[j] = cnt[8 * lane + j]; loc += cc[j]; }
    uint32_t run = lq_warp_incl_scan(loc) - loc;
    #pragma unroll
    for (int j = 0; j < 8; ++j) { start[8 * lane + j] = run; head[8 * lane + j] = run; run += cc[j]; }
    __syncwarp();
    for (uint32_t p = lane; p < n; p += 32) {
        const uint64_t k = kx[p];
        const uint32_t slot = atomicAdd(&head[(uint32_t)(k >> a.shift) & 255u], 1u);
        s_k[slot] = k; s_i[slot] = idx[p];
    }
    __syncwarp();
    uint32_t mid = 0;
    for (uint32_t d = lane; d < 256; d += 32) {
        const uint32_t c = cnt[d];
        const bool big_ = c > LQ_RS_MIN && a.shift > 0;
        af_append(a, beg + start[d], c, big_);
        if (big_ || c < 2) {}
        else if (c > 8 && c <= LQ_RS_MIN) mid |= 1u << (d >> 5);
        else lq_af_insertion_kv(s_k + start[d], s_i + start[d], c);
    }
    __syncwarp();
    for (uint32_t l = 0; l < 32; ++l) {
        uint32_t m = __shfl_sync(0xffffffffu, mid, l);
        while (m) { const uint32_t d = (uint32_t)(__ffs(m) - 1) * 32 + l; m &= m - 1; af_warp_ranksort(s_k + start[d], s_i + start[d], cnt[d], lane); }
    }
    uint64_t *kd = a.kx2 + beg; uint32_t *id = a.idx2 + beg;
    for (uint32_t p = lane; p < n; p += 32) { kd[p] = s_k[p]; id[p] = s_i[p]; }                 /* groups that go on are read from here */
    if (!a.dst_primary) { uint64_t *kp = a.kxp + beg; uint32_t *ip = a.idxp + beg; for (uint32_t p = lane; p < n; p += 32) { kp[p] = s_k[p]; ip[p] = s_i[p]; } }
    __syncwarp();
    return true;
}

__global__ void __launch_bounds__(AF_WARPS * 32) lq_af_level_k(AfArgs a)
{
    __shared__ uint32_t s_cnt[AF_WARPS][256], s_start[AF_WARPS][256], s_head[AF_WARPS][256];
    __shared__ uint64_t s_sk[AF_WARPS][AF_SN]; __shared__ uint32_t s_si[AF_WARPS][AF_SN];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, lt = (1u << lane) - 1;
    uint32_t *cnt = s_cnt[wid], *start = s_start[wid], *head = s_head[wid];
    const uint32_t nb_total = *a.n_cur;
    const uint32_t grab = nb_total > 200000u ? 16u : 1u;   /* many small buckets: fetch them 16 at a time (one same-address atomic per fetch) */
    uint32_t b = 0, b_end = 0;
    for (;;) {
        if (b >= b_end) {
            if (lane == 0) b = atomicAdd(a.cursor, grab);
            b = __shfl_sync(0xffffffffu, b, 0);
            b_end = b + grab < nb_total ? b + grab : nb_total;
            if (b >= nb_total) break;
        }
        const uint32_t bcur = b++;
        const uint32_t beg = a.cur[bcur].beg, n = a.cur[bcur].end - beg;
        uint32_t *idx = a.idx + beg, *idx2 = a.idx2 + beg, *dest = a.dest + beg;
        uint64_t *kx = a.kx + beg;
        uint8_t *dig = a.dig + beg;
        if (lane == 0) atomicAdd(a.n_elem, (unsigned long long)n);
        if (n <= AF_SN && af_small_untied(a, beg, n, lane, s_sk[wid], s_si[wid], cnt, start, head)) continue;
        /* 1. digits + histogram */
        for (uint32_t d = lane; d < 256; d += 32) cnt[d] = 0;
        __syncwarp();
        uint32_t tied = 0;
        for (uint32_t p0 = 0; p0 < n; p0 += 32 * AF_U) {   /* AF_U rows per trip: the idx -> key gathers of the rows overlap */
            uint32_t dv[AF_U]; uint64_t kk[AF_U];
            #pragma unroll
            for (int u = 0; u < AF_U; ++u) { const uint32_t p = p0 + u * 32 + lane; kk[u] = p < n ? kx[p] : 0; if (p < n) tied |= idx[p] >> 31; }
            #pragma unroll
            for (int u = 0; u < AF_U; ++u) dv[u] = (uint32_t)(kk[u] >> a.shift) & 255u;
            #pragma unroll
            for (int u = 0; u < AF_U; ++u) {
                const uint32_t p = p0 + u * 32 + lane; const bool ok = p < n;
                const uint32_t act = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    dig[p] = (uint8_t)dv[u];
                    const uint32_t peers = __match_any_sync(act, dv[u]);
                    if ((peers & lt) == 0) cnt[dv[u]] += __popc(peers);
                }
                __syncwarp();
            }
        }
        /* 2. region starts; lane owns digits 8*lane .. 8*lane+7 */
        uint32_t loc = 0, ne = 0;
        #pragma unroll
        for (int j = 0; j < 8; ++j) { const uint32_t c = cnt[8 * lane + j]; loc += c; if (c) ++ne; }
        uint32_t inc = lq_warp_incl_scan(loc), run = inc - loc;
        #pragma unroll
        for (int j = 0; j < 8; ++j) { start[8 * lane + j] = run; run += cnt[8 * lane + j]; }
        const uint32_t nb = lq_warp_sum(ne);
        tied = __any_sync(0xffffffffu, tied);
        __syncwarp();
        if (nb > 1 && !tied) {
            /* no two keys of this bucket are equal, so the sorted result does not depend on how the reference permutes:
             * stable counting partition, rows of 32 in order (head[] = running count per digit) */
            for (uint32_t d = lane; d < 256; d += 32) head[d] = 0;
            __syncwarp();
            for (uint32_t p0 = 0; p0 < n; p0 += 32 * AF_U) {
                uint32_t dv[AF_U];
                #pragma unroll
                for (int u = 0; u < AF_U; ++u) { const uint32_t p = p0 + u * 32 + lane; dv[u] = p < n ? dig[p] : 0; }
                #pragma unroll
                for (int u = 0; u < AF_U; ++u) {
                    const uint32_t p = p0 + u * 32 + lane; const bool ok = p < n;
                    const uint32_t act = __ballot_sync(0xffffffffu, ok);
                    uint32_t peers = 0;
                    if (ok) { peers = __match_any_sync(act, dv[u]); dest[p] = start[dv[u]] + head[dv[u]] + __popc(peers & lt); }
                    __syncwarp();
                    if (ok && (peers & lt) == 0) head[dv[u]] += __popc(peers);
                    __syncwarp();
                }
            }
        } else if (nb == 2) {
            /* closed form (lq_af_two_dest): d0 < d1 are the two non-empty digits */
            uint32_t d0 = 256, d1 = 0;
            #pragma unroll
            for (int j = 0; j < 8; ++j) if (cnt[8 * lane + j]) { d0 = min(d0, 8 * lane + j); d1 = max(d1, 8 * lane + j); }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) { d0 = min(d0, __shfl_xor_sync(0xffffffffu, d0, o)); d1 = max(d1, __shfl_xor_sync(0xffffffffu, d1, o)); }
            const uint32_t n0 = cnt[d0];
            uint32_t *P = idx2, *Z = idx2 + n0;
            uint32_t runP = 0, runZ = 0;
            for (uint32_t p0 = 0; p0 < n0; p0 += 32 * AF_U) {   /* region 0: foreign == digit d1 */
                uint32_t dv[AF_U];
                #pragma unroll
                for (int u = 0; u < AF_U; ++u) { const uint32_t p = p0 + u * 32 + lane; dv[u] = p < n0 ? dig[p] : 256u; }
                #pragma unroll
                for (int u = 0; u < AF_U; ++u) {
                    const uint32_t p = p0 + u * 32 + lane; const bool fr = dv[u] == d1;
                    const uint32_t m = __ballot_sync(0xffffffffu, fr), rk = runP + __popc(m & lt);
                    if (p < n0) dest[p] = rk;
                    if (fr) P[rk] = p;
                    runP += __popc(m);
                }
            }
            for (uint32_t p0 = n0; p0 < n; p0 += 32 * AF_U) {   /* region 1: foreign == digit d0 */
                uint32_t dv[AF_U];
                #pragma unroll
                for (int u = 0; u < AF_U; ++u) { const uint32_t p = p0 + u * 32 + lane; dv[u] = p < n ? dig[p] : 256u; }
                #pragma unroll
                for (int u = 0; u < AF_U; ++u) {
                    const uint32_t p = p0 + u * 32 + lane; const bool fr = dv[u] == d0;
                    const uint32_t m = __ballot_sync(0xffffffffu, fr), rk = runZ + __popc(m & lt);
                    if (p < n) dest[p] = rk;
                    if (fr) Z[rk] = p;
                    runZ += __popc(m);
                }
            }
            __syncwarp();
            for (uint32_t p0 = lane; p0 < n; p0 += 32 * AF_U) {
                uint32_t dv[AF_U], rk[AF_U];
                #pragma unroll
                for (int u = 0; u < AF_U; ++u) { const uint32_t p = p0 + u * 32; if (p < n) { dv[u] = dig[p]; rk[u] = dest[p]; } }
                #pragma unroll
                for (int u = 0; u < AF_U; ++u) {
                    const uint32_t p = p0 + u * 32;
                    if (p < n) { const int fr = p < n0 ? dv[u] == d1 : dv[u] == d0; dest[p] = lq_af_two_dest(p, n0, fr, rk[u], runP, P, Z); }
                }
            }
            __syncwarp();
        } else if (nb > 2) {
            /* tied keys and more than two digits: the sequential walk, done by lq_af_walk_k with a digit cache */
            uint32_t hi = 0;                                     /* any element with a digit above 15? */
            if (lane >= 2) {
                #pragma unroll
                for (int j = 0; j < 8; ++j) hi |= cnt[8 * lane + j];
            }
            hi = __any_sync(0xffffffffu, hi != 0);
            if (lane == 0) af_push_walk(a, beg, n, !hi);
            __syncwarp();
            continue;
        }
        af_finish_bucket<AF_SN>(a, beg, n, nb, cnt, start, dest, lane, s_sk[wid], s_si[wid]);
    }
}

/* ---- one level for a LONG bucket: a whole CTA.  Same rules as lq_af_level_k (digits, histogram, then the exact destination of
 *      every element), with block-wide primitives so that a 300 k-element bucket is not one warp's serial loop:
 *        no tied key in the bucket    any partition by digit gives the same final order (all keys distinct): counters per digit
 *        tied keys, two digits        the closed form lq_af_two_dest(), its ranks from block scans over tiles taken in order
 *        tied keys, more digits       handed to lq_af_walk_k (wlist), which also finishes the bucket ---- */
#define AFB_THREADS 512
#define AFB_V 4
/* sub-buckets of a finished level (ksort.h:124-133), in the destination buffer; what is finished here must end in the primary one */
__device__ __forceinline__ void afb_subbuckets(const AfArgs &a, uint32_t beg, uint32_t n, const uint32_t *s_cnt, const uint32_t *s_start)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint64_t *kx2 = a.kx2 + beg, *kp = a.kxp + beg; uint32_t *idx2 = a.idx2 + beg, *ip = a.idxp + beg;
    if (a.shift > 0) {
        if (tid < 256) {
            const uint32_t c = s_cnt[tid], s0 = s_start[tid];
            const bool big_ = c > LQ_RS_MIN;
            af_append(a, beg + s0, c, big_);
            if (c >= 1 && c <= 8) {
                if (c > 1) lq_af_insertion_kv(kx2 + s0, idx2 + s0, c);
                if (!a.dst_primary) for (uint32_t x = 0; x < c; ++x) { kp[s0 + x] = kx2[s0 + x]; ip[s0 + x] = idx2[s0 + x]; }
            }
        }
        for (uint32_t d = wid; d < 256; d += AFB_THREADS / 32) {   /* 9..64: a warp each, so that no thread sorts alone while 511 wait */
            const uint32_t c = s_cnt[d], s0 = s_start[d];
            if (c > 8 && c <= LQ_RS_MIN) af_warp_ranksort(kx2 + s0, idx2 + s0, c, lane, a.dst_primary ? (uint64_t*)0 : kp + s0, a.dst_primary ? (uint32_t*)0 : ip + s0);
        }
    } else if (!a.dst_primary) {
        #pragma unroll 4
        for (uint32_t p = tid; p < n; p += AFB_THREADS) { ip[p] = idx2[p]; kp[p] = kx2[p]; }
    }
}

static int g_walk_stats = getenv("LQCOV_WALK_STATS") ? atoi(getenv("LQCOV_WALK_STATS")) : 0;
/* CTAs per SM working on big buckets at a time: the passes over a bucket re-read it, so the buckets in flight should fit the L2 */
/* words of the two-digit levels' shared-memory bitmap (LQCOV_AFB_BITW=0: the levels go through the digit and destination arrays) */
static unsigned g_afb_bitw = getenv("LQCOV_AFB_BITW") ? (unsigned)atoi(getenv("LQCOV_AFB_BITW")) : 24576u;
static unsigned g_afb_ctas = getenv("LQCOV_AFB_CTAS") && atoi(getenv("LQCOV_AFB_CTAS")) > 0 ? (unsigned)atoi(getenv("LQCOV_AFB_CTAS")) : 4u;

extern __shared__ uint32_t afb_bits[];   /* a.bitw words: one bit per element of a two-digit bucket ("its digit is not the first element's") */

__global__ void __launch_bounds__(AFB_THREADS) lq_af_big_k(AfArgs a)
{
    __shared__ uint32_t s_cnt[256], s_start[257], s_head[256], scan_sm[33], s_wt[2][AFB_THREADS / 32];
    __shared__ uint32_t s_b, s_tied, s_nb, s_d0, s_d1, s_few;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, lt = (1u << lane) - 1;
    const uint32_t nbig = *a.n_curb;
    const AfBkt *list = a.curb;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_b = atomicAdd(a.cursorb, 1u);
        __syncthreads();
        const uint32_t b = s_b;
        if (b >= nbig) break;
        const uint32_t beg = list[b].beg, n = list[b].end - beg;
        uint32_t *idx = a.idx + beg, *idx2 = a.idx2 + beg, *dest = a.dest + beg;
        uint64_t *kx = a.kx + beg, *kx2 = a.kx2 + beg;
        uint8_t *dig = a.dig + beg;
        if (tid < 256) { s_cnt[tid] = 0; s_head[tid] = 0; }
        if (tid == 0) { s_tied = 0; atomicAdd(a.n_elem, (unsigned long long)n); }
        __syncthreads();
        /* 1. digits + histogram (warp-aggregated shared-memory atomics: a level may have only two digits).
         * On a two-digit level (a.two) the digits of a bucket that fits stay in shared memory as ONE BIT each and nothing else is
         * written: the bucket is then read once more and permuted by the closed form, whose ranks come straight from the bits
         * (keys 8 B + pairs 12 B in, 12 B out, instead of 13 + 14 + 28 through the digit and destination arrays). */
        const bool fused = a.two && n <= a.bitw * 32u;
        const uint32_t dA = fused ? (uint32_t)(kx[0] >> a.shift) & 255u : 0u;
        uint32_t tied = 0;
        for (uint32_t r0 = 0; r0 < n; r0 += AFB_THREADS * AFB_V) {
            uint32_t dv[AFB_V]; uint64_t kk[AFB_V];
            #pragma unroll
            for (int u = 0; u < AFB_V; ++u) { const uint32_t p = r0 + u * AFB_THREADS + tid; kk[u] = p < n ? kx[p] : 0; if (!fused && p < n) tied |= idx[p] >> 31; }
            #pragma unroll
            for (int u = 0; u < AFB_V; ++u) dv[u] = (uint32_t)(kk[u] >> a.shift) & 255u;
            #pragma unroll
            for (int u = 0; u < AFB_V; ++u) {
                const uint32_t p = r0 + u * AFB_THREADS + tid; const bool ok = p < n;
                const uint32_t act = __ballot_sync(0xffffffffu, ok);
                if (fused) { const uint32_t bits = __ballot_sync(0xffffffffu, ok && dv[u] != dA); if (lane == 0 && (p >> 5) < a.bitw) afb_bits[p >> 5] = bits; }
                if (ok) {
                    if (!fused) dig[p] = (uint8_t)dv[u];
                    const uint32_t peers = __match_any_sync(act, dv[u]);
                    if ((peers & lt) == 0) atomicAdd(&s_cnt[dv[u]], (uint32_t)__popc(peers));
                }
            }
        }
        if (tied) s_tied = 1;
        __syncthreads();
        /* 2. region starts */
        if (wid == 0) {
            uint32_t loc = 0, ne = 0, d0 = 256, d1 = 0, hi = 0;
            #pragma unroll
            for (int j = 0; j < 8; ++j) { const uint32_t c = s_cnt[8 * lane + j]; loc += c; if (c) { ++ne; d0 = min(d0, 8 * lane + j); d1 = max(d1, 8 * lane + j); if (lane >= 2) hi = 1; } }
            const uint32_t inc = lq_warp_incl_scan(loc);
            uint32_t run = inc - loc;
            #pragma unroll
            for (int j = 0; j < 8; ++j) { s_start[8 * lane + j] = run; run += s_cnt[8 * lane + j]; }
            ne = lq_warp_sum(ne);
            hi = __any_sync(0xffffffffu, hi != 0);
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) { d0 = min(d0, __shfl_xor_sync(0xffffffffu, d0, o)); d1 = max(d1, __shfl_xor_sync(0xffffffffu, d1, o)); }
            if (lane == 0) { s_start[256] = n; s_nb = ne; s_d0 = d0; s_d1 = d1; s_few = !hi; }
        }
        __syncthreads();
        const uint32_t nb = s_nb;
        if (fused && nb > 2) {   /* the level was not a two-digit one for this bucket after all: what the general forms need */
            for (uint32_t p = tid; p < n; p += AFB_THREADS) { dig[p] = (uint8_t)((uint32_t)(kx[p] >> a.shift) & 255u); tied |= idx[p] >> 31; }
            if (tied) s_tied = 1;
            __syncthreads();
        }
        const bool tiedb = s_tied != 0;
        if (fused && nb == 2) {
            /* 3f. two digits d0 < d1, closed form (lq_af_two_dest) for tied and untied buckets alike (any partition by digit serves
             * the latter).  b1 = "digit is d1"; foreign = b1 below n0, !b1 from n0 on.  P / Z = the foreign positions of the two regions. */
            const uint32_t d0 = s_d0, n0 = s_cnt[d0], inv = dA == d0 ? 0u : 0xffffffffu, nw = (n + 31) >> 5;
            uint32_t *P = dest, *Z = dest + n0;
            uint32_t runP = 0, runZ = 0;
            for (uint32_t w0 = 0; w0 < nw; w0 += AFB_THREADS) {                      /* the lists, from the bits alone */
                const uint32_t w = w0 + tid, pw = w << 5;
                uint32_t f0 = 0, f1 = 0;
                if (w < nw) {
                    const uint32_t b1 = afb_bits[w] ^ inv;
                    const uint32_t vm = n - pw >= 32 ? 0xffffffffu : (1u << (n - pw)) - 1u;              /* positions below n */
                    const uint32_t lo = pw >= n0 ? 0u : n0 - pw >= 32 ? 0xffffffffu : (1u << (n0 - pw)) - 1u;   /* positions below n0 */
                    f0 = b1 & lo & vm; f1 = ~b1 & ~lo & vm;
                }
                uint32_t tot;
                const uint32_t ex = lq_block_excl_scan<uint32_t>(__popc(f0) | __popc(f1) << 16, scan_sm, &tot);
                uint32_t r0 = runP + (ex & 0xffffu), r1 = runZ + (ex >> 16);
                while (f0) { P[r0++] = pw + (__ffs(f0) - 1); f0 &= f0 - 1; }
                while (f1) { Z[r1++] = pw + (__ffs(f1) - 1); f1 &= f1 - 1; }
                runP += tot & 0xffffu; runZ += tot >> 16;
            }
            __syncthreads();
            /* the permutation: a warp takes 32 AFB_V consecutive elements, 32 at a time (one word of the bitmap each) */
            uint32_t cumP = 0, cumZ = 0; int par = 0;
            for (uint32_t t0 = 0; t0 < n; t0 += AFB_THREADS * AFB_V, par ^= 1) {
                const uint32_t wbase = t0 + wid * 32 * AFB_V;
                uint32_t f0[AFB_V], f1[AFB_V], c0 = 0, c1 = 0; uint64_t kk[AFB_V]; uint32_t ii[AFB_V];
                #pragma unroll
                for (int u = 0; u < AFB_V; ++u) {
                    const uint32_t pw = wbase + u * 32, p = pw + lane;
                    f0[u] = 0; f1[u] = 0;
                    if (pw < n) {
                        const uint32_t b1 = afb_bits[pw >> 5] ^ inv;
                        const uint32_t vm = n - pw >= 32 ? 0xffffffffu : (1u << (n - pw)) - 1u;
                        const uint32_t lo = pw >= n0 ? 0u : n0 - pw >= 32 ? 0xffffffffu : (1u << (n0 - pw)) - 1u;
                        f0[u] = b1 & lo & vm; f1[u] = ~b1 & ~lo & vm;
                    }
                    if (p < n) { kk[u] = kx[p]; ii[u] = idx[p]; }
                    c0 += __popc(f0[u]); c1 += __popc(f1[u]);
                }
                if (lane == 0) s_wt[par][wid] = c0 | c1 << 16;
                __syncthreads();
                const uint32_t wv = lane < AFB_THREADS / 32 ? s_wt[par][lane] : 0u;
                const uint32_t before = lq_warp_sum(lane < wid ? wv : 0u), all = lq_warp_sum(wv);
                uint32_t r0 = cumP + (before & 0xffffu), r1 = cumZ + (before >> 16);
                #pragma unroll
                for (int u = 0; u < AFB_V; ++u) {
                    const uint32_t p = wbase + u * 32 + lane;
                    if (p < n) {
                        const bool low = p < n0;
                        const uint32_t fm = low ? f0[u] : f1[u];
                        const uint32_t rk = (low ? r0 : r1) + __popc(fm & lt);
                        const uint32_t d = lq_af_two_dest(p, n0, (fm >> lane) & 1u, rk, runP, P, Z);
                        idx2[d] = ii[u]; kx2[d] = kk[u];
                    }
                    r0 += __popc(f0[u]); r1 += __popc(f1[u]);
                }
                cumP += all & 0xffffu; cumZ += all >> 16;
            }
            __syncthreads();
            afb_subbuckets(a, beg, n, s_cnt, s_start);
            continue;
        }
        if (nb > 2 && tiedb) {   /* the sequential walk: lq_af_walk3_k / lq_af_walkf_k, then lq_af_place_k */
            if (tid == 0) af_push_walk(a, beg, n, s_few != 0);
            continue;
        }
        if (nb > 1) {
            if (tiedb) {
                /* 3a. two digits d0 < d1, tied keys: closed form.  rk = foreign positions before p in p's own region */
                const uint32_t d0 = s_d0, d1 = s_d1, n0 = s_cnt[d0];
                uint32_t *P = idx2, *Z = idx2 + n0;
                uint32_t runP = 0, runZ = 0;
                for (uint32_t t0 = 0; t0 < n; t0 += AFB_THREADS * AFB_V) {
                    const uint32_t pb = t0 + tid * AFB_V;
                    uint32_t fr[AFB_V], f0 = 0, f1 = 0;
                    #pragma unroll
                    for (int j = 0; j < AFB_V; ++j) {
                        const uint32_t p = pb + j;
                        fr[j] = 0;
                        if (p < n) { const uint32_t d = dig[p]; fr[j] = p < n0 ? d == d1 : d == d0; if (p < n0) f0 += fr[j]; else f1 += fr[j]; }
                    }
                    uint32_t tot;
                    const uint32_t ex = lq_block_excl_scan<uint32_t>(f0 | f1 << 16, scan_sm, &tot);
                    uint32_t r0 = runP + (ex & 0xffffu), r1 = runZ + (ex >> 16);
                    #pragma unroll
                    for (int j = 0; j < AFB_V; ++j) {
                        const uint32_t p = pb + j;
                        if (p < n) {
                            if (p < n0) { dest[p] = r0; if (fr[j]) P[r0++] = p; }
                            else { dest[p] = r1; if (fr[j]) Z[r1++] = p; }
                        }
                    }
                    runP += tot & 0xffffu; runZ += tot >> 16;
                }
                __syncthreads();
                for (uint32_t p = tid; p < n; p += AFB_THREADS) {
                    const uint32_t d = dig[p], rk = dest[p];
                    const int f = p < n0 ? d == d1 : d == d0;
                    dest[p] = lq_af_two_dest(p, n0, f, rk, runP, P, Z);
                }
                __syncthreads();
            } else {
                /* 3b. no tied key: any partition by digit */
                for (uint32_t r0 = 0; r0 < n; r0 += AFB_THREADS * AFB_V) {
                    uint32_t dv[AFB_V];
                    #pragma unroll
                    for (int u = 0; u < AFB_V; ++u) { const uint32_t p = r0 + u * AFB_THREADS + tid; dv[u] = p < n ? dig[p] : 0; }
                    #pragma unroll
                    for (int u = 0; u < AFB_V; ++u) {
                        const uint32_t p = r0 + u * AFB_THREADS + tid; const bool ok = p < n;
                        const uint32_t act = __ballot_sync(0xffffffffu, ok);
                        if (ok) {
                            const uint32_t peers = __match_any_sync(act, dv[u]);
                            const int leader = __ffs(peers) - 1;
                            uint32_t b0 = 0;
                            if ((int)lane == leader) b0 = atomicAdd(&s_head[dv[u]], (uint32_t)__popc(peers));
                            b0 = __shfl_sync(peers, b0, leader);
                            dest[p] = s_start[dv[u]] + b0 + __popc(peers & lt);
                        }
                    }
                }
                __syncthreads();
            }
            /* 4. permute the payload */
            #pragma unroll 4
            for (uint32_t p = tid; p < n; p += AFB_THREADS) { const uint32_t d = dest[p]; idx2[d] = idx[p]; kx2[d] = kx[p]; }
            __syncthreads();
        } else {   /* one digit: nothing moves, but the bucket changes buffers like every other */
            #pragma unroll 4
            for (uint32_t p = tid; p < n; p += AFB_THREADS) { idx2[p] = idx[p]; kx2[p] = kx[p]; }
            __syncthreads();
        }
        /* 5. sub-buckets */
        afb_subbuckets(a, beg, n, s_cnt, s_start);
    }
}

/* ---- the sequential walk of long tied buckets (>= AFW_SMALL elements, more than two digits), one LANE per bucket ----
 * lq_afsort_core.h: the only thing the walk must compute sequentially is the pick-up DIGIT stream; slots, source positions and the
 * payload permutation follow from it in parallel (lq_af_place_k).  A lane therefore writes one byte per pick-up (a 32-bit store
 * every four) plus a phase entry whenever the outer-loop region changes -- no positions, no per-step global traffic.
 * lq_af_walk3_k (any digits): a CTA takes up to AFS_WALKERS buckets at a time ("generation"):
 *   setup   the whole CTA builds the digit histograms (shared-memory atomics), a warp per bucket turns them into region starts (kept
 *           in a global scratch row) and empty per-region states
 *   rounds  all threads refill the 16-byte region states (position + next 11 digits) of every bucket, then the walker warps walk,
 *           lane = bucket, until every lane is done or out of digits somewhere (lq_afq_run: one LDS.128 per step on the critical path)
 * The states are laid out REGION-MAJOR, st[region][lane]: the lanes of a walker warp, each somewhere else in its own bucket, then read
 * consecutive 16-byte words.  (Bucket-major -- 4 KB between lanes -- put all 28 lanes on the same banks: every step of the warp was
 * 28 serialised shared-memory transactions.) */
#define AFS_WPW 28                 /* walker lanes per walker warp */
#define AFS_WW 2                   /* walker warps: they interleave on the SM, hiding each other's load latency */
#define AFS_WALKERS (AFS_WPW * AFS_WW)
#define AFS_THREADS 512
#define AFS_GRID 148               /* one CTA per SM (224 KB of shared memory each) */
#define AFS_ROW 260                /* u32 per bucket in the global scratch: 257 region starts */
#define AFS_LOW 10                 /* regions holding more cached digits than this are not refilled (5: twice the rounds, measured slower) */
__device__ __forceinline__ uint4 afw_load16(const uint8_t *addr)   /* 16 bytes from an arbitrary address (reads up to 31 bytes past it: the arena is padded) */
{
    const uintptr_t A = (uintptr_t)addr & ~(uintptr_t)15; const uint32_t o = (uint32_t)((uintptr_t)addr & 15), q = o >> 2, sh = (o & 3) * 8;
    const uint4 lo = *(const uint4*)A, hi = *(const uint4*)(A + 16);
    const uint32_t t0 = q == 0 ? lo.x : q == 1 ? lo.y : q == 2 ? lo.z : lo.w;
    const uint32_t t1 = q == 0 ? lo.y : q == 1 ? lo.z : q == 2 ? lo.w : hi.x;
    const uint32_t t2 = q == 0 ? lo.z : q == 1 ? lo.w : q == 2 ? hi.x : hi.y;
    const uint32_t t3 = q == 0 ? lo.w : q == 1 ? hi.x : q == 2 ? hi.y : hi.z;
    const uint32_t t4 = q == 0 ? hi.x : q == 1 ? hi.y : q == 2 ? hi.z : hi.w;
    return make_uint4(__funnelshift_r(t0, t1, sh), __funnelshift_r(t1, t2, sh), __funnelshift_r(t2, t3, sh), __funnelshift_r(t3, t4, sh));
}

struct AfsMeta { uint32_t beg, n; };

extern __shared__ __align__(16) uint8_t afs_smem[];

__global__ void __launch_bounds__(AFS_THREADS, 1) lq_af_walk3_k(AfArgs a, uint32_t *gstart /* gridDim.x * AFS_WALKERS * AFS_ROW */)
{
    lq_afp_st *state = (lq_afp_st*)afs_smem;                      /* [256][AFS_WALKERS] */
    __shared__ AfsMeta meta[AFS_WALKERS];
    __shared__ uint32_t s_base, s_alive[AFS_WW];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t nw = *a.n_wlist;
    uint32_t gsize = (nw + gridDim.x - 1) / gridDim.x;            /* buckets per generation: all CTAs busy, at most AFS_WALKERS */
    gsize = gsize < 1 ? 1 : gsize > AFS_WALKERS ? AFS_WALKERS : gsize;
    uint32_t *gs = gstart + (size_t)blockIdx.x * AFS_WALKERS * AFS_ROW;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_base = atomicAdd(a.wcursor, gsize);
        __syncthreads();
        const uint32_t base = s_base;
        if (base >= nw) break;
        const uint32_t nbk = nw - base < gsize ? nw - base : gsize;
        /* ---- setup: histograms by the whole CTA (bucket w counts in the first KB-s of the state area, bucket-major) ---- */
        uint32_t *cnts = (uint32_t*)afs_smem;
        for (uint32_t w = tid; w < nbk; w += AFS_THREADS) { const AfBkt b = a.wlist[base + w]; meta[w].beg = b.beg; meta[w].n = b.end - b.beg; }
        for (uint32_t e = tid; e < nbk * 256; e += AFS_THREADS) cnts[e] = 0;
        for (uint32_t e = tid; e < nbk * 256; e += AFS_THREADS) {   /* phase rows of these buckets: "never opened" */
            lq_afq_phase ph; ph.t = 0xffffffffu; ph.p = 0;
            a.wph[(size_t)(base + (e >> 8)) * 256 + (e & 255)] = ph;
        }
        __syncthreads();
        for (uint32_t w = 0; w < nbk; ++w) {
            const uint32_t n = meta[w].n;
            const uint8_t *dig = a.dig + meta[w].beg;
            uint32_t *cnt = cnts + w * 256;
            uint32_t head = (uint32_t)((16 - ((uintptr_t)dig & 15)) & 15);
            if (head > n) head = n;
            const uint32_t nch = (n - head) >> 4, tail0 = head + (nch << 4);
            if (tid < head) atomicAdd(&cnt[dig[tid]], 1u);
            if (tail0 + tid < n) atomicAdd(&cnt[dig[tail0 + tid]], 1u);
            const uint4 *body = (const uint4*)(dig + head);
            #pragma unroll 2
            for (uint32_t ch = tid; ch < nch; ch += AFS_THREADS) {
                const uint4 v = body[ch];
                const uint32_t ww[4] = { v.x, v.y, v.z, v.w };
                #pragma unroll
                for (int j = 0; j < 4; ++j) {
                    atomicAdd(&cnt[ww[j] & 255u], 1u); atomicAdd(&cnt[(ww[j] >> 8) & 255u], 1u);
                    atomicAdd(&cnt[(ww[j] >> 16) & 255u], 1u); atomicAdd(&cnt[ww[j] >> 24], 1u);
                }
            }
        }
        __syncthreads();
        /* ---- setup: region starts (a warp per bucket, up to four buckets per warp) -> registers; then, behind a barrier (the states
         *      overwrite the counts of OTHER buckets), the empty states ---- */
        uint32_t st0[4][8];
        #pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t w = wid + q * (AFS_THREADS / 32);
            if (w < nbk) {
                const uint32_t *cnt = cnts + w * 256;
                uint32_t loc = 0, cc[8];
                #pragma unroll
                for (int j = 0; j < 8; ++j) { cc[j] = cnt[8 * lane + j]; loc += cc[j]; }
                uint32_t run = lq_warp_incl_scan(loc) - loc;
                #pragma unroll
                for (int j = 0; j < 8; ++j) { st0[q][j] = run; gs[w * AFS_ROW + 8 * lane + j] = run; run += cc[j]; }
                if (lane == 31) gs[w * AFS_ROW + 256] = meta[w].n;
            }
        }
        __syncthreads();
        #pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t w = wid + q * (AFS_THREADS / 32);
            if (w < nbk) {
                #pragma unroll
                for (int j = 0; j < 8; ++j) { lq_afp_st e; e.x = st0[q][j]; e.y = e.z = e.w = 0; state[(8 * lane + j) * AFS_WALKERS + w] = e; }
            }
        }
        __syncthreads();
        /* ---- rounds: everybody refills the regions that moved, then the walker warps walk until every lane is done or out of digits ---- */
        const uint32_t wl = wid * AFS_WPW + lane;                 /* this thread's bucket when it is a walker lane */
        const bool walker = wid < AFS_WW && lane < AFS_WPW && wl < nbk;
        lq_afq_walk ws; bool fin = true; uint32_t my_n = 0; uint32_t *my_seq = 0; const uint32_t *my_start = gs; lq_afq_phase *my_ph = a.wph;
        if (walker) {
            my_start = gs + wl * AFS_ROW; my_ph = a.wph + (size_t)(base + wl) * 256;
            lq_afq_init(&ws, my_start, my_ph); fin = false; my_n = meta[wl].n; my_seq = a.ord + meta[wl].beg;
        }
        for (;;) {
            /* refill: consecutive threads, consecutive state words.  Only regions that are running low are topped up (a region that
             * still holds more than AFS_LOW digits would fetch the same sector again a round later: measured 10.5 GB of DRAM reads
             * for 0.32 GB of digits when every region that had moved was refilled).  All 16 word loads of a thread are issued before
             * the first is used. */
            for (uint32_t e0 = tid; e0 < AFS_WALKERS * 256; e0 += 4 * AFS_THREADS) {
                lq_afp_st S[4]; bool need[4]; uint32_t wv[4][4], sh[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t e = e0 + u * AFS_THREADS;
                    need[u] = false;
                    if (e < AFS_WALKERS * 256 && e % AFS_WALKERS < nbk) { S[u] = state[e]; need[u] = (S[u].w >> 24) <= AFS_LOW; }
                }
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint8_t *p = a.dig + meta[need[u] ? (e0 + u * AFS_THREADS) % AFS_WALKERS : 0].beg + (need[u] ? S[u].x : 0);
                    const uint32_t *q = (const uint32_t*)((uintptr_t)p & ~(uintptr_t)3);
                    sh[u] = (uint32_t)((uintptr_t)p & 3) * 8;
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) wv[u][j] = need[u] ? __ldg(q + j) : 0;   /* the arena is padded: 16 bytes from q stay inside it */
                }
                #pragma unroll
                for (int u = 0; u < 4; ++u) if (need[u]) {
                    S[u].y = __funnelshift_r(wv[u][0], wv[u][1], sh[u]); S[u].z = __funnelshift_r(wv[u][1], wv[u][2], sh[u]);
                    /* bit 7 of the count: the cached digits reach the region's end (lq_afp_refill_host) */
                    const uint32_t e = e0 + u * AFS_THREADS, end = gs[(e % AFS_WALKERS) * AFS_ROW + e / AFS_WALKERS + 1];
                    S[u].w = (__funnelshift_r(wv[u][2], wv[u][3], sh[u]) & 0x00ffffffu) | (uint32_t)(LQ_AFP_DIG | (S[u].x + LQ_AFP_DIG >= end ? 0x80u : 0u)) << 24;
                    state[e] = S[u];
                }
            }
            __syncthreads();
            if (wid < AFS_WW) {
                if (!fin) fin = lq_afq_run(&ws, my_n, my_start, state + wl, AFS_WALKERS, my_seq, my_ph) != 0;
                const uint32_t alive = __ballot_sync(0xffffffffu, !fin);
                if (lane == 0) s_alive[wid] = alive;
            }
            __syncthreads();
            uint32_t alive = 0;
            #pragma unroll
            for (int j = 0; j < AFS_WW; ++j) alive |= s_alive[j];
            if (!alive) break;
        }
        if (tid == 0) {
            unsigned long long tot = 0;
            for (uint32_t w = 0; w < nbk; ++w) tot += meta[w].n;
            atomicAdd(a.n_walk, nbk); atomicAdd(a.n_elem, tot);
        }
    }
}

/* lq_af_walkf_k: the same for buckets all of whose digits are below 16 (lq_afr_*: the rid >> 16 byte once a part holds more than
 * 131 072 reads -- every multi-GPU run, and every 4 G-base part of reads shorter than 30 kb).  Such walks are few (one per tied query and
 * strand) and long (10^5..10^6 pick-ups), so they are spread thinly: AFF_LANES buckets per CTA, three CTAs per SM.  A region caches
 * 392 digits in shared memory ([region][word][lane]: window, queue cursor, queue origin, 57 queue words of 7 four-bit digits); a
 * pick-up is one shared-memory load on the critical path (lq_afsort_core.h, lq_afr_run). */
#define AFF_LANES 16
#define AFF_THREADS 256
#define AFF_GRID (148 * 3)
/* explicit shared-space word access for lq_afr_run: no generic-address conversion inside the walk loop */
struct LqSmemWords {
    uint32_t base;   /* shared-space byte address of this lane's word 0 */
    __device__ __forceinline__ uint32_t ld(uint32_t i) const { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + 4u * i) : "memory"); return v; }
    __device__ __forceinline__ void st(uint32_t i, uint32_t v) const { asm volatile("st.shared.u32 [%0], %1;" :: "r"(base + 4u * i), "r"(v) : "memory"); }
};
struct AffSmem { uint32_t blk[LQ_AFR_R * LQ_AFR_BLK][AFF_LANES]; uint32_t start[LQ_AFR_R + 1][AFF_LANES]; };   /* blk: per region window, queue cursor, queue origin, 57 queue words */

__global__ void __launch_bounds__(AFF_THREADS) lq_af_walkf_k(AfArgs a)
{
    extern __shared__ __align__(16) uint8_t aff_raw[];
    AffSmem &S = *(AffSmem*)aff_raw;
    __shared__ AfsMeta meta[AFF_LANES];
    __shared__ uint32_t s_base, s_alive, s_cnt[AFF_LANES][LQ_AFR_R];
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t nw = *a.n_wlist_w;
    uint32_t gsize = (nw + gridDim.x - 1) / gridDim.x;
    gsize = gsize < 1 ? 1 : gsize > AFF_LANES ? AFF_LANES : gsize;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_base = atomicAdd(a.wcursor_w, gsize);
        __syncthreads();
        const uint32_t base = s_base;
        if (base >= nw) break;
        const uint32_t nbk = nw - base < gsize ? nw - base : gsize;
        if (tid < nbk) { const AfBkt b = a.wlist_w[base + tid]; meta[tid].beg = b.beg; meta[tid].n = b.end - b.beg; }
        for (uint32_t e = tid; e < AFF_LANES * LQ_AFR_R; e += AFF_THREADS) s_cnt[e / LQ_AFR_R][e % LQ_AFR_R] = 0;
        for (uint32_t e = tid; e < nbk * 256; e += AFF_THREADS) {
            lq_afq_phase ph; ph.t = 0xffffffffu; ph.p = 0;
            a.wph[(size_t)(a.wph_cap + base + (e >> 8)) * 256 + (e & 255)] = ph;
        }
        __syncthreads();
        /* histograms: 16 counters per bucket, private per thread, then merged */
        for (uint32_t w = 0; w < nbk; ++w) {
            const uint32_t n = meta[w].n; const uint8_t *dig = a.dig + meta[w].beg;
            uint32_t c[LQ_AFR_R];
            #pragma unroll
            for (int r = 0; r < LQ_AFR_R; ++r) c[r] = 0;
            for (uint32_t p = tid; p < n; p += AFF_THREADS) {
                const uint32_t d = dig[p] & 15u;
                #pragma unroll
                for (int r = 0; r < LQ_AFR_R; ++r) c[r] += d == (uint32_t)r;
            }
            #pragma unroll
            for (int r = 0; r < LQ_AFR_R; ++r) { const uint32_t v = lq_warp_sum(c[r]); if (lane == 0 && v) atomicAdd(&s_cnt[w][r], v); }
        }
        __syncthreads();
        if (tid < nbk) {
            uint32_t run = 0;
            for (int r = 0; r < LQ_AFR_R; ++r) {   /* nothing consumed, queues empty (lq_afr_cache_init_host) */
                S.start[r][tid] = run; S.blk[r * LQ_AFR_BLK][tid] = 0; S.blk[r * LQ_AFR_BLK + 1][tid] = 0; S.blk[r * LQ_AFR_BLK + 2][tid] = run;
                run += s_cnt[tid][r];
            }
            S.start[LQ_AFR_R][tid] = run;
        }
        __syncthreads();
        const bool walker = wid == 0 && lane < nbk;
        lq_afr_walk ws; bool fin = true; uint32_t my_n = 0; uint32_t *my_seq = 0; lq_afq_phase *my_ph = a.wph; uint32_t my_start[LQ_AFR_R + 1];
        #pragma unroll
        for (int r = 0; r <= LQ_AFR_R; ++r) my_start[r] = 0;
        if (walker) {
            #pragma unroll
            for (int r = 0; r <= LQ_AFR_R; ++r) my_start[r] = S.start[r][lane];
            my_ph = a.wph + (size_t)(a.wph_cap + base + lane) * 256;
            lq_afr_init(&ws, my_start, my_ph); fin = false; my_n = meta[lane].n; my_seq = a.ord + meta[lane].beg;
        }
        for (;;) {
            /* refill (lq_afr_refill_host): every region's queue restarts at its next unread position */
            for (uint32_t e = tid; e < nbk * LQ_AFR_R; e += AFF_THREADS) {
                const uint32_t w = e % nbk, r = e / nbk;
                S.blk[r * LQ_AFR_BLK + 2][w] = lq_afr_next(S.blk[r * LQ_AFR_BLK][w], S.blk[r * LQ_AFR_BLK + 1][w], S.blk[r * LQ_AFR_BLK + 2][w], S.start[r + 1][w]);
            }
            __syncthreads();
            for (uint32_t e0 = tid; e0 < nbk * LQ_AFR_R * (LQ_AFR_QW + 1); e0 += 4 * AFF_THREADS) {   /* 8 word loads in flight per thread */
                uint32_t lo[4], hi[4], th[4], sh[4]; bool fast[4];
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t e = e0 + u * AFF_THREADS;
                    fast[u] = false; lo[u] = hi[u] = th[u] = sh[u] = 0;
                    if (e < nbk * LQ_AFR_R * (LQ_AFR_QW + 1)) {
                        const uint32_t w = e % nbk, rw = e / nbk, r = rw / (LQ_AFR_QW + 1), j = rw % (LQ_AFR_QW + 1);
                        const uint32_t end = S.start[r + 1][w], p = S.blk[r * LQ_AFR_BLK + 2][w] + 7 * j;
                        const uint8_t *dig = a.dig + meta[w].beg;
                        if (j < LQ_AFR_QW && p + 7 <= end) {
                            const uint32_t *q = (const uint32_t*)((uintptr_t)(dig + p) & ~(uintptr_t)3);
                            sh[u] = (uint32_t)((uintptr_t)(dig + p) & 3) * 8; fast[u] = true;
                            lo[u] = __ldg(q); hi[u] = __ldg(q + 1); th[u] = __ldg(q + 2);   /* seven bytes from offset 0..3: three aligned words (the arena is padded) */
                        }
                    }
                }
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t e = e0 + u * AFF_THREADS;
                    if (e < nbk * LQ_AFR_R * (LQ_AFR_QW + 1)) {
                        const uint32_t w = e % nbk, rw = e / nbk, r = rw / (LQ_AFR_QW + 1), j = rw % (LQ_AFR_QW + 1);
                        uint32_t v;
                        if (fast[u]) {   /* seven bytes from an arbitrary address -> seven nibbles under the marker */
                            const uint32_t b0 = __funnelshift_r(lo[u], hi[u], sh[u]), b1 = __funnelshift_r(hi[u], th[u], sh[u]);
                            v = (b0 & 15u) | ((b0 >> 4) & 0xf0u) | ((b0 >> 8) & 0xf00u) | ((b0 >> 12) & 0xf000u)
                              | ((b1 & 15u) << 16) | (((b1 >> 8) & 15u) << 20) | (((b1 >> 16) & 15u) << 24) | (1u << 28);
                        } else v = lq_afr_queue_word(a.dig + meta[w].beg, S.blk[r * LQ_AFR_BLK + 2][w], S.start[r + 1][w], j);
                        S.blk[r * LQ_AFR_BLK + 3 + j][w] = v;
                    }
                }
            }
            __syncthreads();
            for (uint32_t e = tid; e < nbk * LQ_AFR_R; e += AFF_THREADS) {
                const uint32_t w = e % nbk, r = e / nbk;
                S.blk[r * LQ_AFR_BLK][w] = S.blk[r * LQ_AFR_BLK + 3][w]; S.blk[r * LQ_AFR_BLK + 1][w] = 1;
            }
            __syncthreads();
            if (wid == 0) {
                if (!fin) { LqSmemWords mw; mw.base = (uint32_t)__cvta_generic_to_shared(&S.blk[0][lane]); fin = lq_afr_run(&ws, my_n, my_start, mw, AFF_LANES, my_seq, my_ph) != 0; }
                const uint32_t alive = __ballot_sync(0xffffffffu, !fin);
                if (lane == 0) s_alive = alive;
            }
            __syncthreads();
            if (!s_alive) break;
        }
        if (tid == 0) {
            unsigned long long tot = 0;
            for (uint32_t w = 0; w < nbk; ++w) tot += meta[w].n;
            atomicAdd(a.n_walk, nbk); atomicAdd(a.n_elem, tot);
        }
    }
}

/* lq_af_place_k: everything about a walked bucket that is NOT sequential (lq_afq_expand), a CTA per bucket at full occupancy:
 *   slot[t] = start[d_t] + (digit-d_t pick-ups before t)        stable ranking of the digit stream, tiles of 4096 pick-ups
 *   ord[t]  = slot[t-1] (+1 behind a pick-up that closed a cycle of the outer-loop region), or the recorded position of a phase
 *   payload[slot[t]] = payload[ord[t]]                          in pick-up order: both sides advance sequentially inside each region
 * then the sub-buckets are handed on like after any other level. */
#define AFP_TILE 4096
#define AFP_ROWS (AFP_TILE / AFB_THREADS)          /* 8 rows of 32 per warp */
__global__ void __launch_bounds__(AFB_THREADS) lq_af_place_k(AfArgs a)
{
    __shared__ uint32_t s_cnt[256], s_start[257], s_run[256], s_slot[AFP_TILE];
    __shared__ uint32_t s_w[AFB_THREADS / 32][256];
    __shared__ uint32_t s_pt[256], s_pp[256], s_pk[256], s_scan[8];
    __shared__ uint32_t s_b, s_np, s_carry_slot, s_carry_d;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, lt = (1u << lane) - 1;
    const uint32_t n1 = *a.n_wlist, n2 = *a.n_wlist_w;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_b = atomicAdd(a.wcursor_f, 1u);
        __syncthreads();
        const uint32_t b = s_b;
        if (b >= n1 + n2) break;
        const AfBkt bk = b < n1 ? a.wlist[b] : a.wlist_w[b - n1];
        const lq_afq_phase *ph = a.wph + (size_t)(b < n1 ? b : a.wph_cap + (b - n1)) * 256;
        const uint32_t beg = bk.beg, n = bk.end - beg;
        const uint32_t *idx = a.idx + beg; uint32_t *idx2 = a.idx2 + beg;
        const uint64_t *kx = a.kx + beg; uint64_t *kx2 = a.kx2 + beg;
        const uint8_t *seq = (const uint8_t*)(a.ord + beg);
        if (tid < 256) { s_cnt[tid] = 0; s_run[tid] = 0; }
        __syncthreads();
        /* region sizes: histogram of the digit stream (the same multiset as the bucket's digits) */
        for (uint32_t r0 = 0; r0 < n; r0 += AFB_THREADS * AFB_V) {
            uint32_t dv[AFB_V];
            #pragma unroll
            for (int u = 0; u < AFB_V; ++u) { const uint32_t p = r0 + u * AFB_THREADS + tid; dv[u] = p < n ? seq[p] : 0; }
            #pragma unroll
            for (int u = 0; u < AFB_V; ++u) {
                const uint32_t p = r0 + u * AFB_THREADS + tid; const bool ok = p < n;
                const uint32_t act = __ballot_sync(0xffffffffu, ok);
                if (ok) { const uint32_t peers = __match_any_sync(act, dv[u]); if ((peers & lt) == 0) atomicAdd(&s_cnt[dv[u]], (uint32_t)__popc(peers)); }
            }
        }
        /* the phases that happened, in time order (== ascending region) */
        if (tid < 256) {
            const lq_afq_phase p = ph[tid];
            const bool ok = p.t != 0xffffffffu;
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) s_scan[wid] = __popc(m);
            __syncwarp();
            /* finished below, behind the barrier */
            if (ok) { s_pk[tid] = __popc(m & lt); s_pt[tid] = p.t; s_pp[tid] = p.p; } else s_pk[tid] = 0xffffffffu;
        }
        __syncthreads();
        uint32_t my_pt = 0, my_pp = 0, my_at = 0xffffffffu;
        if (tid < 256 && s_pk[tid] != 0xffffffffu) {
            uint32_t basew = 0;
            for (uint32_t w = 0; w < wid; ++w) basew += s_scan[w];
            my_at = basew + s_pk[tid]; my_pt = s_pt[tid]; my_pp = s_pp[tid];
        }
        if (wid == 0) {
            uint32_t loc = 0;
            #pragma unroll
            for (int j = 0; j < 8; ++j) loc += s_cnt[8 * lane + j];
            uint32_t run = lq_warp_incl_scan(loc) - loc;
            #pragma unroll
            for (int j = 0; j < 8; ++j) { s_start[8 * lane + j] = run; run += s_cnt[8 * lane + j]; }
            if (lane == 31) s_start[256] = n;
        }
        __syncthreads();
        if (tid == 0) { uint32_t np = 0; for (uint32_t w = 0; w < 8; ++w) np += s_scan[w]; s_np = np; s_carry_slot = 0; s_carry_d = 0; }
        if (my_at != 0xffffffffu) { s_pt[my_at] = my_pt; s_pp[my_at] = my_pp; s_pk[my_at] = tid; }
        __syncthreads();
        const uint32_t np = s_np;
        for (uint32_t t0 = 0; t0 < n; t0 += AFP_TILE) {
            const uint32_t wbase = t0 + wid * (AFP_ROWS * 32);
            /* a. rank of every pick-up among the equal digits of its warp's stretch (rows in order), and the stretch's digit counts */
            #pragma unroll
            for (int j = 0; j < 8; ++j) s_w[wid][lane + 32 * j] = 0;
            __syncwarp();
            uint32_t dv[AFP_ROWS], lr[AFP_ROWS];
            #pragma unroll
            for (int r = 0; r < AFP_ROWS; ++r) { const uint32_t t = wbase + r * 32 + lane; dv[r] = t < n ? seq[t] : 0; }
            #pragma unroll
            for (int r = 0; r < AFP_ROWS; ++r) {
                const uint32_t t = wbase + r * 32 + lane; const bool ok = t < n;
                const uint32_t act = __ballot_sync(0xffffffffu, ok);
                uint32_t peers = 0; lr[r] = 0;
                if (ok) { peers = __match_any_sync(act, dv[r]); lr[r] = s_w[wid][dv[r]] + __popc(peers & lt); }
                __syncwarp();
                if (ok && (peers & lt) == 0) s_w[wid][dv[r]] += __popc(peers);
                __syncwarp();
            }
            __syncthreads();
            /* b. digit d: exclusive prefix over the warps; the tile's total is added to the running count afterwards */
            uint32_t tot_d = 0;
            if (tid < 256) {
                #pragma unroll
                for (int w = 0; w < AFB_THREADS / 32; ++w) { const uint32_t c = s_w[w][tid]; s_w[w][tid] = tot_d; tot_d += c; }
            }
            __syncthreads();
            /* c. slots */
            #pragma unroll
            for (int r = 0; r < AFP_ROWS; ++r) {
                const uint32_t t = wbase + r * 32 + lane;
                if (t < n) s_slot[t - t0] = s_start[dv[r]] + s_run[dv[r]] + s_w[wid][dv[r]] + lr[r];
            }
            __syncthreads();
            /* d. source positions, payload */
            #pragma unroll
            for (int r = 0; r < AFP_ROWS; ++r) {
                const uint32_t t = wbase + r * 32 + lane;
                if (t < n) {
                    /* phase in force at t: the last one with start <= t */
                    uint32_t lo = 0, hi = np;
                    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_pt[mid] <= t) lo = mid; else hi = mid; }
                    uint32_t ord;
                    if (s_pt[lo] == t) ord = s_pp[lo];
                    else {
                        const uint32_t ps = t == t0 ? s_carry_slot : s_slot[t - 1 - t0];
                        const uint32_t pdd = t == t0 ? s_carry_d : (uint32_t)seq[t - 1];
                        /* the phase in force at t-1 is `lo` as well: a phase that opens at t is the case above */
                        ord = ps + (pdd == s_pk[lo] ? 1u : 0u);
                    }
                    const uint32_t sl = s_slot[t - t0];
                    idx2[sl] = idx[ord]; kx2[sl] = kx[ord];
                }
            }
            __syncthreads();
            if (tid < 256) s_run[tid] += tot_d;
            if (tid == 0) { const uint32_t last = (t0 + AFP_TILE <= n ? AFP_TILE : n - t0) - 1; s_carry_slot = s_slot[last]; s_carry_d = seq[t0 + last]; }
            __syncthreads();
        }
        /* sub-buckets */
        afb_subbuckets(a, beg, n, s_cnt, s_start);
    }
}

/* Short walks (the tied sub-buckets of the lower levels: ~10^2 elements, ~10^5 of them): a warp per bucket, lane 0 chases the pointer
 * (lq_afw_run).  Per walk 7 KB of shared memory: region starts, {pos, base} per region, 16 cached digits per region; a step is one
 * 8-byte and one 1-byte shared-memory load; when a region's cached digits run out the whole warp refills every region that moved. */
#define AFW_WARPS 2
#define AFW_SN 320                 /* staging capacity of the finish phase inside the 4 KB digit cache: 320 keys + 320 indices */
struct AfwSmem { uint32_t start[260]; lq_afw_pb pb[256]; uint4 cache[256]; };

__global__ void __launch_bounds__(AFW_WARPS * 32) lq_af_walk_small_k(AfArgs a)
{
    __shared__ __align__(16) AfwSmem s_w[AFW_WARPS];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, lt = (1u << lane) - 1;
    AfwSmem &S = s_w[wid];
    uint32_t *start = S.start, *cnt = (uint32_t*)S.pb;   /* cnt[] shares the pb[] storage: histogram before the walk, sizes after it */
    const uint32_t nw = *a.n_wlist_s;
    for (;;) {
        uint32_t b = 0;
        if (lane == 0) b = atomicAdd(a.wcursor_s, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= nw) break;
        const uint32_t beg = a.wlist_s[b].beg, n = a.wlist_s[b].end - beg;
        uint32_t *dest = a.dest + beg;
        const uint8_t *dig = a.dig + beg;
        for (uint32_t d = lane; d < 256; d += 32) cnt[d] = 0;
        __syncwarp();
        for (uint32_t p0 = 0; p0 < n; p0 += 32 * AF_U) {   /* histogram from the digits the level kernel stored */
            uint32_t dv[AF_U];
            #pragma unroll
            for (int u = 0; u < AF_U; ++u) { const uint32_t p = p0 + u * 32 + lane; dv[u] = p < n ? dig[p] : 0; }
            #pragma unroll
            for (int u = 0; u < AF_U; ++u) {
                const uint32_t p = p0 + u * 32 + lane; const bool ok = p < n;
                const uint32_t act = __ballot_sync(0xffffffffu, ok);
                if (ok) { const uint32_t peers = __match_any_sync(act, dv[u]); if ((peers & lt) == 0) cnt[dv[u]] += __popc(peers); }
                __syncwarp();
            }
        }
        uint32_t loc = 0, ne = 0, cc[8];
        #pragma unroll
        for (int j = 0; j < 8; ++j) { cc[j] = cnt[8 * lane + j]; loc += cc[j]; if (cc[j]) ++ne; }
        uint32_t inc = lq_warp_incl_scan(loc), run = inc - loc;
        #pragma unroll
        for (int j = 0; j < 8; ++j) { start[8 * lane + j] = run; run += cc[j]; }
        if (lane == 31) start[256] = n;
        const uint32_t nb = lq_warp_sum(ne);
        __syncwarp();                                    /* every lane has read its counts: pb[] may overwrite them */
        for (uint32_t r = lane; r < 256; r += 32) { lq_afw_pb e; e.x = start[r]; e.y = start[r] - LQ_AFW_CACHE; S.pb[r] = e; }
        __syncwarp();
        lq_afw_state ws;
        lq_afw_init_state(&ws, start);
        for (;;) {
            #pragma unroll
            for (int i = 0; i < 8; ++i) {                /* refill every region that moved since its last refill */
                const uint32_t r = lane + 32 * i;
                const lq_afw_pb e = S.pb[r];
                if (e.x < start[r + 1] && e.x != e.y) { S.cache[r] = afw_load16(dig + e.x); S.pb[r].y = e.x; }
            }
            __syncwarp();
            int done = 0;
            if (lane == 0) done = lq_afw_run(&ws, n, start, S.pb, (const uint8_t*)S.cache, dest);
            done = __shfl_sync(0xffffffffu, done, 0);
            if (done) break;
        }
        if (lane == 0) { atomicAdd(a.n_walk, 1u); atomicAdd(a.n_elem, (unsigned long long)n); }
        __syncwarp();
        for (uint32_t d = lane; d < 256; d += 32) cnt[d] = start[d + 1] - start[d];
        __syncwarp();
        af_finish_bucket<AFW_SN>(a, beg, n, nb, cnt, start, dest, lane, (uint64_t*)S.cache, (uint32_t*)S.cache + 2 * AFW_SN);
    }
}

/* the keys are sorted in place (kx == ax); the rest of a seed follows through its number */
__global__ void lq_gather_k(uint64_t n, const uint32_t *__restrict__ idx, SeedArrays s, uint32_t *__restrict__ aq, uint32_t *__restrict__ am)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = idx[i] & 0x7fffffffu;
    aq[i] = s.sq[j] & 0x7fffffffu; am[i] = s.sm[j];
}

/* ------------------------------------------------------------------ K6/K7: groups, chaining, accounting */

/* head[i] = 1 when sorted seed i starts a (query, strand, target) run */
/* The runs that can hold a chain, straight from the sorted keys.  A run = maximal stretch of one query's seeds with the same
 * (strand, target) = the high word of the key.  A chain needs >= min_cnt anchors (chain.c:116-119) and its score cannot exceed the
 * sum of its anchors' spans (chain.c:58: each link adds at most q_span), so only runs of T = max(min_cnt, ceil(min_sc / max_span))
 * or more anchors matter -- ~1 % of them; the others are never listed.  Thread i: is seed i the first of its run (key differs from
 * seed i-1; a query's first seed is handled by lq_runs_q_k) and does the run reach i+T-1?  Only then the run's end is located
 * (galloping + binary search, clipped to the query) and the run appended: runs of <= CH_SMALL anchors to the front of `runs` (one
 * thread each, lq_chain_small_k), longer ones to the back (a warp each, lq_chain_k). */
#define CH_SMALL 16
struct RunArgs {
    const uint64_t *ax; uint64_t n; const uint64_t *qoff; uint32_t nqb; uint32_t thr;
    uint2 *runs; uint32_t cap; uint32_t *n_small, *n_large; unsigned long long *n_heads;
    uint32_t *cand, *wcount; uint64_t chunk; uint32_t cand_stride;   /* run starts that reach the threshold, listed per scanning warp (lq_runs_k) */
};
__device__ __forceinline__ uint32_t run_query_of(const RunArgs &a, uint64_t i)   /* batch-relative query owning seed i */
{
    uint32_t lo = 0, hi = a.nqb;
    while (hi - lo > 1) { const uint32_t mid = lo + ((hi - lo) >> 1); if (a.qoff[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}
__device__ __forceinline__ uint64_t run_end(const RunArgs &a, uint64_t i, uint64_t lim /* end of the query */)
{
    const uint32_t key = (uint32_t)(a.ax[i] >> 32);
    uint64_t lo = i + a.thr - 1, step = 1, hi;          /* lo: known inside the run */
    for (;;) { hi = lo + step; if (hi >= lim) { hi = lim; break; } if ((uint32_t)(a.ax[hi] >> 32) != key) break; lo = hi; step <<= 1; }
    while (hi - lo > 1) { const uint64_t mid = lo + ((hi - lo) >> 1); if ((uint32_t)(a.ax[mid] >> 32) == key) lo = mid; else hi = mid; }
    return hi;
}
/* append the run [i, hi) (hi == 0: nothing); called by whole warps: one counter update per warp and list */
__device__ __forceinline__ void run_append(const RunArgs &a, uint64_t i, uint64_t hi)
{
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1;
    const bool sm = hi != 0 && hi - i <= CH_SMALL, lg = hi != 0 && hi - i > CH_SMALL;
    const uint32_t ms = __ballot_sync(0xffffffffu, sm), ml = __ballot_sync(0xffffffffu, lg);
    uint32_t bs = 0, bl = 0;
    if (lane == 0) { if (ms) bs = atomicAdd(a.n_small, (uint32_t)__popc(ms)); if (ml) bl = atomicAdd(a.n_large, (uint32_t)__popc(ml)); }
    bs = __shfl_sync(0xffffffffu, bs, 0); bl = __shfl_sync(0xffffffffu, bl, 0);
    if (sm) a.runs[bs + __popc(ms & lt)] = make_uint2((uint32_t)i, (uint32_t)hi);
    if (lg) a.runs[a.cap - 1 - (bl + __popc(ml & lt))] = make_uint2((uint32_t)i, (uint32_t)hi);
}
#define RUN_GRID (148 * 8)
#define RUN_THREADS 256
#define RUN_WARPS (RUN_GRID * RUN_THREADS / 32)
/* warp gw scans the contiguous chunk [gw*chunk, (gw+1)*chunk) of the seeds and lists its candidates in its own stretch of `cand`
 * (at most chunk/thr + 1 of them: candidate runs are disjoint and >= thr long), so that listing takes no atomic at all */
__global__ void __launch_bounds__(RUN_THREADS) lq_runs_k(RunArgs a)
{
    const uint32_t lane = threadIdx.x & 31, gw = (blockIdx.x * RUN_THREADS + threadIdx.x) >> 5;
    const uint64_t beg = (uint64_t)gw * a.chunk, end = beg + a.chunk < a.n ? beg + a.chunk : a.n;
    uint32_t *mine = a.cand + (size_t)gw * a.cand_stride;
    uint32_t my_heads = 0, cnt = 0;
    for (uint64_t i0 = beg; i0 < end; i0 += 32) {       /* a.chunk is a multiple of 32 */
        const uint64_t i = i0 + lane;
        bool head = false, cand = false;
        if (i < end) {
            const uint32_t key = (uint32_t)(a.ax[i] >> 32);
            head = i > 0 && (uint32_t)(a.ax[i - 1] >> 32) != key;
            cand = head && i + a.thr - 1 < a.n && (uint32_t)(a.ax[i + a.thr - 1] >> 32) == key;
        }
        my_heads += head;
        const uint32_t m = __ballot_sync(0xffffffffu, cand);
        if (cand) mine[cnt + __popc(m & ((1u << lane) - 1))] = (uint32_t)i;
        cnt += __popc(m);
    }
    if (lane == 0) a.wcount[gw] = cnt;
    my_heads = lq_warp_sum(my_heads);                    /* one counter update per warp of the whole grid, not per 32 seeds */
    if (lane == 0 && my_heads) atomicAdd(a.n_heads, (unsigned long long)my_heads);
}
/* locating a run's end is serial pointer chasing: a thread per candidate, so that no lane waits on another's search */
__global__ void __launch_bounds__(RUN_THREADS) lq_runs_emit_k(RunArgs a)
{
    const uint32_t lane = threadIdx.x & 31, gw = (blockIdx.x * RUN_THREADS + threadIdx.x) >> 5;
    const uint32_t *mine = a.cand + (size_t)gw * a.cand_stride;
    const uint32_t nc = a.wcount[gw];
    for (uint32_t c0 = 0; c0 < nc; c0 += 32) {
        uint64_t i = 0, hi = 0;
        if (c0 + lane < nc) {
            i = mine[c0 + lane];
            const uint32_t q = run_query_of(a, i);
            if (a.qoff[q] != i && i + a.thr - 1 < a.qoff[q + 1]) hi = run_end(a, i, a.qoff[q + 1]);   /* a query's first seed: lq_runs_q_k */
        }
        run_append(a, i, hi);
    }
}
/* the run that starts a query */
__global__ void lq_runs_q_k(RunArgs a)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;      /* the grid is a whole number of warps */
    uint64_t i = 0, hi = 0;
    if (q < a.nqb) {
        i = a.qoff[q];
        const uint64_t lim = a.qoff[q + 1];
        if (i < lim) {
            if (i == 0 || (uint32_t)(a.ax[i - 1] >> 32) == (uint32_t)(a.ax[i] >> 32)) atomicAdd(a.n_heads, 1ULL);   /* not counted by lq_runs_k */
            if (i + a.thr - 1 < lim && (uint32_t)(a.ax[i + a.thr - 1] >> 32) == (uint32_t)(a.ax[i] >> 32)) hi = run_end(a, i, lim);
        }
    }
    run_append(a, i, hi);
}

struct ChainArgs {
    const uint64_t *ax; const uint32_t *aq, *am;
    int32_t *f, *p, *v, *t; uint64_t *uend; uint32_t *vl, *s_lo, *s_hi;
    const uint2 *runs; uint32_t big_cap; const uint32_t *n_small, *n_groups; uint32_t *cursor;   /* lq_runs_k: short runs from the front, long ones from the back */
    uint32_t nqb, q0; const uint64_t *qoff;
    const LqQStat *qstat; const uint32_t *qlen, *tlen; const uint64_t *first;
    uint64_t *lambda, *lambda2; uint32_t *mcnt;
    LqOvl *ovl; uint32_t *n_ovl; uint32_t ovl_cap; uint32_t *n_chains;
    LqMapOpt o;
};

#define CH_WARPS 4
#define CH_RING 128
__global__ void __launch_bounds__(CH_WARPS * 32) lq_chain_k(ChainArgs a)
{
    __shared__ uint32_t s_r[CH_WARPS][CH_RING];
    __shared__ int32_t s_q[CH_WARPS][CH_RING], s_f[CH_WARPS][CH_RING], s_p[CH_WARPS][CH_RING], s_v[CH_WARPS][CH_RING], s_t[CH_WARPS][CH_RING];
    const uint32_t lane = threadIdx.x & 31, wid_ = threadIdx.x >> 5;
    uint32_t *ring_r = s_r[wid_]; int32_t *ring_q = s_q[wid_], *ring_f = s_f[wid_], *ring_p = s_p[wid_], *ring_v = s_v[wid_], *ring_t = s_t[wid_];
    const uint32_t ng = *a.n_groups;
    for (;;) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(a.cursor, 1u);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= ng) break;
        const uint2 run = a.runs[a.big_cap - 1 - g];
        const int32_t gb = (int32_t)run.x, ge = (int32_t)run.y, n = ge - gb;
        if (n < a.o.min_cnt) continue; /* a chain needs min_cnt anchors of one (strand, target) run (chain.c:116-119) */
        /* owning query */
        uint32_t lo = 0, hi = a.nqb;
        while (hi - lo > 1) { const uint32_t mid = lo + ((hi - lo) >> 1); if (a.qoff[mid] <= (uint64_t)gb) lo = mid; else hi = mid; }
        const uint32_t q = a.q0 + lo;
        const float avg_span = a.qstat[q].avg_span;

        /* ---- DP (chain.c:41-80), 32 predecessors per step.  The last CH_RING anchors (target pos, query pos, f, p, v and the
         *      t[] stamps) live in a per-warp shared-memory ring, so the inner loop does not wait on global memory; anchors
         *      further back than the ring (dense repeats) are read from global memory. ---- */
        int32_t st = gb;
        for (int e = lane; e < CH_RING; e += 32) ring_t[e] = -1;
        __syncwarp();
        uint64_t nx = a.ax[gb]; uint32_t nq_ = a.aq[gb], nm_ = a.am[gb];     /* next anchor, fetched one iteration ahead */
        for (int32_t i = gb; i < ge; ++i) {
            const uint32_t ri = (uint32_t)nx;
            const int32_t qi = (int32_t)nq_, span = (int32_t)(nm_ >> 24);
            if (i + 1 < ge) { nx = a.ax[i + 1]; nq_ = a.aq[i + 1]; nm_ = a.am[i + 1]; }
            if (lane == 0) { ring_r[i & (CH_RING - 1)] = ri; ring_q[i & (CH_RING - 1)] = qi; }
            int32_t best = span, best_j = -1, n_skip = 0;
            while (st < i) { /* advance the window start (uniform) */
                const int32_t j = st + (int32_t)lane;
                bool far = false;
                if (j < i) { const uint32_t rj = i - j < CH_RING ? ring_r[j & (CH_RING - 1)] : (uint32_t)a.ax[j]; far = (uint64_t)ri - (uint64_t)rj > (uint64_t)a.o.max_dist; }
                const uint32_t m = __ballot_sync(0xffffffffu, far);
                const int adv = m == 0xffffffffu ? 32 : __ffs(~m) - 1; /* rpos ascending: `far` is a prefix */
                st += adv;
                if (adv < 32) break;
            }
            bool stop = false;
            for (int32_t jb = i - 1; jb >= st && !stop; jb -= 32) {
                const int32_t j = jb - (int32_t)lane;
                const bool inr = i - j < CH_RING;                 /* anchor j still in the ring */
                const int js = j & (CH_RING - 1);
                bool valid = false; int32_t sc = INT32_MIN, pj = -1;
                if (j >= st) {
                    const uint32_t rj = inr ? ring_r[js] : (uint32_t)a.ax[j];
                    const int32_t qj = inr ? ring_q[js] : (int32_t)a.aq[j];
                    int32_t gain;
                    if (lq_chain_gain((int64_t)ri - (int64_t)rj, qi - qj, span, a.o.max_dist, a.o.max_dist, a.o.bw, avg_span, &gain)) {
                        valid = true; sc = gain + (inr ? ring_f[js] : a.f[j]); pj = inr ? ring_p[js] : a.p[j];
                    }
                }
                if (valid && pj >= 0) { if (i - pj < CH_RING) ring_t[pj & (CH_RING - 1)] = i; else a.t[pj] = i; }   /* chain.c:77, every lane: stamps past a break are never read */
                __syncwarp();
                const bool tflag = valid && (inr ? ring_t[js] : a.t[j]) == i;
                /* exclusive prefix max over lanes, seeded with `best` */
                int32_t pm = sc;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int32_t o = __shfl_up_sync(0xffffffffu, pm, d); if (lane >= (uint32_t)d) pm = max(pm, o); }
                int32_t ex = __shfl_up_sync(0xffffffffu, pm, 1);
                if (lane == 0) ex = INT32_MIN;
                ex = max(ex, best);
                const bool is_new = valid && sc > ex;
                const uint32_t newmask = __ballot_sync(0xffffffffu, is_new);
                const uint32_t incmask = __ballot_sync(0xffffffffu, valid && !is_new && tflag);
                uint32_t ev = newmask | incmask; int stop_lane = 32;
                while (ev) { /* replay chain.c:70-76 on the event stream (uniform across the warp) */
                    const int l = __ffs(ev) - 1; ev &= ev - 1;
                    if (newmask >> l & 1) { if (n_skip > 0) --n_skip; }
                    else if (++n_skip > a.o.max_skip) { stop_lane = l; break; }
                }
                const uint32_t nm = stop_lane < 32 ? newmask & ((1u << stop_lane) - 1) : newmask;
                if (nm) { const int l = 31 - __clz(nm); best = __shfl_sync(0xffffffffu, sc, l); best_j = jb - l; }
                if (stop_lane < 32) stop = true;
            }
            if (lane == 0) {
                int32_t vb = best;
                if (best_j >= 0) { const int32_t vj = i - best_j < CH_RING ? ring_v[best_j & (CH_RING - 1)] : a.v[best_j]; if (vj > best) vb = vj; }
                a.f[i] = best; a.p[i] = best_j; a.v[i] = vb;
                const int is_ = i & (CH_RING - 1);
                ring_f[is_] = best; ring_p[is_] = best_j; ring_v[is_] = vb; ring_t[is_] = -1;
            }
            __syncwarp();
        }

        /* ---- chain ends (chain.c:82-101) ---- */
        for (int32_t i = gb + lane; i < ge; i += 32) a.t[i] = 0;
        __syncwarp();
        for (int32_t i = gb + lane; i < ge; i += 32) if (a.p[i] >= 0) a.t[a.p[i]] = 1;
        __syncwarp();
        uint32_t n_end = 0;
        for (int32_t i0 = gb; i0 < ge; i0 += 32) {
            const int32_t i = i0 + (int32_t)lane;
            bool is_end = i < ge && a.t[i] == 0 && a.v[i] >= a.o.min_sc;
            uint64_t u = 0;
            if (is_end) {
                int32_t j = i;
                while (j >= 0 && a.f[j] < a.v[j]) j = a.p[j];
                if (j < 0) j = i;
                u = (uint64_t)(uint32_t)a.f[j] << 32 | (uint32_t)j;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, is_end);
            if (is_end) a.uend[gb + n_end + __popc(m & ((1u << lane) - 1))] = u;
            n_end += __popc(m);
        }
        __syncwarp();
        if (n_end == 0) continue;
        /* ---- order by (score, index) descending (chain.c:102-106): rank sort, staged through
         *      the group's slices of the head/gid arrays (dead once the group list exists) ---- */
        uint64_t *ue = a.uend + gb;
        if (n_end > 1) {
            uint32_t *s_lo = a.s_lo + gb, *s_hi = a.s_hi + gb;
            for (uint32_t e = lane; e < n_end; e += 32) {
                const uint64_t k = ue[e]; uint32_t r = 0;
                for (uint32_t o = 0; o < n_end; ++o) r += ue[o] > k || (ue[o] == k && o < e); /* two ends may share a peak: equal keys */
                s_lo[r] = (uint32_t)k; s_hi[r] = (uint32_t)(k >> 32);
            }
            __syncwarp();
            for (uint32_t e = lane; e < n_end; e += 32) ue[e] = (uint64_t)s_hi[e] << 32 | s_lo[e];
            __syncwarp();
        }
        /* ---- backtrack (chain.c:108-125) + regions (hit.c:23-38) + accounting (esterr.c:99-139) ---- */
        for (int32_t i = gb + lane; i < ge; i += 32) a.t[i] = 0;
        __syncwarp();
        const int32_t qlen = (int32_t)a.qlen[q];
        int32_t n_v = 0;
        for (uint32_t e = 0; e < n_end; ++e) {
            int32_t cnt = 0, score = 0, keepc = 0;
            if (lane == 0) {
                const int32_t n_v0 = n_v;
                int32_t j = (int32_t)(uint32_t)ue[e];
                do { a.vl[gb + n_v++] = (uint32_t)j; a.t[j] = 1; j = a.p[j]; } while (j >= 0 && a.t[j] == 0);
                cnt = n_v - n_v0;
                if (j < 0) { score = (int32_t)(ue[e] >> 32); keepc = cnt >= a.o.min_cnt; }
                else if ((int32_t)(ue[e] >> 32) - a.f[j] >= a.o.min_sc) { score = (int32_t)(ue[e] >> 32) - a.f[j]; keepc = cnt >= a.o.min_cnt; }
                if (!keepc) n_v = n_v0;
            }
            keepc = __shfl_sync(0xffffffffu, keepc, 0);
            if (!keepc) continue;
            cnt = __shfl_sync(0xffffffffu, cnt, 0); score = __shfl_sync(0xffffffffu, score, 0);
            n_v = __shfl_sync(0xffffffffu, n_v, 0);
            __syncwarp();
            const uint32_t *anch = a.vl + gb + (n_v - cnt);      /* descending index: anch[cnt-1] is the first anchor */
            const uint32_t i_first = anch[cnt - 1], i_last = anch[0];
            const uint64_t x0 = a.ax[i_first];
            const int32_t span0 = (int32_t)(a.am[i_first] >> 24), rev = (int32_t)(x0 >> 63), rid = (int32_t)(x0 << 1 >> 33);
            const int32_t rs = (int32_t)(uint32_t)x0 + 1 > span0 ? (int32_t)(uint32_t)x0 + 1 - span0 : 0;
            const int32_t re = (int32_t)(uint32_t)a.ax[i_last] + 1;
            int32_t qs, qe;
            if (!rev) { qs = (int32_t)a.aq[i_first] + 1 - span0; qe = (int32_t)a.aq[i_last] + 1; }
            else { qs = qlen - ((int32_t)a.aq[i_last] + 1); qe = qlen - ((int32_t)a.aq[i_first] + 1 - span0); }
            if (lane == 0) atomicAdd(a.n_chains, 1u);
            /* esterr.c:112-119 (unsigned arithmetic, double compare) */
            const uint32_t uqs = (uint32_t)qs, uqe = (uint32_t)qe, urs = (uint32_t)rs, ure = (uint32_t)re, rl = a.tlen[rid];
            const uint32_t h5 = uqs < urs ? uqs : urs;
            const uint32_t h3 = (uint32_t)qlen - uqe < rl - ure ? (uint32_t)qlen - uqe : rl - ure;
            if ((double)(uqe - uqs) < (double)(uqe - uqs + h5 + h3) * a.o.min_ratio || h5 > (uint32_t)a.o.max_overhang || h3 > (uint32_t)a.o.max_overhang)
                continue;
            const uint32_t flag = score >= (int32_t)(uint16_t)a.o.min_sc_med ? 2u : 0u;
            if (lane == 0) {
                atomicAdd((unsigned long long*)&a.lambda[q], (unsigned long long)(uqe - uqs + 1));
                const uint32_t at = atomicAdd(a.n_ovl, 1u);
                if (at < a.ovl_cap) { a.ovl[at].q = q; a.ovl[at].start = uqs << 3 | flag; a.ovl[at].end = uqe << 3 | flag | 1u; }
            }
            if (score < (int32_t)(uint16_t)a.o.min_sc_good) continue;
            if (lane == 0) atomicAdd((unsigned long long*)&a.lambda2[q], (unsigned long long)(uqe - uqs + 1));
            /* esterr.c:130-137: every anchor of the chain is a kept minimizer of the query */
            uint32_t *mc = a.mcnt + a.first[q];
            for (int32_t c = lane; c < cnt; c += 32) atomicAdd(&mc[a.am[anch[c]] & 0xffffffu], 1u);
        }
        __syncwarp();
    }
}


/* ---- runs of <= CH_SMALL anchors: one thread per run, the reference's loops as they are (chain.c:41-125), then the same
 *      region/accounting code as the warp kernel.  Most of these runs are chance hits that end without a chain. ---- */
__global__ void __launch_bounds__(128) lq_chain_small_k(ChainArgs a)
{
    const uint32_t ns = *a.n_small;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < ns; b += gridDim.x * blockDim.x) {
        const uint2 run = a.runs[b];
        const int32_t gb = (int32_t)run.x, n = (int32_t)run.y - gb;
        uint32_t lo = 0, hi = a.nqb;
        while (hi - lo > 1) { const uint32_t mid = lo + ((hi - lo) >> 1); if (a.qoff[mid] <= (uint64_t)gb) lo = mid; else hi = mid; }
        const uint32_t q = a.q0 + lo;
        const float avg_span = a.qstat[q].avg_span;
        uint32_t r[CH_SMALL]; int32_t qp[CH_SMALL], f[CH_SMALL], p[CH_SMALL], v[CH_SMALL], t[CH_SMALL]; uint32_t am[CH_SMALL];
        for (int i = 0; i < n; ++i) { r[i] = (uint32_t)a.ax[gb + i]; qp[i] = (int32_t)a.aq[gb + i]; am[i] = a.am[gb + i]; t[i] = -1; }
        int st = 0, peak = 0;
        for (int i = 0; i < n; ++i) {
            const int32_t span = (int32_t)(am[i] >> 24);
            int32_t best = span, best_j = -1, n_skip = 0;
            while (st < i && (uint64_t)r[i] - (uint64_t)r[st] > (uint64_t)a.o.max_dist) ++st;
            for (int j = i - 1; j >= st; --j) {
                int32_t sc;
                if (!lq_chain_gain((int64_t)r[i] - (int64_t)r[j], qp[i] - qp[j], span, a.o.max_dist, a.o.max_dist, a.o.bw, avg_span, &sc)) continue;
                sc += f[j];
                if (sc > best) { best = sc; best_j = j; if (n_skip > 0) --n_skip; }
                else if (t[j] == i) { if (++n_skip > a.o.max_skip) break; }
                if (p[j] >= 0) t[p[j]] = i;
            }
            f[i] = best; p[i] = best_j;
            v[i] = best_j >= 0 && v[best_j] > best ? v[best_j] : best;
            if (v[i] > peak) peak = v[i];
        }
        if (peak < a.o.min_sc) continue;                      /* no chain end can qualify (chain.c:86) */
        /* chain ends, best first (chain.c:82-106) */
        for (int i = 0; i < n; ++i) t[i] = 0;
        for (int i = 0; i < n; ++i) if (p[i] >= 0) t[p[i]] = 1;
        uint64_t u[CH_SMALL]; int n_u = 0;
        for (int i = 0; i < n; ++i)
            if (t[i] == 0 && v[i] >= a.o.min_sc) {
                int j = i;
                while (j >= 0 && f[j] < v[j]) j = p[j];
                if (j < 0) j = i;
                const uint64_t key = (uint64_t)(uint32_t)f[j] << 32 | (uint32_t)j;
                int e = n_u++;
                while (e > 0 && u[e - 1] < key) { u[e] = u[e - 1]; --e; }   /* descending; equal keys (shared peak) stay adjacent */
                u[e] = key;
            }
        /* backtrack (chain.c:108-125) and account */
        for (int i = 0; i < n; ++i) t[i] = 0;
        const int32_t qlen = (int32_t)a.qlen[q];
        int32_t vl[CH_SMALL]; int n_v = 0;
        for (int e = 0; e < n_u; ++e) {
            const int n_v0 = n_v;
            int j = (int)(uint32_t)u[e];
            do { vl[n_v++] = j; t[j] = 1; j = p[j]; } while (j >= 0 && t[j] == 0);
            const int cnt = n_v - n_v0;
            int32_t score; bool keep;
            if (j < 0) { score = (int32_t)(u[e] >> 32); keep = cnt >= a.o.min_cnt; }
            else if ((int32_t)(u[e] >> 32) - f[j] >= a.o.min_sc) { score = (int32_t)(u[e] >> 32) - f[j]; keep = cnt >= a.o.min_cnt; }
            else { score = 0; keep = false; }
            if (!keep) { n_v = n_v0; continue; }
            atomicAdd(a.n_chains, 1u);
            const int i_first = vl[n_v - 1], i_last = vl[n_v0];
            const uint64_t x0 = a.ax[gb + i_first];
            const int32_t span0 = (int32_t)(am[i_first] >> 24), rev = (int32_t)(x0 >> 63), rid = (int32_t)(x0 << 1 >> 33);
            const int32_t rs = (int32_t)r[i_first] + 1 > span0 ? (int32_t)r[i_first] + 1 - span0 : 0;
            const int32_t re = (int32_t)r[i_last] + 1;
            int32_t qs, qe;
            if (!rev) { qs = qp[i_first] + 1 - span0; qe = qp[i_last] + 1; }
            else { qs = qlen - (qp[i_last] + 1); qe = qlen - (qp[i_first] + 1 - span0); }
            const uint32_t uqs = (uint32_t)qs, uqe = (uint32_t)qe, urs = (uint32_t)rs, ure = (uint32_t)re, rl = a.tlen[rid];
            const uint32_t h5 = uqs < urs ? uqs : urs;
            const uint32_t h3 = (uint32_t)qlen - uqe < rl - ure ? (uint32_t)qlen - uqe : rl - ure;
            if ((double)(uqe - uqs) < (double)(uqe - uqs + h5 + h3) * a.o.min_ratio || h5 > (uint32_t)a.o.max_overhang || h3 > (uint32_t)a.o.max_overhang)
                continue;
            const uint32_t flag = score >= (int32_t)(uint16_t)a.o.min_sc_med ? 2u : 0u;
            atomicAdd((unsigned long long*)&a.lambda[q], (unsigned long long)(uqe - uqs + 1));
            const uint32_t at = atomicAdd(a.n_ovl, 1u);
            if (at < a.ovl_cap) { a.ovl[at].q = q; a.ovl[at].start = uqs << 3 | flag; a.ovl[at].end = uqe << 3 | flag | 1u; }
            if (score < (int32_t)(uint16_t)a.o.min_sc_good) continue;
            atomicAdd((unsigned long long*)&a.lambda2[q], (unsigned long long)(uqe - uqs + 1));
            uint32_t *mc = a.mcnt + a.first[q];
            for (int c = n_v0; c < n_v; ++c) atomicAdd(&mc[am[vl[c]] & 0xffffffu], 1u);
        }
    }
}

/* sorted by key (stable, so equal keys of one query are adjacent): flag (key, query, strand) groups of size > 1 */
__global__ void lq_dupflag_k(uint64_t n, const uint32_t *__restrict__ skey, const uint64_t *__restrict__ smi, const uint64_t *__restrict__ qy, uint8_t *__restrict__ dup)
{
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t key = skey[j]; const uint64_t mi = smi[j], y = qy[mi];
    const uint32_t q = (uint32_t)(y >> 32), strand = (uint32_t)y & 1;
    bool d = false;
    for (uint64_t o = j; o-- > 0 && skey[o] == key; ) { const uint64_t yo = qy[smi[o]]; if ((uint32_t)(yo >> 32) != q) break; if (((uint32_t)yo & 1) == strand) { d = true; break; } }
    for (uint64_t o = j + 1; !d && o < n && skey[o] == key; ++o) { const uint64_t yo = qy[smi[o]]; if ((uint32_t)(yo >> 32) != q) break; if (((uint32_t)yo & 1) == strand) d = true; }
    dup[mi] = d;
}
__global__ void lq_iota64_k(uint64_t *a, uint64_t n) { const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = i; }

/* minimap2-coverage.c:552-562 */
__global__ void lq_nmatch_k(uint32_t nq, const uint64_t *__restrict__ first, const uint32_t *__restrict__ mcnt, const uint32_t *__restrict__ npre,
                            uint32_t *__restrict__ n_match, uint32_t *__restrict__ sat)
{
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nq) return;
    const uint64_t b = first[q];
    uint64_t e = first[q + 1];
    uint32_t n_div = (uint32_t)(e - b);
    if (npre) { n_div = npre[q]; if (b + n_div < e) e = b + n_div; }   /* the reference sums its first n counters, n from the command line's k / w */
    uint32_t sum = 0, big = 0;
    for (uint64_t i = b + lane; i < e; i += 32) { sum += mcnt[i]; big |= mcnt[i] >= 65535u; }
    sum = lq_warp_sum(sum); big = __any_sync(0xffffffffu, big);
    uint32_t nm = 0;
    if (n_div > 0) {
        const uint32_t mean = sum / n_div;
        for (uint64_t i = b + lane; i < e; i += 32) nm += mcnt[i] > mean;
        nm = lq_warp_sum(nm);
    }
    if (lane == 0) { n_match[q] = nm; if (big) atomicAdd(sat, 1u); }
}

/* ------------------------------------------------------------------ host orchestration */

static int upload_u32(LqDevBuf &b, const uint32_t *h, size_t n, cudaStream_t st)
{
    LQ_TRY(b.ensure((n + 1) * 4));
    if (n) LQ_CUDA_OK(cudaMemcpyAsync(b.p, h, n * 4, cudaMemcpyHostToDevice, st));
    return 0;
}

struct BatchPtrs {
    SeedArrays s; uint64_t *kx2; uint32_t *idx, *idx2, *dest, *ord; uint8_t *dig;   /* arena1, sort phase (s.sx == ax: the keys are sorted in place in arena2) */
    int32_t *f, *p, *v, *t; uint64_t *uend; uint32_t *vl, *head, *gid;  /* arena1, chain phase (aliases) */
    uint64_t *ax; uint32_t *aq, *am;                           /* arena2 */
};

static inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

static int carve(LqMapScratch *sc, uint64_t nb, BatchPtrs *b)
{
    const size_t n = (size_t)nb + 64;
    /* sort phase: kx2 8, sq 4, sm 4, idx 4, idx2 4, dest 4, ord 4, dig 1 */
    const size_t sort_bytes = al(n * 8) + 6 * al(n * 4) + al(n);
    /* chain phase: f,p,v,t 4 each, uend 8, vl 4, head 4, gid 4 */
    const size_t chain_bytes = 7 * al(n * 4) + al(n * 8) + 256;
    LQ_TRY(sc->arena1.ensure(std::max(sort_bytes, chain_bytes)));
    LQ_TRY(sc->arena2.ensure(al(n * 8) + 2 * al(n * 4)));
    char *p = (char*)sc->arena1.p;
    b->kx2 = (uint64_t*)p; p += al(n * 8);
    b->s.sq = (uint32_t*)p; p += al(n * 4);
    b->s.sm = (uint32_t*)p; p += al(n * 4);
    b->idx = (uint32_t*)p; p += al(n * 4);
    b->idx2 = (uint32_t*)p; p += al(n * 4);
    b->dest = (uint32_t*)p; p += al(n * 4);
    b->ord = (uint32_t*)p; p += al(n * 4);
    b->dig = (uint8_t*)p;
    p = (char*)sc->arena1.p;
    b->uend = (uint64_t*)p; p += al(n * 8);
    b->f = (int32_t*)p; p += al(n * 4);
    b->p = (int32_t*)p; p += al(n * 4);
    b->v = (int32_t*)p; p += al(n * 4);
    b->t = (int32_t*)p; p += al(n * 4);
    b->vl = (uint32_t*)p; p += al(n * 4);
    b->head = (uint32_t*)p; p += al(n * 4);
    b->gid = (uint32_t*)p;
    p = (char*)sc->arena2.p;
    b->ax = (uint64_t*)p; p += al(n * 8);
    b->aq = (uint32_t*)p; p += al(n * 4);
    b->am = (uint32_t*)p;
    b->s.sx = b->ax; b->s.idx = b->idx;
    return 0;
}

static void fill_flargs(FlArgs *fa, LqQueryDev *qd, const LqIndexDev *ix, const MapTables &mt, uint32_t thr, uint32_t q0, uint32_t q1);

/* seeds of queries [q0,q1) -> sorted arrays in arena2.  h_qoff: nqb+1 batch-relative seed offsets. */
static int seed_and_sort(LqQueryDev *qd, const LqIndexDev *ix, const MapTables &mt, uint32_t filter_thr, uint32_t q0, uint32_t q1, const std::vector<uint64_t> &h_first,
                         uint64_t seed_base, uint64_t nb, const std::vector<uint64_t> &h_qoff, LqMapScratch *sc, BatchPtrs *b, uint64_t **d_qoff_out,
                         LqMapStats *stats, cudaStream_t st)
{
    const uint32_t nqb = q1 - q0;
    LQ_TRY(carve(sc, nb, b));
    /* misc: qoff (nqb+1 u64) | counters (16 u32) */
    LQ_TRY(sc->misc.ensure(al((size_t)(nqb + 2) * 8) + 512));
    uint64_t *d_qoff = sc->misc.as<uint64_t>();
    uint32_t *ctr = (uint32_t*)((char*)sc->misc.p + al((size_t)(nqb + 2) * 8));
    LQ_CUDA_OK(cudaMemcpyAsync(d_qoff, h_qoff.data(), (size_t)(nqb + 1) * 8, cudaMemcpyHostToDevice, st));
    LQ_CUDA_OK(cudaMemsetAsync(ctr, 0, 512, st));   /* ctr[0..15], ctr[48..55]: u32 counters; ctr[16..47]: 16 u64 element counters (8 levels, 8 long walks); ctr[64..79]: 8 more (short walks) */
    *d_qoff_out = d_qoff;
    if (nb == 0) return 0;
    if (nb >= (1ULL << 31)) { fprintf(stderr, "[lqcov] a batch of %llu seeds exceeds the 31-bit seed numbers of the sort\n", (unsigned long long)nb); return -1; }
    const uint64_t mi0 = h_first[q0], mi1 = h_first[q1];
    if (mi1 > mi0) {
        LqProfScope ps("seed_fill", st, 1, (mi1 - mi0) * 32 + nb * 24);
        lq_fill_k<<<lq_grid((mi1 - mi0) * 32, 256), 256, 0, st>>>(mi0, mi1, qd->mins.key.as<uint32_t>(), qd->mins.y.as<uint64_t>(),
            qd->mins.has_span ? qd->mins.span.as<uint8_t>() : 0, ix->k, qd->first.as<uint64_t>(), qd->keep.as<uint32_t>(), qd->neff.as<uint32_t>(),
            qd->krank.as<uint32_t>(), qd->soff.as<uint64_t>(), seed_base, ix->counts.as<uint32_t>(), ix->offs.as<uint64_t>(), ix->rec.y.as<uint64_t>(),
            mt, qd->reads.len.as<uint32_t>(), qd->dup.as<uint8_t>(), filter_thr ? qd->qtied.as<uint8_t>() : (const uint8_t*)0, b->s);
        LQ_CUDA_OK(cudaGetLastError());
    }
    if (mi1 > mi0 && filter_thr) {
        FlArgs fa;
        fill_flargs(&fa, qd, ix, mt, filter_thr, q0, q1);
        fa.seed_base = seed_base; fa.s = b->s;
        LqProfScope ps("seed_fill_filtered", st, 1, 0);
        lq_fill_masked_k<<<lq_grid(mi1 - mi0, 128), 128, 0, st>>>(fa, mi0, mi1);
        LQ_CUDA_OK(cudaGetLastError());
    }
    /* bucket lists: at most nb/65 + nqb live buckets per level */
    const size_t bcap = (size_t)(nb / (LQ_RS_MIN + 1)) + nqb + 16, bcapb = (size_t)(nb / AFB_MIN) + nqb + 16;
    const size_t bcapw = (size_t)(nb / AFW_SMALL) + nqb + 16;   /* long walks: buckets of >= AFW_SMALL elements */
    LQ_TRY(sc->bkt.ensure((4 * bcap + 2 * bcapb + bcapw) * sizeof(AfBkt)));
    LQ_TRY(sc->wph.ensure(2 * bcapw * 256 * sizeof(lq_afq_phase)));
    LQ_TRY(sc->wst.ensure((size_t)AFS_GRID * AFS_WALKERS * AFS_ROW * 4));
    LQ_CUDA_OK(cudaFuncSetAttribute(lq_af_walk3_k, cudaFuncAttributeMaxDynamicSharedMemorySize, AFS_WALKERS * 4096));
    LQ_CUDA_OK(cudaFuncSetAttribute(lq_af_walkf_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AffSmem)));
    LQ_CUDA_OK(cudaFuncSetAttribute(lq_af_big_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(g_afb_bitw * 4)));
    AfBkt *bk[2] = { sc->bkt.as<AfBkt>(), sc->bkt.as<AfBkt>() + bcap };
    AfBkt *wl = sc->bkt.as<AfBkt>() + 2 * bcap;
    AfBkt *wls = sc->bkt.as<AfBkt>() + 3 * bcap;                                                     /* short walks: ctr[14] = count, ctr[15] = cursor */
    AfBkt *bkb[2] = { sc->bkt.as<AfBkt>() + 4 * bcap, sc->bkt.as<AfBkt>() + 4 * bcap + bcapb };   /* long buckets: ctr[11], ctr[12] = counts, ctr[13] = cursor */
    AfBkt *wlw = sc->bkt.as<AfBkt>() + 4 * bcap + 2 * bcapb;                                         /* few-region walks: ctr[52] = count, ctr[53] = cursor */
    /* counters: ctr[0],ctr[1] = bucket counts of the two lists, ctr[2] = cursor, ctr[3] = walks */
    lq_prof_count_launch(1);
    lq_af_init_k<<<lq_grid(nqb, 128), 128, 0, st>>>(nqb, d_qoff, b->ax, b->idx, bk[0], ctr + 0, bkb[0], ctr + 11);
    LQ_CUDA_OK(cudaGetLastError());
    static const char *lvl_name[8] = { "seed_sort_s0", "seed_sort_s8", "seed_sort_s16", "seed_sort_s24", "seed_sort_s32", "seed_sort_s40", "seed_sort_s48", "seed_sort_s56" };
    static const char *plc_name[8] = { "seed_place_s0", "seed_place_s8", "seed_place_s16", "seed_place_s24", "seed_place_s32", "seed_place_s40", "seed_place_s48", "seed_place_s56" };
    static const char *wls_name[8] = { "seed_walksmall_s0", "seed_walksmall_s8", "seed_walksmall_s16", "seed_walksmall_s24", "seed_walksmall_s32", "seed_walksmall_s40", "seed_walksmall_s48", "seed_walksmall_s56" };
    static const char *wlk_name[8] = { "seed_walk_s0", "seed_walk_s8", "seed_walk_s16", "seed_walk_s24", "seed_walk_s32", "seed_walk_s40", "seed_walk_s48", "seed_walk_s56" };
    int cur = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        AfArgs a;
        const int odd = ((56 - shift) >> 3) & 1;           /* even levels read the primary buffer (ax, idx) and leave their buckets in (kx2, idx2); odd levels the other way */
        a.kx = odd ? b->kx2 : b->ax; a.kx2 = odd ? b->ax : b->kx2; a.idx = odd ? b->idx2 : b->idx; a.idx2 = odd ? b->idx : b->idx2;
        a.kxp = b->ax; a.idxp = b->idx; a.dst_primary = odd;
        a.dest = b->dest; a.ord = b->ord; a.dig = b->dig;
        a.cur = bk[cur]; a.n_cur = ctr + cur; a.nxt = bk[cur ^ 1]; a.n_nxt = ctr + (cur ^ 1); a.cursor = ctr + 2; a.n_walk = ctr + 3; a.shift = shift;
        a.wlist = wl; a.n_wlist = ctr + 9; a.wcursor = ctr + 10;
        a.wlist_s = wls; a.n_wlist_s = ctr + 14; a.wcursor_s = ctr + 15; a.wcursor_f = ctr + 48;
        a.wlist_w = wlw; a.n_wlist_w = ctr + 52; a.wcursor_w = ctr + 53; a.wph = sc->wph.as<lq_afq_phase>(); a.wph_cap = (uint32_t)bcapw;
        a.bitw = g_afb_bitw; a.two = g_afb_bitw && ((shift == 56 && ix->n_seq <= (1u << 24)) || (shift == 48 && ix->n_seq <= (1u << 17)));
        LQ_CUDA_OK(cudaMemsetAsync(ctr + 14, 0, 8, st));
        LQ_CUDA_OK(cudaMemsetAsync(ctr + 48, 0, 4, st));
        LQ_CUDA_OK(cudaMemsetAsync(ctr + 52, 0, 8, st));
        a.curb = bkb[cur]; a.n_curb = ctr + 11 + cur; a.nxtb = bkb[cur ^ 1]; a.n_nxtb = ctr + 11 + (cur ^ 1); a.cursorb = ctr + 13;
        LQ_CUDA_OK(cudaMemsetAsync(ctr + 11 + (cur ^ 1), 0, 4, st));
        LQ_CUDA_OK(cudaMemsetAsync(ctr + 13, 0, 4, st));
        LQ_CUDA_OK(cudaMemsetAsync(ctr + (cur ^ 1), 0, 4, st));
        LQ_CUDA_OK(cudaMemsetAsync(ctr + 2, 0, 4, st));
        LQ_CUDA_OK(cudaMemsetAsync(ctr + 9, 0, 8, st));
        unsigned long long *n_elem = (unsigned long long*)(ctr + 16);
        a.n_elem = n_elem + (shift >> 3);
        { LqProfScope ps(lvl_name[shift >> 3], st, 2, 0);
          lq_af_big_k<<<148 * g_afb_ctas, AFB_THREADS, g_afb_bitw * 4, st>>>(a);
          lq_af_level_k<<<148 * 16, AF_WARPS * 32, 0, st>>>(a); }
        a.n_elem = n_elem + 8 + (shift >> 3);
        { LqProfScope ps(wlk_name[shift >> 3], st, 2, 0);
          lq_af_walk3_k<<<AFS_GRID, AFS_THREADS, AFS_WALKERS * 4096, st>>>(a, sc->wst.as<uint32_t>());
          lq_af_walkf_k<<<AFF_GRID, AFF_THREADS, sizeof(AffSmem), st>>>(a); }
        if (g_walk_stats) {   /* LQCOV_WALK_STATS=1: the walks of this level (count, elements, longest) on stderr -- the longest one is the critical path */
            uint32_t cnt[2] = { 0, 0 };
            cudaStreamSynchronize(st);
            cudaMemcpy(&cnt[0], ctr + 9, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cnt[1], ctr + 52, 4, cudaMemcpyDeviceToHost);
            const AfBkt *lists[2] = { a.wlist, a.wlist_w };
            for (int f = 0; f < 2; ++f) {
                std::vector<AfBkt> h(cnt[f]);
                if (cnt[f]) cudaMemcpy(h.data(), lists[f], (size_t)cnt[f] * sizeof(AfBkt), cudaMemcpyDeviceToHost);
                uint64_t tot = 0; uint32_t mx = 0;
                for (size_t i = 0; i < h.size(); ++i) { const uint32_t n = h[i].end - h[i].beg; tot += n; if (n > mx) mx = n; }
                if (cnt[f]) fprintf(stderr, "[lqcov] walk stats: shift %d %s walks %u elements %llu longest %u\n", shift, f ? "few-region" : "256-region", cnt[f], (unsigned long long)tot, mx);
            }
        }
        { LqProfScope ps(plc_name[shift >> 3], st, 1, 0);
          lq_af_place_k<<<148 * 3, AFB_THREADS, 0, st>>>(a); }
        a.n_elem = (unsigned long long*)(ctr + 64) + (shift >> 3);
        { LqProfScope ps(wls_name[shift >> 3], st, 1, 0);
          lq_af_walk_small_k<<<148 * 16, AFW_WARPS * 32, 0, st>>>(a); }
        LQ_CUDA_OK(cudaGetLastError());
        cur ^= 1;
    }
    LqProfScope psg("seed_gather", st, 1, nb * 20);
    lq_gather_k<<<lq_grid(nb, 256), 256, 0, st>>>(nb, b->idx, b->s, b->aq, b->am);
    LQ_CUDA_OK(cudaGetLastError());
    if (stats) {
        uint32_t w = 0; unsigned long long ne[16], nes[8];
        LQ_CUDA_OK(cudaMemcpyAsync(&w, ctr + 3, 4, cudaMemcpyDeviceToHost, st));
        LQ_CUDA_OK(cudaMemcpyAsync(ne, ctr + 16, sizeof ne, cudaMemcpyDeviceToHost, st));
        LQ_CUDA_OK(cudaMemcpyAsync(nes, ctr + 64, sizeof nes, cudaMemcpyDeviceToHost, st));
        LQ_CUDA_OK(cudaStreamSynchronize(st));
        stats->n_walk_buckets += w;
        if (lq_prof_on()) {
            /* algorithmic bytes per element and level: the 12-byte (key, seed number) pair read once and written once.
             * Digits, destinations and pick-up lists are this implementation's own traffic and are not counted. */
            for (int l = 0; l < 8; ++l) {
                lq_prof_add_bytes(lvl_name[l], ne[l] * 24ULL);
                lq_prof_add_bytes(wlk_name[l], ne[8 + l] * 2ULL);      /* the sequential part: one digit byte in, one out */
                lq_prof_add_bytes(plc_name[l], ne[8 + l] * 24ULL);
                lq_prof_add_bytes(wls_name[l], nes[l] * 24ULL);
            }
        }
    }
    return 0;
}

static void fill_flargs(FlArgs *fa, LqQueryDev *qd, const LqIndexDev *ix, const MapTables &mt, uint32_t thr, uint32_t q0, uint32_t q1)
{
    fa->qkey = qd->mins.key.as<uint32_t>(); fa->qy = qd->mins.y.as<uint64_t>(); fa->qspan = qd->mins.has_span ? qd->mins.span.as<uint8_t>() : 0; fa->k = ix->k;
    fa->first = qd->first.as<uint64_t>(); fa->keep = qd->keep.as<uint32_t>(); fa->neff = qd->neff.as<uint32_t>(); fa->qtied = qd->qtied.as<uint8_t>();
    fa->counts = ix->counts.as<uint32_t>(); fa->offs = ix->offs.as<uint64_t>(); fa->pos = ix->rec.y.as<uint64_t>();
    fa->t = mt; fa->qlen = qd->reads.len.as<uint32_t>(); fa->thr = thr;
    fa->krank = qd->krank.as<uint32_t>(); fa->soff = qd->soff.as<uint64_t>(); fa->seed_base = 0; fa->s.sx = 0; fa->s.sq = 0; fa->s.sm = 0; fa->s.idx = 0;
    fa->q0 = q0; fa->q1 = q1;
    fa->mask = qd->fmask.as<uint32_t>(); fa->mstride = qd->fmask_stride;
}

/* T of the pre-filter, 0 = off (debug hook, thresholds beyond the 4-bit counters, LQCOV_NO_SEED_FILTER=1) */
static uint32_t filter_threshold(const LqQueryDev *qd, const LqIndexDev *ix, const LqMapOpt *opt, int mid_occ)
{
    /* the survivor bits take ceil(mid_occ/32) words per query minimizer: not worth it for absurd thresholds */
    if (mid_occ <= 0 || (uint64_t)qd->n_min * (uint64_t)((mid_occ + 31) / 32) * 4 > (8ULL << 30)) return 0;
    static const int off = getenv("LQCOV_NO_SEED_FILTER") ? atoi(getenv("LQCOV_NO_SEED_FILTER")) : 0;
    if (off) return 0;
    const int max_span = qd->mins.has_span ? 255 : ix->k;
    const int by_score = (std::max(opt->min_sc, 0) + max_span - 1) / max_span;
    const int thr = std::max(std::max(opt->min_cnt, by_score), 1);
    return thr >= 2 && thr <= 15 ? (uint32_t)thr : 0;
}

static int run_lookup(LqQueryDev *qd, const LqIndexDev *ix, const LqMapOpt *opt, int mid_occ, MapTables *mt, uint32_t filter_thr,
                      const uint32_t *h_self_off, const uint32_t *h_self_list, const uint32_t *h_qrank, const uint32_t *h_trank,
                      LqMapScratch *sc, std::vector<LqQStat> *h_stat, cudaStream_t st)
{
    const uint32_t nq = qd->nq; const uint64_t nm = qd->n_min;
    LQ_TRY(upload_u32(qd->self_off, h_self_off, (size_t)nq + 1, st));
    LQ_TRY(upload_u32(qd->self_list, h_self_list, h_self_off[nq], st));
    if (opt->ava) { LQ_TRY(upload_u32(qd->qrank, h_qrank, nq, st)); LQ_TRY(upload_u32(qd->trank, h_trank, ix->n_seq, st)); }
    mt->no_self = opt->no_self; mt->ava = opt->ava;
    mt->self_off = qd->self_off.as<uint32_t>(); mt->self_list = qd->self_list.as<uint32_t>();
    mt->qrank = qd->qrank.as<uint32_t>(); mt->trank = qd->trank.as<uint32_t>();
    LQ_TRY(qd->keep.ensure((nm + 1) * 4)); LQ_TRY(qd->neff.ensure((nm + 1) * 4));
    LQ_TRY(qd->krank.ensure((nm + 2) * 4)); LQ_TRY(qd->soff.ensure((nm + 2) * 8));
    LQ_TRY(qd->qstat.ensure(((size_t)nq + 1) * sizeof(LqQStat)));
    h_stat->resize(nq);
    if (nm == 0) { for (uint32_t q = 0; q < nq; ++q) memset(&(*h_stat)[q], 0, sizeof(LqQStat)); return 0; }
    lq_prof_count_launch(1); lq_prof_h2d(((uint64_t)nq + 1 + h_self_off[nq]) * 4); lq_prof_d2h((uint64_t)nq * sizeof(LqQStat));
    { LqProfScope ps("seed_lookup", st, 1, nm * (12 + 4 + 8 + 8 * 7 + 8));   /* minimizer 12 B, count 4, offset 8, ~7 probes of the self-hit search, 8 written */
    lq_lookup_k<<<lq_grid(nm, 256), 256, 0, st>>>(qd->mins.key.as<uint32_t>(), qd->mins.y.as<uint64_t>(), nm, ix->counts.as<uint32_t>(), ix->offs.as<uint64_t>(),
        ix->rec.y.as<uint64_t>(), mid_occ, *mt, qd->lambda.as<uint64_t>(), qd->reads.len.as<uint32_t>(), opt->covt, qd->keep.as<uint32_t>(), qd->neff.as<uint32_t>()); }
    LQ_CUDA_OK(cudaGetLastError());
    LQ_TRY((lq_exclusive_scan<uint32_t, uint32_t>(qd->keep.as<uint32_t>(), qd->krank.as<uint32_t>(), nm, 1, sc->ws, st)));
    lq_qstat_k<<<lq_grid((size_t)nq * 32, 256), 256, 0, st>>>(nq, qd->first.as<uint64_t>(), qd->keep.as<uint32_t>(), qd->neff.as<uint32_t>(),
        qd->mins.has_span ? qd->mins.span.as<uint8_t>() : 0, ix->k, qd->lambda.as<uint64_t>(), qd->reads.len.as<uint32_t>(), opt->covt, qd->qstat.as<LqQStat>());
    LQ_CUDA_OK(cudaGetLastError());
    if (filter_thr) { /* seeds of runs too short to chain are not written for queries without tied keys */
        FlArgs fa;
        qd->fmask_stride = (uint32_t)((mid_occ + 31) / 32);   /* kept minimizers occur < mid_occ times */
        LQ_TRY(qd->fmask.ensure((size_t)nm * qd->fmask_stride * 4 + 64));
        fill_flargs(&fa, qd, ix, *mt, filter_thr, 0, nq);
        LQ_CUDA_OK(cudaFuncSetAttribute(lq_filter_count_k, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_WORDS * 4));
        LqProfScope ps("seed_filter", st, 1, 0);
        lq_filter_count_k<<<nq < 148 * 1 ? nq : 148, FL_THREADS, FL_WORDS * 4, st>>>(fa);
    }
    LQ_CUDA_OK(cudaGetLastError());
    LQ_TRY((lq_exclusive_scan<uint32_t, uint64_t>(qd->neff.as<uint32_t>(), qd->soff.as<uint64_t>(), nm, 1, sc->ws, st)));
    lq_qseeds_k<<<lq_grid(nq, 256), 256, 0, st>>>(nq, qd->first.as<uint64_t>(), qd->soff.as<uint64_t>(), qd->qstat.as<LqQStat>());
    lq_prof_count_launch(1);
    LQ_CUDA_OK(cudaGetLastError());
    LQ_CUDA_OK(cudaMemcpyAsync(h_stat->data(), qd->qstat.p, (size_t)nq * sizeof(LqQStat), cudaMemcpyDeviceToHost, st));
    LQ_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int lq_map_part(LqQueryDev *qd, const LqIndexDev *ix, const LqMapOpt *opt, int mid_occ,
                const uint32_t *h_self_off, const uint32_t *h_self_list, const uint32_t *h_qrank, const uint32_t *h_trank,
                uint64_t seed_cap, LqMapScratch *sc, std::vector<LqOvl> *ovl_out, std::vector<LqQStat> *h_stat,
                LqMapStats *stats, cudaStream_t st)
{
    const uint32_t nq = qd->nq;
    MapTables mt;
    if (nq == 0) return 0;
    const uint32_t filter_thr = filter_threshold(qd, ix, opt, mid_occ);
    LQ_TRY(run_lookup(qd, ix, opt, mid_occ, &mt, filter_thr, h_self_off, h_self_list, h_qrank, h_trank, sc, h_stat, st));
    if (filter_thr && lq_prof_on()) {   /* algorithmic bytes of the pre-filter: both kernels stream the 8-byte occurrence lists of the filtered queries twice */
        uint64_t occ = 0, kept = 0;
        for (uint32_t q = 0; q < nq; ++q) if ((*h_stat)[q].n_sorted != (*h_stat)[q].n_seeds) { occ += (*h_stat)[q].n_seeds; kept += (*h_stat)[q].n_sorted; }
        lq_prof_add_bytes("seed_filter", occ * 16); lq_prof_add_bytes("seed_fill_filtered", occ * 16 + kept * 16);
    }
    std::vector<uint64_t> h_first((size_t)nq + 1);
    LQ_CUDA_OK(cudaMemcpyAsync(h_first.data(), qd->first.p, ((size_t)nq + 1) * 8, cudaMemcpyDeviceToHost, st));
    LQ_CUDA_OK(cudaStreamSynchronize(st));
    if (seed_cap > 0x70000000ULL) seed_cap = 0x70000000ULL; /* 32-bit seed indices inside a batch */
    uint64_t seed_base = 0;
    for (uint32_t q0 = 0; q0 < nq; ) {
        /* batch = consecutive queries whose seeds fit the budget */
        uint32_t q1 = q0; uint64_t nb = 0;
        std::vector<uint64_t> h_qoff; h_qoff.push_back(0);
        while (q1 < nq && (q1 == q0 || nb + (*h_stat)[q1].n_sorted <= seed_cap)) { nb += (*h_stat)[q1].n_sorted; h_qoff.push_back(nb); ++q1; }
        if (nb > 0x7fffff00ULL) { fprintf(stderr, "[lqcov] query %u alone has %llu seeds in one part: beyond the 2^31 batch limit\n", q0, (unsigned long long)nb); return -1; }
        const uint32_t nqb = q1 - q0;
        BatchPtrs b; uint64_t *d_qoff = 0;
        LQ_TRY(seed_and_sort(qd, ix, mt, filter_thr, q0, q1, h_first, seed_base, nb, h_qoff, sc, &b, &d_qoff, stats, st));
        if (nb > 0) {
            uint32_t *ctr = (uint32_t*)((char*)sc->misc.p + al((size_t)(nqb + 2) * 8));
            /* runs that can chain (ctr[4] = long runs, ctr[8] = short runs, ctr[5] = cursor, ctr[6] = n_ovl, ctr[7] = n_chains; ctr[50..51] = u64 number of runs) */
            const int max_span = qd->mins.has_span ? 255 : ix->k;
            const uint32_t run_thr = (uint32_t)std::max(std::max(opt->min_cnt, 1), (std::max(opt->min_sc, 0) + max_span - 1) / max_span);
            const size_t big_cap = (size_t)(nb / run_thr) + 64;
            const uint64_t run_chunk = (((nb + RUN_WARPS - 1) / RUN_WARPS) + 31) & ~31ULL;
            const uint32_t cand_stride = (uint32_t)(run_chunk / run_thr + 2);
            LQ_TRY(sc->grp.ensure(big_cap * sizeof(uint2) + ((size_t)RUN_WARPS * cand_stride + RUN_WARPS + 16) * 4));
            const uint32_t ovl_cap = (uint32_t)std::min<uint64_t>(nb / (uint64_t)std::max(opt->min_cnt, 1) + 1024, 0x7fffffffULL);
            LQ_TRY(sc->ovl.ensure((size_t)ovl_cap * sizeof(LqOvl)));
            LQ_CUDA_OK(cudaMemsetAsync(ctr + 4, 0, 16, st));
            LQ_CUDA_OK(cudaMemsetAsync(ctr + 8, 0, 4, st));
            LQ_CUDA_OK(cudaMemsetAsync(ctr + 50, 0, 8, st));
            RunArgs ra;
            ra.ax = b.ax; ra.n = nb; ra.qoff = d_qoff; ra.nqb = nqb; ra.thr = run_thr; ra.runs = sc->grp.as<uint2>(); ra.cap = (uint32_t)big_cap;
            ra.n_small = ctr + 8; ra.n_large = ctr + 4; ra.n_heads = (unsigned long long*)(ctr + 50);
            ra.cand = (uint32_t*)(sc->grp.as<uint2>() + big_cap); ra.wcount = ra.cand + (size_t)RUN_WARPS * cand_stride; ra.chunk = run_chunk; ra.cand_stride = cand_stride;
            { LqProfScope ps("runs", st, 3, nb * 8);
              lq_runs_k<<<RUN_GRID, RUN_THREADS, 0, st>>>(ra);
              lq_runs_emit_k<<<RUN_GRID, RUN_THREADS, 0, st>>>(ra);
              lq_runs_q_k<<<lq_grid(nqb, 128), 128, 0, st>>>(ra); }
            LQ_CUDA_OK(cudaGetLastError());
            ChainArgs a;
            a.ax = b.ax; a.aq = b.aq; a.am = b.am; a.f = b.f; a.p = b.p; a.v = b.v; a.t = b.t; a.uend = b.uend; a.vl = b.vl; a.s_lo = b.head; a.s_hi = b.gid;
            a.runs = sc->grp.as<uint2>(); a.big_cap = (uint32_t)big_cap; a.n_small = ctr + 8; a.n_groups = ctr + 4; a.cursor = ctr + 5;
            a.nqb = nqb; a.q0 = q0; a.qoff = d_qoff; a.qstat = qd->qstat.as<LqQStat>(); a.qlen = qd->reads.len.as<uint32_t>(); a.tlen = ix->tlen.as<uint32_t>();
            a.first = qd->first.as<uint64_t>(); a.lambda = qd->lambda.as<uint64_t>(); a.lambda2 = qd->lambda2.as<uint64_t>(); a.mcnt = qd->mcnt.as<uint32_t>();
            a.ovl = sc->ovl.as<LqOvl>(); a.n_ovl = ctr + 6; a.ovl_cap = ovl_cap; a.n_chains = ctr + 7; a.o = *opt;
            /* the DP reads t[] before writing it only through `t[j] == i` with i a seed index of this batch: clear it */
            LQ_CUDA_OK(cudaMemsetAsync(b.t, 0xff, (size_t)nb * 4, st));
            { LqProfScope ps("chain_small", st, 1, 0);
              lq_chain_small_k<<<148 * 16, 128, 0, st>>>(a); }
            { LqProfScope ps("chain", st, 1, nb * 32);
              lq_chain_k<<<148 * 16, CH_WARPS * 32, 0, st>>>(a); }
            LQ_CUDA_OK(cudaGetLastError());
            uint32_t h_ctr[4]; unsigned long long ng = 0;
            LQ_CUDA_OK(cudaMemcpyAsync(h_ctr, ctr + 4, 16, cudaMemcpyDeviceToHost, st));
            LQ_CUDA_OK(cudaMemcpyAsync(&ng, ctr + 50, 8, cudaMemcpyDeviceToHost, st));
            LQ_CUDA_OK(cudaStreamSynchronize(st));
            const uint32_t n_ovl = h_ctr[2];
            if (n_ovl > ovl_cap) { fprintf(stderr, "[lqcov] overlap list overflow (%u > %u)\n", n_ovl, ovl_cap); return -1; }
            const size_t old = ovl_out->size();
            ovl_out->resize(old + n_ovl);
            if (n_ovl) LQ_CUDA_OK(cudaMemcpyAsync(ovl_out->data() + old, sc->ovl.p, (size_t)n_ovl * sizeof(LqOvl), cudaMemcpyDeviceToHost, st));
            lq_prof_d2h((uint64_t)n_ovl * sizeof(LqOvl) + 32); lq_prof_h2d((uint64_t)(nqb + 1) * 8);
            LQ_CUDA_OK(cudaStreamSynchronize(st));
            if (stats) { stats->n_seeds += nb; stats->n_groups += ng; stats->n_chains += h_ctr[3]; stats->n_ovl += n_ovl; stats->n_batches += 1; }
        }
        seed_base += nb;
        q0 = q1;
    }
    return 0;
}

int lq_map_flag_dups(LqQueryDev *qd, int key_bits, LqDevBuf &ws, cudaStream_t st)
{
    const uint64_t n = qd->n_min;
    LQ_TRY(qd->dup.ensure(n + 16)); LQ_TRY(qd->qtied.ensure((size_t)qd->nq + 16));
    LQ_CUDA_OK(cudaMemsetAsync(qd->qtied.p, 0, (size_t)qd->nq + 16, st));
    if (n == 0) return 0;
    LqMinimizers &tmp = qd->dup_tmp; LqDevBuf &tk = qd->dup_tk, &ty = qd->dup_ty, &ts = qd->dup_ts, &hist = qd->dup_hist;
    int rc = -1;
    if (tmp.key.ensure(n * 4) == 0 && tmp.y.ensure(n * 8) == 0) {
        tmp.n = n; tmp.has_span = 0;
        cudaMemcpyAsync(tmp.key.p, qd->mins.key.p, n * 4, cudaMemcpyDeviceToDevice, st);
        lq_iota64_k<<<lq_grid(n, 256), 256, 0, st>>>(tmp.y.as<uint64_t>(), n);
        if (lq_sort_by_key(&tmp, key_bits, tk, ty, ts, hist, ws, st, 1) == 0) {
            const uint32_t *skey = tmp.key.as<uint32_t>();
            lq_dupflag_k<<<lq_grid(n, 256), 256, 0, st>>>(n, skey, tmp.y.as<uint64_t>(), qd->mins.y.as<uint64_t>(), qd->dup.as<uint8_t>());
            lq_qtied_k<<<lq_grid((size_t)qd->nq * 32, 256), 256, 0, st>>>(qd->nq, qd->first.as<uint64_t>(), qd->dup.as<uint8_t>(), qd->qtied.as<uint8_t>());
            lq_prof_count_launch(3);
            rc = cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess ? 0 : -1;
        }
    }
    return rc;
}

int lq_map_nmatch(LqQueryDev *qd, const uint32_t *h_npre, std::vector<uint32_t> *n_match, cudaStream_t st)
{
    const uint32_t nq = qd->nq;
    n_match->assign(nq, 0);
    if (nq == 0) return 0;
    LqDevBuf &out = qd->nmatch_buf; LQ_TRY(out.ensure(((size_t)nq + 2) * 4 + (h_npre ? (size_t)nq * 4 : 0)));
    LQ_CUDA_OK(cudaMemsetAsync(out.p, 0, ((size_t)nq + 2) * 4, st));
    uint32_t *d_npre = 0;
    if (h_npre) { d_npre = out.as<uint32_t>() + nq + 2; LQ_CUDA_OK(cudaMemcpyAsync(d_npre, h_npre, (size_t)nq * 4, cudaMemcpyHostToDevice, st)); }
    lq_prof_count_launch(1); lq_prof_d2h((uint64_t)nq * 4);
    lq_nmatch_k<<<lq_grid((size_t)nq * 32, 256), 256, 0, st>>>(nq, qd->first.as<uint64_t>(), qd->mcnt.as<uint32_t>(), d_npre, out.as<uint32_t>(), out.as<uint32_t>() + nq);
    LQ_CUDA_OK(cudaGetLastError());
    uint32_t sat = 0;
    LQ_CUDA_OK(cudaMemcpyAsync(n_match->data(), out.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    LQ_CUDA_OK(cudaMemcpyAsync(&sat, out.as<uint32_t>() + nq, 4, cudaMemcpyDeviceToHost, st));
    LQ_CUDA_OK(cudaStreamSynchronize(st));
    if (sat) fprintf(stderr, "[lqcov] WARNING: %u queries have a minimizer matched >= 65535 times: the reference's uint16 counters saturate "
                             "in an order-dependent way there (esterr.c:130-137); column 8 of those rows is outside the parity domain\n", sat);
    return 0;
}

int lq_map_debug_sorted_seeds(LqQueryDev *qd, const LqIndexDev *ix, const LqMapOpt *opt, int mid_occ, uint32_t q,
                              const uint32_t *h_self_off, const uint32_t *h_self_list, const uint32_t *h_qrank, const uint32_t *h_trank,
                              LqMapScratch *sc, std::vector<lq_mm128> *unsorted, std::vector<lq_mm128> *sorted, cudaStream_t st)
{
    MapTables mt; std::vector<LqQStat> hs;
    LQ_TRY(run_lookup(qd, ix, opt, mid_occ, &mt, 0, h_self_off, h_self_list, h_qrank, h_trank, sc, &hs, st));
    std::vector<uint64_t> h_first((size_t)qd->nq + 1);
    LQ_CUDA_OK(cudaMemcpyAsync(h_first.data(), qd->first.p, ((size_t)qd->nq + 1) * 8, cudaMemcpyDeviceToHost, st));
    LQ_CUDA_OK(cudaStreamSynchronize(st));
    uint64_t base = 0; for (uint32_t i = 0; i < q; ++i) base += hs[i].n_sorted;
    const uint64_t nb = hs[q].n_sorted;
    std::vector<uint64_t> h_qoff(2); h_qoff[0] = 0; h_qoff[1] = nb;
    BatchPtrs b; uint64_t *d_qoff = 0;
    LQ_TRY(seed_and_sort(qd, ix, mt, 0, q, q + 1, h_first, base, nb, h_qoff, sc, &b, &d_qoff, 0, st));
    std::vector<uint64_t> x(nb), x2(nb); std::vector<uint32_t> sq(nb), sm(nb), aq(nb), am(nb), pi(nb);
    if (nb) {
        LQ_CUDA_OK(cudaMemcpyAsync(pi.data(), b.idx, nb * 4, cudaMemcpyDeviceToHost, st));   /* sorted position -> seed number */
        LQ_CUDA_OK(cudaMemcpyAsync(sq.data(), b.s.sq, nb * 4, cudaMemcpyDeviceToHost, st));
        LQ_CUDA_OK(cudaMemcpyAsync(sm.data(), b.s.sm, nb * 4, cudaMemcpyDeviceToHost, st));
        LQ_CUDA_OK(cudaMemcpyAsync(x2.data(), b.ax, nb * 8, cudaMemcpyDeviceToHost, st));
        LQ_CUDA_OK(cudaMemcpyAsync(aq.data(), b.aq, nb * 4, cudaMemcpyDeviceToHost, st));
        LQ_CUDA_OK(cudaMemcpyAsync(am.data(), b.am, nb * 4, cudaMemcpyDeviceToHost, st));
        LQ_CUDA_OK(cudaStreamSynchronize(st));
    }
    unsorted->resize(nb); sorted->resize(nb);
    for (uint64_t i = 0; i < nb; ++i) x[pi[i] & 0x7fffffffu] = x2[i];   /* the keys were sorted in place */
    for (uint64_t i = 0; i < nb; ++i) {
        (*unsorted)[i].x = x[i]; (*unsorted)[i].y = (uint64_t)(sm[i] >> 24) << 32 | (sq[i] & 0x7fffffffu);
        (*sorted)[i].x = x2[i];  (*sorted)[i].y = (uint64_t)(am[i] >> 24) << 32 | aq[i];
    }
    return 0;
}
