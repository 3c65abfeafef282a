/* lq_sketch_lane_core.h -- mm_sketch() (reference sketch.c:76-142), 16 bases per lane.
 *
 * The rolling kernel (lq_sketch.cu, rk_scan) lets a thread run the reference scan over 64 bases after a 32-base warm-up that
 * only serves to bring its state to the reference's.  Here a thread owns ONE packed word (16 bases) and starts from the state
 * itself, handed over by the thread before it -- no warm-up, no divergence between lanes beyond the scan's own branches:
 *
 *   (1) the candidates of its 16 bases (hash of the canonical k-mer, strand, "not its own reverse complement") depend only on
 *       the word and on the k bases before it: every thread computes them independently (lq_rl_cands);
 *   (2) the scan state before base i0 is { ring = the last w pushes, running minimum = rightmost minimum of the ring
 *       (lq_sketch_core.h, fact 1), run counter saturated } whenever
 *           A. the read has at least 64 bases before i0 and none of [i0-64, end of the segment) is ambiguous,
 *           B. at least w+k of the 32 bases before i0 push (are not palindromic k-mers): the run counter is then >= w+k at i0,
 *              every gate of sketch.c:116-137 is open, and each of the last w pushes was made with >= k pushes behind it
 *              in the same stretch, i.e. carries a real candidate,
 *           C. at least w of the 16 bases before i0 push: the ring is then the tail of the PREVIOUS thread's own candidates
 *       (lq_rl_inject_ok).  The previous thread's tail (lq_rl_tail) comes over by warp shuffle; two halo lanes per warp
 *       recompute the 32 bases before the warp's first segment;
 *   (3) lq_rl_steady() then is the reference's steady state (sketch.c:122-137 with every gate open) over <= 16 bases.
 * A segment that fails A-C (read starts, the neighbourhood of an N, runs of palindromic k-mers: ~1 % of real reads) takes the
 * rolling scan with its certified warm-up instead.  2k <= 30 (k <= 15) so that a candidate fits 32 bits next to the sentinel.
 *
 * Host+device: tests/test_hostcheck.py runs these functions segment by segment against the oracle.
 */
#ifndef LQ_SKETCH_LANE_CORE_H
#define LQ_SKETCH_LANE_CORE_H

#include "lq_sketch_core.h"

#define LQ_RL_SEG 16
#define LQ_RL_BACK 64               /* unambiguous bases required before a segment that starts from a handed-over state */
#define LQ_RL_MAXH 0xffffffffu

/* reverse the order of the 2-bit groups of a 32-bit word */
LQ_HD uint32_t lq_rev2_32(uint32_t v)
{
    v = (v >> 2 & 0x33333333u) | (v & 0x33333333u) << 2;
    v = (v >> 4 & 0x0F0F0F0Fu) | (v & 0x0F0F0F0Fu) << 4;
    v = (v >> 8 & 0x00FF00FFu) | (v & 0x00FF00FFu) << 8;
    return v >> 16 | v << 16;
}

/* candidates of the 16 bases packed in `cur` (base j in bits 2j..2j+1), given `prev` = the 16 bases before them.
 * cx[j] = hash of the canonical k-mer ending at base j; bit j of *zmask = its strand (sketch.c:109);
 * bit j of *okmask = the k-mer differs from its reverse complement (sketch.c:107: it pushes). */
LQ_HD void lq_rl_cands(uint32_t prev, uint32_t cur, int k, uint32_t *cx, uint32_t *zmask, uint32_t *okmask)
{
    const uint32_t mask = (1u << 2 * k) - 1, top = 2u * (uint32_t)(k - 1);
    const uint32_t le = prev >> (2 * (16 - k));             /* the k bases before the segment, oldest in the low bits */
    uint32_t rv = ~le & mask;                               /* sketch.c:106 */
    uint32_t fw = lq_rev2_32(le) >> (32 - 2 * k);           /* sketch.c:105: newest base in the low bits */
    uint32_t zm = 0, ok = 0;
#ifdef __CUDA_ARCH__
    #pragma unroll
#endif
    for (int j = 0; j < LQ_RL_SEG; ++j) {
        const uint32_t c = (cur >> (2 * j)) & 3u;
        fw = (fw << 2 | c) & mask;
        rv = rv >> 2 | (3u ^ c) << top;
        const uint32_t z = fw < rv ? 0u : 1u;
        cx[j] = lq_hash32(z ? rv : fw, mask);
        zm |= z << j;
        ok |= (fw != rv ? 1u : 0u) << j;
    }
    *zmask = zm; *okmask = ok;
}

LQ_HD int lq_rl_popc16(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return __popc(v & 0xffffu);
#else
    int n = 0; v &= 0xffffu; while (v) { v &= v - 1; ++n; } return n;
#endif
}

/* the last W pushes of a segment, oldest first, as (hash, pos<<1|strand); meaningful when the segment has >= W pushes.
 * i0 = read position of the segment's first base. */
template <int W>
LQ_HD void lq_rl_tail(const uint32_t *cx, uint32_t zmask, uint32_t okmask, int i0, uint32_t *tx, uint32_t *tp)
{
    if ((okmask & 0xffffu) == 0xffffu) {
#ifdef __CUDA_ARCH__
        #pragma unroll
#endif
        for (int t = 0; t < W; ++t) { const int j = LQ_RL_SEG - W + t; tx[t] = cx[j]; tp[t] = (uint32_t)(i0 + j) << 1 | ((zmask >> j) & 1u); }
    } else {
#ifdef __CUDA_ARCH__
        #pragma unroll
#endif
        for (int t = 0; t < W; ++t) { tx[t] = LQ_RL_MAXH; tp[t] = LQ_RL_MAXH; }
#ifdef __CUDA_ARCH__
        #pragma unroll
#endif
        for (int j = 0; j < LQ_RL_SEG; ++j) if ((okmask >> j) & 1u) {
#ifdef __CUDA_ARCH__
            #pragma unroll
#endif
            for (int t = 0; t + 1 < W; ++t) { tx[t] = tx[t + 1]; tp[t] = tp[t + 1]; }
            tx[W - 1] = cx[j]; tp[W - 1] = (uint32_t)(i0 + j) << 1 | ((zmask >> j) & 1u);
        }
    }
}

/* conditions A-C above.  g = global base index of the segment's first base, nseg = its bases inside the read,
 * ok_prev1 / ok_prev2 = the push masks of the one / two segments before it. */
LQ_HD int lq_rl_inject_ok(const uint32_t *nm, uint64_t g, int i0, int nseg, uint32_t ok_prev1, uint32_t ok_prev2, int w, int k)
{
    if (i0 < LQ_RL_BACK || k > 15 || w > LQ_RL_SEG) return 0;
    if (lq_amb_any(nm, g - LQ_RL_BACK, g + (uint64_t)nseg - 1)) return 0;
    const int p1 = lq_rl_popc16(ok_prev1), p2 = lq_rl_popc16(ok_prev2);
    return p1 >= w && p1 + p2 >= w + k;
}

/* sketch.c:122-137 with every gate open over the (<= 16) bases of a segment.  cx[j * cxs]: the candidates (the device keeps them
 * in shared memory so that this loop stays a loop: unrolled 16 times, with its rarely-taken branches, it outgrows the instruction
 * cache).  wx/wp: the ring on entry, oldest first (in/out); okmask must already be clipped to the bases inside the read;
 * `last`: the segment holds the read's last base (sketch.c:140). */
template <int W, class Sink>
LQ_HD void lq_rl_steady(const uint32_t *cx, int cxs, uint32_t zmask, uint32_t okmask, int i0, uint32_t *wx, uint32_t *wp, int last, Sink &sink)
{
    uint32_t mx = LQ_RL_MAXH, mp = LQ_RL_MAXH; int mi = 0;
#ifdef __CUDA_ARCH__
    #pragma unroll
#endif
    for (int t = 0; t < W; ++t) if (wx[t] <= mx) { mx = wx[t]; mp = wp[t]; mi = t; }   /* rightmost minimum of the ring */
#ifdef __CUDA_ARCH__
    #pragma unroll 1
#endif
    for (int j = 0; j < LQ_RL_SEG; ++j) {
        if (!((okmask >> j) & 1u)) continue;                                         /* sketch.c:107, or past the read's end */
        const uint32_t c = cx[j * cxs], cp = (uint32_t)(i0 + j) << 1 | ((zmask >> j) & 1u);
        if (c <= mx) {                                                               /* sketch.c:122-124 */
            sink(mx, mp);
            mx = c; mp = cp; mi = W;
        } else if (mi == 0) {                                                        /* sketch.c:125-137: the minimum leaves the window */
            sink(mx, mp);
            mx = LQ_RL_MAXH; mp = LQ_RL_MAXH;
#ifdef __CUDA_ARCH__
            #pragma unroll
#endif
            for (int t = 1; t < W; ++t) if (mx >= wx[t]) { mx = wx[t]; mp = wp[t]; mi = t; }
            if (mx >= c) { mx = c; mp = cp; mi = W; }
#ifdef __CUDA_ARCH__
            #pragma unroll
#endif
            for (int t = 1; t < W; ++t) if (wx[t] == mx && wp[t] != mp) sink(wx[t], wp[t]);
            if (c == mx && cp != mp) sink(c, cp);
        }
#ifdef __CUDA_ARCH__
        #pragma unroll
#endif
        for (int t = 0; t + 1 < W; ++t) { wx[t] = wx[t + 1]; wp[t] = wp[t + 1]; }
        wx[W - 1] = c; wp[W - 1] = cp;
        --mi;
    }
    if (last && mx != LQ_RL_MAXH) sink(mx, mp);                                       /* sketch.c:140-141 */
}

#endif
