/* lq_chain_core.h -- scoring pieces of mm_chain_dp (reference chain.c:15-20, 52-68) shared by the
 * chaining kernel and its CPU test harness. */
#ifndef LQ_CHAIN_CORE_H
#define LQ_CHAIN_CORE_H
#include "lq_common.h"

LQ_HD int lq_ilog2_32(uint32_t v) /* chain.c:15-20, v > 0 */
{
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}

/* Score of extending the chain ending at anchor j by anchor i (chain.c:52-68, n_segs == 1, !is_cdna).
 * dr = target distance (same read and strand), dq = query distance.  Returns 0 when j is skipped
 * (`continue`), else 1 with *sc = gain (f[j] not yet added). avg_span is the float of chain.c:38. */
LQ_HD int lq_chain_gain(int64_t dr, int32_t dq, int32_t span_i, int max_dist_x, int max_dist_y, int bw, float avg_span, int32_t *sc)
{
    if (dr == 0 || dq <= 0) return 0;
    if (dq > max_dist_y || dq > max_dist_x) return 0;
    const int32_t dd = dr > dq ? (int32_t)(dr - dq) : (int32_t)(dq - dr);
    if (dd > bw) return 0;
    const int32_t min_d = dq < dr ? dq : (int32_t)dr;
    int32_t s = min_d > span_i ? span_i : min_d;
    const int32_t lg = dd ? lq_ilog2_32((uint32_t)dd) : 0;
    s -= (int)(dd * .01 * avg_span) + (lg >> 1); /* double * double * (double)float, truncated toward zero */
    *sc = s;
    return 1;
}
#endif
