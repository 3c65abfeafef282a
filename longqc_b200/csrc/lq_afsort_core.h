/* lq_afsort_core.h -- the reference's seed sort, reproduced exactly.
 *
 * lq_map_frag_mod() sorts the seeds of a query with radix_sort_128x (reference lqmap.c:238,
 * misc.c:125-126, ksort.h:84-134): an IN-PLACE, UNSTABLE MSD radix sort (8-bit digits from bit 56
 * down, insertion sort for buckets of <= 64).  Seeds with equal keys end in an order that depends on
 * the whole array, and mm_chain_dp is sensitive to it, so the permutation itself must
 * be reproduced, not just "sorted by x".
 *
 * One level of that sort, seen from above.  The array is cut into 256 destination regions R_0..R_255
 * (sizes = digit histogram).  ksort.h:109-122 repeatedly takes the element at the head of the current
 * region; the element's digit d says which region it belongs to; it is dropped at the head of R_d and the
 * element that was there is picked up in turn.  Hence:
 *   * elements are picked up from every region strictly in the region's original order;
 *   * after picking an element with digit d the next pick is from the head of R_d; if R_d is exhausted
 *     (only possible for the region k of the outer loop) the walk moves to the next non-exhausted region;
 *   * an element lands in R_d at slot (number of digit-d elements picked before it).
 * So the level is: "pick-up order = a deterministic walk over 256 queues", result = stable partition of
 * the pick-up order by digit.  lq_af_walk() runs that walk on the DIGITS ALONE (1 byte per element) and
 * returns each element's destination; the payload is then permuted in parallel.  With only two non-empty
 * regions the walk has a closed form (lq_af_two_dest(), prefix sums only).
 * Buckets of <= 64 elements are finished by insertion sort, which is stable, so their final order is
 * "by key, ties by position on entry".
 */
#ifndef LQ_AFSORT_CORE_H
#define LQ_AFSORT_CORE_H

#include "lq_common.h"

#define LQ_RS_MIN 64

/* dest[p] (0-based inside the bucket) for p in [0,n).  cnt/start: digit histogram and its exclusive
 * scan; head[256] scratch (zeroed here). */
LQ_HD void lq_af_walk(const uint8_t *dig, uint32_t n, const uint32_t *cnt, const uint32_t *start, uint32_t *head, uint32_t *dest)
{
    uint32_t k = 0, c, step, arrived_k = 0; /* arrivals into the outer-loop region lag its pick-ups by the open hole */
    for (c = 0; c < 256; ++c) head[c] = 0;
    while (k < 256 && cnt[k] == 0) ++k;
    c = k;
    for (step = 0; step < n; ++step) {
        const uint32_t p = start[c] + head[c];
        const uint32_t d = dig[p];
        ++head[c];
        /* slot = number of digit-d elements that arrived before: for d != k that is head[d] (every arrival
         * also picked one up); for d == k it is the separate arrival counter */
        if (d == k) dest[p] = start[k] + arrived_k++;
        else dest[p] = start[d] + head[d];
        c = d;
        if (c == k && head[k] == cnt[k]) { /* region k complete: open the next non-exhausted region */
            do { ++k; } while (k < 256 && head[k] == cnt[k]);
            if (k < 256) { c = k; arrived_k = head[k]; }
        }
    }
}

/* ---- the same walk with the per-region state packed for a short dependent chain (what the device runs) ----
 * The walk is one thread chasing a pointer through 256 queues, so its speed is the latency of one step.  State per region c:
 *   pb[c]    = { pos, base }: next unread position of region c (bucket-relative), and the position cache[c][0] holds
 *   cache[c] = the digits of positions base .. base+15
 * One step = one 8-byte load (pb[c]) and one byte load (the digit); the digit's own pb[] entry is both this element's
 * destination and the next step's position.  lq_afw_run() walks until every element is placed or a region runs out of cached
 * digits; the caller then refills every region's cache (on the device: the whole warp, in parallel) and calls it again. */
typedef struct
#ifdef __CUDACC__
__align__(8)
#endif
{ uint32_t x, y; } lq_afw_pb;
typedef struct { uint32_t k, c, arrived, step; } lq_afw_state;
#define LQ_AFW_CACHE 16

/* start[0..256]: region starts, start[256] = n.  Sets the state to the first non-empty region and marks every cache empty. */
LQ_HD void lq_afw_init_state(lq_afw_state *s, const uint32_t *start)
{
    uint32_t k = 0;
    while (k < 256 && start[k + 1] == start[k]) ++k;
    s->k = k; s->c = k; s->arrived = 0; s->step = 0;
}

/* returns 1 when all n elements are placed, 0 when region s->c needs a refill */
LQ_HD int lq_afw_run(lq_afw_state *s, uint32_t n, const uint32_t *start, lq_afw_pb *pb, const uint8_t *cache, uint32_t *dest)
{
    uint32_t k = s->k, c = s->c, arrived = s->arrived, step = s->step;
    uint32_t start_k = k < 256 ? start[k] : 0, end_k = k < 256 ? start[k + 1] : 0;
    lq_afw_pb cur = pb[c];
    int done = 1;
    while (step < n) {
        const uint32_t j = cur.x - cur.y;
        if (j >= LQ_AFW_CACHE) { done = 0; break; }
        const uint32_t d = cache[c * LQ_AFW_CACHE + j], p = cur.x;
        pb[c].x = p + 1;
        ++step;
        if (d != k) { cur = pb[d]; dest[p] = cur.x; c = d; }   /* slot = number of digit-d elements picked before = the next pick of region d */
        else {
            dest[p] = start_k + arrived++;                     /* arrivals into the outer-loop region lag its pick-ups by the open hole */
            c = k;
            if (pb[k].x == end_k) {                            /* region k complete: open the next non-exhausted region */
                do { ++k; } while (k < 256 && pb[k].x == start[k + 1]);
                if (k < 256) { c = k; start_k = start[k]; end_k = start[k + 1]; arrived = pb[k].x - start_k; }
                else c = 0;
            }
            cur = pb[c];
        }
    }
    s->k = k; s->c = c; s->arrived = arrived; s->step = step;
    return done;
}

/* host form of the refill (the device does it with two aligned 16-byte loads per region) */
LQ_HD void lq_afw_refill_host(const uint8_t *dig, uint32_t n, const uint32_t *start, lq_afw_pb *pb, uint8_t *cache)
{
    for (uint32_t r = 0; r < 256; ++r) {
        const uint32_t p = pb[r].x;
        if (p >= start[r + 1] || p == pb[r].y) continue;
        for (uint32_t j = 0; j < LQ_AFW_CACHE; ++j) cache[r * LQ_AFW_CACHE + j] = p + j < n ? dig[p + j] : 0;
        pb[r].y = p;
    }
}

/* ---- packed form: position and upcoming digits of a region in ONE 16-byte word, so that a step is a single load ----
 * st[c] = { x: next unread position of region c (bucket-relative),
 *           y, z, w: the digits of positions x, x+1, ... (byte 0 of y first); top byte of w = how many of them are valid (<= 11) }
 * The digit's own entry gives the element's destination AND is the next step's load.  This is what lq_af_walk_k runs with one LANE
 * per bucket: the dependent chain of a step is load -> and -> address. */
typedef struct
#ifdef __CUDACC__
__align__(16)
#endif
{ uint32_t x, y, z, w; } lq_afp_st;
typedef struct { uint32_t k, c, arrived, step, start_k, end_k; } lq_afp_walk;
#define LQ_AFP_DIG 11
#ifdef __CUDA_ARCH__
#define LQ_FUNNEL_R8(lo_, hi_) __funnelshift_r((lo_), (hi_), 8)
#else
#define LQ_FUNNEL_R8(lo_, hi_) ((uint32_t)((((uint64_t)(hi_) << 32) | (uint64_t)(lo_)) >> 8))
#endif

LQ_HD void lq_afp_init(lq_afp_walk *s, const uint32_t *start)
{
    uint32_t k = 0;
    while (k < 256 && start[k + 1] == start[k]) ++k;
    s->k = k; s->c = k < 256 ? k : 0; s->arrived = 0; s->step = 0;
    s->start_k = k < 256 ? start[k] : 0; s->end_k = k < 256 ? start[k + 1] : 0;
}

/* returns 1 when all n elements are placed, 0 when region s->c has no cached digit left (refill, then call again).
 * Output is the walk itself: the t-th pick-up took position ord[t] and dropped it at slot[t].  Both streams are written in pick-up
 * order, and in that order the reads and the writes of the payload permutation advance sequentially inside each of the 256
 * regions -- which is what lets the permutation run at full sector efficiency afterwards (lq_af_big_k<true>).
 * The state of the digit's region is loaded as soon as the digit is known (Sn), before this step's bookkeeping: the dependent chain
 * of a step is load -> and -> address -> load, everything else sits in the shadow of the load.  Sn is stale only when the digit
 * names the region just read (d == c); the updated copy in registers is used then. */
LQ_HD int lq_afp_run(lq_afp_walk *s, uint32_t n, const uint32_t *start, lq_afp_st *st, uint32_t *ord, uint32_t *slot)
{
    uint32_t k = s->k, c = s->c, arrived = s->arrived, step = s->step, start_k = s->start_k, end_k = s->end_k;
    int done = 1;
    if (step < n) {
        lq_afp_st S = st[c];
        for (;;) {
            const uint32_t left = S.w >> 24;
            if (left == 0) { done = 0; break; }
            const uint32_t d = S.y & 255u, p = S.x;
            const lq_afp_st Sn = st[d];
            S.x = p + 1; S.y = LQ_FUNNEL_R8(S.y, S.z); S.z = LQ_FUNNEL_R8(S.z, S.w); S.w = ((S.w >> 8) & 0xffffu) | (left - 1) << 24;
            st[c] = S;
            ord[step] = p;
            if (d != k) {
                if (d != c) S = Sn;
                slot[step] = S.x;                                  /* lands where its region's next pick-up is taken from */
                c = d;
            } else {
                slot[step] = start_k + arrived++;                  /* arrivals into the outer-loop region lag its pick-ups by the open hole */
                if (c != k) S = Sn;
                c = k;
                if (S.x == end_k) {                                /* region k complete: open the next non-exhausted region */
                    do { ++k; } while (k < 256 && st[k].x == start[k + 1]);
                    if (k < 256) { c = k; start_k = start[k]; end_k = start[k + 1]; arrived = st[k].x - start_k; }
                    else c = 0;
                    S = st[c];
                }
            }
            if (++step >= n) break;
        }
    }
    s->k = k; s->c = c; s->arrived = arrived; s->step = step; s->start_k = start_k; s->end_k = end_k;
    return done;
}

/* refill rule of one region (host form; the device loads the bytes with two aligned 16-byte loads): a region that moved since its
 * last refill gets the LQ_AFP_DIG digits from its position on.  Digits past the region's end are cached too (the next region's, or
 * padding): the walk never reads them, because a region is picked from exactly as often as it has elements. */
LQ_HD void lq_afp_refill_host(const uint8_t *dig, const uint32_t *start, lq_afp_st *st, uint32_t r)
{
    lq_afp_st S = st[r];
    const uint32_t n = start[256];
    if ((S.w >> 24) >= LQ_AFP_DIG) return;
    uint8_t b[12]; uint32_t j;
    for (j = 0; j < 12; ++j) b[j] = j < LQ_AFP_DIG && S.x + j < n ? dig[S.x + j] : 0;
    S.y = (uint32_t)b[0] | (uint32_t)b[1] << 8 | (uint32_t)b[2] << 16 | (uint32_t)b[3] << 24;
    S.z = (uint32_t)b[4] | (uint32_t)b[5] << 8 | (uint32_t)b[6] << 16 | (uint32_t)b[7] << 24;
    S.w = (uint32_t)b[8] | (uint32_t)b[9] << 8 | (uint32_t)b[10] << 16 | (uint32_t)LQ_AFP_DIG << 24;
    st[r] = S;
}

/* Closed form for exactly two non-empty digits d0 < d1 (regions [0,n0) and [n0,n)).
 * fr[p]  = 1 if p is "foreign" (p < n0 with digit d1, or p >= n0 with digit d0)
 * rk[p]  = number of foreign positions before p within p's own region (exclusive rank)
 * P[i]/Z[i] = i-th foreign position of region 0 / region 1.
 *   native of region 0            -> stays
 *   P[i]                          -> Z[i-1] + 1   (Z[-1] = n0 - 1)
 *   Z[i]                          -> P[i]
 *   native q of region 1          -> q + 1 if some Z lies after q, else q */
LQ_HD uint32_t lq_af_two_dest(uint32_t p, uint32_t n0, int foreign, uint32_t rk, uint32_t n_for, const uint32_t *P, const uint32_t *Z)
{
    if (p < n0) {
        if (!foreign) return p;
        return rk == 0 ? n0 : Z[rk - 1] + 1;
    }
    if (foreign) return P[rk];
    return rk < n_for ? p + 1 : p;
}

/* stable insertion sort of idx[0..n) by key[idx] (ksort.h:88-98) */
LQ_HD void lq_af_insertion(uint32_t *idx, uint32_t n, const uint64_t *key)
{
    for (uint32_t i = 1; i < n; ++i) {
        const uint32_t t = idx[i]; const uint64_t kt = key[t];
        if (kt < key[idx[i - 1]]) {
            uint32_t j = i;
            while (j > 0 && kt < key[idx[j - 1]]) { idx[j] = idx[j - 1]; --j; }
            idx[j] = t;
        }
    }
}

/* the same on (key, value) pairs held in place: what the device runs once the keys travel with the elements */
LQ_HD void lq_af_insertion_kv(uint64_t *key, uint32_t *val, uint32_t n)
{
    for (uint32_t i = 1; i < n; ++i) {
        const uint64_t kt = key[i]; const uint32_t vt = val[i];
        if (kt < key[i - 1]) {
            uint32_t j = i;
            while (j > 0 && kt < key[j - 1]) { key[j] = key[j - 1]; val[j] = val[j - 1]; --j; }
            key[j] = kt; val[j] = vt;
        }
    }
}

#endif
