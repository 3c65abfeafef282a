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
#define LQ_UNLIKELY(x) __builtin_expect(!!(x), 0)   /* rare blocks out of line: the walk's common path then has no taken branch */

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
LQ_HD void lq_afp_refill_host(const uint8_t *dig, const uint32_t *start, lq_afp_st *st, uint32_t r)   /* st: the states with stride 1; st + r * (stride - 1) for strided ones */
{
    lq_afp_st S = st[r];
    const uint32_t n = start[256];
    if ((S.w >> 24) >= LQ_AFP_DIG) return;
    uint8_t b[12]; uint32_t j;
    for (j = 0; j < 12; ++j) b[j] = j < LQ_AFP_DIG && S.x + j < n ? dig[S.x + j] : 0;
    S.y = (uint32_t)b[0] | (uint32_t)b[1] << 8 | (uint32_t)b[2] << 16 | (uint32_t)b[3] << 24;
    S.z = (uint32_t)b[4] | (uint32_t)b[5] << 8 | (uint32_t)b[6] << 16 | (uint32_t)b[7] << 24;
    /* bit 7 of the count: the cached digits reach the region's end -- it never needs another refill (and never counts down to 0) */
    S.w = (uint32_t)b[8] | (uint32_t)b[9] << 8 | (uint32_t)b[10] << 16 | (uint32_t)(LQ_AFP_DIG | (S.x + LQ_AFP_DIG >= start[r + 1] ? 0x80u : 0u)) << 24;
    st[r] = S;
}

/* ---- the walk reduced to what only it can compute: the pick-up DIGIT stream (what lq_af_walk3_k / lq_af_walkf_k run) ----
 * Everything else about a level follows from that stream in parallel (lq_afq_expand, on the device lq_af_place_k):
 *   - the t-th pick-up, of digit d, lands at slot start[d] + (number of digit-d pick-ups before t): a stable partition by digit;
 *   - it is taken from the slot the pick-up before it was dropped at (in-place exchange), one further when that one closed a cycle
 *     of the outer-loop region k (the hole k left open lags its pick-ups by one), or from the recorded position when step t opens
 *     a new outer-loop region.
 * So the sequential part writes ONE BYTE per element (four pick-ups per 32-bit store) and no positions at all; the phase list
 * (when each outer-loop region opened, and where) has at most 256 entries per bucket. */
typedef struct { uint32_t t, p; } lq_afq_phase;          /* region k became the outer-loop region at pick-up t, its next unread position was p; t = ~0: never */
typedef struct { uint32_t k, c, step, rem_k, acc; } lq_afq_walk;   /* rem_k: elements of the outer-loop region k still to be picked up */

LQ_HD void lq_afq_init(lq_afq_walk *s, const uint32_t *start, lq_afq_phase *ph /* [256], all t = ~0 on entry */)
{
    uint32_t k = 0;
    while (k < 256 && start[k + 1] == start[k]) ++k;
    s->k = k; s->c = k < 256 ? k : 0; s->step = 0; s->acc = 0; s->rem_k = k < 256 ? start[k + 1] - start[k] : 0;
    if (k < 256) { ph[k].t = 0; ph[k].p = start[k]; }
}

/* st[r * stride]: packed state of region r as in lq_afp_st (x = next unread position, then up to 11 digits and their number; bit 7
 * of the number: the cached digits reach the region's end).  Every region with elements left holds at least one digit when this is
 * called (the refill tops up every region that is running low).  seq32: the digit stream, four pick-ups per word (little endian).
 * Returns 1 when all n elements are picked, 0 when a region has just used its last cached digit (refill, call again).
 * The critical path of a pick-up is ONE 16-byte load: the state of the digit's region is requested as soon as the digit is known,
 * before the state just read is shifted, stored and checked (a digit naming the region just read takes the shifted copy instead);
 * "region k complete" is a counter, not a comparison on the freshly loaded state. */
LQ_HD int lq_afq_run(lq_afq_walk *s, uint32_t n, const uint32_t *start, lq_afp_st *st, uint32_t stride, uint32_t *seq32, lq_afq_phase *ph)
{
    uint32_t k = s->k, c = s->c, step = s->step, rem_k = s->rem_k, acc = s->acc;
    uint32_t need = 0;
    if (step < n) {
        lq_afp_st S = st[c * stride];
#define LQ_AFQ_STEP(J) { \
            const uint32_t d = S.y & 255u; \
            lq_afp_st Sn = st[d * stride]; \
            const uint32_t left = (S.w >> 24) - 1u; \
            S.x += 1; S.y = LQ_FUNNEL_R8(S.y, S.z); S.z = LQ_FUNNEL_R8(S.z, S.w); S.w = ((S.w >> 8) & 0xffffu) | left << 24; \
            st[c * stride] = S; \
            need |= (uint32_t)(left == 0u); \
            rem_k -= (uint32_t)(c == k); \
            acc |= d << (8 * (J)); \
            if ((J) == 3) { seq32[step >> 2] = acc; acc = 0; } \
            if (d == c) Sn = S; \
            c = d; S = Sn; \
            if (LQ_UNLIKELY(rem_k == 0u && d == k)) {             /* region k complete: open the next non-exhausted region */ \
                do { ++k; } while (k < 256 && st[k * stride].x == start[k + 1]); \
                if (k < 256) { c = k; S = st[c * stride]; rem_k = start[k + 1] - S.x; ph[k].t = step + 1; ph[k].p = S.x; } \
                else { c = 0; S = st[0]; } \
            } \
            ++step; \
            if (LQ_UNLIKELY(need || step >= n)) break; }
        for (;;) {
            switch (step & 3u) {      /* resume in the middle of a word after a refill */
            case 0: LQ_AFQ_STEP(0)
            /* fall through */
            case 1: LQ_AFQ_STEP(1)
            /* fall through */
            case 2: LQ_AFQ_STEP(2)
            /* fall through */
            default: LQ_AFQ_STEP(3)
            }
            if (need || step >= n) break;
        }
#undef LQ_AFQ_STEP
        if (step >= n && (step & 3u)) seq32[step >> 2] = acc;     /* the last, partial word */
    }
    s->k = k; s->c = c; s->step = step; s->rem_k = rem_k; s->acc = acc;
    return step >= n;
}

/* ord[t] / slot[t] of every pick-up from the digit stream (host reference of lq_af_place_k); run[256] scratch */
LQ_HD void lq_afq_expand(const uint8_t *seq, uint32_t n, const uint32_t *start, const lq_afq_phase *ph, uint32_t *run, uint32_t *ord, uint32_t *slot)
{
    uint32_t t, k = 0, kn, prev_slot = 0, prev_d = 0, prev_k = 0;
    for (t = 0; t < 256; ++t) run[t] = 0;
    while (k < 256 && ph[k].t != 0) ++k;                           /* the region the walk starts in */
    kn = k + 1; while (kn < 256 && ph[kn].t == 0xffffffffu) ++kn;   /* the next region that ever opens (they open in ascending order) */
    for (t = 0; t < n; ++t) {
        const uint32_t d = seq[t];
        if (kn < 256 && ph[kn].t == t) { k = kn; ord[t] = ph[k].p; ++kn; while (kn < 256 && ph[kn].t == 0xffffffffu) ++kn; }
        else if (t == 0) ord[t] = ph[k].p;
        else ord[t] = prev_slot + (prev_d == prev_k ? 1u : 0u);
        slot[t] = start[d] + run[d]++;
        prev_slot = slot[t]; prev_d = d; prev_k = k;
    }
}

/* ---- the same walk for levels with FEW regions (all digits < 16: the rid >> 16 byte once a part holds more than 131 072 reads).
 * Such walks are long (a whole (query, strand) bucket: 10^5..10^6 pick-ups) and nothing hides the latency of a lone walker, so the
 * step is cut down to ONE load on the critical path.  Per region, in the walker's cache (words, interleaved by `stride`):
 *   word 0   the WINDOW: the region's next (up to) 7 digits, 4 bits each, lowest first, under a marker bit -- 1 = empty
 *   word 1   how many queue words have been moved into the window
 *   word 2   bucket-relative position of the first digit of queue word 0
 *   word 3.. the QUEUE: LQ_AFR_QW words in window format, then a sentinel: 0 = "more digits in global memory", 1 = "region ends here"
 * A pick-up is: load the window of region c, d = w & 15, store w >> 4 -- or, when that leaves the window empty, the next queue word
 * (eagerly: the next visit again finds its digit with one load).  A sentinel 0 ends the round: every region's queue is then rebuilt
 * from its next unread position (lq_afr_refill_*).  Digits cached beyond a region's end are never consumed: a region is visited
 * exactly as often as it has elements. */
#define LQ_AFR_R 16
#define LQ_AFR_BLK 60
#define LQ_AFR_QW (LQ_AFR_BLK - 4)          /* 56 queue words = 392 digits per region and round */
typedef struct { uint32_t k, c, step, acc, rem_k; } lq_afr_walk;

LQ_HD uint32_t lq_afr_digits_left(uint32_t win)   /* digits under the marker bit (win >= 1) */
{
#ifdef __CUDA_ARCH__
    return (31u - (uint32_t)__clz((int)win)) >> 2;
#else
    uint32_t n = 0; while (win > 15u) { win >>= 4; ++n; } return n;
#endif
}

/* next unread position of region r (whose elements end at `end`) from its three state words */
LQ_HD uint32_t lq_afr_next(uint32_t win, uint32_t q, uint32_t qbase, uint32_t end)
{
    if (win == 0u) return qbase + 7u * q;                         /* the whole queue has been consumed */
    const uint32_t wstart = qbase + 7u * (q - 1u);                /* first digit of the word now in the window */
    const uint32_t cntw = wstart < end ? (end - wstart < 7u ? end - wstart : 7u) : 0u;
    return wstart + cntw - lq_afr_digits_left(win);
}

LQ_HD void lq_afr_init(lq_afr_walk *s, const uint32_t *start, lq_afq_phase *ph)
{
    uint32_t k = 0;
    while (k < LQ_AFR_R && start[k + 1] == start[k]) ++k;
    s->k = k; s->c = k < LQ_AFR_R ? k : 0; s->step = 0; s->acc = 0; s->rem_k = k < LQ_AFR_R ? start[k + 1] - start[k] : 0;
    if (k < LQ_AFR_R) { ph[k].t = 0; ph[k].p = start[k]; }
}

/* queue word j (0..LQ_AFR_QW) of a region whose next unread position is nxt and whose elements end at `end` */
LQ_HD uint32_t lq_afr_queue_word(const uint8_t *dig, uint32_t nxt, uint32_t end, uint32_t j)
{
    const uint32_t p = nxt + 7u * j;
    if (j == LQ_AFR_QW) return p < end ? 0u : 1u;
    const uint32_t cnt = p < end ? (end - p < 7u ? end - p : 7u) : 0u;
    uint32_t v = 1u << (4u * cnt);
    for (uint32_t b = 0; b < cnt; ++b) v |= (uint32_t)(dig[p + b] & 15u) << (4u * b);
    return v;
}

/* M: word access to the walk's cache, M::ld(i) / M::st(i, v) with i = (r * LQ_AFR_BLK + j) * stride -- on the device explicit
 * shared-space loads and stores (LqSmemWords in lq_map.cu), on the host a plain array.  Returns 1 when the bucket is finished,
 * 0 when a queue ran dry (refill every region and call again). */
template <class M>
LQ_HD int lq_afr_run(lq_afr_walk *s, uint32_t n, const uint32_t *start, M mem, uint32_t stride, uint32_t *seq32, lq_afq_phase *ph)
{
    uint32_t k = s->k, c = s->c, step = s->step, acc = s->acc, rem_k = s->rem_k;
    const uint32_t RB = LQ_AFR_BLK * stride;
    uint32_t need = 0;
#define LQ_AFR_STEP(J, T) { \
        const uint32_t ob = c * RB; \
        const uint32_t w = mem.ld(ob); \
        const uint32_t d = w & 15u; \
        uint32_t wn = w >> 4; \
        rem_k -= (c == k); \
        acc |= d << (8 * (J)); \
        if (LQ_UNLIKELY(wn == 1u)) {                              /* window empty: the next queue word, now */ \
            const uint32_t q = mem.ld(ob + stride); \
            wn = mem.ld(ob + (3u + q) * stride); \
            if (wn != 0u) mem.st(ob + stride, q + 1u); else need = 1u; \
        } \
        mem.st(ob, wn); \
        c = d; \
        if (LQ_UNLIKELY(rem_k == 0u && d == k)) {                 /* region k complete: open the next non-exhausted region */ \
            for (;;) { \
                ++k; \
                if (k >= LQ_AFR_R) break; \
                const uint32_t kb = k * RB; \
                const uint32_t nxt = lq_afr_next(mem.ld(kb), mem.ld(kb + stride), mem.ld(kb + 2u * stride), start[k + 1]); \
                if (nxt < start[k + 1]) { c = k; rem_k = start[k + 1] - nxt; ph[k].t = (T) + 1; ph[k].p = nxt; break; } \
            } \
            if (k >= LQ_AFR_R) c = 0; \
        } }
    while (step < n && !need) {
        if ((step & 3u) == 0u && step + 4 <= n) {                 /* whole words of the digit stream: static byte lanes */
            LQ_AFR_STEP(0, step)
            if (LQ_UNLIKELY(need)) { step += 1; break; }
            LQ_AFR_STEP(1, step + 1)
            if (LQ_UNLIKELY(need)) { step += 2; break; }
            LQ_AFR_STEP(2, step + 2)
            if (LQ_UNLIKELY(need)) { step += 3; break; }
            LQ_AFR_STEP(3, step + 3)
            seq32[step >> 2] = acc; acc = 0;
            step += 4;
        } else {
            switch (step & 3u) {
                case 0: LQ_AFR_STEP(0, step) break;
                case 1: LQ_AFR_STEP(1, step) break;
                case 2: LQ_AFR_STEP(2, step) break;
                default: LQ_AFR_STEP(3, step) break;
            }
            ++step;
            if ((step & 3u) == 0u) { seq32[(step - 1) >> 2] = acc; acc = 0; }
        }
    }
#undef LQ_AFR_STEP
    const int done = step >= n;
    if (done && (step & 3u)) seq32[step >> 2] = acc;
    s->k = k; s->c = c; s->step = step; s->acc = acc; s->rem_k = rem_k;
    return done;
}

/* plain-array word access (host check) */
struct lq_afr_host_words { uint32_t *a; LQ_HD uint32_t ld(uint32_t i) const { return a[i]; } LQ_HD void st(uint32_t i, uint32_t v) const { a[i] = v; } };

/* the cache before the first refill: nothing consumed, queues empty */
LQ_HD void lq_afr_cache_init_host(uint32_t *blk, const uint32_t *start)
{
    for (uint32_t r = 0; r < LQ_AFR_R; ++r) { blk[r * LQ_AFR_BLK] = 0; blk[r * LQ_AFR_BLK + 1] = 0; blk[r * LQ_AFR_BLK + 2] = start[r]; }
}

/* host form of the few-region refill: every region's queue restarts at its next unread position */
LQ_HD void lq_afr_refill_host(const uint8_t *dig, const uint32_t *start, uint32_t *blk)
{
    for (uint32_t r = 0; r < LQ_AFR_R; ++r) {
        uint32_t *b = blk + r * LQ_AFR_BLK;
        const uint32_t end = start[r + 1], nxt = lq_afr_next(b[0], b[1], b[2], end);
        for (uint32_t j = 0; j <= LQ_AFR_QW; ++j) b[3 + j] = lq_afr_queue_word(dig, nxt, end, j);
        b[2] = nxt; b[1] = 1; b[0] = b[3];
    }
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
