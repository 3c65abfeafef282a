/* lq_sdust_core.h -- symmetric DUST masked length of one read (reference sdust.c:72-185), host+device.
 *
 * The state is local to the last W bases: triplet deque (<= W-2 words), counts cw/cv of the whole
 * window / of its longest "clean" suffix, running scores rw/rv, and the list P of perfect intervals of
 * the current window (descending start).  Only the MASKED LENGTH is needed by the sdust table
 * (sdust.c:205-209), so merged intervals are folded into a running sum instead of a result vector.
 * N / end of read flushes P but does NOT clear the deque or the counts (sdust.c:158-162).
 */
#ifndef LQ_SDUST_CORE_H
#define LQ_SDUST_CORE_H
#include "lq_common.h"
#include "lq_sketch_core.h" /* lq_nt4 */

/* Perfect intervals alive at one time: at most (W-2)^2.  Proof: an entry is inserted by find_perfect with start' = ii + start,
 * 0 <= ii <= W-3 (one entry per ii and call), and lives while start' >= the current `start`.  `start` is frozen for the first W-2 calls
 * after a reset of l (sdust.c:149-151: l <= W) and then grows by one per call, so a call made d starts ago still owns at most W-2-d
 * entries: the total peaks at (W-2)^2 when the frozen stretch ends.  The bound is reached in practice only after an N, when the
 * deque keeps its stale words (sdust.c:158-162): 'A'*200 + 'N' + 'A'*200 needs 3596 of the 3844 entries at W = 64.
 * The caller provides the storage (4 ints per entry) so that device threads can keep it in a global scratch slice. */
#define LQ_SD_PCAP(W) (((W) - 2) * ((W) - 2) + 2)

template <class Sink>
struct lq_sd_state {
    uint8_t win[64], cw[64], cv[64]; int w_front, w_n;
    int rw, rv, L;
    int *Ps, *Pf, *Pr, *Pl, nP, capP;
    int overflow;
    int now;            /* the base being processed: sinks that attribute an interval to the step that gave it up read it */
    Sink *sink;
};

template <class Sink>
LQ_HD void lq_sd_save(lq_sd_state<Sink> *s, int start) /* sdust.c:94-108 */
{
    if (s->nP == 0 || s->Ps[s->nP - 1] >= start) return;
    (*s->sink)(s->Ps[s->nP - 1], s->Pf[s->nP - 1], s->now);
    int i = s->nP - 1;
    while (i >= 0 && s->Ps[i] < start) --i;
    s->nP = i + 1;
}

/* The scan of bases [from, to) of a read (to > l_seq: followed by the end-of-read flush), starting from an EMPTY state at `from`.
 * from == 0 is sdust_core() itself.  A later `from` reproduces every step from `lo` on exactly when the cold start lies far enough
 * before lo (lq_sd_warm_from): the deque of <= W-2 triplets, the counts of the window and of its clean suffix are functions of the
 * last W-2 triplets pushed, and the interval list only holds intervals created during the last 2 W bases.  Intervals are handed to
 * the sink in the order sdust.c:94-108 gives them up, with the step that did it (host-checked). */
template <class Sink>
LQ_HD void lq_sdust_scan(const uint8_t *seq, int l_seq, int T, int W, int from, int to, int *pbuf /* 4*capP ints */, int capP, Sink *sink, int *overflow)
{
    lq_sd_state<Sink> s;
    s.Ps = pbuf; s.Pf = pbuf + capP; s.Pr = pbuf + 2 * capP; s.Pl = pbuf + 3 * capP; s.capP = capP; s.sink = sink;
    int i, l = 0, start;
    unsigned t = 0;
    for (i = 0; i < 64; ++i) { s.cw[i] = 0; s.cv[i] = 0; }
    s.w_front = s.w_n = 0; s.rw = s.rv = s.L = 0; s.nP = 0; s.overflow = 0;
    for (i = from; i <= l_seq && i < to; ++i) {
        const int b = i < l_seq ? (int)lq_nt4(seq[i], 1) : 4;
        s.now = i;
        if (b < 4) {
            ++l; t = (t << 2 | (unsigned)b) & 63u;
            if (l >= 3) {
                int x;
                start = (l - W > 0 ? l - W : 0) + (i + 1 - l);
                lq_sd_save(&s, start);
                /* shift_window (sdust.c:72-92) */
                if (s.w_n >= W - 2) {
                    x = s.win[s.w_front]; s.w_front = (s.w_front + 1) & 63; --s.w_n;
                    s.rw -= --s.cw[x];
                    if (s.L > s.w_n) { --s.L; s.rv -= --s.cv[x]; }
                }
                s.win[(s.w_front + s.w_n++) & 63] = (uint8_t)t;
                ++s.L;
                s.rw += s.cw[t]++; s.rv += s.cv[t]++;
                if (s.cv[t] * 10 > T << 1) {
                    do { x = s.win[(s.w_front + s.w_n - s.L) & 63]; s.rv -= --s.cv[x]; --s.L; } while (x != (int)t);
                }
                if (s.rw * 10 > s.L * T) {
                    /* find_perfect (sdust.c:110-134), with its two inner loops made linear.  The reference, for every suffix ii of the
                     * window (longest last), rescans the list from the top down to the first entry starting before ii + start, folding
                     * the best score ratio seen, and memmoves the tail to insert.  The bound ii + start only falls and every entry
                     * inserted on the way starts at or above the next bound, so (1) the scan can go on where the previous suffix's
                     * stopped -- folding an entry twice changes nothing, and the inserted entries carry the running best itself --
                     * and (2) the insertion points are non-decreasing in the ORIGINAL list: they are collected (at most W-2 per call)
                     * and merged in with one backward pass.  Same list, O(|P| + W) instead of O(W |P|) steps per base. */
                    uint8_t c[64]; int r = s.rv, ii, max_r = 0, max_l = 0, q, j = 0, nnew = 0;
                    uint16_t nat[64]; uint32_t nrl[64];
                    for (q = 0; q < 64; ++q) c[q] = s.cv[q];
                    for (ii = s.w_n - s.L - 1; ii >= 0; --ii) {
                        const int tt = s.win[(s.w_front + ii) & 63];
                        int new_r, new_l;
                        r += c[tt]++;
                        new_r = r; new_l = s.w_n - ii - 1;
                        if (new_r * 10 > T * new_l) {
                            for (; j < s.nP && s.Ps[j] >= ii + start; ++j)
                                if (max_r == 0 || s.Pr[j] * max_l > max_r * s.Pl[j]) { max_r = s.Pr[j]; max_l = s.Pl[j]; }
                            if (max_r == 0 || new_r * max_l >= max_r * new_l) {
                                max_r = new_r; max_l = new_l;
                                nat[nnew] = (uint16_t)j; nrl[nnew] = (uint32_t)new_r << 8 | (uint32_t)new_l; ++nnew;
                            }
                        }
                    }
                    if (s.nP + nnew > s.capP) s.overflow = 1;
                    else if (nnew) {
                        int dst = s.nP + nnew - 1, src = s.nP - 1, m;
                        const int fin = s.w_n + 2 + start;
                        for (m = nnew - 1; m >= 0; --m) {
                            const int at = (int)nat[m], nl = (int)(nrl[m] & 255u);
                            for (; src >= at; --src, --dst) { s.Ps[dst] = s.Ps[src]; s.Pf[dst] = s.Pf[src]; s.Pr[dst] = s.Pr[src]; s.Pl[dst] = s.Pl[src]; }
                            s.Ps[dst] = s.w_n - 1 - nl + start; s.Pf[dst] = fin; s.Pr[dst] = (int)(nrl[m] >> 8); s.Pl[dst] = nl;
                            --dst;
                        }
                        s.nP += nnew;
                    }
                }
            }
        } else {
            start = (l - W + 1 > 0 ? l - W + 1 : 0) + (i + 1 - l);
            while (s.nP) lq_sd_save(&s, start++);
            l = 0; t = 0;
        }
    }
    if (overflow) *overflow = s.overflow;
}

/* what happens to an interval the scan gives up (sdust.c:94-108 appends / merges it into the result vector) */
struct lq_sd_merge_sink {   /* merged length, as sdust.c:205-209 sums the merged vector */
    int have_last, last_s, last_f; int64_t masked;
    LQ_HD void init() { have_last = 0; last_s = last_f = 0; masked = 0; }
    LQ_HD void operator()(int ps, int pf, int) { add(ps, pf); }
    LQ_HD void add(int ps, int pf)
    {
        if (have_last && ps <= last_f) { if (pf > last_f) last_f = pf; }
        else { if (have_last) masked += (int64_t)(last_f - last_s); have_last = 1; last_s = ps; last_f = pf; }
    }
    LQ_HD int64_t total() { if (have_last) { masked += (int64_t)(last_f - last_s); have_last = 0; } return masked; }
};

/* masked length of a whole read */
LQ_HD int64_t lq_sdust_masked(const uint8_t *seq, int l_seq, int T, int W, int *pbuf /* 4*capP ints */, int capP, int *overflow)
{
    lq_sd_merge_sink sink; sink.init();
    lq_sdust_scan(seq, l_seq, T, W, 0, 0x7fffffff, pbuf, capP, &sink, overflow);
    return sink.total();
}

/* ---- segment form: the steps [lo, hi) of a read (hi > l_seq: with the end-of-read flush) on their own ----
 * cold start: far enough before lo that (1) the deque holds only real triplets again when (2) the oldest interval that can still be
 * alive at lo was created: 2 W bases (an interval lives while `start` has not passed it: W bases of frozen start after an N, then
 * W-2 more) + W-2 triplets, counted as triplets because an ambiguous base costs three bases before the next one */
LQ_HD int lq_sd_warm_from(const uint8_t *seq, int lo, int W)
{
    int p = lo - 2 * W - 2, words = 0;
    if (p <= 0) return 0;
    while (p > 2 && words < W) {
        --p;
        if (lq_nt4(seq[p], 1) < 4 && lq_nt4(seq[p - 1], 1) < 4 && lq_nt4(seq[p - 2], 1) < 4) ++words;
    }
    return p > 2 ? p - 2 : 0;
}
/* the intervals given up during steps [lo, hi), folded as sdust.c:94-108 folds them: runs[] receives the merged runs in order (at most
 * cap; *n_runs counts all of them).  Folding the runs of all segments of a read one after the other with the same rule gives the read's
 * merged vector: inside a run every "overlaps the one before" test holds whatever came before the segment (a larger running end only
 * makes it hold more), and a run's first interval is tested against the true running end when the segments are joined. */
struct lq_sd_run { int s, f; };
struct lq_sd_seg_sink {
    int lo, have, cur_s, cur_f, n, cap; lq_sd_run *runs;
    LQ_HD void init(int lo_, lq_sd_run *r, int cap_) { lo = lo_; have = 0; cur_s = cur_f = 0; n = 0; cap = cap_; runs = r; }
    LQ_HD void flush() { if (have) { if (n < cap) { runs[n].s = cur_s; runs[n].f = cur_f; } ++n; have = 0; } }
    LQ_HD void operator()(int ps, int pf, int now)
    {
        if (now < lo) return;                   /* given up during the warm-up: another segment's */
        if (have && ps <= cur_f) { if (pf > cur_f) cur_f = pf; }
        else { flush(); have = 1; cur_s = ps; cur_f = pf; }
    }
};
LQ_HD int lq_sdust_segment(const uint8_t *seq, int l_seq, int T, int W, int lo, int hi, int *pbuf, int capP, lq_sd_run *runs, int cap_runs, int *overflow)
{
    lq_sd_seg_sink sink; sink.init(lo, runs, cap_runs);
    lq_sdust_scan(seq, l_seq, T, W, lq_sd_warm_from(seq, lo, W), hi, pbuf, capP, &sink, overflow);
    sink.flush();
    return sink.n;
}
#endif
