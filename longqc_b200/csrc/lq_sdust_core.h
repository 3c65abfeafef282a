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

struct lq_sd_state {
    int win[64], w_front, w_n;
    int cw[64], cv[64], rw, rv, L;
    int *Ps, *Pf, *Pr, *Pl, nP, capP;
    int have_last, last_s, last_f;   /* the open (still mergeable) result interval */
    int64_t masked;
    int overflow;
};

LQ_HD void lq_sd_save(lq_sd_state *s, int start) /* sdust.c:94-108 */
{
    if (s->nP == 0 || s->Ps[s->nP - 1] >= start) return;
    const int ps = s->Ps[s->nP - 1], pf = s->Pf[s->nP - 1];
    if (s->have_last && ps <= s->last_f) { if (pf > s->last_f) s->last_f = pf; }
    else {
        if (s->have_last) s->masked += (int64_t)(s->last_f - s->last_s);
        s->have_last = 1; s->last_s = ps; s->last_f = pf;
    }
    int i = s->nP - 1;
    while (i >= 0 && s->Ps[i] < start) --i;
    s->nP = i + 1;
}

/* masked length = sum(finish - start) over the merged intervals */
LQ_HD int64_t lq_sdust_masked(const uint8_t *seq, int l_seq, int T, int W, int *pbuf /* 4*capP ints */, int capP, int *overflow)
{
    lq_sd_state s;
    s.Ps = pbuf; s.Pf = pbuf + capP; s.Pr = pbuf + 2 * capP; s.Pl = pbuf + 3 * capP; s.capP = capP;
    int i, l = 0, start;
    unsigned t = 0;
    for (i = 0; i < 64; ++i) { s.cw[i] = 0; s.cv[i] = 0; }
    s.w_front = s.w_n = 0; s.rw = s.rv = s.L = 0; s.nP = 0; s.have_last = 0; s.last_s = s.last_f = 0; s.masked = 0; s.overflow = 0;
    for (i = 0; i <= l_seq; ++i) {
        const int b = i < l_seq ? (int)lq_nt4(seq[i], 1) : 4;
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
                s.win[(s.w_front + s.w_n++) & 63] = (int)t;
                ++s.L;
                s.rw += s.cw[t]++; s.rv += s.cv[t]++;
                if (s.cv[t] * 10 > T << 1) {
                    do { x = s.win[(s.w_front + s.w_n - s.L) & 63]; s.rv -= --s.cv[x]; --s.L; } while (x != (int)t);
                }
                if (s.rw * 10 > s.L * T) { /* find_perfect (sdust.c:110-134) */
                    int c[64], r = s.rv, ii, max_r = 0, max_l = 0, q;
                    for (q = 0; q < 64; ++q) c[q] = s.cv[q];
                    for (ii = s.w_n - s.L - 1; ii >= 0; --ii) {
                        const int tt = s.win[(s.w_front + ii) & 63];
                        int j, new_r, new_l;
                        r += c[tt]++;
                        new_r = r; new_l = s.w_n - ii - 1;
                        if (new_r * 10 > T * new_l) {
                            for (j = 0; j < s.nP && s.Ps[j] >= ii + start; ++j)
                                if (max_r == 0 || s.Pr[j] * max_l > max_r * s.Pl[j]) { max_r = s.Pr[j]; max_l = s.Pl[j]; }
                            if (max_r == 0 || new_r * max_l >= max_r * new_l) {
                                max_r = new_r; max_l = new_l;
                                if (s.nP >= s.capP) { s.overflow = 1; }
                                else {
                                    for (q = s.nP; q > j; --q) { s.Ps[q] = s.Ps[q-1]; s.Pf[q] = s.Pf[q-1]; s.Pr[q] = s.Pr[q-1]; s.Pl[q] = s.Pl[q-1]; }
                                    ++s.nP;
                                    s.Ps[j] = ii + start; s.Pf[j] = s.w_n + 2 + start; s.Pr[j] = new_r; s.Pl[j] = new_l;
                                }
                            }
                        }
                    }
                }
            }
        } else {
            start = (l - W + 1 > 0 ? l - W + 1 : 0) + (i + 1 - l);
            while (s.nP) lq_sd_save(&s, start++);
            l = 0; t = 0;
        }
    }
    if (s.have_last) s.masked += (int64_t)(s.last_f - s.last_s);
    if (overflow) *overflow = s.overflow;
    return s.masked;
}
#endif
