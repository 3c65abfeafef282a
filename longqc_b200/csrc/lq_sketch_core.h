/* lq_sketch_core.h -- (w,k)-minimizer sketch of one read, position-parallel and exact.
 *
 * Replaces mm_sketch() (reference minimap2-coverage/sketch.c:76-142).  The reference is a
 * sequential scan with a ring of the last w candidates and a running minimum.  Two facts make
 * an exact position-parallel form possible (checked against the compiled reference on
 * adversarial inputs by tests/test_sketch_core.py):
 *
 *   (1) after every step the running minimum equals the RIGHTMOST minimum of the ring
 *       (`<=` at sketch.c:122, `>=` at :128-129), so the scan state is a pure function of the
 *       last w ring pushes, the k-mer registers and the run counter l (saturating at w+k);
 *   (2) a base whose k-mer equals its reverse complement pushes nothing (`continue`, :107).
 *
 * So for a position i whose last w+k accepted bases are contiguous (no ambiguous base, no
 * palindromic k-mer, not the read start) the records emitted *while processing i* are a
 * closed-form function of the candidates at i-w..i ("fast path").  Every other position
 * ("slow path", <1% of real reads) replays the reference state machine from a restart point
 * at which the state is provably identical (see sk_replay_from()).
 *
 * All functions are host+device so the same code is unit-tested on the CPU.
 */
#ifndef LQ_SKETCH_CORE_H
#define LQ_SKETCH_CORE_H

#include "lq_common.h"

/* sketch.c:27-37, 64-bit form (any k<=28) */
LQ_HD uint64_t lq_hash64(uint64_t key, uint64_t mask)
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}
/* same function when 2k <= 32: every step is taken modulo 2^(2k), so 32-bit lanes suffice */
LQ_HD uint32_t lq_hash32(uint32_t key, uint32_t mask)
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = (key * 265u) & mask;
    key = key ^ key >> 14;
    key = (key * 21u) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

LQ_HD uint32_t lq_base_at(const uint32_t *b2, uint64_t g) { return (b2[g >> 4] >> ((uint32_t)(g & 15) * 2)) & 3u; }
LQ_HD uint32_t lq_amb_at(const uint32_t *nm, uint64_t g) { return (nm[g >> 5] >> (uint32_t)(g & 31)) & 1u; }

/* reverse the order of the 2-bit groups of a 64-bit word */
LQ_HD uint64_t lq_rev2_64(uint64_t v)
{
    v = (v >> 2 & 0x3333333333333333ULL) | (v & 0x3333333333333333ULL) << 2;
    v = (v >> 4 & 0x0F0F0F0F0F0F0F0FULL) | (v & 0x0F0F0F0F0F0F0F0FULL) << 4;
    v = (v >> 8 & 0x00FF00FF00FF00FFULL) | (v & 0x00FF00FF00FF00FFULL) << 8;
    v = (v >> 16 & 0x0000FFFF0000FFFFULL) | (v & 0x0000FFFF0000FFFFULL) << 16;
    return v >> 32 | v << 32;
}

/* One scan candidate: x = hash<<8|span as in the reference; strand in bit 0 of `z`. */
typedef struct { uint64_t x; uint32_t pos; uint32_t z; } lq_cand;

/* k-mer registers of the reference scan after consuming base i of a read (sketch.c:105-106):
 * fw holds the last k unambiguous bases (newest in the low bits), rv their complement in
 * reverse; ambiguous bases leave both untouched, and before k bases have been seen the
 * missing ones contribute zero bits to BOTH registers (initial value {0,0}).
 * Generic (slow) form: walk back over at most k unambiguous bases. */
LQ_HD void lq_regs_at(const uint32_t *b2, const uint32_t *nm, uint64_t g0, int i, int k, uint64_t *fw, uint64_t *rv)
{
    uint64_t f = 0, r = 0;
    int got = 0, j;
    for (j = i; j >= 0 && got < k; --j) {
        if (lq_amb_at(nm, g0 + (uint64_t)j)) continue;
        {
            uint64_t c = lq_base_at(b2, g0 + (uint64_t)j);
            f |= c << (2 * got);
            r |= (3ULL ^ c) << (2 * (k - 1 - got));
            ++got;
        }
    }
    *fw = f; *rv = r;
}

/* The reference state machine, restartable.  Runs steps [from, to] (inclusive, clipped to the
 * read) and hands every record pushed while processing a step s with emit_lo <= s <= emit_hi to
 * `sink(x, y)`; the end-of-read push (sketch.c:140-141) counts as step `len`.
 *   fresh != 0 : true start-of-read state (from must be 0)
 *   fresh == 0 : restart: registers are rebuilt exactly from the bases before `from`, the ring is
 *                all-empty and the run counter starts saturated.  The caller must make sure this is
 *                equivalent (sk_replay_from()); `*accepted` returns the number of ring pushes made
 *                by steps < emit_lo so that it can check.
 * Handles HPC (sketch.c:93-104) only with fresh != 0. */
template <class Sink>
LQ_HD void lq_sketch_replay(const uint32_t *b2, const uint32_t *nm, uint64_t g0, int len, int w, int k, uint32_t rid, int is_hpc,
                            int from, int fresh, int to, int emit_lo, int emit_hi, int *accepted, Sink &sink)
{
    const uint64_t mask = (1ULL << 2 * k) - 1;
    const int top = 2 * (k - 1);
    uint64_t fw = 0, rv = 0, ring_x[LQ_MAX_W], ring_y[LQ_MAX_W], best_x = LQ_U64MAX, best_y = LQ_U64MAX;
    int slot = 0, best_slot = 0, run = 0, span = 0, i, j, npush = 0;
    int hq[32], hq_front = 0, hq_n = 0;
    for (j = 0; j < w; ++j) ring_x[j] = ring_y[j] = LQ_U64MAX;
    if (!fresh) {
        if (from > 0) lq_regs_at(b2, nm, g0, from - 1, k, &fw, &rv);
        run = 1 << 28; /* every gate of sketch.c:116-137 is open */
    }
    if (to >= len) to = len - 1;
    for (i = from; i <= to; ++i) {
        const int step = i; /* HPC moves i to the end of the run; records belong to the step that started it */
        int c = lq_amb_at(nm, g0 + (uint64_t)i) ? 4 : (int)lq_base_at(b2, g0 + (uint64_t)i);
        uint64_t cx = LQ_U64MAX, cy = LQ_U64MAX;
        const bool out = step >= emit_lo && step <= emit_hi;
        if (c < 4) {
            int strand;
            if (is_hpc) {
                int rl = 1;
                while (i + rl < len && !lq_amb_at(nm, g0 + (uint64_t)(i + rl)) && (int)lq_base_at(b2, g0 + (uint64_t)(i + rl)) == c) ++rl;
                i += rl - 1;
                hq[(hq_front + hq_n++) & 31] = rl;
                span += rl;
                if (hq_n > k) { span -= hq[hq_front]; hq_front = (hq_front + 1) & 31; --hq_n; }
            } else span = run + 1 < k ? run + 1 : k;
            fw = (fw << 2 | (uint64_t)c) & mask;
            rv = rv >> 2 | (uint64_t)(3 ^ c) << top;
            if (fw == rv) continue;
            strand = fw < rv ? 0 : 1;
            ++run;
            if (run >= k && span < 256) {
                cx = lq_hash64(strand ? rv : fw, mask) << 8 | (uint64_t)span;
                cy = (uint64_t)rid << 32 | (uint64_t)((uint32_t)i << 1) | (uint64_t)strand;
            }
        } else { run = 0; hq_n = hq_front = 0; span = 0; }
        if (step < emit_lo) ++npush;
        ring_x[slot] = cx; ring_y[slot] = cy;
        if (run == w + k - 1 && best_x != LQ_U64MAX) {
            for (j = slot + 1; j < w; ++j) if (ring_x[j] == best_x && ring_y[j] != best_y && out) sink(ring_x[j], ring_y[j]);
            for (j = 0; j < slot; ++j)     if (ring_x[j] == best_x && ring_y[j] != best_y && out) sink(ring_x[j], ring_y[j]);
        }
        if (cx <= best_x) {
            if (run >= w + k && best_x != LQ_U64MAX && out) sink(best_x, best_y);
            best_x = cx; best_y = cy; best_slot = slot;
        } else if (slot == best_slot) {
            if (run >= w + k - 1 && best_x != LQ_U64MAX && out) sink(best_x, best_y);
            best_x = LQ_U64MAX;
            for (j = slot + 1; j < w; ++j) if (best_x >= ring_x[j]) { best_x = ring_x[j]; best_y = ring_y[j]; best_slot = j; }
            for (j = 0; j <= slot; ++j)    if (best_x >= ring_x[j]) { best_x = ring_x[j]; best_y = ring_y[j]; best_slot = j; }
            if (run >= w + k - 1 && best_x != LQ_U64MAX) {
                for (j = slot + 1; j < w; ++j) if (ring_x[j] == best_x && ring_y[j] != best_y && out) sink(ring_x[j], ring_y[j]);
                for (j = 0; j <= slot; ++j)    if (ring_x[j] == best_x && ring_y[j] != best_y && out) sink(ring_x[j], ring_y[j]);
            }
        }
        if (++slot == w) slot = 0;
    }
    if (to == len - 1 && len >= emit_lo && len <= emit_hi && best_x != LQ_U64MAX) sink(best_x, best_y);
    if (accepted) *accepted = npush;
}

/* Records the reference emits while processing step i (and the end-of-read record when
 * i == len-1), obtained by replaying from the nearest provably-equivalent restart point:
 *   p0 = i - LB with LB = 2(w+k)+16.  The restart is exact when [p0-k, i] holds no ambiguous base
 *   and at least w+k ring pushes happen in [p0, i): then the true run counter at i is >= w+k, the
 *   ring holds only candidates pushed after p0, all of them valid, and the registers are exact.
 *   Otherwise replay from the start of the read (always exact). */
struct lq_sk_buf {
    uint64_t x[2 * LQ_MAX_W + 2], y[2 * LQ_MAX_W + 2];
    int n;
    LQ_HD void operator()(uint64_t x_, uint64_t y_) { if (n < 2 * LQ_MAX_W + 2) { x[n] = x_; y[n] = y_; } ++n; }
};

template <class Sink>
LQ_HD void lq_sketch_slow_at(const uint32_t *b2, const uint32_t *nm, uint64_t g0, int len, int w, int k, uint32_t rid, int i, Sink &sink)
{
    const int LB = 2 * (w + k) + 16;
    const int hi = (i == len - 1) ? len : i;
    const int p0 = i - LB;
    if (p0 - k > 0) {
        int j, clean = 1;
        for (j = p0 - k; j <= i; ++j) if (lq_amb_at(nm, g0 + (uint64_t)j)) { clean = 0; break; }
        if (clean) {
            lq_sk_buf buf; int pushes = 0;
            buf.n = 0;
            lq_sketch_replay(b2, nm, g0, len, w, k, rid, 0, p0, 0, i, i, hi, &pushes, buf);
            if (pushes >= w + k) { /* ring pushes made by steps p0..i-1 */
                for (j = 0; j < buf.n; ++j) sink(buf.x[j], buf.y[j]);
                return;
            }
        }
    }
    lq_sketch_replay(b2, nm, g0, len, w, k, rid, 0, 0, 1, i, i, hi, (int*)0, sink);
}

/* Fast path.  cand[0..w] = candidates at positions i-w..i (all valid, contiguous, gates open).
 * Emits what sketch.c:122-137 pushes for step i, plus the end-of-read record if `last`. */
template <class Sink>
LQ_HD void lq_sketch_fast_at(const uint64_t *cx /* w+1 values hash<<8|span, oldest first */, int w, uint32_t rid, int i,
                             const uint32_t *cz /* strand per candidate */, int last, Sink &sink)
{
    /* rightmost minimum of the old window cx[0..w-1] */
    int m = 0, j;
    for (j = 1; j < w; ++j) if (cx[j] <= cx[m]) m = j;
    #define LQ_Y(j_) ((uint64_t)rid << 32 | (uint64_t)((uint32_t)(i - w + (j_)) << 1) | (uint64_t)cz[j_])
    if (cx[w] <= cx[m]) {
        sink(cx[m], LQ_Y(m));
        if (last) sink(cx[w], LQ_Y(w));
    } else if (m == 0) { /* the minimum was the oldest candidate: it leaves the window now */
        int m2 = 1;
        sink(cx[0], LQ_Y(0));
        for (j = 2; j <= w; ++j) if (cx[j] <= cx[m2]) m2 = j;
        for (j = 1; j <= w; ++j) if (j != m2 && cx[j] == cx[m2]) sink(cx[j], LQ_Y(j));
        if (last) sink(cx[m2], LQ_Y(m2));
    } else if (last) sink(cx[m], LQ_Y(m));
    #undef LQ_Y
}


/* ---- clean-window candidate: bases [g-k+1, g] are all unambiguous and inside the read ---- */

/* the k bases starting at base index g_first, oldest in the low bits (needs k <= 28: 3 words) */
LQ_HD uint64_t lq_kmer_le(const uint32_t *b2, uint64_t g_first, int k)
{
    const uint64_t wi = g_first >> 4;
    const uint32_t sh = (uint32_t)(g_first & 15) * 2;
    uint64_t lo = (uint64_t)b2[wi] | (uint64_t)b2[wi + 1] << 32;
    uint64_t v = lo >> sh;
    if (sh && 2 * k + (int)sh > 64) v |= (uint64_t)b2[wi + 2] << (64 - sh);
    return v & ((1ULL << 2 * k) - 1);
}

/* candidate of the clean window ending at base index g: returns 0 if the k-mer is its own
 * reverse complement (no push), else 1 with *hash = hash64(min(fw,rv)) and *strand */
LQ_HD int lq_cand_clean(const uint32_t *b2, uint64_t g, int k, uint64_t *hash, uint32_t *strand)
{
    const uint64_t mask = (1ULL << 2 * k) - 1;
    const uint64_t le = lq_kmer_le(b2, g - (uint64_t)(k - 1), k);
    const uint64_t rv = ~le & mask;                    /* complement, oldest base in the low bits == sketch.c:106 */
    const uint64_t fw = lq_rev2_64(le) >> (64 - 2 * k);  /* newest base in the low bits == sketch.c:105 */
    if (fw == rv) return 0;
    *strand = fw < rv ? 0u : 1u;
    *hash = lq_hash64(fw < rv ? fw : rv, mask);
    return 1;
}

/* any ambiguous base in [g_lo, g_hi] (inclusive) */
LQ_HD int lq_amb_any(const uint32_t *nm, uint64_t g_lo, uint64_t g_hi)
{
    uint64_t wlo = g_lo >> 5, whi = g_hi >> 5, wj;
    for (wj = wlo; wj <= whi; ++wj) {
        uint32_t m = nm[wj];
        if (wj == wlo) m &= 0xffffffffu << (uint32_t)(g_lo & 31);
        if (wj == whi) m &= 0xffffffffu >> (31 - (uint32_t)(g_hi & 31));
        if (m) return 1;
    }
    return 0;
}

/* does base i of the read push a candidate into the ring (unambiguous and not palindromic)? exact for every i */
LQ_HD int lq_pos_ok(const uint32_t *b2, const uint32_t *nm, uint64_t g0, int i, int k)
{
    if (lq_amb_at(nm, g0 + (uint64_t)i)) return 0;
    if (i >= k - 1 && !lq_amb_any(nm, g0 + (uint64_t)(i - k + 1), g0 + (uint64_t)i)) {
        uint64_t h; uint32_t z;
        return lq_cand_clean(b2, g0 + (uint64_t)i, k, &h, &z);
    } else {
        uint64_t fw, rv;
        lq_regs_at(b2, nm, g0, i, k, &fw, &rv);
        return fw != rv;
    }
}

/* Everything the reference pushes while processing base i of a read (non-HPC), position-parallel. */
template <class Sink>
LQ_HD void lq_sketch_at(const uint32_t *b2, const uint32_t *nm, uint64_t g0, int len, int w, int k, uint32_t rid, int i, Sink &sink)
{
    const int need = w + k - 1;
    int j, fast = i >= need;
    if (fast) for (j = i - need; j <= i; ++j) if (!lq_pos_ok(b2, nm, g0, j, k)) { fast = 0; break; }
    if (fast) {
        uint64_t cx[LQ_MAX_W + 1]; uint32_t cz[LQ_MAX_W + 1];
        for (j = 0; j <= w; ++j) {
            uint64_t h = 0; uint32_t z = 0;
            lq_cand_clean(b2, g0 + (uint64_t)(i - w + j), k, &h, &z);
            cx[j] = h << 8 | (uint64_t)k; cz[j] = z;
        }
        lq_sketch_fast_at(cx, w, rid, i, cz, i == len - 1, sink);
    } else lq_sketch_slow_at(b2, nm, g0, len, w, k, rid, i, sink);
}

/* ---- windowed fast path: palindromic k-mers and ambiguous bases inside the look-back are handled in closed form ----
 *
 * okw / ambw describe the 64 bases ending at base i of the read: bit 63-d is set iff base i-d pushes a candidate
 * (unambiguous and not palindromic) / is ambiguous; bits of bases before the read start are 0.
 * Let l' = number of pushes after the last ambiguous base of the window (the read start counts as one).  The true run
 * counter at i is >= l'.  If base i pushes and l' >= w+k then every gate of sketch.c:116-137 is open, the last w+1 ring
 * pushes are exactly the last w+1 set bits of okw (all made with run >= k, i.e. valid, and none of them a MAX push of an
 * ambiguous base), and the running minimum is the rightmost minimum of the w older ones: the closed form applies with
 * NON-contiguous candidates.  Returns 0 when the bounded replay is needed instead (first w+k pushes of a run; the
 * end-of-read record of a read whose last base does not push). */
#ifdef __CUDA_ARCH__
#define LQ_CLZ64(x) __clzll((long long)(x))
#define LQ_POPC64(x) __popcll(x)
#else
#define LQ_CLZ64(x) __builtin_clzll(x)
#define LQ_POPC64(x) __builtin_popcountll(x)
#endif

/* WC > 0: w is the compile-time constant WC (== WT), so the candidate arrays stay in registers */
template <int WT, int WC, class Fetch, class Sink>
LQ_HD int lq_sketch_fast_win(uint64_t okw, uint64_t ambw, int w_rt, int k, uint32_t rid, int i, int last, Fetch &fetch, Sink &sink)
{
    const int w = WC ? WC : w_rt;
    if (!(okw >> 63)) return last ? 0 : 1;           /* base i pushes nothing valid and emits nothing (sketch.c:107 / :114 with l = 0) */
    uint64_t m = okw;
    if (ambw) { const int hb = 63 - LQ_CLZ64(ambw); m &= ~((2ULL << hb) - 1ULL); }  /* pushes after the last ambiguous base */
    if (LQ_POPC64(m) < w + k) return 0;
    /* the w+1 most recent pushes, newest first while scanning, stored oldest first */
    uint32_t cx[WT + 1], cz[WT + 1]; int cd[WT + 1];
    #pragma unroll
    for (int j = 0; j <= WT; ++j) {
        if (j <= w) {
            const int b = 63 - LQ_CLZ64(m);          /* highest set bit */
            m &= ~(1ULL << b);
            const int d = 63 - b;                    /* distance back from i */
            uint32_t h, z; fetch(d, &h, &z);
            cx[w - j] = h; cz[w - j] = z; cd[w - j] = d;
        }
    }
    #define LQ_EMIT(j_) sink((uint64_t)cx[j_] << 8 | (uint64_t)k, (uint64_t)rid << 32 | (uint64_t)((uint32_t)(i - cd[j_]) << 1) | (uint64_t)cz[j_])
    /* rightmost minimum of the old window cx[0..w-1] and of the new one cx[1..w] */
    int mi = 0, m2 = 1; uint32_t v1 = cx[0], v2 = cx[1];   /* values carried along: no dynamically indexed register array */
    #pragma unroll
    for (int j = 1; j < WT; ++j) if (j < w && cx[j] <= v1) { mi = j; v1 = cx[j]; }
    #pragma unroll
    for (int j = 2; j <= WT; ++j) if (j <= w && cx[j] <= v2) { m2 = j; v2 = cx[j]; }
    const bool c1 = cx[w] <= v1;            /* sketch.c:122: the newcomer takes over, the old minimum is written */
    const bool c2 = !c1 && mi == 0;             /* sketch.c:125: the minimum was the oldest candidate and leaves the window */
    if (c1 | c2) {                              /* one record, selected without a divergent branch per case */
        const int sel = c1 ? mi : 0;
        uint32_t ex = cx[0], ez = cz[0]; int ed = cd[0];
        #pragma unroll
        for (int j = 1; j < WT; ++j) if (j < w && j == sel) { ex = cx[j]; ez = cz[j]; ed = cd[j]; }
        sink((uint64_t)ex << 8 | (uint64_t)k, (uint64_t)rid << 32 | (uint64_t)((uint32_t)(i - ed) << 1) | (uint64_t)ez);
    }
    if (c2) {                                   /* twins of the new minimum (sketch.c:131-136): equal hashes inside one window, rare */
        bool any = false;
        #pragma unroll
        for (int j = 1; j <= WT; ++j) if (j <= w && j != m2 && cx[j] == v2) any = true;
        if (any) {
            #pragma unroll
            for (int j = 1; j <= WT; ++j) if (j <= w && j != m2 && cx[j] == v2) LQ_EMIT(j);
        }
    }
    if (last) {                                 /* sketch.c:140-141: the minimum of the final window */
        const int fin = c1 ? w : (c2 ? m2 : mi);
        #pragma unroll
        for (int j = 0; j <= WT; ++j) if (j <= w && j == fin) LQ_EMIT(j);
    }
    #undef LQ_EMIT
    return 1;
}

/* host/test form of the same rule: windows and candidates recomputed from the packed read */
struct lq_fetch_packed {
    const uint32_t *b2; uint64_t g; int k;
    LQ_HD void operator()(int d, uint32_t *h, uint32_t *z) const { uint64_t hh = 0; lq_cand_clean(b2, g - (uint64_t)d, k, &hh, z); *h = (uint32_t)hh; }
};

template <class Sink>
LQ_HD void lq_sketch_at_win(const uint32_t *b2, const uint32_t *nm, uint64_t g0, int len, int w, int k, uint32_t rid, int i, Sink &sink)
{
    uint64_t okw = 0, ambw = 0;
    for (int d = 0; d < 64 && d <= i; ++d) {
        if (lq_amb_at(nm, g0 + (uint64_t)(i - d))) ambw |= 1ULL << (63 - d);
        else if (lq_pos_ok(b2, nm, g0, i - d, k)) okw |= 1ULL << (63 - d);
    }
    lq_fetch_packed f; f.b2 = b2; f.g = g0 + (uint64_t)i; f.k = k;
    if (k <= 16 && w <= LQ_MAX_W && lq_sketch_fast_win<LQ_MAX_W, 0>(okw, ambw, w, k, rid, i, i == len - 1, f, sink)) return;
    lq_sketch_slow_at(b2, nm, g0, len, w, k, rid, i, sink);
}

/* ASCII -> reference base code (sketch.c:8-25; sdust.c:26-43 when sdust_tbl: U is not a base there) */
LQ_HD uint32_t lq_nt4(uint32_t c, int sdust_tbl)
{
    const uint32_t u = c & 0xDFu;
    if (c < 4) return c;
    if ((c | 0x20u) < 'a' || (c | 0x20u) > 'z') return 4;
    if (u == 'A') return 0;
    if (u == 'C') return 1;
    if (u == 'G') return 2;
    if (u == 'T') return 3;
    if (u == 'U') return sdust_tbl ? 4u : 3u;
    return 4;
}

#endif
