/* lq_sketch_pk_core.h -- (w,k)-minimizers of one 64-base segment of a read, everything in registers.
 *
 * Replaces the inner loop of mm_sketch() (reference minimap2-coverage/sketch.c:76-142) for k <= 15.  Facts used
 * (lq_sketch_core.h has the first two):
 *   (1) the running minimum of the reference scan is the RIGHTMOST minimum of the ring, so the scan state is a function of
 *       the last w ring pushes once every gate of sketch.c:116-137 is open (run counter l >= w+k);
 *   (2) a k-mer equal to its reverse complement pushes nothing (sketch.c:107);
 *   (3) hash64 is a bijection of the 2k-bit k-mer, so "equal hash" means "equal canonical k-mer";
 *   (4) the record written at a step is the minimum that stops being one -- because a newcomer is <= it (sketch.c:122) or
 *       because it leaves the window (:125) -- so a step writes a record exactly when the rightmost minimum of the new window
 *       is another ring entry than that of the old one, and what it writes is the old one.  Only the second case can write
 *       more (:131-136: the entries of the new window that share the new minimum's hash, oldest first).
 * A candidate is ONE integer: hash in the high bits, 255 - (2*p + strand) in the low byte, p = the base's position in the
 * thread's 128-base frame (64 bases of look-back, 64 of its own).  The plain unsigned minimum of such keys IS the rightmost
 * minimum, and the position of the minimum comes back out of the key: a step is two 3-input minima and one comparison.
 * k-mers come out of the packed words by funnel shifts with compile-time amounts (a segment starts on a word boundary):
 * the forward k-mer from the 2-bit-group-reversed words, the complement one from the words as they are.
 * The segment is scanned in four blocks of 16 bases by ONE unrolled copy of the step (lq_pk_fast): the frame of the words moves
 * with the block, the codes stay positions in the segment's frame.  The unrolled step is straight-line and knows neither twins
 * (fact 4, second case) nor gates; it only records the smallest XOR between a candidate and the entries of its window, and a block
 * that saw one below 256 (two equal k-mers in one window, fact 3) is undone and scanned again by the general step (lq_pk_slow,
 * run-time position, writes the twins), as is the block after it.  A read's first and last blocks run a gated copy of the unrolled
 * step (lq_pk_fast_g).
 *
 * A thread may start in the middle of a read ("case A") when the look-back is free of ambiguous bases and holds >= w+k pushes
 * (counted: palindromic k-mers do not push), the last w of them inside the part that is hashed -- then l >= w+k, the ring is
 * exactly those w pushes.  At the start of a read ("case B") the reference's gates are functions of the position as long as
 * nothing but plain pushes happens before the first full window.  Everything else -- an ambiguous base in reach, too many
 * palindromes in the look-back, a palindrome or an equal pair inside a read's first window (sketch.c:116-121) -- returns
 * "not here" and the caller runs the general state machine for the segment.  Host + device: tests/test_hostcheck.py runs
 * this file against the oracle segment by segment.
 */
#ifndef LQ_SKETCH_PK_CORE_H
#define LQ_SKETCH_PK_CORE_H

#include "lq_sketch_core.h"

#define LQ_PK_SEG 64       /* bases per thread */
#define LQ_PK_PM 8         /* palindromes tolerated in the counted look-back */

template <bool WIDE> struct lq_pk_tr;
template <> struct lq_pk_tr<false> {          /* 2k <= 24: hash and code share 32 bits */
    typedef uint32_t key;
    static LQ_HD key mk(uint32_t h, uint32_t code) { return (h << 8) + code; }   /* code < 256; a sum folds into the hash's last multiply-add */
    static LQ_HD bool same(key d /* a ^ b */) { return d < 256u; }
    static LQ_HD uint32_t hash(key a) { return a >> 8; }
    static LQ_HD uint32_t code(key a) { return a & 255u; }
    static LQ_HD key none() { return 0xffffffffu; }
    static LQ_HD key fill() { return 0xfffffeffu; }      /* neither an entry nor the twin of one (nor of none()) */
    static LQ_HD key min2(key a, key b) { return a < b ? a : b; }
#ifdef __CUDA_ARCH__
    static LQ_HD key min3(key a, key b, key c) { return __vimin3_u32(a, b, c); }
#else
    static LQ_HD key min3(key a, key b, key c) { return min2(min2(a, b), c); }
#endif
};
template <> struct lq_pk_tr<true> {           /* 2k <= 32: hash in the high word */
    typedef uint64_t key;
    static LQ_HD key mk(uint32_t h, uint32_t code) { return ((uint64_t)h << 32) + code; }
    static LQ_HD bool same(key d) { return (d >> 32) == 0; }
    static LQ_HD uint32_t hash(key a) { return (uint32_t)(a >> 32); }
    static LQ_HD uint32_t code(key a) { return (uint32_t)a & 255u; }
    static LQ_HD key none() { return LQ_U64MAX; }
    static LQ_HD key fill() { return 0xfffffffeffffffffULL; }
    static LQ_HD key min2(key a, key b) { return a < b ? a : b; }
    static LQ_HD key min3(key a, key b, key c) { return min2(min2(a, b), c); }
};

LQ_HD uint32_t lq_pk_rev2(uint32_t x)         /* the sixteen 2-bit groups of a word in reverse order */
{
#ifdef __CUDA_ARCH__
    const uint32_t t = __brev(x);
#else
    uint32_t t = x;
    t = (t >> 16) | (t << 16);
    t = (t >> 8 & 0x00ff00ffu) | (t & 0x00ff00ffu) << 8;
    t = (t >> 4 & 0x0f0f0f0fu) | (t & 0x0f0f0f0fu) << 4;
    t = (t >> 2 & 0x33333333u) | (t & 0x33333333u) << 2;
    t = (t >> 1 & 0x55555555u) | (t & 0x55555555u) << 1;
#endif
    return (t >> 1 & 0x55555555u) | (t & 0x55555555u) << 1;
}
LQ_HD uint32_t lq_fsr(uint32_t lo, uint32_t hi, uint32_t sh)   /* bits sh..sh+31 of hi:lo, sh < 32 */
{
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? lo >> sh | hi << (32 - sh) : lo;
#endif
}

/* the thread's state: ring of the last W pushes (r[0] oldest), the current minimum, and the smallest difference seen between a
 * candidate and an entry of its window (below 256 / 2^32: two equal canonical k-mers have shared a window) */
template <int W, bool WIDE> struct lq_pk_state {
    typename lq_pk_tr<WIDE>::key r[W], om, d;
    uint32_t bad;    /* not a segment for this form */
};

/* forward and complement k-mer registers after base P of the frame (sketch.c:105-106), P >= K-1 */
template <int K>
LQ_HD void lq_pk_regs(const uint32_t *Lw /* [9] */, const uint32_t *Rw /* [9] */, const int P, uint32_t *fw, uint32_t *rv)
{
    const uint32_t mask = (1u << 2 * K) - 1;
    const int a = P - K + 1, b = 127 - P;
    *rv = ~lq_fsr(Lw[a >> 4], Lw[(a >> 4) + 1], 2 * (a & 15)) & mask;
    *fw = lq_fsr(Rw[b >> 4], Rw[(b >> 4) + 1], 2 * (b & 15)) & mask;
}

template <int W, bool WIDE>
LQ_HD void lq_pk_window(const typename lq_pk_tr<WIDE>::key kk, lq_pk_state<W, WIDE> &s, typename lq_pk_tr<WIDE>::key *nm_out)
{
    typedef lq_pk_tr<WIDE> T;
    typename T::key d = s.d, nm = kk;
    int j = 1;
    for (; j + 1 < W; j += 2) { nm = T::min3(nm, s.r[j], s.r[j + 1]); d = T::min3(d, kk ^ s.r[j], kk ^ s.r[j + 1]); }
    if (j < W) { nm = T::min2(nm, s.r[j]); d = T::min2(d, kk ^ s.r[j]); }
    s.d = d; *nm_out = nm;
}

/* one base, the common case, straight-line: every gate of sketch.c:116-137 is open, the base exists, P is a compile-time
 * constant after unrolling.  sink.put(key, yes) stores unconditionally and keeps the record only if `yes`.  A palindromic
 * k-mer (no push) leaves the state alone; its key still enters s.d, which is harmless: it cannot equal a pushed k-mer. */
template <int W, int K, bool EMIT, class Sink>
LQ_HD void lq_pk_fast(const uint32_t *Lw, const uint32_t *Rw, const int P, lq_pk_state<W, (K > 12)> &s, Sink &sink, const uint32_t cb = 255u /* 255 - 32 * block: the code is made of the position in the SEGMENT's frame */)
{
    typedef lq_pk_tr<(K > 12)> T;
    typedef typename T::key key;
    const uint32_t mask = (1u << 2 * K) - 1;
    uint32_t fw, rv;
    lq_pk_regs<K>(Lw, Rw, P, &fw, &rv);
    const bool push = fw != rv;
    const uint32_t z = fw < rv ? 0u : 1u;
    const key kk = T::mk(lq_hash32(z ? rv : fw, mask), cb - 2u * (uint32_t)P - z);
    key nm;
    lq_pk_window<W, (K > 12)>(kk, s, &nm);
    if (EMIT) sink.put(s.om, push && nm != s.om);
    s.om = push ? nm : s.om;
    for (int j = 0; j + 1 < W; ++j) s.r[j] = push ? s.r[j + 1] : s.r[j];
    s.r[W - 1] = push ? kk : s.r[W - 1];
}

/* the same with the gates of a read's first and last blocks: a base counts from plo on (at a read start: once k bases have
 * been seen) and below phi (the read's end), records leave from pef on (sketch.c:123, :126: l >= w+k at a read start).  A
 * palindromic k-mer before pef makes the segment one for the general state machine (l is no longer the position there). */
template <int W, int K, class Sink>
LQ_HD void lq_pk_fast_g(const uint32_t *Lw, const uint32_t *Rw, const int P, lq_pk_state<W, (K > 12)> &s, Sink &sink, const uint32_t cb,
                        const int plo, const int phi, const int pef)
{
    typedef lq_pk_tr<(K > 12)> T;
    typedef typename T::key key;
    const uint32_t mask = (1u << 2 * K) - 1;
    uint32_t fw, rv;
    lq_pk_regs<K>(Lw, Rw, P, &fw, &rv);
    const bool live = P >= plo && P < phi;
    const bool push = live && fw != rv;
    if (live && fw == rv && P < pef) s.bad = 1;
    const uint32_t z = fw < rv ? 0u : 1u;
    const key kk = live ? T::mk(lq_hash32(z ? rv : fw, mask), cb - 2u * (uint32_t)P - z) : T::fill();   /* padding must not look like an equal pair */
    key nm;
    lq_pk_window<W, (K > 12)>(kk, s, &nm);
    sink.put(s.om, push && nm != s.om && P >= pef);
    s.om = push ? nm : s.om;
    for (int j = 0; j + 1 < W; ++j) s.r[j] = push ? s.r[j + 1] : s.r[j];
    s.r[W - 1] = push ? kk : s.r[W - 1];
}

/* one base, the general case of this form, P a run-time value (the words are indexed: on the device `lw8` is shared memory):
 * honours lo <= P < hi (bases that exist and, at a read start, have seen k bases), writes only from P >= efrom on
 * (sketch.c:123, :126: l >= w+k at a read start), and writes the twins of a minimum that took over by age (sketch.c:131-136). */
template <int W, int K, class Sink>
LQ_HD void lq_pk_slow(const uint32_t *lw8, const int P, lq_pk_state<W, (K > 12)> &s, const int lo, const int hi, const int efrom, Sink &sink)
{
    typedef lq_pk_tr<(K > 12)> T;
    typedef typename T::key key;
    const uint32_t mask = (1u << 2 * K) - 1;
    const int a = P - K + 1;
    if (P < lo || P >= hi) return;
    const uint32_t le = lq_fsr(lw8[a >> 4], (a >> 4) + 1 < 8 ? lw8[(a >> 4) + 1] : 0u, 2 * (a & 15)) & mask;
    const uint32_t rv = ~le & mask, fw = lq_pk_rev2(le) >> (32 - 2 * K);
    if (fw == rv) { if (P < efrom) s.bad = 1; return; }          /* a palindrome before a read's first full window: l is no longer the position */
    const uint32_t z = fw < rv ? 0u : 1u;
    const key kk = T::mk(lq_hash32(z ? rv : fw, mask), (255u - 2u * (uint32_t)P) - z);
    key nm;
    lq_pk_window<W, (K > 12)>(kk, s, &nm);
    if (T::same(s.d) && P < efrom) s.bad = 1;                    /* sketch.c:116-121 is not restated */
    if (nm != s.om) {
        if (P >= efrom) {
            sink.push(s.om);
            if (s.om == s.r[0] && T::hash(kk) > T::hash(s.om)) {
                for (int j = 1; j < W; ++j) if (s.r[j] != nm && s.r[j] != T::none() && T::same(s.r[j] ^ nm)) sink.push(s.r[j]);
            }
        }
        s.om = nm;
    }
    for (int j = 0; j + 1 < W; ++j) s.r[j] = s.r[j + 1];
    s.r[W - 1] = kk;
}

/* lw8: the packed words of the 64 bases before the segment and of the segment; nw4: their ambiguity bits; i0: the segment's
 * first base in its read (a multiple of 64); nseg: its bases (1..64); is_last: the read ends with it.  The sink receives the
 * records in the reference's order as keys (put: at most one per base from an unrolled block, NOT checked -- a sink of fixed
 * size must have 64 entries of slack behind it and is full when room() < 0 at the end; push: checked); lq_pk_p2z() turns a key's code into pos<<1|strand.  Returns 0; 1 = not here; 2 = the sink is full
 * (in both cases whatever the sink holds is to be dropped). */
LQ_HD uint32_t lq_pk_p2z(int i0, uint32_t code) { return (uint32_t)(2 * i0 + 127) - code; }

template <int W, int K, class Sink>
LQ_HD int lq_pk_segment(const uint32_t *lw8, const uint32_t *nw4, const int i0, const int nseg, const bool is_last, Sink &sink)
{
    typedef lq_pk_tr<(K > 12)> T;
    const int LBC = W + K + LQ_PK_PM;                      /* counted look-back: positions 64-LBC .. 63 */
    const int PW = W + 3;                                  /* of which the last PW are hashed and fill the ring */
    static_assert(K >= 2 && K <= 15 && W >= 2 && W <= 16, "lq_pk_segment: (w,k) outside the form");
    static_assert(64 - LBC - (K - 1) >= 0, "lq_pk_segment: look-back does not fit the frame");
    uint32_t Lw[9], Rw[9], nw[4];
#ifdef __CUDA_ARCH__                                       /* the caller's pointers are 16- and 8-byte aligned */
    { const uint4 w0 = *(const uint4*)lw8, w1 = *(const uint4*)(lw8 + 4); const uint2 m0 = *(const uint2*)nw4, m1 = *(const uint2*)(nw4 + 2);
      Lw[0] = w0.x; Lw[1] = w0.y; Lw[2] = w0.z; Lw[3] = w0.w; Lw[4] = w1.x; Lw[5] = w1.y; Lw[6] = w1.z; Lw[7] = w1.w;
      nw[0] = m0.x; nw[1] = m0.y; nw[2] = m1.x; nw[3] = m1.y; }
#else
    for (int j = 0; j < 8; ++j) Lw[j] = lw8[j];
    for (int j = 0; j < 4; ++j) nw[j] = nw4[j];
#endif
    Lw[8] = 0;
    for (int j = 0; j < 8; ++j) Rw[j] = lq_pk_rev2(Lw[7 - j]);
    Rw[8] = 0;
    const uint64_t amb_own = ((uint64_t)nw[2] | (uint64_t)nw[3] << 32) & (nseg >= 64 ? LQ_U64MAX : (1ULL << nseg) - 1);
    if (amb_own) return 1;
    lq_pk_state<W, (K > 12)> s;
    for (int j = 0; j < W; ++j) s.r[j] = T::none();
    s.om = T::none(); s.d = T::none(); s.bad = 0;
    int lo = 0, efrom = 0;
    const int hi = 64 + nseg;
    if (i0 > 0) {                                          /* case A */
        const uint64_t amb_lb = ((uint64_t)nw[0] | (uint64_t)nw[1] << 32) >> (64 - LBC - (K - 1));
        if (amb_lb) return 1;
        int cnt = 0;
#ifdef __CUDA_ARCH__
        #pragma unroll
#endif
        for (int P = 64 - LBC; P < 64 - PW; ++P) { uint32_t fw, rv; lq_pk_regs<K>(Lw, Rw, P, &fw, &rv); cnt += fw != rv; }
        int cw = 0;
#ifdef __CUDA_ARCH__
        #pragma unroll
#endif
        for (int P = 64 - PW; P < 64; ++P) {
            uint32_t fw, rv; lq_pk_regs<K>(Lw, Rw, P, &fw, &rv); cw += fw != rv;
            lq_pk_fast<W, K, false>(Lw, Rw, P, s, sink);
        }
        if (cw < W || cnt + cw < W + K) return 1;
    } else {                                               /* case B: the first K-1 bases push nothing valid, records from l = w+k on */
        lo = 64 + K - 1; efrom = 64 + W + K - 1;
    }
    bool dup = T::same(s.d);                               /* an equal pair in the block before: this block looks for twins */
    {
        /* One copy of the unrolled block: the frame of the WORDS moves with the block (its 16 bases are always bit positions
         * 64..79, the words it needs always Lw[3], Lw[4] and their reversals); the codes stay positions in the segment's frame. */
#ifdef __CUDA_ARCH__
        #pragma unroll 1
#endif
        for (int B = 0; B * 16 < nseg; ++B) {
            if (B) { Lw[3] = Lw[4]; Lw[4] = lw8[4 + B]; Rw[4] = Rw[3]; Rw[3] = lq_pk_rev2(Lw[4]); }
            const uint32_t cb = 255u - 32u * (uint32_t)B;
            const bool first = i0 == 0 && B * 16 < W + K - 1, gated = first || nseg < (B + 1) * 16;
            bool slow = dup;
            s.d = T::none();
            if (!slow) {
                const lq_pk_state<W, (K > 12)> keep = s;
                const typename Sink::mark_t m = sink.mark();
                if (!gated) {
#ifdef __CUDA_ARCH__
                    #pragma unroll
#endif
                    for (int j = 0; j < 16; ++j) lq_pk_fast<W, K, true>(Lw, Rw, 64 + j, s, sink, cb);
                } else {
                    const int plo = lo - 16 * B, phi = hi - 16 * B, pef = efrom - 16 * B;
#ifdef __CUDA_ARCH__
                    #pragma unroll
#endif
                    for (int j = 0; j < 16; ++j) lq_pk_fast_g<W, K>(Lw, Rw, 64 + j, s, sink, cb, plo, phi, pef);
                }
                if (T::same(s.d)) {
                    if (first) return 1;                         /* an equal pair around a read's first window (sketch.c:116-121 is not restated) */
                    s = keep; s.d = T::none(); sink.rewind(m); slow = true;
                }
            }
            if (slow) {
#ifdef __CUDA_ARCH__
                #pragma unroll 1
#endif
                for (int j = 0; j < 16; ++j) lq_pk_slow<W, K>(lw8, 64 + B * 16 + j, s, lo, hi, efrom, sink);
            }
            dup = T::same(s.d);
        }
    }
    if (s.bad) return 1;
    if (is_last && s.om != T::none()) sink.push(s.om);     /* sketch.c:140-141 */
    return sink.room() < 0 ? 2 : 0;
}

#endif
