/* lq_sketch.cu -- K0 (ASCII -> 2-bit + ambiguity planes) and K1 ((w,k)-minimizer sketch).
 *
 * Replaces, on the device, mm_sketch() as called for every target read (reference index.c:291-302)
 * and every query read (minimap2-coverage.c:419, lqmap.c:131).  Four exact forms of K1, by (w, k):
 *   lq_sketch_pk_k    w = 5, k = 12 / 15 (LongQC's overlap runs): 64 bases per thread in registers, candidates as single
 *                     integers, tile words by cp.async.bulk -- lq_sketch_pk_core.h.  The default and the fast one.
 *   lq_sketch_roll_k  w = 5 / 10, other k <= 15: a thread runs the reference scan over 64 bases after a certified warm-up
 *                     (rk_scan, also the general state machine the packed-key kernel falls back on for the segments it declines)
 *   lq_sketch_k       any other (w, k <= 15): position-parallel closed form over a tile of 1024 bases (lq_sketch_core.h):
 *                       stage 0  the tile's 2-bit words (+256-base halo) and ambiguity words go to shared memory
 *                       stage A  every base of tile+128-base halo gets its candidate: 2k-bit hash of min(fw,rv), strand bit,
 *                                and an "ok" bit (pushes into the reference's ring: unambiguous and not palindromic)
 *                       stage B  every base evaluates what the reference scan pushes while processing it: closed form over
 *                                the candidates at i-w..i when the last w+k bases are all ok, else a bounded replay
 *   lq_sketch_seq_k   HPC mode (-H) and k > 15: the restartable state machine itself, one thread per read
 * All of them write their records in base order (block scan of the per-thread counts + the tile's place from a decoupled
 * look-back over the tiles): the output is globally ordered by (read, position) == ascending y, which the index build needs.
 */
#include "lq_cuda.cuh"
#include "lq_sketch_core.h"
#include "lq_sketch_pk_core.h"
#include "lq_device.h"

#define SK_THREADS 256
#define SK_PER_THREAD 4
#define SK_TILE (SK_THREADS * SK_PER_THREAD)   /* 1024 bases */
#define SK_HALO 128                            /* candidates kept before the tile */
#define SK_HALO_W 256                          /* packed words kept before the tile */
#define SK_NPOS (SK_TILE + SK_HALO)

/* ------------------------------------------------------------------ K0: pack */

__global__ void lq_pack_k(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ seq_off, const uint64_t *__restrict__ slot0,
                          const uint32_t *__restrict__ len, uint32_t n_reads, uint64_t n_slots, int sdust_tbl,
                          uint32_t *__restrict__ b2, uint32_t *__restrict__ nm, uint32_t *__restrict__ slot_read,
                          const uint8_t *seq_lo, const uint8_t *seq_hi /* the bases of all reads lie in [seq_lo, seq_hi) */,
                          uint64_t slot_begin /* this launch packs slots [slot_begin, n_slots) */)
{
    /* one thread per 32 bases: two 2-bit words and one ambiguity word */
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + slot_begin * 4;
    const uint64_t slot = t >> 2;
    if (slot >= n_slots) return;
    /* owner read of the slot: last r with slot0[r] <= slot */
    uint32_t lo = 0, hi = n_reads;
    while (hi - lo > 1) { uint32_t mid = lo + ((hi - lo) >> 1); if (slot0[mid] <= slot) lo = mid; else hi = mid; }
    const uint32_t rd = lo;
    if ((t & 3) == 0) slot_read[slot] = rd;
    const uint64_t i0 = (slot - slot0[rd]) * LQ_SLOT + (t & 3) * 32;
    const uint64_t L = len[rd];
    const uint8_t *s = seq + seq_off[rd];
    uint32_t w0 = 0, w1 = 0, m = 0;
    const uint8_t *A = s + i0;
    const uintptr_t A16 = (uintptr_t)A & ~(uintptr_t)15;
    if (i0 + 32 <= L && A16 >= (uintptr_t)seq_lo && A16 + 48 <= (uintptr_t)seq_hi) {
        /* 32 whole bases: three aligned 16-byte loads cover them wherever the read starts; realigned with funnel shifts */
        const uint4 *P = (const uint4*)A16;
        const uint4 x0 = P[0], x1 = P[1], x2 = P[2];
        const uint32_t o = (uint32_t)((uintptr_t)A & 15), q = o >> 2, sh = (o & 3) * 8;
        const uint32_t w[12] = { x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w };
        uint32_t sel[9];
        #pragma unroll
        for (int u = 0; u < 9; ++u) sel[u] = q == 0 ? w[u] : q == 1 ? w[u + 1] : q == 2 ? w[u + 2] : w[u + 3];
        #pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t v = __funnelshift_r(sel[u], sel[u + 1], sh);
            #pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 4 * u + e;
                const uint32_t c = lq_nt4((v >> (8 * e)) & 255u, sdust_tbl);
                if (c < 4) { if (j < 16) w0 |= c << (2 * j); else w1 |= c << (2 * (j - 16)); }
                else m |= 1u << j;
            }
        }
    } else {
        #pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            uint32_t c = 4;
            if (i0 + j < L) c = lq_nt4(s[i0 + j], sdust_tbl);
            if (c < 4) { if (j < 16) w0 |= c << (2 * j); else w1 |= c << (2 * (j - 16)); }
            else m |= 1u << j;
        }
    }
    b2[t * 2] = w0; b2[t * 2 + 1] = w1; nm[t] = m;
}

/* ------------------------------------------------------------------ K1: sketch */

struct SkCount { int n; __device__ __forceinline__ void operator()(uint64_t, uint64_t) { ++n; } };
/* counts, and keeps the first two records of a base in registers (more than two is rare: that base is evaluated again) */
struct SkHold {
    uint32_t k0, p0, k1, p1; int n;
    __device__ __forceinline__ void operator()(uint64_t x_, uint64_t y_) {
        if (n == 0) { k0 = (uint32_t)(x_ >> 8); p0 = (uint32_t)y_; } else if (n == 1) { k1 = (uint32_t)(x_ >> 8); p1 = (uint32_t)y_; }
        ++n;
    }
};
struct SkWrite {
    uint32_t *key; uint64_t *yy; uint64_t at;
    __device__ __forceinline__ void operator()(uint64_t x_, uint64_t y_) { key[at] = (uint32_t)(x_ >> 8); yy[at] = y_; ++at; }
};

struct SkArgs {
    const uint32_t *b2, *nm, *slot_read, *len;
    const uint64_t *slot0;
    uint64_t n_slots;
    int w, k;
    uint32_t rid_base;
    /* single-pass ordering of the output (decoupled look-back over the tiles) */
    uint32_t *ticket;            /* dynamic tile id: tiles are numbered in the order their CTAs start */
    unsigned long long *state;   /* per tile: flag<<62 | count; flag 1 = tile total, 2 = inclusive prefix */
    uint64_t cap;                /* capacity of out_key / out_y in records */
    uint32_t *err;               /* bit 0: look-back timed out */
    uint32_t *out_key; uint64_t *out_y;
    /* chunked launches (lq_upload_sketch_pipelined): this launch covers the packed bases from g_begin to slot n_slots; its records
     * follow the *base_in of the launches before it, and its last tile leaves the running total in *base_out */
    uint64_t g_begin; const unsigned long long *base_in; unsigned long long *base_out;
};

/* candidate of smem position p (hash of min(fw,rv), strand) */
struct SkFetch {
    const uint32_t *cand, *zb; int idx;
    __device__ __forceinline__ void operator()(int d, uint32_t *h, uint32_t *z) const { const int p = idx - d; *h = cand[p]; *z = (zb[p >> 5] >> (p & 31)) & 1u; }
};

__device__ __forceinline__ uint64_t sk_bits64(const uint32_t *w, int lo) /* bits lo..lo+63 of a bit array (lo >= 0) */
{
    const uint32_t wi = (uint32_t)lo >> 5, sh = (uint32_t)lo & 31;
    uint64_t v = ((uint64_t)w[wi] | (uint64_t)w[wi + 1] << 32) >> sh;
    if (sh) v |= (uint64_t)w[wi + 2] << (64 - sh);
    return v;
}

/* the bounded replay, out of line: it is rare, and inlining it four times per thread costs registers on the common path */
__device__ __noinline__ void sk_slow(const uint32_t *b2, const uint32_t *nm, uint64_t g0, int L, int w, int k, uint32_t rid, int i, lq_sk_buf *buf)
{
    buf->n = 0;
    lq_sketch_slow_at(b2, nm, g0, L, w, k, rid, i, *buf);
}

/* what the reference pushes while processing base g (tile-relative candidate index idx); returns the read's rid */
template <int WT, int WC, class Sink>
__device__ __forceinline__ uint32_t sk_eval(const SkArgs &a, const uint32_t *cand, const uint32_t *okb, const uint32_t *zb, const uint32_t *s_nm,
                                            int idx, uint64_t g, Sink &sink)
{
    const uint64_t slot = g >> 7;
    if (slot >= a.n_slots) return 0;
    const uint32_t rd = a.slot_read[slot];
    const uint64_t s0 = a.slot0[rd];
    const int L = (int)a.len[rd];
    const int i = (int)(slot - s0) * LQ_SLOT + (int)(g & 127);
    if (i >= L) return 0;
    uint64_t okw = sk_bits64(okb, idx - 63);
    uint64_t ambw = sk_bits64(s_nm, idx + (SK_HALO_W - SK_HALO) - 63);   /* s_nm is indexed from W0 = T0 - SK_HALO_W, cand/okb from T0 - SK_HALO */
    if (i < 63) { const uint64_t keep = ~0ULL << (63 - i); okw &= keep; ambw &= keep; }   /* nothing before the read start */
    SkFetch f; f.cand = cand; f.zb = zb; f.idx = idx;
    if (!lq_sketch_fast_win<WT, WC>(okw, ambw, a.w, a.k, a.rid_base + rd, i, i == L - 1, f, sink)) {
        lq_sk_buf buf;
        sk_slow(a.b2, a.nm, s0 * LQ_SLOT, L, a.w, a.k, a.rid_base + rd, i, &buf);
        for (int j = 0; j < buf.n; ++j) sink(buf.x[j], buf.y[j]);
    }
    return a.rid_base + rd;
}

/* 32-bit k-mer / hash math for k <= 16 (the common case: LongQC uses 12 and 15) */
__device__ __forceinline__ int sk_cand32(const uint32_t *b2, uint64_t g, int k, uint32_t *hash, uint32_t *strand)
{
    const uint64_t first = g - (uint64_t)(k - 1);
    const uint32_t wi = (uint32_t)(first >> 4), sh = (uint32_t)(first & 15) * 2;
    const uint32_t mask = k == 16 ? 0xffffffffu : ((1u << 2 * k) - 1);
    const uint32_t le = __funnelshift_r(b2[wi], b2[wi + 1], sh) & mask;     /* oldest base in the low bits */
    const uint32_t rv = ~le & mask;                                         /* sketch.c:106 */
    uint32_t v = __brev(le);                                                /* reverse the 2-bit groups: bit reversal, then swap within pairs */
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    const uint32_t fw = v >> (32 - 2 * k);                                  /* sketch.c:105 */
    if (fw == rv) return 0;
    *strand = fw < rv ? 0u : 1u;
    *hash = lq_hash32(fw < rv ? fw : rv, mask);
    return 1;
}

template <int WT, int WC>
__global__ void __launch_bounds__(SK_THREADS) lq_sketch_k(SkArgs a)
{
    __shared__ uint32_t s_b2[(SK_TILE + SK_HALO_W) / 16 + 4];
    __shared__ uint32_t s_nm[(SK_TILE + SK_HALO_W) / 32 + 2];
    __shared__ uint32_t cand[SK_NPOS];
    __shared__ uint32_t okb[SK_NPOS / 32 + 2], zb[SK_NPOS / 32 + 2];
    __shared__ uint64_t scan_sm[33];
    __shared__ uint32_t s_tile;
    __shared__ uint4 s_hold[SK_PER_THREAD][SK_THREADS];
    __shared__ uint32_t s_rid[SK_PER_THREAD][SK_THREADS];
    __shared__ uint64_t s_base;

    const int tid = threadIdx.x;
    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t T0 = (int64_t)tile * SK_TILE;          /* first base of the tile (global base index) */
    const int64_t W0 = T0 - SK_HALO_W;                    /* first base held in shared memory */
    const int64_t n_bases = (int64_t)a.n_slots * LQ_SLOT;

    /* stage 0 */
    for (int j = tid; j < (SK_TILE + SK_HALO_W) / 16 + 4; j += SK_THREADS) {
        int64_t wi = W0 / 16 + j;
        s_b2[j] = (wi >= 0 && wi < n_bases / 16) ? a.b2[wi] : 0u;
    }
    for (int j = tid; j < (SK_TILE + SK_HALO_W) / 32 + 2; j += SK_THREADS) {
        int64_t wi = W0 / 32 + j;
        s_nm[j] = (wi >= 0 && wi < n_bases / 32) ? a.nm[wi] : 0xffffffffu;
    }
    if (tid < 2) { okb[SK_NPOS / 32 + tid] = 0; zb[SK_NPOS / 32 + tid] = 0; }
    __syncthreads();

    /* stage A: candidates for bases T0-128 .. T0+1023 (idx 0..1151); a warp covers 32 consecutive idx = one quarter of a slot,
     * so the owning read is looked up once per warp */
    for (int idx = tid; idx < SK_NPOS; idx += SK_THREADS) {
        const int64_t g = T0 - SK_HALO + idx;
        uint32_t ok = 0, z = 0, h32 = 0;
        if (g >= 0 && (g >> 7) < (int64_t)a.n_slots) {
            const uint64_t slot = (uint64_t)g >> 7;
            const uint32_t rd = a.slot_read[slot];
            const uint64_t s0 = a.slot0[rd];
            const int L = (int)a.len[rd];
            const int i = (int)(slot - s0) * LQ_SLOT + (int)((uint32_t)g & 127);
            const uint32_t sg = (uint32_t)(idx + (SK_HALO_W - SK_HALO));          /* index relative to the shared copies */
            if (i < L && !((s_nm[sg >> 5] >> (sg & 31)) & 1u)) {
                /* ambiguity bits of the k-1 bases before: bits sg-k+1 .. sg-1 */
                const uint32_t lo = sg - (uint32_t)(a.k - 1);
                const uint32_t wv = __funnelshift_r(s_nm[lo >> 5], s_nm[(lo >> 5) + 1], lo & 31) & ((1u << (a.k - 1)) - 1);
                if (i >= a.k - 1 && wv == 0) {
                    ok = (uint32_t)sk_cand32(s_b2, sg, a.k, &h32, &z);
                } else { /* k-mer registers carry bits from before an ambiguous base / the read start */
                    uint64_t fw, rv;
                    lq_regs_at(a.b2, a.nm, s0 * LQ_SLOT, i, a.k, &fw, &rv);
                    ok = fw != rv;
                }
            }
        }
        cand[idx] = h32;
        const uint32_t okm = __ballot_sync(0xffffffffu, ok), zm = __ballot_sync(0xffffffffu, z);
        if ((tid & 31) == 0) { okb[idx >> 5] = okm; zb[idx >> 5] = zm; }
    }
    __syncthreads();

    /* stage B: evaluate every base of the tile once; rows r = 0..3, base = T0 + r*256 + tid.  The (at most two) records of
     * a base wait in shared memory until the tile's output offset is known. */
    int cnt[SK_PER_THREAD];
    #pragma unroll
    for (int r = 0; r < SK_PER_THREAD; ++r) {
        SkHold h; h.n = 0; h.k0 = h.p0 = h.k1 = h.p1 = 0;
        const uint32_t rid = sk_eval<WT, WC>(a, cand, okb, zb, s_nm, SK_HALO + r * SK_THREADS + tid, (uint64_t)(T0 + r * SK_THREADS + tid), h);
        cnt[r] = h.n;
        s_hold[r][tid] = make_uint4(h.k0, h.p0, h.k1, h.p1);
        s_rid[r][tid] = rid;
    }
    /* one block scan for the four rows: 16 bits per row (<= 256*(2w+2) < 65536 records per row) */
    const uint64_t packed = (uint64_t)cnt[0] | (uint64_t)cnt[1] << 16 | (uint64_t)cnt[2] << 32 | (uint64_t)cnt[3] << 48;
    uint64_t tot;
    const uint64_t ex = lq_block_excl_scan(packed, scan_sm, &tot);
    const uint64_t tile_total = (tot & 0xffff) + ((tot >> 16) & 0xffff) + ((tot >> 32) & 0xffff) + (tot >> 48);

    /* decoupled look-back (warp 0): exclusive prefix of the tile totals, tiles in ticket order */
    if (tid < 32) {
        const unsigned long long FLAG_AGG = 1ULL << 62, FLAG_PRE = 2ULL << 62, VMASK = (1ULL << 62) - 1;
        uint64_t base = 0;
        if (tile == 0) { if (tid == 0) atomicExch(&a.state[0], FLAG_PRE | tile_total); }
        else {
            if (tid == 0) atomicExch(&a.state[tile], FLAG_AGG | tile_total);
            int64_t hi = (int64_t)tile - 1; uint32_t spins = 0;
            for (;;) {   /* lanes look at tiles hi, hi-1, ..., hi-31 */
                const int64_t j = hi - tid;
                unsigned long long sv = FLAG_PRE;                     /* tiles before the first one: an empty prefix */
                if (j >= 0) sv = *(volatile unsigned long long*)&a.state[j];
                const uint32_t ready = __ballot_sync(0xffffffffu, (sv >> 62) != 0);
                const uint32_t pre = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
                /* usable lanes: 0 .. first PREFIX lane, all of which must be ready */
                const int stop = pre ? __ffs(pre) - 1 : 31;
                const uint32_t need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1);
                if ((ready & need) != need) { __nanosleep(spins < 5 ? 32u << spins : 1024u); if (++spins > (1u << 22)) { if (tid == 0) atomicOr(a.err, 1u); break; } continue; }
                uint64_t v = (tid <= stop) ? (sv & VMASK) : 0;
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                base += v;
                if (pre) break;
                hi -= 32;
            }
            if (tid == 0) atomicExch(&a.state[tile], FLAG_PRE | (base + tile_total));
        }
        if (tid == 0) s_base = base;
    }
    __syncthreads();
    uint64_t base = s_base;
    if (base + tile_total > a.cap) return;     /* output buffer too small: the host re-runs with the exact size */

    /* write in base order */
    #pragma unroll
    for (int r = 0; r < SK_PER_THREAD; ++r) {
        const uint64_t at = base + ((ex >> (16 * r)) & 0xffff);
        if (cnt[r] > 2) {   /* rare (equal minimizers inside one window): evaluate again, writing directly */
            SkWrite wr; wr.key = a.out_key; wr.yy = a.out_y; wr.at = at;
            sk_eval<WT, WC>(a, cand, okb, zb, s_nm, SK_HALO + r * SK_THREADS + tid, (uint64_t)(T0 + r * SK_THREADS + tid), wr);
        } else if (cnt[r] > 0) {
            const uint4 h = s_hold[r][tid]; const uint64_t hi = (uint64_t)s_rid[r][tid] << 32;
            a.out_key[at] = h.x; a.out_y[at] = hi | h.y;
            if (cnt[r] > 1) { a.out_key[at + 1] = h.z; a.out_y[at + 1] = hi | h.w; }
        }
        base += (tot >> (16 * r)) & 0xffff;
    }
}


/* ------------------------------------------------------------------ K1, rolling form (w = 5 or 10, k <= 15: LongQC's settings)
 *
 * One thread owns RK_SEG = 64 consecutive bases of a read and runs the reference scan itself, with the k-mer registers
 * rolled base by base and the ring of the last W candidates held in registers as a shift register (oldest first).
 * It starts RK_WU = 32 bases early with an empty state and emits nothing until its state is CERTIFIED equal to the
 * reference's:
 *   - a segment that starts within RK_WU bases of the read start begins at base 0 in the true initial state;
 *   - otherwise `good` counts the ring pushes made since the registers became exact (k unambiguous bases consumed) and
 *     since the last ambiguous base: once good >= w+k the true run counter is >= w+k (all gates of sketch.c:116-137 open),
 *     the ring holds only pushes made in that stretch, all valid -- the state the reference has (lq_sketch_core.h, fact 1);
 *   - an ambiguous base met with exact registers resets the run counter to 0 in both, which is exact from then on.
 * Bases of the segment reached before certification (a palindromic k-mer or an N inside the warm-up: rare) take the
 * bounded replay (sk_slow).  Records wait in shared memory (transposed, conflict-free) until the CTA's output offset is
 * known from the look-back; a thread that emits more than RK_CAP records makes the CTA re-run the scan writing directly. */
#define RK_SEG 64
#define RK_WU 32
#define RK_THREADS 128
#define RK_CAP 32
#define RK_TILE (RK_SEG * RK_THREADS)

template <int W, int MODE /* 0: buffer in smem, 1: write directly at `wat` */, int THREADS = RK_THREADS, int CAP = RK_CAP /* staging geometry of the calling kernel */, bool INL = false /* the replay inline: no call in the kernel */>
__device__ __forceinline__ int rk_scan(const SkArgs &a, uint32_t rd, uint64_t g0, int L, int i0, int i1 /* segment [i0,i1) */,
                                       uint2 *s_rec, int tid, uint64_t wat)
{
    const int w = W, k = a.k;
    const uint32_t mask = (1u << 2 * k) - 1, top = 2 * (k - 1);
    const uint32_t MAXH = 0xffffffffu;
    const uint32_t rid = a.rid_base + rd;
    uint32_t wx[W], wp[W];                 /* ring as a shift register: [0] oldest ... [W-1] newest; wp = pos<<1|strand */
    #pragma unroll
    for (int j = 0; j < W; ++j) { wx[j] = MAXH; wp[j] = MAXH; }
    uint32_t mx = MAXH, mp = MAXH; int mi = 0;   /* running minimum (copy) and its index in the shift register */
    uint32_t fw = 0, rv = 0;
    const int p0 = i0 - RK_WU > 0 ? i0 - RK_WU : 0;
    int run = 0;                           /* the reference's l when exact, else a lower bound */
    int nb = p0 == 0 ? 1 << 20 : 0;        /* unambiguous bases consumed (registers exact once nb >= k) */
    int good = p0 == 0 ? 1 << 20 : 0;
    bool cert = p0 == 0, lexact = p0 == 0;
    int n = 0;
    bool last_slow = false;
    lq_sk_buf sbuf;
    #define RK_EMIT(H_, P_) do { if (MODE == 0) { if (n < CAP) s_rec[n * THREADS + tid] = make_uint2((H_), (P_)); } \
                                 else { a.out_key[wat + n] = (H_); a.out_y[wat + n] = (uint64_t)rid << 32 | (P_); } ++n; } while (0)
    uint32_t wb = 0, wn = 0;               /* current packed words */
    int i = p0;
    if (i < i1) { wb = a.b2[(g0 + (uint64_t)i) >> 4]; wn = a.nm[(g0 + (uint64_t)i) >> 5]; }
    while (i < i1) {
        /* ---- steady state: certified, every gate open, no ambiguous base: the common case, kept lean ---- */
        if (cert && (lexact ? run >= w + k : good >= w + k)) {
            while (i < i1) {
                if ((i & 15) == 0) wb = a.b2[(g0 + (uint64_t)i) >> 4];
                if ((i & 31) == 0) wn = a.nm[(g0 + (uint64_t)i) >> 5];
                if ((wn >> (i & 31)) & 1u) break;                       /* ambiguous base: back to the general step */
                const uint32_t c = (wb >> ((i & 15) * 2)) & 3u;
                fw = (fw << 2 | c) & mask;
                rv = rv >> 2 | (3u ^ c) << top;
                if (fw != rv) {                                          /* else sketch.c:107: no push */
                    const uint32_t z = fw < rv ? 0u : 1u;
                    const uint32_t cx = lq_hash32(z ? rv : fw, mask), cp = (uint32_t)i << 1 | z;
                    const bool em = i >= i0;
                    if (cx <= mx) {
                        if (em) RK_EMIT(mx, mp);
                        mx = cx; mp = cp; mi = W;
                    } else if (mi == 0) {
                        if (em) RK_EMIT(mx, mp);
                        mx = MAXH; mp = MAXH;
                        #pragma unroll
                        for (int j = 1; j < W; ++j) if (mx >= wx[j]) { mx = wx[j]; mp = wp[j]; mi = j; }
                        if (mx >= cx) { mx = cx; mp = cp; mi = W; }
                        bool twin = cx == mx && cp != mp;
                        #pragma unroll
                        for (int j = 1; j < W; ++j) twin |= wx[j] == mx && wp[j] != mp;
                        if (twin && em) {
                            #pragma unroll
                            for (int j = 1; j < W; ++j) if (wx[j] == mx && wp[j] != mp) RK_EMIT(wx[j], wp[j]);
                            if (cx == mx && cp != mp) RK_EMIT(cx, cp);
                        }
                    }
                    #pragma unroll
                    for (int j = 0; j + 1 < W; ++j) { wx[j] = wx[j + 1]; wp[j] = wp[j + 1]; }
                    wx[W - 1] = cx; wp[W - 1] = cp;
                    --mi; ++run; ++good;
                }
                ++nb; ++i;
            }
            if (i >= i1) break;
        }
        /* ---- general step (warm-up, first windows of a run, ambiguous bases) ---- */
        const uint64_t g = g0 + (uint64_t)i;
        if ((i & 15) == 0) wb = a.b2[g >> 4];
        if ((i & 31) == 0) wn = a.nm[g >> 5];
        const uint32_t c = (wb >> ((i & 15) * 2)) & 3u;
        const bool amb = (wn >> (i & 31)) & 1u;
        const bool out = i >= i0;
        const bool use_slow = out && !cert;
        if (i == i1 - 1) last_slow = use_slow;
        if (use_slow) {                    /* not certified yet: this base's records come from the bounded replay */
            if (INL) { sbuf.n = 0; lq_sketch_slow_at(a.b2, a.nm, g0, L, w, k, rid, i, sbuf); }
            else sk_slow(a.b2, a.nm, g0, L, w, k, rid, i, &sbuf);
            for (int j = 0; j < sbuf.n; ++j) RK_EMIT((uint32_t)(sbuf.x[j] >> 8), (uint32_t)sbuf.y[j]);
        }
        uint32_t cx = MAXH, cp = MAXH;
        bool push = true;
        if (!amb) {
            fw = (fw << 2 | c) & mask;
            rv = rv >> 2 | (3u ^ c) << top;
            ++nb;
            if (fw == rv) push = false;                /* sketch.c:107 */
            else {
                const uint32_t z = fw < rv ? 0u : 1u;
                ++run;
                if (nb >= k) ++good; else good = 0;
                if (run >= k) { cx = lq_hash32(z ? rv : fw, mask); cp = (uint32_t)i << 1 | z; }
            }
        } else {
            if (nb >= k) { cert = true; lexact = true; }   /* registers exact: the reset is exact */
            run = 0; good = 0;
        }
        if (push) {
            if (!lexact && good >= w + k) cert = true;
            const int l = (lexact || !cert) ? run : (run > w + k ? run : w + k);   /* certified by `good`: every gate is open */
            const bool em = out && !use_slow;      /* certified at the top of this step */
            /* the slot being overwritten is wx[0]; (A) first full window, sketch.c:116-121: entries older than the newcomer */
            if (l == w + k - 1 && mx != MAXH) {
                #pragma unroll
                for (int j = 1; j < W; ++j) if (wx[j] == mx && wp[j] != mp && em) RK_EMIT(wx[j], wp[j]);
            }
            if (cx <= mx) {                                 /* sketch.c:122-124 */
                if (l >= w + k && mx != MAXH && em) RK_EMIT(mx, mp);
                mx = cx; mp = cp; mi = W;                   /* index after the shift below: W-1 */
            } else if (mi == 0) {                           /* sketch.c:125-137: the minimum's slot is overwritten */
                if (l >= w + k - 1 && mx != MAXH && em) RK_EMIT(mx, mp);
                mx = MAXH; mp = MAXH; mi = 1;
                #pragma unroll
                for (int j = 1; j < W; ++j) if (mx >= wx[j]) { mx = wx[j]; mp = wp[j]; mi = j; }
                if (mx >= cx) { mx = cx; mp = cp; mi = W; }
                if (l >= w + k - 1 && mx != MAXH) {
                    #pragma unroll
                    for (int j = 1; j < W; ++j) if (wx[j] == mx && wp[j] != mp && em) RK_EMIT(wx[j], wp[j]);
                    if (cx == mx && cp != mp && em) RK_EMIT(cx, cp);
                }
            }
            #pragma unroll
            for (int j = 0; j + 1 < W; ++j) { wx[j] = wx[j + 1]; wp[j] = wp[j + 1]; }
            wx[W - 1] = cx; wp[W - 1] = cp;
            --mi;
        }
        ++i;
    }
    if (i1 == L && !last_slow) {           /* sketch.c:140-141; a last base that took the replay already got this record from it */
        if (mx != MAXH) RK_EMIT(mx, mp);
    }
    #undef RK_EMIT
    return n;
}

/* decoupled look-back over the tiles (called by warp 0 of the CTA): publishes this tile's record count, returns through *s_base
 * the number of records before it (those of the launches before this one included) */
__device__ __forceinline__ void sk_tile_base(const SkArgs &a, const uint32_t tile, const uint64_t tot, const int tid, uint64_t *s_base_p)
{
    uint64_t &s_base = *s_base_p;
    {
        const unsigned long long FLAG_AGG = 1ULL << 62, FLAG_PRE = 2ULL << 62, VMASK = (1ULL << 62) - 1;
        uint64_t base = 0;
        if (tile == 0) { if (tid == 0) atomicExch(&a.state[0], FLAG_PRE | tot); }
        else {
            if (tid == 0) atomicExch(&a.state[tile], FLAG_AGG | tot);
            int64_t hi = (int64_t)tile - 1; uint32_t spins = 0;
            for (;;) {   /* lanes look at tiles hi, hi-1, ..., hi-31 (rounds of 128 were measured: no faster -- the wait is for the
                          * predecessors to finish, not for the walk) */
                const int64_t j = hi - tid;
                unsigned long long sv = FLAG_PRE;
                if (j >= 0) sv = *(volatile unsigned long long*)&a.state[j];
                const uint32_t ready = __ballot_sync(0xffffffffu, (sv >> 62) != 0);
                const uint32_t pre = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
                const int stop = pre ? __ffs(pre) - 1 : 31;
                const uint32_t need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1);
                if ((ready & need) != need) { __nanosleep(spins < 5 ? 32u << spins : 1024u); if (++spins > (1u << 22)) { if (tid == 0) atomicOr(a.err, 1u); break; } continue; }
                uint64_t v = (tid <= stop) ? (sv & VMASK) : 0;
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                base += v;
                if (pre) break;
                hi -= 32;
            }
            if (tid == 0) atomicExch(&a.state[tile], FLAG_PRE | (base + tot));
        }
        if (tid == 0) {
            const uint64_t cb = a.base_in ? *a.base_in : 0;                  /* records of the launches before this one */
            s_base = cb + base;
            if (a.base_out && tile == gridDim.x - 1) *a.base_out = cb + base + tot;
        }
    }
}

template <int W>
__global__ void __launch_bounds__(RK_THREADS) lq_sketch_roll_k(SkArgs a)
{
    __shared__ uint2 s_rec[RK_CAP * RK_THREADS];       /* 32 KB */
    __shared__ uint64_t scan_sm[33];
    __shared__ uint32_t s_tile; __shared__ uint64_t s_base; __shared__ int s_over;
    const int tid = threadIdx.x;
    if (tid == 0) { s_tile = atomicAdd(a.ticket, 1u); s_over = 0; }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t g = a.g_begin + ((uint64_t)tile * RK_THREADS + tid) * RK_SEG;   /* global base index of the segment */
    const uint64_t slot = g >> 7;
    uint32_t rd = 0; uint64_t g0 = 0; int L = 0, i0 = 0, i1 = 0;
    if (slot < a.n_slots) {
        rd = a.slot_read[slot];
        const uint64_t s0 = a.slot0[rd];
        L = (int)a.len[rd]; g0 = s0 * LQ_SLOT;
        i0 = (int)(slot - s0) * LQ_SLOT + (int)(g & 127);
        i1 = i0 + RK_SEG < L ? i0 + RK_SEG : L;
    }
    int n = 0;
    if (i0 < i1) n = rk_scan<W, 0>(a, rd, g0, L, i0, i1, s_rec, tid, 0);
    if (n > RK_CAP) s_over = 1;
    uint64_t tot;
    const uint64_t ex = lq_block_excl_scan((uint64_t)n, scan_sm, &tot);
    if (tid < 32) sk_tile_base(a, tile, tot, tid, &s_base);   /* decoupled look-back (warp 0) */
    __syncthreads();
    const uint64_t at = s_base + ex;
    if (s_base + tot > a.cap || n == 0) return;
    if (!s_over) {
        for (int j = 0; j < n; ++j) { const uint2 r = s_rec[j * RK_THREADS + tid]; a.out_key[at + j] = r.x; a.out_y[at + j] = (uint64_t)(a.rid_base + rd) << 32 | r.y; }
    } else rk_scan<W, 1>(a, rd, g0, L, i0, i1, s_rec, tid, at);   /* low-complexity tile: scan again, writing in place */
}

/* ------------------------------------------------------------------ K1, packed-key form (w = 5, k <= 15): the default
 *
 * lq_sketch_pk_core.h: a thread scans 64 bases of a read out of two 16-byte shared-memory loads, its candidates are single
 * integers (hash | inverted position), the window minimum two 3-input minima, the look-back 25-28 bases of which 8 are hashed.
 * The tile's packed words (64 slots and the one before: 2 KB + 1 KB) are brought into shared memory by two bulk asynchronous
 * copies (cp.async.bulk, completion on an mbarrier) issued by one thread while the others fetch their read's geometry.
 * Records wait in shared memory as keys, a row per thread (odd row stride: no bank conflicts when the lanes of a warp write
 * their rows, none when the warp later reads ONE row); after the look-back has given the tile its place, each warp writes the
 * rows of its threads one after the other, lanes along the row: runs of ~21 consecutive records per store.
 * Segments the form declines (ambiguous bases, palindrome-rich look-back, equal k-mers inside a read's first window) run
 * rk_scan into a small pool; a tile with more of those than the pool holds, or with a row that overflows, runs rk_scan
 * everywhere, writing in place. */
#define PK_CAP 32                        /* records a thread can stage: 64 bases write 21.5 on average, more than 30 with probability 1e-5.
                                          * The unrolled blocks store unchecked: a row that overflows runs into the next one (into the
                                          * pool behind the last one), which only ever happens in a tile that is then scanned again in place */
#define PK_STRIDE (PK_CAP + 1)
#ifndef PK_MIN_CTAS
#define PK_MIN_CTAS 8                    /* 64 registers: with 23 KB of shared memory per CTA, 8 to 9 CTAs per SM */
#endif
#define PK_FB 6
#define PK_SLOTS (RK_TILE / LQ_SLOT)     /* 64 slots per tile */
#define PK_OFF_NM ((PK_SLOTS + 1) * LQ_SLOT_W2 * 4)
#define PK_OFF_STAGE (PK_OFF_NM + (PK_SLOTS + 1) * LQ_SLOT_WN * 4)
#define PK_SMEM(KEYBYTES) (PK_OFF_STAGE + RK_THREADS * PK_STRIDE * (KEYBYTES) + PK_FB * PK_CAP * 8)

__device__ __forceinline__ uint32_t lq_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lq_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void lq_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
/* global -> shared bulk copy (TMA engine, no tensor map): 16-byte aligned on both sides, bytes a multiple of 16 */
__device__ __forceinline__ void lq_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t lq_mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}

/* a thread's row of staged keys, addressed in the shared window (32-bit arithmetic): put() is the unrolled blocks'
 * store-always / keep-if, push() checks */
__device__ __forceinline__ void lq_sts(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void lq_sts(uint32_t addr, uint64_t v) { asm volatile("st.shared.b64 [%0], %1;" :: "r"(addr), "l"(v) : "memory"); }
template <class KEY> struct PkStage {
    typedef uint32_t mark_t;
    uint32_t w, row;                   /* shared-window byte addresses */
    __device__ __forceinline__ void put(KEY kk, bool yes) { lq_sts(w, kk); w += yes ? (uint32_t)sizeof(KEY) : 0u; }
    __device__ __forceinline__ void push(KEY kk) { if (w - row < PK_CAP * sizeof(KEY)) lq_sts(w, kk); w += (uint32_t)sizeof(KEY); }
    __device__ __forceinline__ int count() const { return (int)((w - row) / (uint32_t)sizeof(KEY)); }
    __device__ __forceinline__ int room() const { return PK_CAP - count(); }
    __device__ __forceinline__ mark_t mark() const { return w; }
    __device__ __forceinline__ void rewind(mark_t m) { w = m; }
};
/* the general state machine for one segment (inline, replay included: a call in this kernel costs the common path spills) */
template <int W>
__device__ __forceinline__ int pk_general(const SkArgs &a, uint32_t rd, uint64_t g0, int L, int i0, int i1, uint2 *rec)
{
    return rk_scan<W, 0, 1, PK_CAP, true>(a, rd, g0, L, i0, i1, rec, 0, 0);
}
template <int W>
__device__ __forceinline__ void pk_general_inplace(const SkArgs &a, uint32_t rd, uint64_t g0, int L, int i0, int i1, uint64_t at)
{
    rk_scan<W, 1, RK_THREADS, RK_CAP, true>(a, rd, g0, L, i0, i1, (uint2*)0, 0, at);
}

extern __shared__ __align__(16) unsigned char pk_smem[];

template <int W, int K, int MINB>
__global__ void __launch_bounds__(RK_THREADS, MINB) lq_sketch_pk_k(SkArgs a, int use_bulk)
{
    typedef lq_pk_tr<(K > 12)> T;
    typedef typename T::key key;
    uint32_t *s_b2 = (uint32_t*)pk_smem;                                    /* the slot before the tile, then the tile */
    uint32_t *s_nm = (uint32_t*)(pk_smem + PK_OFF_NM);
    key *s_stage = (key*)(pk_smem + PK_OFF_STAGE);
    uint2 *s_fb = (uint2*)(pk_smem + PK_OFF_STAGE + RK_THREADS * PK_STRIDE * sizeof(key));
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint64_t scan_sm[33];
    __shared__ uint32_t s_tile, s_nfb; __shared__ uint64_t s_base; __shared__ int s_over;
    const int tid = threadIdx.x;
    const uint32_t bar = lq_smem_u32(&s_bar);
    if (tid == 0) { s_tile = atomicAdd(a.ticket, 1u); s_over = 0; s_nfb = 0; if (use_bulk) lq_mbar_init(bar, 1); }
    __syncthreads();
    const uint32_t tile = s_tile;
    {
        const uint64_t t0 = (a.g_begin >> 7) + (uint64_t)tile * PK_SLOTS;              /* first slot of the tile */
        const uint64_t c0 = t0 ? t0 - 1 : 0, c1 = t0 + PK_SLOTS < a.n_slots ? t0 + PK_SLOTS : a.n_slots;
        const uint32_t off = t0 ? 0u : 1u, ns = (uint32_t)(c1 - c0);                    /* slots copied, and where the first one lands */
        if (use_bulk) {
            if (tid == 0) {
                lq_mbar_expect_tx(bar, ns * (LQ_SLOT_W2 + LQ_SLOT_WN) * 4);
                lq_bulk_g2s(lq_smem_u32(s_b2 + off * LQ_SLOT_W2), a.b2 + c0 * LQ_SLOT_W2, ns * LQ_SLOT_W2 * 4, bar);
                lq_bulk_g2s(lq_smem_u32(s_nm + off * LQ_SLOT_WN), a.nm + c0 * LQ_SLOT_WN, ns * LQ_SLOT_WN * 4, bar);
            }
        } else {
            const uint4 *gb = (const uint4*)(a.b2 + c0 * LQ_SLOT_W2), *gn = (const uint4*)(a.nm + c0 * LQ_SLOT_WN);
            uint4 *sb = (uint4*)(s_b2 + off * LQ_SLOT_W2), *sn = (uint4*)(s_nm + off * LQ_SLOT_WN);
            for (uint32_t i = tid; i < ns * 2; i += RK_THREADS) sb[i] = gb[i];
            for (uint32_t i = tid; i < ns; i += RK_THREADS) sn[i] = gn[i];
        }
    }
    const uint64_t g = a.g_begin + ((uint64_t)tile * RK_THREADS + tid) * RK_SEG;   /* global base index of the segment */
    const uint64_t slot = g >> 7;
    uint32_t rd = 0; uint64_t g0 = 0; int L = 0, i0 = 0, i1 = 0;
    if (slot < a.n_slots) {
        rd = a.slot_read[slot];
        const uint64_t s0 = a.slot0[rd];
        L = (int)a.len[rd]; g0 = s0 * LQ_SLOT;
        i0 = (int)(slot - s0) * LQ_SLOT + (int)(g & 127);
        i1 = i0 + RK_SEG < L ? i0 + RK_SEG : L;
    }
    if (use_bulk) {
        uint32_t spins = 0;
        while (!lq_mbar_try_wait(bar, 0)) { if (++spins > (1u << 22)) { atomicOr(a.err, 2u); break; } }
    } else __syncthreads();
    int n = 0, fbi = -1;
    if (i0 < i1) {
        PkStage<key> sink; sink.row = sink.w = lq_smem_u32(s_stage + tid * PK_STRIDE);
        const int r = lq_pk_segment<W, K>(s_b2 + 4 + tid * 4, s_nm + 2 + tid * 2, i0, i1 - i0, i1 == L, sink);
        if (r == 0) n = sink.count();
        else {
            fbi = (int)atomicAdd(&s_nfb, 1u);
            if (fbi >= PK_FB || r == 2) s_over = 1;
        }
    }
    __syncthreads();
    if (s_nfb) {                                              /* rare: some segments of the tile take the general state machine */
        if (fbi >= 0) {
            n = pk_general<W>(a, rd, g0, L, i0, i1, s_fb + (fbi < PK_FB ? fbi : 0) * PK_CAP);   /* past the pool: only the count matters */
            if (n > PK_CAP) s_over = 1;
        }
    }
    uint64_t tot;
    const uint64_t ex = lq_block_excl_scan((uint64_t)n, scan_sm, &tot);
    if (tid < 32) sk_tile_base(a, tile, tot, tid, &s_base);   /* decoupled look-back (warp 0) */
    __syncthreads();
    if (s_base + tot > a.cap) return;
    const uint64_t at = s_base + ex;
    if (s_over) { if (n) pk_general_inplace<W>(a, rd, g0, L, i0, i1, at); return; }   /* low-complexity tile */
    /* A warp's records are one stretch of the output.  Every thread leaves what a writer needs to know about its row where the
     * packed words were (all warps are past them); then lane l of round i writes record 32 i + l of the stretch: it finds the
     * row by bisection over the rows' first records (5 shared loads), and the warp's stores are whole sectors. */
    const uint32_t lane = tid & 31, wb = tid & ~31u;
    const uint32_t wex0 = __shfl_sync(0xffffffffu, (uint32_t)ex, 0), wtot = __shfl_sync(0xffffffffu, (uint32_t)ex + (uint32_t)n, 31) - wex0;
    uint4 *s_meta = (uint4*)pk_smem + wb;
    s_meta[lane] = make_uint4((uint32_t)ex - wex0, (uint32_t)(fbi + 1), (uint32_t)(2 * i0 + 127), a.rid_base + rd);
    __syncwarp();
    uint32_t *const ok = a.out_key + (s_base + wex0); uint64_t *const oy = a.out_y + (s_base + wex0);
    #pragma unroll 1
    for (uint32_t o = lane; o < wtot; o += 32) {
        uint32_t t = 0;
        #pragma unroll
        for (uint32_t st = 16; st > 0; st >>= 1) if (s_meta[t + st].x <= o) t += st;
        const uint4 m = s_meta[t];
        const uint32_t j = o - m.x;
        uint32_t h, p;
        if (m.y == 0) { const key kk = s_stage[(wb + t) * PK_STRIDE + j]; h = T::hash(kk); p = m.z - T::code(kk); }
        else { const uint2 r = s_fb[(m.y - 1) * PK_CAP + j]; h = r.x; p = r.y; }
        ok[o] = h; oy[o] = (uint64_t)m.w << 32 | p;
    }
}

/* ------------------------------------------------------------------ HPC sketch: one thread per read (spike-in run, reference sketch.c:93-104) */

struct SkWriteSpan {
    uint32_t *key; uint64_t *key64; uint64_t *yy; uint8_t *span; uint64_t at;
    __device__ __forceinline__ void operator()(uint64_t x_, uint64_t y_)
    {
        if (key64) key64[at] = x_ >> 8; else key[at] = (uint32_t)(x_ >> 8);
        yy[at] = y_; if (span) span[at] = (uint8_t)x_; ++at;
    }
};

template <int WRITE>
__global__ void lq_sketch_seq_k(SkArgs a, uint32_t n_reads, int is_hpc, uint32_t *read_count, const uint64_t *read_base, uint8_t *out_span, uint64_t *out_key64)
{
    const uint32_t rd = blockIdx.x * blockDim.x + threadIdx.x;
    if (rd >= n_reads) return;
    const int L = (int)a.len[rd];
    const uint64_t g0 = a.slot0[rd] * LQ_SLOT;
    if (!WRITE) {
        SkCount c; c.n = 0;
        if (L > 0) lq_sketch_replay(a.b2, a.nm, g0, L, a.w, a.k, a.rid_base + rd, is_hpc, 0, 1, L - 1, 0, L, (int*)0, c);
        read_count[rd] = (uint32_t)c.n;
    } else {
        SkWriteSpan wr; wr.key = a.out_key; wr.key64 = out_key64; wr.yy = a.out_y; wr.span = out_span; wr.at = read_base[rd];
        if (L > 0) lq_sketch_replay(a.b2, a.nm, g0, L, a.w, a.k, a.rid_base + rd, is_hpc, 0, 1, L - 1, 0, L, (int*)0, wr);
    }
}

/* per-read record ranges from the y stream (rid in the high word): first[r] = first record with rid >= r */
__global__ void lq_read_first_k(const uint64_t *__restrict__ y, uint64_t n, uint32_t rid_base, uint32_t n_reads, uint64_t *__restrict__ first)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const int64_t prev = i == 0 ? -1 : (int64_t)((uint32_t)(y[i - 1] >> 32) - rid_base);
    const int64_t cur = i == n ? (int64_t)n_reads : (int64_t)((uint32_t)(y[i] >> 32) - rid_base);
    for (int64_t r = prev + 1; r <= cur; ++r) first[r] = i;
}

/* ------------------------------------------------------------------ host side */

/* test switch: force the tiled position-parallel kernel even where the rolling kernel applies (LQCOV_SKETCH_TILED=1) */
static int g_sketch_tiled = getenv("LQCOV_SKETCH_TILED") ? atoi(getenv("LQCOV_SKETCH_TILED")) : 0;
/* LQCOV_SKETCH_PK: 1 (default) packed-key kernel fed by bulk asynchronous copies; 2 the same fed by plain loads; 0 the rolling kernel */
static int g_sketch_pk = getenv("LQCOV_SKETCH_PK") ? atoi(getenv("LQCOV_SKETCH_PK")) : 1;
/* test switch: 0 the defaults; 1 tiled kernel; 2 packed-key kernel fed by plain loads; 3 rolling kernel where the packed-key one is the default */
extern "C" void lqcov_debug_sketch_tiled(int on) { g_sketch_tiled = on == 1; g_sketch_pk = on == 2 ? 2 : on == 3 ? 0 : 1; }
/* the 64-bases-per-thread kernels share the tile geometry (RK_TILE bases per CTA) */
static void sk_launch_seg64(const SkArgs &a, unsigned nblk, cudaStream_t st)
{
    const int bulk = g_sketch_pk != 2;
    /* 7 / 8 / 9 CTAs per SM (72 / 64 / 56 registers) were measured: 3.71 / 3.64 / 3.76 ms -- occupancy is not what limits the kernel */
    if (a.w == 5 && a.k == 12 && g_sketch_pk) lq_sketch_pk_k<5, 12, PK_MIN_CTAS><<<nblk, RK_THREADS, PK_SMEM(4), st>>>(a, bulk);          /* LongQC's overlap runs */
    else if (a.w == 5 && a.k == 15 && g_sketch_pk) {                                                                                   /* --fast */
        cudaFuncSetAttribute(lq_sketch_pk_k<5, 15, PK_MIN_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM(8));   /* per device: every launch, the executable drives several GPUs from one process */
        lq_sketch_pk_k<5, 15, PK_MIN_CTAS><<<nblk, RK_THREADS, PK_SMEM(8), st>>>(a, bulk);
    }
    else if (a.w == 5) lq_sketch_roll_k<5><<<nblk, RK_THREADS, 0, st>>>(a);
    else lq_sketch_roll_k<10><<<nblk, RK_THREADS, 0, st>>>(a);
}

int lq_reads_upload(LqReadsDev *d, const uint8_t *h_seq, const uint64_t *h_off, uint32_t n_reads, int seq_on_device, int sdust_tbl, cudaStream_t st)
{
    /* h_off: n_reads+1 offsets into h_seq (host array always); h_seq host or device (seq_on_device) */
    d->n_reads = n_reads;
    d->n_bases = h_off[n_reads] - h_off[0];
    d->h_len.resize(n_reads); d->h_slot0.resize((size_t)n_reads + 1);
    uint64_t slots = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        uint64_t L = h_off[r + 1] - h_off[r];
        if (L > 0x7fffffffULL) { fprintf(stderr, "[lqcov] read %u longer than 2^31\n", r); return -1; }
        d->h_len[r] = (uint32_t)L; d->h_slot0[r] = slots;
        slots += (L + LQ_SLOT - 1) / LQ_SLOT;
    }
    d->h_slot0[n_reads] = slots;
    d->n_slots = slots;
    LQ_TRY(d->len.ensure((size_t)(n_reads + 1) * 4));
    LQ_TRY(d->slot0.ensure((size_t)(n_reads + 1) * 8));
    LQ_TRY(d->off.ensure((size_t)(n_reads + 1) * 8));
    LQ_TRY(d->b2.ensure((size_t)(slots * LQ_SLOT_W2 + 16) * 4));
    LQ_TRY(d->nm.ensure((size_t)(slots * LQ_SLOT_WN + 16) * 4));
    LQ_TRY(d->slot_read.ensure((size_t)(slots + 1) * 4));
    if (n_reads == 0) return 0;
    LQ_CUDA_OK(cudaMemcpyAsync(d->len.p, d->h_len.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
    LQ_CUDA_OK(cudaMemcpyAsync(d->slot0.p, d->h_slot0.data(), (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    LQ_CUDA_OK(cudaMemcpyAsync(d->off.p, h_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    const uint8_t *d_seq;
    if (seq_on_device) d_seq = h_seq;
    else {
        LQ_TRY(d->ascii.ensure((size_t)d->n_bases + 16));
        LQ_CUDA_OK(cudaMemcpyAsync(d->ascii.p, h_seq + h_off[0], (size_t)d->n_bases, cudaMemcpyHostToDevice, st));
        d_seq = d->ascii.as<uint8_t>() - h_off[0];
    }
    /* pad words past the last slot (the 3-word k-mer gather may read one word past a read's last slot) */
    LQ_CUDA_OK(cudaMemsetAsync(d->b2.as<uint32_t>() + slots * LQ_SLOT_W2, 0, 16 * 4, st));
    LQ_CUDA_OK(cudaMemsetAsync(d->nm.as<uint32_t>() + slots * LQ_SLOT_WN, 0xff, 16 * 4, st));
    lq_prof_h2d((uint64_t)(n_reads + 1) * 20 + (seq_on_device ? 0 : d->n_bases));
    if (slots) {
        LqProfScope ps("pack", st, 1, d->n_bases + slots * (LQ_SLOT_W2 + LQ_SLOT_WN + 1) * 4);
        lq_pack_k<<<lq_grid(slots * 4, 256), 256, 0, st>>>(d_seq, d->off.as<uint64_t>(), d->slot0.as<uint64_t>(), d->len.as<uint32_t>(),
                                                           n_reads, slots, sdust_tbl, d->b2.as<uint32_t>(), d->nm.as<uint32_t>(), d->slot_read.as<uint32_t>(),
                                                           d_seq + h_off[0], d_seq + h_off[n_reads], 0);
        LQ_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int lq_sketch_run(const LqReadsDev *rd, int w, int k, int is_hpc, uint32_t rid_base, LqMinimizers *out, LqDevBuf &ws, cudaStream_t st);

/* ------------------------------------------------------------------ upload + pack + sketch, pipelined (host buffers)
 *
 * A part's bases come over PCIe (~15 ms for 0.8 GB) while nothing else runs.  Here the reads are cut into chunks of ~48 MB: chunk c+1
 * is copied on a second stream while chunk c is packed and sketched (rolling kernel, launched over the chunk's slots; its records
 * follow the running total the previous launch left in device memory, so the output is the same contiguous, ordered array).
 * Returns 1 (nothing done) where the rolling kernel does not apply or the input is small: the caller takes the plain sequence. */
static int g_pipeline = getenv("LQCOV_NO_PIPELINE") ? 0 : 1;
#define PL_CHUNK (48ull << 20)

int lq_upload_sketch_pipelined(LqReadsDev *d, const uint8_t *h_seq, const uint64_t *h_off, uint32_t n_reads, int w, int k, int is_hpc, uint32_t rid_base,
                               LqMinimizers *out, LqDevBuf &ws, cudaStream_t st)
{
    const bool roll = !is_hpc && (w == 5 || w == 10) && k <= 15 && !g_sketch_tiled;
    if (!g_pipeline || !roll || n_reads == 0 || h_off[n_reads] - h_off[0] < 2 * PL_CHUNK) return 1;
    static cudaStream_t st_copy = 0;
    if (!st_copy) LQ_CUDA_OK(cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking));
    /* layout, as lq_reads_upload() */
    d->n_reads = n_reads; d->n_bases = h_off[n_reads] - h_off[0];
    d->h_len.resize(n_reads); d->h_slot0.resize((size_t)n_reads + 1);
    uint64_t slots = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        const uint64_t L = h_off[r + 1] - h_off[r];
        if (L > 0x7fffffffULL) { fprintf(stderr, "[lqcov] read %u longer than 2^31\n", r); return -1; }
        d->h_len[r] = (uint32_t)L; d->h_slot0[r] = slots;
        slots += (L + LQ_SLOT - 1) / LQ_SLOT;
    }
    d->h_slot0[n_reads] = slots; d->n_slots = slots;
    LQ_TRY(d->len.ensure((size_t)(n_reads + 1) * 4)); LQ_TRY(d->slot0.ensure((size_t)(n_reads + 1) * 8)); LQ_TRY(d->off.ensure((size_t)(n_reads + 1) * 8));
    LQ_TRY(d->b2.ensure((size_t)(slots * LQ_SLOT_W2 + 16) * 4)); LQ_TRY(d->nm.ensure((size_t)(slots * LQ_SLOT_WN + 16) * 4));
    LQ_TRY(d->slot_read.ensure((size_t)(slots + 1) * 4)); LQ_TRY(d->ascii.ensure((size_t)d->n_bases + 16));
    LQ_CUDA_OK(cudaMemcpyAsync(d->len.p, d->h_len.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
    LQ_CUDA_OK(cudaMemcpyAsync(d->slot0.p, d->h_slot0.data(), (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    LQ_CUDA_OK(cudaMemcpyAsync(d->off.p, h_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    LQ_CUDA_OK(cudaMemsetAsync(d->b2.as<uint32_t>() + slots * LQ_SLOT_W2, 0, 16 * 4, st));
    LQ_CUDA_OK(cudaMemsetAsync(d->nm.as<uint32_t>() + slots * LQ_SLOT_WN, 0xff, 16 * 4, st));
    lq_prof_h2d((uint64_t)(n_reads + 1) * 20 + d->n_bases);
    const uint8_t *d_seq = d->ascii.as<uint8_t>() - h_off[0];
    /* chunks = read ranges of >= PL_CHUNK bases */
    std::vector<uint32_t> cut; cut.push_back(0);
    for (uint32_t r = 0; r < n_reads; ) {
        uint64_t sz = 0;
        while (r < n_reads && sz < PL_CHUNK) { sz += d->h_len[r]; ++r; }
        cut.push_back(r);
    }
    const size_t nc = cut.size() - 1;
    /* output + per-chunk look-back state + running totals */
    out->n = 0; out->has_span = 0;
    const uint64_t cap = (uint64_t)((double)d->n_bases * 2.6 / (w + 1)) + 4096;
    LQ_TRY(out->key.ensure((size_t)(cap + 1) * 4)); LQ_TRY(out->y.ensure((size_t)(cap + 1) * 8));
    std::vector<size_t> soff(nc + 1, 0);
    for (size_t c = 0; c < nc; ++c) {
        const uint64_t ns = d->h_slot0[cut[c + 1]] - d->h_slot0[cut[c]];
        soff[c + 1] = soff[c] + (size_t)((ns * LQ_SLOT + RK_TILE - 1) / RK_TILE) + 2;       /* tiles + {ticket, err} */
    }
    LQ_TRY(out->blk.ensure((soff[nc] + nc + 4) * 8 + 64));
    unsigned long long *state = out->blk.as<unsigned long long>(), *totals = state + soff[nc];
    LQ_CUDA_OK(cudaMemsetAsync(state, 0, (soff[nc] + nc + 4) * 8, st));
    std::vector<cudaEvent_t> ev(nc);
    cudaEvent_t ev0;                                       /* the copies must not start before this call's place in the stream */
    LQ_CUDA_OK(cudaEventCreateWithFlags(&ev0, cudaEventDisableTiming));
    LQ_CUDA_OK(cudaEventRecord(ev0, st));
    LQ_CUDA_OK(cudaStreamWaitEvent(st_copy, ev0, 0));
    for (size_t c = 0; c < nc; ++c) {
        LQ_CUDA_OK(cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming));
        const uint64_t b0 = h_off[cut[c]], b1 = h_off[cut[c + 1]];
        LQ_CUDA_OK(cudaMemcpyAsync(d->ascii.as<uint8_t>() + (b0 - h_off[0]), h_seq + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, st_copy));
        LQ_CUDA_OK(cudaEventRecord(ev[c], st_copy));
    }
    SkArgs a;
    a.b2 = d->b2.as<uint32_t>(); a.nm = d->nm.as<uint32_t>(); a.slot_read = d->slot_read.as<uint32_t>(); a.len = d->len.as<uint32_t>();
    a.slot0 = d->slot0.as<uint64_t>(); a.w = w; a.k = k; a.rid_base = rid_base; a.cap = cap;
    a.out_key = out->key.as<uint32_t>(); a.out_y = out->y.as<uint64_t>();
    for (size_t c = 0; c < nc; ++c) {
        const uint64_t s0 = d->h_slot0[cut[c]], s1 = d->h_slot0[cut[c + 1]], bases = h_off[cut[c + 1]] - h_off[cut[c]];
        LQ_CUDA_OK(cudaStreamWaitEvent(st, ev[c], 0));
        if (s1 == s0) { LQ_CUDA_OK(cudaMemcpyAsync(totals + c + 1, totals + c, 8, cudaMemcpyDeviceToDevice, st)); continue; }
        { LqProfScope ps("pack", st, 1, bases + (s1 - s0) * (LQ_SLOT_W2 + LQ_SLOT_WN + 1) * 4);
          lq_pack_k<<<lq_grid((s1 - s0) * 4, 256), 256, 0, st>>>(d_seq, d->off.as<uint64_t>(), d->slot0.as<uint64_t>(), d->len.as<uint32_t>(), n_reads, s1, 0,
                                                               d->b2.as<uint32_t>(), d->nm.as<uint32_t>(), d->slot_read.as<uint32_t>(), d_seq + h_off[0], d_seq + h_off[n_reads], s0); }
        const unsigned nblk = (unsigned)(((s1 - s0) * LQ_SLOT + RK_TILE - 1) / RK_TILE);
        a.n_slots = s1; a.g_begin = s0 * LQ_SLOT;
        a.state = state + soff[c]; a.ticket = (uint32_t*)(a.state + nblk); a.err = a.ticket + 1;
        a.base_in = totals + c; a.base_out = totals + c + 1;
        { LqProfScope ps("sketch", st, 1, (s1 - s0) * (LQ_SLOT_W2 + LQ_SLOT_WN) * 4 + (uint64_t)((double)bases * 2.0 / (w + 1)) * 12);
          sk_launch_seg64(a, nblk, st); }
        LQ_CUDA_OK(cudaGetLastError());
    }
    unsigned long long total = 0;
    std::vector<unsigned long long> h_state(soff[nc]);
    LQ_CUDA_OK(cudaMemcpyAsync(&total, totals + nc, 8, cudaMemcpyDeviceToHost, st));
    LQ_CUDA_OK(cudaMemcpyAsync(h_state.data(), state, soff[nc] * 8, cudaMemcpyDeviceToHost, st)); lq_prof_d2h(8 + soff[nc] * 8);
    LQ_CUDA_OK(cudaStreamSynchronize(st));
    for (size_t c = 0; c < nc; ++c) cudaEventDestroy(ev[c]);
    cudaEventDestroy(ev0);
    for (size_t c = 0; c < nc; ++c) {                      /* the err word of every chunk */
        const uint64_t ns = d->h_slot0[cut[c + 1]] - d->h_slot0[cut[c]];
        const size_t nblk = (size_t)((ns * LQ_SLOT + RK_TILE - 1) / RK_TILE);
        if (((const uint32_t*)(h_state.data() + soff[c] + nblk))[1]) { fprintf(stderr, "[lqcov] sketch: look-back or bulk copy timed out\n"); return -1; }
    }
    if (total > cap) return lq_sketch_run(d, w, k, is_hpc, rid_base, out, ws, st);   /* low-complexity input: the bases are packed, sketch again with room */
    out->n = total;
    return 0;
}

int lq_sketch_run(const LqReadsDev *rd, int w, int k, int is_hpc, uint32_t rid_base, LqMinimizers *out, LqDevBuf &ws, cudaStream_t st)
{
    out->n = 0; out->has_span = is_hpc;
    if (w < 1 || w > LQ_MAX_W) { fprintf(stderr, "[lqcov] window size %d not supported on the GPU path (1..%d)\n", w, LQ_MAX_W); return -1; }
    if (k < 1 || k > LQ_MAX_K) { fprintf(stderr, "[lqcov] k-mer size %d outside 1..%d\n", k, LQ_MAX_K); return -1; }
    if (rd->n_reads == 0 || rd->n_slots == 0) return 0;
    SkArgs a;
    a.b2 = rd->b2.as<uint32_t>(); a.nm = rd->nm.as<uint32_t>(); a.slot_read = rd->slot_read.as<uint32_t>(); a.len = rd->len.as<uint32_t>();
    a.slot0 = rd->slot0.as<uint64_t>(); a.n_slots = rd->n_slots; a.w = w; a.k = k; a.rid_base = rid_base;
    a.ticket = 0; a.state = 0; a.cap = 0; a.err = 0; a.out_key = 0; a.out_y = 0; a.g_begin = 0; a.base_in = 0; a.base_out = 0;
    uint64_t total = 0;
    out->wide = k > LQ_MAX_K_DIRECT;   /* 2k-bit hashes in key64; one thread per read, as in HPC mode (the general state machine) */
    if (!is_hpc && !out->wide) {
        const bool roll = (w == 5 || w == 10) && k <= 15 && !g_sketch_tiled;
        const unsigned nblk = roll ? (unsigned)((rd->n_slots * LQ_SLOT + RK_TILE - 1) / RK_TILE) : (unsigned)((rd->n_slots * LQ_SLOT + SK_TILE - 1) / SK_TILE);
        LQ_TRY(out->blk.ensure((size_t)(nblk + 2) * 8 + 64));
        unsigned long long *state = out->blk.as<unsigned long long>();
        uint32_t *ticket = (uint32_t*)(state + nblk + 1), *err = ticket + 1;
        /* expected density 2/(w+1) records per base; the exact total is known after the pass, which is repeated only
         * if the guess was too small (low-complexity input) */
        uint64_t cap = (uint64_t)((double)rd->n_bases * 2.6 / (w + 1)) + 4096;
        for (int attempt = 0; attempt < 2; ++attempt) {
            LQ_TRY(out->key.ensure((size_t)(cap + 1) * 4));
            LQ_TRY(out->y.ensure((size_t)(cap + 1) * 8));
            LQ_CUDA_OK(cudaMemsetAsync(state, 0, (size_t)(nblk + 2) * 8 + 16, st));
            a.ticket = ticket; a.state = state; a.cap = cap; a.err = err;
            a.out_key = out->key.as<uint32_t>(); a.out_y = out->y.as<uint64_t>();
            {
                LqProfScope ps("sketch", st, 1, rd->n_slots * (LQ_SLOT_W2 + LQ_SLOT_WN) * 4 + (uint64_t)((double)rd->n_bases * 2.0 / (w + 1)) * 12);
                if (roll) sk_launch_seg64(a, nblk, st);
                else if (w == 5) lq_sketch_k<5, 5><<<nblk, SK_THREADS, 0, st>>>(a);
                else if (w == 10) lq_sketch_k<10, 10><<<nblk, SK_THREADS, 0, st>>>(a);
                else lq_sketch_k<LQ_MAX_W, 0><<<nblk, SK_THREADS, 0, st>>>(a);
            }
            LQ_CUDA_OK(cudaGetLastError());
            unsigned long long last = 0; uint32_t h_err = 0;
            LQ_CUDA_OK(cudaMemcpyAsync(&last, state + (nblk - 1), 8, cudaMemcpyDeviceToHost, st));
            LQ_CUDA_OK(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, st)); lq_prof_d2h(12);
            LQ_CUDA_OK(cudaStreamSynchronize(st));
            if (h_err) { fprintf(stderr, "[lqcov] sketch: look-back or bulk copy timed out\n"); return -1; }
            total = last & ((1ULL << 62) - 1);
            if (total <= cap) break;
            if (attempt == 1) { fprintf(stderr, "[lqcov] sketch: output overflow after resize\n"); return -1; }
            cap = total;
        }
    } else {
        const uint32_t n = rd->n_reads;
        LQ_TRY(out->blk.ensure((size_t)(n + 1) * 4 + (size_t)(n + 2) * 8));
        uint32_t *cnt = out->blk.as<uint32_t>();
        uint64_t *base = (uint64_t*)((char*)out->blk.p + (((size_t)(n + 1) * 4 + 7) & ~(size_t)7));
        lq_prof_count_launch(2);
        lq_sketch_seq_k<0><<<lq_grid(n, 64), 64, 0, st>>>(a, n, is_hpc, cnt, 0, 0, 0);
        LQ_CUDA_OK(cudaGetLastError());
        LQ_TRY((lq_exclusive_scan<uint32_t, uint64_t>(cnt, base, n, 1, ws, st)));
        LQ_CUDA_OK(cudaMemcpyAsync(&total, base + n, 8, cudaMemcpyDeviceToHost, st));
        LQ_CUDA_OK(cudaStreamSynchronize(st));
        LQ_TRY(out->key.ensure((size_t)(total + 1) * 4));
        LQ_TRY(out->y.ensure((size_t)(total + 1) * 8));
        if (is_hpc) LQ_TRY(out->span.ensure((size_t)(total + 1)));
        if (k > LQ_MAX_K_DIRECT) LQ_TRY(out->key64.ensure((size_t)(total + 1) * 8));
        a.out_key = out->key.as<uint32_t>(); a.out_y = out->y.as<uint64_t>();
        lq_sketch_seq_k<1><<<lq_grid(n, 64), 64, 0, st>>>(a, n, is_hpc, 0, base, is_hpc ? out->span.as<uint8_t>() : 0, k > LQ_MAX_K_DIRECT ? out->key64.as<uint64_t>() : 0);
        LQ_CUDA_OK(cudaGetLastError());
    }
    out->n = total; out->has_span = is_hpc;
    return 0;
}

int lq_read_first(const LqMinimizers *m, uint32_t rid_base, uint32_t n_reads, LqDevBuf &first, cudaStream_t st)
{
    LQ_TRY(first.ensure((size_t)(n_reads + 2) * 8));
    lq_prof_count_launch(1);
    lq_read_first_k<<<lq_grid(m->n + 1, 256), 256, 0, st>>>(m->y.as<uint64_t>(), m->n, rid_base, n_reads, first.as<uint64_t>());
    LQ_CUDA_OK(cudaGetLastError());
    return 0;
}

/* ------------------------------------------------------------------ a part arriving in chunks (host staging buffers -> records)
 *
 * The streaming form of lq_upload_sketch_pipelined(): the chunks are not cut from one big host blob but arrive one by one from the
 * reader (lq_ingest.c fills pinned staging buffers while the GPU works on the chunk before), and the part's size is not known in
 * advance -- the packed read store, the per-read arrays and the record arrays grow on demand (LqDevBuf::ensure_keep).  The ASCII
 * bases only pass through a ring of two chunk-sized device buffers. */
void LqPartStream::release()
{
    for (int i = 0; i < LQ_STREAM_RING; ++i) asc[i].release();
    state.release(); totals.release();
    if (ev_made) for (int i = 0; i < LQ_STREAM_RING; ++i) { cudaEventDestroy(ev_copied[i]); cudaEventDestroy(ev_packed[i]); }
    ev_made = false;
    if (st_copy) { cudaStreamDestroy(st_copy); st_copy = 0; }
    for (int i = 0; i < LQ_STREAM_RING; ++i) { if (h_meta[i]) cudaFreeHost(h_meta[i]); h_meta[i] = 0; h_meta_cap[i] = 0; }
}

bool lq_stream_ok(int w, int k, int is_hpc) { return !is_hpc && (w == 5 || w == 10) && k <= 15 && !g_sketch_tiled; }

#define ST_MAX_CHUNKS 65536
int lq_stream_begin(LqPartStream *s, LqReadsDev *d, LqMinimizers *out, int w, int k, uint32_t rid_base, uint64_t expect_bases, cudaStream_t st)
{
    s->d = d; s->out = out; s->w = w; s->k = k; s->rid_base = rid_base; s->st = st;
    if (!s->st_copy) LQ_CUDA_OK(cudaStreamCreateWithFlags(&s->st_copy, cudaStreamNonBlocking));
    if (!s->ev_made) {
        for (int i = 0; i < LQ_STREAM_RING; ++i) {
            LQ_CUDA_OK(cudaEventCreateWithFlags(&s->ev_copied[i], cudaEventDisableTiming));
            LQ_CUDA_OK(cudaEventCreateWithFlags(&s->ev_packed[i], cudaEventDisableTiming));
        }
        s->ev_made = true;
    }
    d->n_reads = 0; d->n_bases = 0; d->n_slots = 0; d->h_len.clear(); d->h_slot0.clear(); d->h_slot0.push_back(0);
    out->n = 0; out->has_span = 0;
    s->state_used = 0; s->n_chunks = 0; s->chunk_state_off.clear(); s->chunk_tiles.clear();
    /* first sizes from the caller's expectation; everything grows when the part turns out larger */
    const uint64_t eb = expect_bases ? expect_bases : (256ull << 20), er = eb / 2000 + 1024, es = eb / LQ_SLOT + er;
    LQ_TRY(d->len.ensure((size_t)(er + 1) * 4)); LQ_TRY(d->slot0.ensure((size_t)(er + 1) * 8)); LQ_TRY(d->off.ensure((size_t)(er + 1) * 8));
    LQ_TRY(d->b2.ensure((size_t)(es * LQ_SLOT_W2 + 16) * 4)); LQ_TRY(d->nm.ensure((size_t)(es * LQ_SLOT_WN + 16) * 4)); LQ_TRY(d->slot_read.ensure((size_t)(es + 1) * 4));
    s->cap_rec = (uint64_t)((double)eb * 2.6 / (w + 1)) + 4096;
    LQ_TRY(out->key.ensure((size_t)(s->cap_rec + 1) * 4)); LQ_TRY(out->y.ensure((size_t)(s->cap_rec + 1) * 8));
    LQ_TRY(s->state.ensure((size_t)(es * LQ_SLOT / RK_TILE + 4096) * 8));
    LQ_TRY(s->totals.ensure((size_t)(ST_MAX_CHUNKS + 2) * 8));
    LQ_CUDA_OK(cudaMemsetAsync(s->totals.p, 0, 8, st));
    s->open = true;
    return 0;
}

int lq_stream_push(LqPartStream *s, const uint8_t *h_seq, const uint64_t *h_off, uint32_t n, cudaEvent_t *copied)
{
    LqReadsDev *d = s->d; LqMinimizers *out = s->out; cudaStream_t st = s->st;
    if (copied) *copied = 0;
    if (n == 0) return 0;
    if (s->n_chunks >= ST_MAX_CHUNKS) { fprintf(stderr, "[lqcov] too many chunks in one index part\n"); return -1; }
    const uint32_t r0 = d->n_reads; const uint64_t s0 = d->n_slots, bases = h_off[n] - h_off[0];
    const int ring = (int)(s->n_chunks % LQ_STREAM_RING);
    /* layout of the new reads; their per-read arrays travel through a pinned block of the ring slot (it was last used two chunks
     * ago: the copy that read it is ordered before ev_packed of that chunk, which the host waits for here) */
    uint64_t slots = s0;
    const size_t meta_bytes = (size_t)n * 20 + 64;
    if (s->n_chunks >= LQ_STREAM_RING) LQ_CUDA_OK(cudaEventSynchronize(s->ev_packed[ring]));
    if (meta_bytes > s->h_meta_cap[ring]) {
        if (s->h_meta[ring]) cudaFreeHost(s->h_meta[ring]);
        s->h_meta[ring] = 0; s->h_meta_cap[ring] = 0;
        LQ_CUDA_OK(cudaHostAlloc(&s->h_meta[ring], meta_bytes * 2, cudaHostAllocDefault));
        s->h_meta_cap[ring] = meta_bytes * 2;
    }
    uint64_t *aoff = (uint64_t*)s->h_meta[ring], *pslot0 = aoff + n; uint32_t *plen = (uint32_t*)(pslot0 + n + 1);   /* offsets inside the ring buffer | slot0 | len */
    d->h_len.resize((size_t)r0 + n); d->h_slot0.resize((size_t)r0 + n + 1);
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t L = h_off[i + 1] - h_off[i];
        if (L > 0x7fffffffULL) { fprintf(stderr, "[lqcov] read longer than 2^31\n"); return -1; }
        d->h_len[r0 + i] = (uint32_t)L; d->h_slot0[r0 + i] = slots; aoff[i] = h_off[i] - h_off[0]; pslot0[i] = slots; plen[i] = (uint32_t)L;
        slots += (L + LQ_SLOT - 1) / LQ_SLOT;
    }
    pslot0[n] = slots;
    d->h_slot0[(size_t)r0 + n] = slots;
    const uint32_t r1 = r0 + n; const uint64_t s1 = slots;
    /* room (growth keeps what earlier chunks left; it waits for the stream, which is fine for a rare event) */
    LQ_TRY(d->len.ensure_keep((size_t)(r1 + 1) * 4, (size_t)r0 * 4, st)); LQ_TRY(d->slot0.ensure_keep((size_t)(r1 + 1) * 8, (size_t)(r0 + 1) * 8, st));
    LQ_TRY(d->off.ensure_keep((size_t)(r1 + 1) * 8, (size_t)r0 * 8, st));
    LQ_TRY(d->b2.ensure_keep((size_t)(s1 * LQ_SLOT_W2 + 16) * 4, (size_t)s0 * LQ_SLOT_W2 * 4, st)); LQ_TRY(d->nm.ensure_keep((size_t)(s1 * LQ_SLOT_WN + 16) * 4, (size_t)s0 * LQ_SLOT_WN * 4, st));
    LQ_TRY(d->slot_read.ensure_keep((size_t)(s1 + 1) * 4, (size_t)s0 * 4, st));
    const uint64_t need_rec = (uint64_t)((double)(d->n_bases + bases) * 2.6 / (s->w + 1)) + 4096;
    if (need_rec > s->cap_rec) {
        /* the records written so far: the running total is on the device; keep the whole old capacity */
        LQ_TRY(out->key.ensure_keep((size_t)(need_rec * 3 / 2 + 1) * 4, (size_t)s->cap_rec * 4, st)); LQ_TRY(out->y.ensure_keep((size_t)(need_rec * 3 / 2 + 1) * 8, (size_t)s->cap_rec * 8, st));
        s->cap_rec = need_rec * 3 / 2;
    }
    const unsigned nblk = (unsigned)(((s1 - s0) * LQ_SLOT + RK_TILE - 1) / RK_TILE);
    LQ_TRY(s->state.ensure_keep((s->state_used + nblk + 4) * 8, s->state_used * 8, st));
    /* the ring buffer must have been packed (two chunks ago) before it is overwritten */
    LQ_TRY(s->asc[ring].ensure((size_t)bases + 64));   /* ensure() frees only when too small: chunks have one size in practice */
    if (s->n_chunks >= LQ_STREAM_RING) LQ_CUDA_OK(cudaStreamWaitEvent(s->st_copy, s->ev_packed[ring], 0));
    else { cudaEvent_t e0; LQ_CUDA_OK(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming)); LQ_CUDA_OK(cudaEventRecord(e0, st)); LQ_CUDA_OK(cudaStreamWaitEvent(s->st_copy, e0, 0)); cudaEventDestroy(e0); }
    LQ_CUDA_OK(cudaMemcpyAsync(s->asc[ring].p, h_seq + h_off[0], (size_t)bases, cudaMemcpyHostToDevice, s->st_copy));
    LQ_CUDA_OK(cudaEventRecord(s->ev_copied[ring], s->st_copy));
    if (copied) *copied = s->ev_copied[ring];
    LQ_CUDA_OK(cudaMemcpyAsync(d->len.as<uint32_t>() + r0, plen, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    LQ_CUDA_OK(cudaMemcpyAsync(d->slot0.as<uint64_t>() + r0, pslot0, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    LQ_CUDA_OK(cudaMemcpyAsync(d->off.as<uint64_t>() + r0, aoff, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    lq_prof_h2d((uint64_t)n * 20 + bases);
    LQ_CUDA_OK(cudaMemsetAsync(d->b2.as<uint32_t>() + s1 * LQ_SLOT_W2, 0, 16 * 4, st));
    LQ_CUDA_OK(cudaMemsetAsync(d->nm.as<uint32_t>() + s1 * LQ_SLOT_WN, 0xff, 16 * 4, st));
    LQ_CUDA_OK(cudaStreamWaitEvent(st, s->ev_copied[ring], 0));
    const uint8_t *d_seq = s->asc[ring].as<uint8_t>();
    if (s1 > s0) {
        { LqProfScope ps("pack", st, 1, bases + (s1 - s0) * (LQ_SLOT_W2 + LQ_SLOT_WN + 1) * 4);
          /* lq_pack_k finds a slot's read by binary search over slot0[0..n_reads): the reads of this chunk are [r0, r1) */
          lq_pack_k<<<lq_grid((s1 - s0) * 4, 256), 256, 0, st>>>(d_seq, d->off.as<uint64_t>(), d->slot0.as<uint64_t>(), d->len.as<uint32_t>(), r1, s1, 0,
                                                               d->b2.as<uint32_t>(), d->nm.as<uint32_t>(), d->slot_read.as<uint32_t>(), d_seq, d_seq + bases, s0); }
        LQ_CUDA_OK(cudaEventRecord(s->ev_packed[ring], st));
        unsigned long long *state = s->state.as<unsigned long long>() + s->state_used;
        LQ_CUDA_OK(cudaMemsetAsync(state, 0, (size_t)(nblk + 2) * 8, st));
        SkArgs a;
        a.b2 = d->b2.as<uint32_t>(); a.nm = d->nm.as<uint32_t>(); a.slot_read = d->slot_read.as<uint32_t>(); a.len = d->len.as<uint32_t>();
        a.slot0 = d->slot0.as<uint64_t>(); a.w = s->w; a.k = s->k; a.rid_base = s->rid_base; a.cap = s->cap_rec;
        a.out_key = out->key.as<uint32_t>(); a.out_y = out->y.as<uint64_t>();
        a.n_slots = s1; a.g_begin = s0 * LQ_SLOT;
        a.state = state; a.ticket = (uint32_t*)(state + nblk); a.err = a.ticket + 1;
        a.base_in = s->totals.as<unsigned long long>() + s->n_chunks; a.base_out = s->totals.as<unsigned long long>() + s->n_chunks + 1;
        { LqProfScope ps("sketch", st, 1, (s1 - s0) * (LQ_SLOT_W2 + LQ_SLOT_WN) * 4 + (uint64_t)((double)bases * 2.0 / (s->w + 1)) * 12);
          sk_launch_seg64(a, nblk, st); }
        LQ_CUDA_OK(cudaGetLastError());
        s->chunk_state_off.push_back(s->state_used); s->chunk_tiles.push_back(nblk);
        s->state_used += nblk + 2;
    } else {
        LQ_CUDA_OK(cudaEventRecord(s->ev_packed[ring], st));
        LQ_CUDA_OK(cudaMemcpyAsync(s->totals.as<unsigned long long>() + s->n_chunks + 1, s->totals.as<unsigned long long>() + s->n_chunks, 8, cudaMemcpyDeviceToDevice, st));
        s->chunk_state_off.push_back(s->state_used); s->chunk_tiles.push_back(0);
    }
    d->n_reads = r1; d->n_slots = s1; d->n_bases += bases;
    ++s->n_chunks;
    return 0;
}

int lq_stream_end(LqPartStream *s, LqDevBuf &ws)
{
    LqReadsDev *d = s->d; LqMinimizers *out = s->out; cudaStream_t st = s->st;
    s->open = false;
    if (s->n_chunks == 0) { out->n = 0; LQ_CUDA_OK(cudaStreamSynchronize(st)); return 0; }
    unsigned long long total = 0;
    std::vector<unsigned long long> h_state(s->state_used);
    LQ_CUDA_OK(cudaMemcpyAsync(&total, s->totals.as<unsigned long long>() + s->n_chunks, 8, cudaMemcpyDeviceToHost, st));
    if (s->state_used) LQ_CUDA_OK(cudaMemcpyAsync(h_state.data(), s->state.p, s->state_used * 8, cudaMemcpyDeviceToHost, st));
    lq_prof_d2h(8 + s->state_used * 8);
    LQ_CUDA_OK(cudaStreamSynchronize(st));
    for (size_t c = 0; c < s->chunk_tiles.size(); ++c)
        if (s->chunk_tiles[c] && ((const uint32_t*)(h_state.data() + s->chunk_state_off[c] + s->chunk_tiles[c]))[1]) { fprintf(stderr, "[lqcov] sketch: look-back or bulk copy timed out\n"); return -1; }
    if (total > s->cap_rec) return lq_sketch_run(d, s->w, s->k, 0, s->rid_base, out, ws, st);   /* low-complexity input: the bases are packed, sketch again with room */
    out->n = total;
    return 0;
}
