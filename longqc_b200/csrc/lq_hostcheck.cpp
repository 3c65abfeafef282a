// lq_hostcheck.cpp -- TEST-ONLY host build of the host/device core headers.
// Lets the CPU-only test tier (-m "not gpu") execute the exact functions the CUDA kernels are
// built from (position-parallel sketch, sort-order walk, chain chunk resolution) against the
// oracle.  Built as liblqcov_hostcheck.so; NEVER linked into liblqcov.so and never a fallback:
// the product path has no CPU implementation.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "lq_common.h"
#include "lq_sketch_core.h"

namespace {
struct VecSink {
    std::vector<lq_mm128> *v;
    void operator()(uint64_t x, uint64_t y) { lq_mm128 e; e.x = x; e.y = y; v->push_back(e); }
};
void pack_read(const char *seq, int len, int sdust_tbl, std::vector<uint32_t> &b2, std::vector<uint32_t> &nm)
{
    int nslot = (len + LQ_SLOT - 1) / LQ_SLOT;
    if (nslot == 0) nslot = 1;
    b2.assign((size_t)nslot * LQ_SLOT_W2 + 4, 0u);  // +4: the 3-word k-mer gather may touch one word past the end
    nm.assign((size_t)nslot * LQ_SLOT_WN + 2, 0xffffffffu);
    for (int i = 0; i < len; ++i) {
        uint32_t c = lq_nt4((unsigned char)seq[i], sdust_tbl);
        if (c < 4) { b2[i >> 4] |= c << ((i & 15) * 2); nm[i >> 5] &= ~(1u << (i & 31)); }
    }
}
}

extern "C" {

// position-parallel sketch of one read (every position evaluated independently, then concatenated)
int lqhc_sketch_parallel(const char *seq, int len, int w, int k, uint32_t rid, lq_mm128 *out, int cap)
{
    std::vector<uint32_t> b2, nm; std::vector<lq_mm128> v; VecSink s; s.v = &v;
    pack_read(seq, len, 0, b2, nm);
    for (int i = 0; i < len; ++i) lq_sketch_at(b2.data(), nm.data(), 0, len, w, k, rid, i, s);
    int n = (int)v.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
    return n;
}

// the windowed fast path (closed form around palindromes / ambiguous bases) at every position
int lqhc_sketch_parallel_win(const char *seq, int len, int w, int k, uint32_t rid, lq_mm128 *out, int cap)
{
    std::vector<uint32_t> b2, nm; std::vector<lq_mm128> v; VecSink s; s.v = &v;
    pack_read(seq, len, 0, b2, nm);
    for (int i = 0; i < len; ++i) lq_sketch_at_win(b2.data(), nm.data(), 0, len, w, k, rid, i, s);
    int n = (int)v.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
    return n;
}

// the restartable state machine run from the start of the read (HPC allowed)
int lqhc_sketch_replay(const char *seq, int len, int w, int k, uint32_t rid, int is_hpc, lq_mm128 *out, int cap)
{
    std::vector<uint32_t> b2, nm; std::vector<lq_mm128> v; VecSink s; s.v = &v;
    pack_read(seq, len, 0, b2, nm);
    lq_sketch_replay(b2.data(), nm.data(), 0, len, w, k, rid, is_hpc, 0, 1, len - 1, 0, len, (int*)0, s);
    int n = (int)v.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
    return n;
}

// only the slow path at every position (stress of the bounded restart)
int lqhc_sketch_slow_everywhere(const char *seq, int len, int w, int k, uint32_t rid, lq_mm128 *out, int cap)
{
    std::vector<uint32_t> b2, nm; std::vector<lq_mm128> v; VecSink s; s.v = &v;
    pack_read(seq, len, 0, b2, nm);
    for (int i = 0; i < len; ++i) lq_sketch_slow_at(b2.data(), nm.data(), 0, len, w, k, rid, i, s);
    int n = (int)v.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
    return n;
}

uint32_t lqhc_hash32(uint32_t key, uint32_t mask) { return lq_hash32(key, mask); }
uint64_t lqhc_hash64(uint64_t key, uint64_t mask) { return lq_hash64(key, mask); }

} // extern "C"



// ---------------------------------------------------------------- seed sort (lq_afsort_core.h)
#include "lq_afsort_core.h"
namespace {
struct Bkt { uint32_t beg, end; };
static int cnt_above_15(const uint32_t *cnt) { int n = 0; for (int d = 16; d < 256; ++d) n += cnt[d] != 0; return n; }
// same orchestration as the device: level by level, buckets > 64 keep going, 2..64 get the stable insertion sort
void afsort_levels(const uint64_t *key, uint32_t n, uint32_t *idx, int use_two)
{
    for (uint32_t i = 0; i < n; ++i) idx[i] = i;
    if (n <= LQ_RS_MIN) { lq_af_insertion(idx, n, key); return; }
    std::vector<Bkt> cur, nxt; Bkt b0 = {0, n}; cur.push_back(b0);
    std::vector<uint8_t> dig(n); std::vector<uint32_t> dest(n), tmp(n), P, Z, rk(n); std::vector<uint8_t> fr(n);
    for (int s = 56; s >= 0 && !cur.empty(); s -= 8) {
        nxt.clear();
        for (size_t bi = 0; bi < cur.size(); ++bi) {
            const uint32_t beg = cur[bi].beg, m = cur[bi].end - cur[bi].beg;
            uint32_t cnt[256], start[256], head[256], nb = 0, acc = 0;
            memset(cnt, 0, sizeof(cnt));
            for (uint32_t p = 0; p < m; ++p) { dig[beg + p] = (uint8_t)(key[idx[beg + p]] >> s); ++cnt[dig[beg + p]]; }
            for (int d = 0; d < 256; ++d) { start[d] = acc; acc += cnt[d]; nb += cnt[d] != 0; }
            if (nb > 1) {
                if (nb == 2 && (use_two & 1)) {
                    int d0 = -1, d1 = -1;
                    for (int d = 0; d < 256; ++d) if (cnt[d]) { if (d0 < 0) d0 = d; else d1 = d; }
                    const uint32_t n0 = cnt[d0];
                    P.clear(); Z.clear();
                    for (uint32_t p = 0; p < m; ++p) {
                        int f = p < n0 ? dig[beg + p] == d1 : dig[beg + p] == d0;
                        fr[p] = (uint8_t)f;
                        if (p < n0) { rk[p] = (uint32_t)P.size(); if (f) P.push_back(p); }
                        else { rk[p] = (uint32_t)Z.size(); if (f) Z.push_back(p); }
                    }
                    for (uint32_t p = 0; p < m; ++p)
                        dest[beg + p] = lq_af_two_dest(p, n0, fr[p], rk[p], (uint32_t)P.size(), P.data(), Z.data());
                } else if ((use_two & 16) && cnt_above_15(cnt) == 0) { // few regions: cached stretches + byte offsets in registers, digit stream out
                    uint32_t st257[257]; std::vector<uint32_t> cache(LQ_AFR_R * LQ_AFR_BLK, 0u); lq_afr_walk ws; lq_afq_phase ph[256];
                    for (int d = 0; d < 256; ++d) { st257[d] = start[d]; ph[d].t = 0xffffffffu; ph[d].p = 0; }
                    st257[256] = m;
                    lq_afr_cache_init_host(cache.data(), st257);
                    lq_afr_init(&ws, st257, ph);
                    std::vector<uint32_t> seq32(m / 4 + 2), ord(m), slot(m); uint32_t run[256];
                    for (;;) {
                        lq_afr_refill_host(dig.data() + beg, st257, cache.data());
                        { lq_afr_host_words hw; hw.a = cache.data(); if (lq_afr_run(&ws, m, st257, hw, 1, seq32.data(), ph)) break; }
                    }
                    lq_afq_expand((const uint8_t*)seq32.data(), m, st257, ph, run, ord.data(), slot.data());
                    for (uint32_t t = 0; t < m; ++t) dest[beg + ord[t]] = slot[t];
                } else if (use_two & 8) { // the digit-stream walk (strided packed states) + parallel-form expansion
                    const uint32_t stride = 3;
                    uint32_t st257[257]; std::vector<lq_afp_st> pst(256 * stride); lq_afq_walk ws; lq_afq_phase ph[256];
                    for (int d = 0; d < 256; ++d) { st257[d] = start[d]; pst[d * stride].x = start[d]; pst[d * stride].y = pst[d * stride].z = pst[d * stride].w = 0; ph[d].t = 0xffffffffu; ph[d].p = 0; }
                    st257[256] = m;
                    lq_afq_init(&ws, st257, ph);
                    std::vector<uint32_t> seq32(m / 4 + 2), ord(m), slot(m); uint32_t run[256];
                    for (;;) {
                        for (uint32_t r = 0; r < 256; ++r) { lq_afp_st tmp[256]; tmp[r] = pst[r * stride]; lq_afp_refill_host(dig.data() + beg, st257, tmp, r); pst[r * stride] = tmp[r]; }
                        if (lq_afq_run(&ws, m, st257, pst.data(), stride, seq32.data(), ph)) break;
                    }
                    lq_afq_expand((const uint8_t*)seq32.data(), m, st257, ph, run, ord.data(), slot.data());
                    for (uint32_t t = 0; t < m; ++t) dest[beg + ord[t]] = slot[t];
                } else if (use_two & 4) { // the packed one-load-per-step form (what the device runs, one lane per bucket)
                    uint32_t st257[257]; lq_afp_st pst[256]; lq_afp_walk ws;
                    for (int d = 0; d < 256; ++d) { st257[d] = start[d]; pst[d].x = start[d]; pst[d].y = pst[d].z = pst[d].w = 0; }
                    st257[256] = m;
                    lq_afp_init(&ws, st257);
                    std::vector<uint32_t> ord(m), slot(m);
                    for (;;) {
                        for (uint32_t r = 0; r < 256; ++r) lq_afp_refill_host(dig.data() + beg, st257, pst, r);
                        if (lq_afp_run(&ws, m, st257, pst, ord.data(), slot.data())) break;
                    }
                    for (uint32_t t = 0; t < m; ++t) dest[beg + ord[t]] = slot[t];
                } else if (use_two & 2) { // the device's form of the walk: packed per-region state, cached digits, refill rounds
                    uint32_t st257[257]; lq_afw_pb pb[256]; uint8_t cache[256 * LQ_AFW_CACHE]; lq_afw_state ws;
                    for (int d = 0; d < 256; ++d) { st257[d] = start[d]; pb[d].x = start[d]; pb[d].y = start[d] - LQ_AFW_CACHE; }
                    st257[256] = m;
                    lq_afw_init_state(&ws, st257);
                    for (;;) {
                        lq_afw_refill_host(dig.data() + beg, m, st257, pb, cache);
                        if (lq_afw_run(&ws, m, st257, pb, cache, dest.data() + beg)) break;
                    }
                } else lq_af_walk(dig.data() + beg, m, cnt, start, head, dest.data() + beg);
                for (uint32_t p = 0; p < m; ++p) tmp[beg + dest[beg + p]] = idx[beg + p];
                memcpy(idx + beg, tmp.data() + beg, (size_t)m * 4);
            }
            if (s > 0) {
                for (int d = 0; d < 256; ++d) {
                    if (cnt[d] > LQ_RS_MIN) { Bkt b = {beg + start[d], beg + start[d] + cnt[d]}; nxt.push_back(b); }
                    else if (cnt[d] > 1) lq_af_insertion(idx + beg + start[d], cnt[d], key);
                }
            }
        }
        cur.swap(nxt);
    }
}
}

extern "C" int lqhc_afsort(const uint64_t *key, uint32_t n, uint32_t *idx, int use_two) { afsort_levels(key, n, idx, use_two); return 0; }

// ---------------------------------------------------------------- chaining DP: chunked (warp-shaped) vs sequential
#include "lq_chain_core.h"
extern "C" {
// plain restatement of chain.c:41-80 on (rpos, qpos, span) of one (strand, target) group
void lqhc_chain_seq(const uint32_t *rpos, const int32_t *qpos, const uint8_t *span, int n, int max_dist, int bw, int max_skip, float avg_span,
                    int32_t *f, int32_t *p, int32_t *v)
{
    std::vector<int32_t> t(n, -1);
    int st = 0;
    for (int i = 0; i < n; ++i) {
        int32_t best = span[i], best_j = -1, n_skip = 0;
        while (st < i && (uint64_t)rpos[i] - rpos[st] > (uint64_t)max_dist) ++st;
        for (int j = i - 1; j >= st; --j) {
            int32_t sc;
            if (!lq_chain_gain((int64_t)rpos[i] - rpos[j], qpos[i] - qpos[j], span[i], max_dist, max_dist, bw, avg_span, &sc)) continue;
            sc += f[j];
            if (sc > best) { best = sc; best_j = j; if (n_skip > 0) --n_skip; }
            else if (t[j] == i) { if (++n_skip > max_skip) break; }
            if (p[j] >= 0) t[p[j]] = i;
        }
        f[i] = best; p[i] = best_j;
        v[i] = best_j >= 0 && v[best_j] > best ? v[best_j] : best;
    }
}
// the kernel's shape: 32 predecessors at a time, speculative stamping, prefix-max, event replay
void lqhc_chain_chunked(const uint32_t *rpos, const int32_t *qpos, const uint8_t *span, int n, int max_dist, int bw, int max_skip, float avg_span,
                        int32_t *f, int32_t *p, int32_t *v)
{
    std::vector<int32_t> t(n, -1);
    int st = 0;
    for (int i = 0; i < n; ++i) {
        int32_t best = span[i], best_j = -1, n_skip = 0;
        bool stop = false;
        while (st < i && (uint64_t)rpos[i] - rpos[st] > (uint64_t)max_dist) ++st;
        for (int jb = i - 1; jb >= st && !stop; jb -= 32) {
            bool valid[32]; int32_t sc[32]; bool tf[32];
            for (int l = 0; l < 32; ++l) {
                int j = jb - l; valid[l] = false; sc[l] = 0;
                if (j < st) continue;
                int32_t g;
                if (!lq_chain_gain((int64_t)rpos[i] - rpos[j], qpos[i] - qpos[j], span[i], max_dist, max_dist, bw, avg_span, &g)) continue;
                valid[l] = true; sc[l] = g + f[j];
            }
            for (int l = 0; l < 32; ++l) { int j = jb - l; if (valid[l] && p[j] >= 0) t[p[j]] = i; } // speculative, all lanes
            for (int l = 0; l < 32; ++l) { int j = jb - l; tf[l] = valid[l] && t[j] == i; }
            int32_t run = best; uint32_t newmask = 0, incmask = 0;
            for (int l = 0; l < 32; ++l) { // exclusive prefix max seeded with `best`
                if (valid[l] && sc[l] > run) { newmask |= 1u << l; run = sc[l]; }
                else if (tf[l]) incmask |= 1u << l;
            }
            uint32_t ev = newmask | incmask; int stop_lane = 32;
            while (ev) {
                int l = __builtin_ctz(ev); ev &= ev - 1;
                if (newmask >> l & 1) { if (n_skip > 0) --n_skip; }
                else if (++n_skip > max_skip) { stop_lane = l; break; }
            }
            uint32_t nm = stop_lane < 32 ? newmask & ((1u << stop_lane) - 1) : newmask;
            if (nm) { int l = 31 - __builtin_clz(nm); best = sc[l]; best_j = jb - l; }
            if (stop_lane < 32) stop = true;
        }
        f[i] = best; p[i] = best_j;
        v[i] = best_j >= 0 && v[best_j] > best ? v[best_j] : best;
    }
}
}

// ---------------------------------------------------------------- sdust core
#include "lq_sdust_core.h"
extern "C" long lqhc_sdust_masked(const char *seq, int len, int T, int W)
{
    int ov = 0;
    std::vector<int> pbuf(4 * (size_t)LQ_SD_PCAP(W));
    long m = (long)lq_sdust_masked((const uint8_t*)seq, len, T, W, pbuf.data(), LQ_SD_PCAP(W), &ov);
    return ov ? -1 : m;
}

// segment form: every segment of S steps scanned on its own after its warm-up, the merged runs of all segments folded in order (what
// lq_sdust_seg_k + lq_sdust_merge_k do); cap_runs small on purpose: -2 = a segment had more runs than fit (the device falls back)
extern "C" long lqhc_sdust_segments(const char *seq, int len, int T, int W, int S, int capP, int cap_runs)
{
    std::vector<int> pbuf(4 * (size_t)capP);
    std::vector<lq_sd_run> runs((size_t)cap_runs);
    lq_sd_merge_sink all; all.init();
    for (int lo = 0; lo <= len; lo += S) {
        int ov = 0;
        const bool last = lo + S > len;
        const int n = lq_sdust_segment((const uint8_t*)seq, len, T, W, lo, last ? 0x7fffffff : lo + S, pbuf.data(), capP, runs.data(), cap_runs, &ov);
        if (ov) return -1;
        if (n > cap_runs) return -2;
        for (int j = 0; j < n; ++j) all.add(runs[j].s, runs[j].f);
        if (last) break;
    }
    return (long)all.total();
}

// ---------------------------------------------------------------- host table code (lq_table.c is plain C: linked in for its self-test)
extern "C" double lqh_q2p(int q);
extern "C" double lqhc_q2p(int q) { return lqh_q2p(q); }

// ---------------------------------------------------------------- index dump (lq_mmi.cpp is plain C++: linked in for the format check)
#include <algorithm>
#include "lq_mmi.h"
// records (key, y) of one part in y order -> the index arrays of lq_index.cu built on the host -> the reference's file image
extern "C" int lqhc_mmi_dump(const char *path, int append, int w, int k, int is_hpc, const lqcov_reads_t *part, const uint32_t *key, const uint64_t *y, uint64_t n)
{
    const uint64_t nk = 1ULL << (2 * k);
    std::vector<uint32_t> counts(nk, 0); std::vector<uint64_t> offs(nk + 1, 0), pos(n + 1), fill;
    for (uint64_t i = 0; i < n; ++i) ++counts[key[i]];
    for (uint64_t q = 0; q < nk; ++q) offs[q + 1] = offs[q] + counts[q];
    fill.assign(offs.begin(), offs.end() - 1);
    for (uint64_t i = 0; i < n; ++i) pos[fill[key[i]]++] = y[i];          // stable: y ascending inside a key
    FILE *fp = fopen(path, append ? "ab" : "wb");
    if (!fp) return -1;
    const int rc = lq_mmi_dump_part(fp, w, k, is_hpc, part, counts.data(), offs.data(), pos.data());
    fclose(fp);
    return rc;
}
// an index file -> number of parts, records and sequences (checks the reader)
extern "C" long lqhc_mmi_load_count(const char *path, long *n_rec, long *n_seq)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return -1;
    LqMmiPart mp; long parts = 0; *n_rec = 0; *n_seq = 0; int r;
    while ((r = lq_mmi_load_part(fp, &mp)) == 1) { ++parts; *n_rec += (long)mp.key.size(); *n_seq += mp.n_seq; }
    fclose(fp);
    return r < 0 ? -1 : parts;
}

// the 64-bases-per-thread packed-key form (lq_sketch_pk_core.h): every segment it accepts comes from it, the others from the
// general state machine; *n_lean counts the segments it accepted.  The words before a read's first segment are garbage on purpose.
#include "lq_sketch_pk_core.h"
namespace {
template <int W, int K>
int sketch_pk(const char *seq, int len, uint32_t rid, lq_mm128 *out, int cap, int *n_lean)
{
    typedef lq_pk_tr<(K > 12)> T;
    std::vector<uint32_t> b2, nm; std::vector<lq_mm128> v; VecSink s; s.v = &v;
    pack_read(seq, len, 0, b2, nm);
    b2.resize(b2.size() + 8, 0u); nm.resize(nm.size() + 4, 0xffffffffu);
    int lean = 0;
    for (int i0 = 0; i0 < len; i0 += LQ_PK_SEG) {
        const int nseg = len - i0 < LQ_PK_SEG ? len - i0 : LQ_PK_SEG;
        const bool is_last = i0 + nseg == len;
        uint32_t lw8[8], nw4[4];
        for (int j = 0; j < 8; ++j) lw8[j] = i0 ? b2[(i0 >> 4) - 4 + j] : (j < 4 ? 0xdeadbeefu * (j + 1) : b2[j - 4]);
        for (int j = 0; j < 4; ++j) nw4[j] = i0 ? nm[(i0 >> 5) - 2 + j] : (j < 2 ? 0x5a5a5a5au : nm[j - 2]);
        struct KeySink {
            typedef size_t mark_t;
            std::vector<lq_mm128> *v; uint32_t rid; int i0;
            void push(typename T::key kk) { lq_mm128 e; e.x = (uint64_t)T::hash(kk) << 8 | (uint64_t)K; e.y = (uint64_t)rid << 32 | lq_pk_p2z(i0, T::code(kk)); v->push_back(e); }
            void put(typename T::key kk, bool yes) { if (yes) push(kk); }
            int room() const { return 1 << 20; }
            mark_t mark() const { return v->size(); }
            void rewind(mark_t m) { v->resize(m); }
        } ks; ks.v = &v; ks.rid = rid; ks.i0 = i0;
        const size_t mark = v.size();
        if (lq_pk_segment<W, K>(lw8, nw4, i0, nseg, is_last, ks) == 0) ++lean;
        else {
            v.resize(mark);
            lq_sketch_replay(b2.data(), nm.data(), 0, len, W, K, rid, 0, 0, 1, i0 + nseg - 1, i0, is_last ? len : i0 + nseg - 1, (int*)0, s);
        }
    }
    if (n_lean) *n_lean = lean;
    int n = (int)v.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
    return n;
}
}
extern "C" int lqhc_sketch_pk(const char *seq, int len, int w, int k, uint32_t rid, int *n_lean, lq_mm128 *out, int cap)
{
    switch (w * 100 + k) {
    case 512: return sketch_pk<5, 12>(seq, len, rid, out, cap, n_lean);
    case 515: return sketch_pk<5, 15>(seq, len, rid, out, cap, n_lean);
    case 511: return sketch_pk<5, 11>(seq, len, rid, out, cap, n_lean);
    case 1015: return sketch_pk<10, 15>(seq, len, rid, out, cap, n_lean);
    case 1012: return sketch_pk<10, 12>(seq, len, rid, out, cap, n_lean);
    case 308: return sketch_pk<3, 8>(seq, len, rid, out, cap, n_lean);
    case 204: return sketch_pk<2, 4>(seq, len, rid, out, cap, n_lean);
    case 713: return sketch_pk<7, 13>(seq, len, rid, out, cap, n_lean);
    }
    return -1;
}
