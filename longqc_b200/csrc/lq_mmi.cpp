/* lq_mmi.cpp -- the on-disk minimizer index of the reference ("MMI\2", index.c:390-479): written by `minimap2-coverage -d FILE`
 * and read back when the target argument is such a file (longQC.py --db: lines 266-277, 439-441).
 *
 * The file is a byte-for-byte image of the reference's in-memory index, part after part:
 *   "MMI\2" | w k b n_seq flag (u32 each) | per sequence: name length (u8), name, length (u32)
 *   per bucket i < 2^b (b = min(14, 2k); a minimizer lives in bucket key & (2^b - 1)):
 *       n (i32) | p[n] (u64: the positions of every minimizer that occurs more than once, minimizer after minimizer in ascending key
 *       order, ascending y inside one) | size (u32: distinct minimizers of the bucket) | size x { hash key, value } (u64 each) IN THE SLOT
 *       ORDER OF THE REFERENCE'S khash: key = (minimizer >> b) << 1 | singleton, value = the position itself (singleton) or
 *       start_in_p << 32 | count
 *   the sequences, 4 bits per base ((sum_len + 7) / 8 u32 words, mmpriv.h:27)
 * To write the same bytes the slot order of khash (klib khash.h: power-of-two table, triangular probing, 0.77 load bound, in-place
 * kick-out rehash on growth) is restated below (KhIdx) and driven exactly as worker_post() drives it (index.c:150-201): kh_resize to the
 * number of keys, then kh_put in ascending key order.  Reading needs none of that: the (key, positions) pairs go back into the
 * direct-address table of lq_index.cu.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>
#include "lq_common.h"
#include "lq_mmi.h"
#include "lq_sketch_core.h"

namespace {

/* khash.h, KHASH_INIT(idx, uint64_t, uint64_t, 1, idx_hash = key >> 1, idx_eq): only what an insert-only table needs */
struct KhIdx {
    uint32_t n_buckets, size, n_occupied, upper_bound;
    std::vector<uint32_t> flags;     /* 2 bits per slot: bit 1 = empty, bit 0 = deleted (khash.h:166-172) */
    std::vector<uint64_t> keys, vals;
    KhIdx() : n_buckets(0), size(0), n_occupied(0), upper_bound(0) {}
    static uint32_t fsize(uint32_t m) { return m < 16 ? 1 : m >> 4; }
    static bool isempty(const std::vector<uint32_t> &f, uint32_t i) { return (f[i >> 4] >> ((i & 0xfU) << 1)) & 2; }
    static bool iseither(const std::vector<uint32_t> &f, uint32_t i) { return (f[i >> 4] >> ((i & 0xfU) << 1)) & 3; }
    static void set_isdel_true(std::vector<uint32_t> &f, uint32_t i) { f[i >> 4] |= 1u << ((i & 0xfU) << 1); }
    static void set_isempty_false(std::vector<uint32_t> &f, uint32_t i) { f[i >> 4] &= ~(2u << ((i & 0xfU) << 1)); }
    static void set_isboth_false(std::vector<uint32_t> &f, uint32_t i) { f[i >> 4] &= ~(3u << ((i & 0xfU) << 1)); }
    static uint32_t hash(uint64_t key) { return (uint32_t)(key >> 1); }

    void resize(uint32_t new_n)      /* kh_resize (khash.h:233-293) */
    {
        --new_n; new_n |= new_n >> 1; new_n |= new_n >> 2; new_n |= new_n >> 4; new_n |= new_n >> 8; new_n |= new_n >> 16; ++new_n;
        if (new_n < 4) new_n = 4;
        if (size >= (uint32_t)(new_n * 0.77 + 0.5)) return;            /* requested size is too small */
        std::vector<uint32_t> nf(fsize(new_n), 0xaaaaaaaau);
        if (n_buckets < new_n) { keys.resize(new_n); vals.resize(new_n); }
        const uint32_t new_mask = new_n - 1;
        for (uint32_t j = 0; j != n_buckets; ++j) {
            if (iseither(flags, j)) continue;
            uint64_t key = keys[j], val = vals[j];
            set_isdel_true(flags, j);
            for (;;) {                                                 /* kick-out process */
                uint32_t i = hash(key) & new_mask, step = 0;
                while (!isempty(nf, i)) i = (i + (++step)) & new_mask;
                set_isempty_false(nf, i);
                if (i < n_buckets && !iseither(flags, i)) {            /* kick out the existing element */
                    std::swap(keys[i], key); std::swap(vals[i], val);
                    set_isdel_true(flags, i);
                } else { keys[i] = key; vals[i] = val; break; }
            }
        }
        if (n_buckets > new_n) { keys.resize(new_n); vals.resize(new_n); }
        flags.swap(nf);
        n_buckets = new_n; n_occupied = size; upper_bound = (uint32_t)(n_buckets * 0.77 + 0.5);
    }
    void put(uint64_t key, uint64_t val)   /* kh_put of a key known to be absent (khash.h:294-334) */
    {
        if (n_occupied >= upper_bound) { if (n_buckets > (size << 1)) resize(n_buckets - 1); else resize(n_buckets + 1); }
        const uint32_t mask = n_buckets - 1;
        uint32_t i = hash(key) & mask, step = 0;
        while (!isempty(flags, i)) i = (i + (++step)) & mask;          /* no deleted slots exist in an insert-only table */
        keys[i] = key; vals[i] = val;
        set_isboth_false(flags, i);
        ++size; ++n_occupied;
    }
};

bool wr(FILE *fp, const void *p, size_t n) { return n == 0 || fwrite(p, 1, n, fp) == n; }

}

/* one part: the index as lq_index.cu holds it (counts, offsets, positions sorted by key), names, lengths, the bases */
int lq_mmi_dump_part(FILE *fp, int w, int k, int is_hpc, const lqcov_reads_t *part, const uint32_t *counts, const uint64_t *offs, const uint64_t *pos)
{
    const int b = 2 * k < 14 ? 2 * k : 14;
    const uint32_t hdr[5] = { (uint32_t)w, (uint32_t)k, (uint32_t)b, part->n, is_hpc ? 1u : 0u };
    bool ok = wr(fp, "MMI\2", 4) && wr(fp, hdr, 20);
    uint64_t sum_len = 0;
    for (uint32_t i = 0; i < part->n && ok; ++i) {
        const size_t nl = (size_t)(part->name_off[i + 1] - part->name_off[i]);
        const uint8_t l = (uint8_t)nl;                                /* index.c:404: a name longer than 255 bytes is cut the same way */
        const uint32_t len = (uint32_t)(part->seq_off[i + 1] - part->seq_off[i]);
        ok = wr(fp, &l, 1) && wr(fp, part->names + part->name_off[i], l) && wr(fp, &len, 4);
        sum_len += len;
    }
    const uint64_t nkeys = 1ULL << (2 * k), per_bucket = nkeys >> b, mask = (1ULL << b) - 1;
    std::vector<uint64_t> p;
    for (uint64_t i = 0; i <= mask && ok; ++i) {
        uint32_t n_keys = 0; uint64_t n_p = 0;
        for (uint64_t hi = 0; hi < per_bucket; ++hi) { const uint32_t c = counts[hi << b | i]; if (c) { ++n_keys; if (c > 1) n_p += c; } }
        const int32_t bn = (int32_t)n_p;
        p.clear(); p.reserve((size_t)n_p);
        KhIdx h;
        if (n_keys) {
            h.resize(n_keys);                                         /* index.c:171 */
            for (uint64_t hi = 0; hi < per_bucket; ++hi) {            /* ascending minimizer == the sorted bucket of worker_post */
                const uint64_t key = hi << b | i; const uint32_t c = counts[key];
                if (c == 0) continue;
                if (c == 1) h.put(hi << 1 | 1, pos[offs[key]]);
                else {
                    const uint64_t start = p.size();
                    p.insert(p.end(), pos + offs[key], pos + offs[key] + c);   /* ascending y already (index.c:188) */
                    h.put(hi << 1, start << 32 | c);
                }
            }
        }
        ok = wr(fp, &bn, 4) && wr(fp, p.data(), p.size() * 8);
        const uint32_t size = n_keys;
        ok = ok && wr(fp, &size, 4);
        if (size == 0) continue;
        for (uint32_t s = 0; s < h.n_buckets && ok; ++s) {
            if (KhIdx::iseither(h.flags, s)) continue;
            const uint64_t x[2] = { h.keys[s], h.vals[s] };
            ok = wr(fp, x, 16);
        }
    }
    if (ok) {                                                         /* 4-bit packed bases: code 0..3, 4 for everything else (index.c:279-284) */
        std::vector<uint32_t> S((size_t)((sum_len + 7) / 8), 0u);
        uint64_t o = 0;
        for (uint32_t i = 0; i < part->n; ++i)
            for (uint64_t j = part->seq_off[i]; j < part->seq_off[i + 1]; ++j, ++o)
                S[o >> 3] |= lq_nt4((uint8_t)part->seq[j], 0) << ((o & 7) << 2);
        ok = wr(fp, S.data(), S.size() * 4);
    }
    fflush(fp);
    if (!ok) fprintf(stderr, "[lqcov] ERROR: writing the index dump failed\n");
    return ok ? 0 : -1;
}

/* 1 = the file starts with the index magic (index.c:481-498), 0 = not, -1 = cannot open */
int lq_mmi_is_index(const char *path)
{
    if (!path || strcmp(path, "-") == 0) return 0;
    FILE *fp = fopen(path, "rb");
    if (!fp) return -1;
    char m[4]; const int is = fread(m, 1, 4, fp) == 4 && memcmp(m, "MMI\2", 4) == 0;
    fclose(fp);
    return is;
}

/* next part of a dump: header values, names/lengths, and the (key, y) records of every minimizer with the positions of a key in
 * ascending y.  Returns 1, 0 at end of file, -1 on a damaged file. */
int lq_mmi_load_part(FILE *fp, LqMmiPart *out)
{
    char magic[4]; uint32_t x[5];
    if (fread(magic, 1, 4, fp) != 4) return 0;
    if (memcmp(magic, "MMI\2", 4) != 0 || fread(x, 4, 5, fp) != 5) return -1;
    out->w = (int)x[0]; out->k = (int)x[1]; out->b = (int)x[2]; out->n_seq = x[3]; out->flag = x[4];
    out->names.clear(); out->name_off.assign(1, 0); out->seq_off.assign(1, 0); out->key.clear(); out->y.clear();
    uint64_t sum_len = 0;
    for (uint32_t i = 0; i < out->n_seq; ++i) {
        uint8_t l; uint32_t len; char nm[256];
        if (fread(&l, 1, 1, fp) != 1 || (l && fread(nm, 1, l, fp) != l) || fread(&len, 4, 1, fp) != 1) return -1;
        out->names.append(nm, l); out->name_off.push_back(out->names.size());
        sum_len += len; out->seq_off.push_back(sum_len);
    }
    if (out->k < 1 || out->k > LQ_MAX_K_DIRECT) { fprintf(stderr, "[lqcov] ERROR: the index was built with k=%d; this build maps k <= %d\n", out->k, LQ_MAX_K_DIRECT); return -1; }
    std::vector<uint64_t> p;
    for (uint64_t i = 0; i < (1ULL << out->b); ++i) {
        int32_t n; uint32_t size;
        if (fread(&n, 4, 1, fp) != 1 || n < 0) return -1;
        p.resize((size_t)n);
        if (n && fread(p.data(), 8, (size_t)n, fp) != (size_t)n) return -1;
        if (fread(&size, 4, 1, fp) != 1) return -1;
        for (uint32_t j = 0; j < size; ++j) {
            uint64_t e[2];
            if (fread(e, 8, 2, fp) != 2) return -1;
            const uint64_t key = (e[0] >> 1) << out->b | i;
            if (e[0] & 1) { out->key.push_back((uint32_t)key); out->y.push_back(e[1]); }
            else {
                const uint64_t start = e[1] >> 32, c = (uint32_t)e[1];
                if (start + c > (uint64_t)n) return -1;
                for (uint64_t q = 0; q < c; ++q) { out->key.push_back((uint32_t)key); out->y.push_back(p[start + q]); }
            }
        }
    }
    if (!(out->flag & 2)) {                                           /* MM_I_NO_SEQ: the packed bases are not needed for mapping */
        const uint64_t words = (sum_len + 7) / 8;
        if (fseeko(fp, (off_t)(words * 4), SEEK_CUR) != 0) return -1;
    }
    return 1;
}
