/* lq_oracle.c -- plain-C CPU restatement of LongQC's minimap2-coverage + sdust.
 *
 * TEST INFRASTRUCTURE ONLY (see lq_oracle.h).  Parity pinned against the compiled
 * reference (oracle/_ref) by tests/test_oracle_vs_reference.py and tests/golden/.
 * Each function cites the reference file:line whose behaviour it restates; the code
 * is written from the behaviour, not transcribed.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <ctype.h>
#include <inttypes.h>
#include <zlib.h>
#include "lq_oracle.h"

#define U64MAX UINT64_MAX

/* ------------------------------------------------------------------ utils */

static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) { fprintf(stderr, "[oracle] out of memory\n"); abort(); } return p; }
static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s ? s : 1); if (!p) { fprintf(stderr, "[oracle] out of memory\n"); abort(); } return p; }
static void *xrealloc(void *q, size_t n) { void *p = realloc(q, n ? n : 1); if (!p) { fprintf(stderr, "[oracle] out of memory\n"); abort(); } return p; }

static void v128_push(lqo_mm128_v *v, lqo_mm128 e)
{
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 64; v->a = (lqo_mm128*)xrealloc(v->a, v->m * sizeof(lqo_mm128)); }
    v->a[v->n++] = e;
}

/* base -> 0..3, anything else 4; U/u count as T (sketch.c:8-25) */
static inline int nt4(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: break;
    }
    return c < 4 ? c : 4; /* bytes 0..3 map to themselves in the reference table */
}
/* sdust's private table does not know U (sdust.c:26-43) */
static inline int nt4_sdust(unsigned char c)
{
    if (c == 'U' || c == 'u') return 4;
    return nt4(c);
}

void lqo_opt_default(lqo_opt *o)
{
    memset(o, 0, sizeof(*o));
    o->k = 12; o->w = 5; o->is_hpc = 0;
    o->batch_size = 4000000000ULL; o->mini_batch_size = 50000000;
    o->no_self = 1; o->ava = 0;
    o->max_gap = 10000; o->min_cnt = 3; o->min_chain_score = 40;
    o->min_score_med = 40; o->min_score_good = 40; o->max_chain_skip = 25; o->bw = 500;
    o->mid_occ_frac = 2e-4f; o->seed = 11;
    o->max_overhang = 2000; o->min_ovlp = 1000; o->min_coverage = 3; o->min_ratio = 0.4; o->filter = 0;
}

/* ------------------------------------------------------------------ sketch */

/* sketch.c:27-37: invertible integer mix, truncated to 2k bits after every multiply step */
uint64_t lqo_hash64(uint64_t key, uint64_t mask)
{
    key = (~key + (key << 21)) & mask;
    key ^= key >> 24;
    key = (key + (key << 3) + (key << 8)) & mask;
    key ^= key >> 14;
    key = (key + (key << 2) + (key << 4)) & mask;
    key ^= key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

/* sketch.c:76-142.  A (w,k) minimizer scan kept as the reference's state machine:
 * ring of the last w candidates, current minimum (a copy) and the ring slot it sits in,
 * `run` = number of accepted k-mer steps since the last ambiguous base. */
void lqo_sketch(const char *str, int len, int w, int k, uint32_t rid, int is_hpc, lqo_mm128_v *p)
{
    const uint64_t mask = (1ULL << 2 * k) - 1;
    const int top = 2 * (k - 1);
    uint64_t fw = 0, rv = 0;
    lqo_mm128 ring[256], best = { U64MAX, U64MAX };
    int slot = 0, best_slot = 0, run = 0, span = 0, i, j;
    int hq[32], hq_front = 0, hq_n = 0; /* run lengths of the last k compressed bases (HPC) */

    for (j = 0; j < w; ++j) ring[j].x = ring[j].y = U64MAX;
    for (i = 0; i < len; ++i) {
        int c = nt4((unsigned char)str[i]);
        lqo_mm128 cand = { U64MAX, U64MAX };
        if (c < 4) {
            int strand;
            if (is_hpc) {
                int rl = 1;
                while (i + rl < len && nt4((unsigned char)str[i + rl]) == c) ++rl;
                i += rl - 1; /* i now indexes the last base of the homopolymer run */
                hq[(hq_front + hq_n++) & 31] = rl;
                span += rl;
                if (hq_n > k) { span -= hq[hq_front]; hq_front = (hq_front + 1) & 31; --hq_n; }
            } else span = run + 1 < k ? run + 1 : k;
            fw = (fw << 2 | (uint64_t)c) & mask;
            rv = rv >> 2 | (uint64_t)(3 ^ c) << top;
            if (fw == rv) continue; /* strand-symmetric k-mer: consumes no window slot (sketch.c:107) */
            strand = fw < rv ? 0 : 1;
            ++run;
            if (run >= k && span < 256) {
                cand.x = lqo_hash64(strand ? rv : fw, mask) << 8 | (uint64_t)span;
                cand.y = (uint64_t)rid << 32 | (uint32_t)i << 1 | (uint64_t)strand;
            }
        } else { run = 0; hq_n = hq_front = 0; span = 0; }
        ring[slot] = cand;
        if (run == w + k - 1 && best.x != U64MAX) { /* first full window: twins of the minimum (sketch.c:116-121) */
            for (j = slot + 1; j < w; ++j)
                if (ring[j].x == best.x && ring[j].y != best.y) v128_push(p, ring[j]);
            for (j = 0; j < slot; ++j)
                if (ring[j].x == best.x && ring[j].y != best.y) v128_push(p, ring[j]);
        }
        if (cand.x <= best.x) { /* new (rightmost) minimum; the old one leaves (sketch.c:122-124) */
            if (run >= w + k && best.x != U64MAX) v128_push(p, best);
            best = cand; best_slot = slot;
        } else if (slot == best_slot) { /* the minimum's slot was just overwritten (sketch.c:125-137) */
            if (run >= w + k - 1 && best.x != U64MAX) v128_push(p, best);
            best.x = U64MAX;
            for (j = slot + 1; j < w; ++j) if (best.x >= ring[j].x) { best = ring[j]; best_slot = j; }
            for (j = 0; j <= slot; ++j)    if (best.x >= ring[j].x) { best = ring[j]; best_slot = j; }
            if (run >= w + k - 1 && best.x != U64MAX) {
                for (j = slot + 1; j < w; ++j)
                    if (ring[j].x == best.x && ring[j].y != best.y) v128_push(p, ring[j]);
                for (j = 0; j <= slot; ++j)
                    if (ring[j].x == best.x && ring[j].y != best.y) v128_push(p, ring[j]);
            }
        }
        if (++slot == w) slot = 0;
    }
    if (best.x != U64MAX) v128_push(p, best); /* sketch.c:140-141 */
}

/* ------------------------------------------------------------------ ksort.h radix sort (unstable, order observable) */

/* ksort.h:84-134.  In-place MSD "American flag" sort, 8-bit digits, insertion sort for <=64 elements.
 * Generic over element size through macros so that the 128x / 64 / 32 instances share one statement. */
#define LQO_RS_MIN 64
#define LQO_DEFINE_RADIX(NAME, TYPE, KEY) \
static void ins_##NAME(TYPE *a, long n) \
{ \
    long i, j; \
    for (i = 1; i < n; ++i) \
        if (KEY(a[i]) < KEY(a[i-1])) { \
            TYPE t = a[i]; \
            for (j = i; j > 0 && KEY(t) < KEY(a[j-1]); --j) a[j] = a[j-1]; \
            a[j] = t; \
        } \
} \
static void af_##NAME(TYPE *a, long n, int s) \
{ \
    long head[256], tail[256], cnt[256], i; \
    int d; \
    memset(cnt, 0, sizeof(cnt)); \
    for (i = 0; i < n; ++i) ++cnt[(KEY(a[i]) >> s) & 255]; \
    for (d = 0, i = 0; d < 256; ++d) { head[d] = i; i += cnt[d]; tail[d] = i; } \
    for (d = 0; d < 256; ) { \
        if (head[d] == tail[d]) { ++d; continue; } \
        { \
            int e = (int)((KEY(a[head[d]]) >> s) & 255); \
            if (e == d) { ++head[d]; continue; } \
            { /* cycle: carry the displaced element to its bucket until one for bucket d turns up */ \
                TYPE carry = a[head[d]]; \
                do { \
                    TYPE t = a[head[e]]; a[head[e]++] = carry; carry = t; \
                    e = (int)((KEY(carry) >> s) & 255); \
                } while (e != d); \
                a[head[d]++] = carry; \
            } \
        } \
    } \
    if (s) { \
        int s2 = s > 8 ? s - 8 : 0; \
        for (d = 0; d < 256; ++d) { \
            long st = tail[d] - cnt[d]; \
            if (cnt[d] > LQO_RS_MIN) af_##NAME(a + st, cnt[d], s2); \
            else if (cnt[d] > 1) ins_##NAME(a + st, cnt[d]); \
        } \
    } \
} \
static void radix_##NAME(TYPE *beg, TYPE *end, int top_shift) \
{ \
    if (end - beg <= LQO_RS_MIN) ins_##NAME(beg, end - beg); \
    else af_##NAME(beg, end - beg, top_shift); \
}

#define KEY128(e) ((e).x)
#define KEYSELF(e) (e)
LQO_DEFINE_RADIX(k128, lqo_mm128, KEY128)
LQO_DEFINE_RADIX(k64, uint64_t, KEYSELF)
LQO_DEFINE_RADIX(k32, uint32_t, KEYSELF)

void lqo_radix_sort_128x(lqo_mm128 *beg, lqo_mm128 *end) { radix_k128(beg, end, 56); }
void lqo_radix_sort_64(uint64_t *beg, uint64_t *end) { radix_k64(beg, end, 56); }
void lqo_radix_sort_32(uint32_t *beg, uint32_t *end) { radix_k32(beg, end, 24); }

/* ------------------------------------------------------------------ index (one part) */

typedef struct {
    int k, w, is_hpc;
    uint32_t n_seq;
    char **name;       /* borrowed */
    uint32_t *len;
    /* CSR over distinct minimizer keys, keys ascending, positions ascending in y (index.c:150-201) */
    size_t n_keys, n_pos;
    uint64_t *key;     /* n_keys */
    uint64_t *off;     /* n_keys + 1 */
    uint64_t *pos;     /* n_pos: rid<<32 | lastPos<<1 | strand */
} lqo_index;

static void index_free(lqo_index *ix)
{
    free(ix->name); free(ix->len); free(ix->key); free(ix->off); free(ix->pos);
    memset(ix, 0, sizeof(*ix));
}

/* stable LSD radix on (x>>8) -- input is already ascending in y, so stability gives the order
 * index.c:188 obtains with radix_sort_64 on each key's positions */
static void sort_by_key_stable(lqo_mm128 *a, size_t n, int key_bits)
{
    lqo_mm128 *b = (lqo_mm128*)xmalloc(n * sizeof(lqo_mm128)), *src = a, *dst = b, *t;
    int sh;
    size_t i;
    for (sh = 0; sh < key_bits; sh += 16) {
        size_t *cnt = (size_t*)xcalloc(65537, sizeof(size_t));
        for (i = 0; i < n; ++i) ++cnt[(((src[i].x >> 8) >> sh) & 0xffff) + 1];
        for (i = 0; i < 65536; ++i) cnt[i + 1] += cnt[i];
        for (i = 0; i < n; ++i) dst[cnt[((src[i].x >> 8) >> sh) & 0xffff]++] = src[i];
        free(cnt);
        t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(lqo_mm128));
    free(b);
}

/* index.c:238-309 (sketch every target, rid = order in the part) + index.c:150-201 (group by key) */
static void index_build(lqo_index *ix, const lqo_opt *opt, char **name, char **seq, const int *len, uint32_t n_seq)
{
    lqo_mm128_v mv = { 0, 0, 0 };
    size_t i, nk;
    uint32_t r;
    memset(ix, 0, sizeof(*ix));
    ix->k = opt->k; ix->w = opt->w; ix->is_hpc = opt->is_hpc; ix->n_seq = n_seq;
    ix->name = (char**)xmalloc(n_seq * sizeof(char*));
    ix->len = (uint32_t*)xmalloc(n_seq * sizeof(uint32_t));
    for (r = 0; r < n_seq; ++r) {
        ix->name[r] = name[r]; ix->len[r] = (uint32_t)len[r];
        if (len[r] > 0) lqo_sketch(seq[r], len[r], opt->w, opt->k, r, opt->is_hpc, &mv); /* index.c:295-296 */
    }
    sort_by_key_stable(mv.a, mv.n, 2 * opt->k);
    for (i = 0, nk = 0; i < mv.n; ++i)
        if (i == 0 || mv.a[i].x >> 8 != mv.a[i-1].x >> 8) ++nk;
    ix->n_keys = nk; ix->n_pos = mv.n;
    ix->key = (uint64_t*)xmalloc(nk * 8);
    ix->off = (uint64_t*)xmalloc((nk + 1) * 8);
    ix->pos = (uint64_t*)xmalloc(mv.n * 8);
    for (i = 0, nk = 0; i < mv.n; ++i) {
        if (i == 0 || mv.a[i].x >> 8 != mv.a[i-1].x >> 8) { ix->key[nk] = mv.a[i].x >> 8; ix->off[nk] = i; ++nk; }
        ix->pos[i] = mv.a[i].y;
    }
    ix->off[nk] = mv.n;
    free(mv.a);
}

/* index.c:69-86 */
static const uint64_t *index_get(const lqo_index *ix, uint64_t key, int *n)
{
    size_t lo = 0, hi = ix->n_keys;
    *n = 0;
    while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (ix->key[mid] < key) lo = mid + 1; else hi = mid;
    }
    if (lo == ix->n_keys || ix->key[lo] != key) return 0;
    *n = (int)(ix->off[lo + 1] - ix->off[lo]);
    return ix->pos + ix->off[lo];
}

static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x < y ? -1 : x > y;
}

/* index.c:123-144: occurrence count at the (1-f) quantile of distinct minimizers, plus one */
static int32_t index_mid_occ(const lqo_index *ix, float f)
{
    uint32_t *c, thres;
    size_t i, n = ix->n_keys, kk;
    if (f <= 0.) return INT32_MAX;
    if (n == 0) return 1; /* reference reads uninitialised memory here; no seeds exist either way */
    c = (uint32_t*)xmalloc(n * 4);
    for (i = 0; i < n; ++i) c[i] = (uint32_t)(ix->off[i + 1] - ix->off[i]);
    qsort(c, n, 4, cmp_u32);
    kk = (uint32_t)((1. - f) * n);
    thres = c[kk < n ? kk : n - 1] + 1;
    free(c);
    return (int32_t)thres;
}

/* ------------------------------------------------------------------ seeds (lqmap.c:140-205) */

#define LQO_SEED_TANDEM (1ULL << 42)

static lqo_mm128 *collect_seeds(const lqo_opt *opt, int max_occ, const lqo_index *ix, const char *qname,
                                const lqo_mm128_v *mv, int qlen, int64_t *n_a, int *n_mini_pos, uint64_t **mini_pos)
{
    size_t i;
    int64_t cap = 0, na = 0;
    lqo_mm128 *a;
    *n_mini_pos = 0;
    *mini_pos = (uint64_t*)xmalloc(mv->n * 8);
    for (i = 0; i < mv->n; ++i) {
        int t; index_get(ix, mv->a[i].x >> 8, &t);
        if (t < max_occ) cap += t;
    }
    a = (lqo_mm128*)xmalloc((size_t)cap * sizeof(lqo_mm128));
    for (i = 0; i < mv->n; ++i) {
        const lqo_mm128 *p = &mv->a[i];
        int t, j, span = (int)(p->x & 0xff), tandem = 0;
        uint32_t qpos2 = (uint32_t)p->y; /* lastPos<<1 | strand */
        const uint64_t *r = index_get(ix, p->x >> 8, &t);
        if (t >= max_occ) continue; /* high-frequency minimizer (lqmap.c:166-173) */
        (*mini_pos)[(*n_mini_pos)++] = (uint64_t)span << 32 | qpos2 >> 1;
        if (i > 0 && p->x >> 8 == mv->a[i-1].x >> 8) tandem = 1;
        if (i + 1 < mv->n && p->x >> 8 == mv->a[i+1].x >> 8) tandem = 1;
        for (j = 0; j < t; ++j) {
            int32_t rpos = (int32_t)((uint32_t)r[j] >> 1);
            lqo_mm128 s;
            if (qname && (opt->no_self || opt->ava)) {
                int cmp = strcmp(qname, ix->name[r[j] >> 32]);
                if (opt->no_self && cmp == 0 && rpos == (int32_t)(qpos2 >> 1)) continue; /* the diagonal */
                if (opt->ava && cmp > 0) continue;
            }
            if ((r[j] & 1) == (qpos2 & 1)) { /* same strand */
                s.x = (r[j] & 0xffffffff00000000ULL) | (uint32_t)rpos;
                s.y = (uint64_t)span << 32 | qpos2 >> 1;
            } else {
                s.x = 1ULL << 63 | (r[j] & 0xffffffff00000000ULL) | (uint32_t)rpos;
                s.y = (uint64_t)span << 32 | (uint32_t)(qlen - ((int32_t)(qpos2 >> 1) + 1 - span) - 1);
            }
            if (tandem) s.y |= LQO_SEED_TANDEM;
            a[na++] = s;
        }
    }
    *n_a = na;
    return a;
}

/* ------------------------------------------------------------------ chaining (chain.c:22-157) */

static inline int ilog2_u32(uint32_t v) /* chain.c:15-20: floor(log2 v), v > 0 */
{
    int r = 0;
    while (v >>= 1) ++r;
    return r;
}

lqo_mm128 *lqo_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int min_cnt, int min_sc,
                        int64_t n, lqo_mm128 *a, int *n_u_, uint64_t **u_)
{
    int32_t *f, *p, *t, *v, n_u, n_v, k;
    int64_t i, j, st = 0;
    uint64_t *u, sum_span = 0;
    float avg_span;
    lqo_mm128 *b, *w;

    *u_ = 0; *n_u_ = 0;
    f = (int32_t*)xmalloc(n * 4); p = (int32_t*)xmalloc(n * 4);
    t = (int32_t*)xcalloc(n, 4);  v = (int32_t*)xmalloc(n * 4);
    for (i = 0; i < n; ++i) sum_span += a[i].y >> 32 & 0xff;
    avg_span = (float)sum_span / n; /* chain.c:38: float division */

    for (i = 0; i < n; ++i) { /* chain.c:41-80 */
        uint64_t ri = a[i].x;
        int32_t qi = (int32_t)a[i].y, span = (int32_t)(a[i].y >> 32 & 0xff);
        int32_t best = span, n_skip = 0;
        int64_t best_j = -1;
        while (st < i && ri - a[st].x > (uint64_t)max_dist_x) ++st;
        for (j = i - 1; j >= st; --j) {
            int64_t dr = (int64_t)(ri - a[j].x);
            int32_t dq = qi - (int32_t)a[j].y, dd, sc, lg;
            if (dr == 0 || dq <= 0) continue;
            if (dq > max_dist_y || dq > max_dist_x) continue;
            dd = dr > dq ? (int32_t)(dr - dq) : (int32_t)(dq - dr);
            if (dd > bw) continue;
            sc = dq < dr ? dq : (int32_t)dr;
            if (sc > span) sc = span;
            lg = dd ? ilog2_u32((uint32_t)dd) : 0;
            sc -= (int)(dd * .01 * avg_span) + (lg >> 1); /* double * double * (float->double), truncated */
            sc += f[j];
            if (sc > best) {
                best = sc; best_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[j] == i) {
                if (++n_skip > max_skip) break;
            }
            if (p[j] >= 0) t[p[j]] = (int32_t)i;
        }
        f[i] = best; p[i] = (int32_t)best_j;
        v[i] = best_j >= 0 && v[best_j] > best ? v[best_j] : best;
    }

    /* chain ends (chain.c:82-106) */
    memset(t, 0, n * 4);
    for (i = 0; i < n; ++i) if (p[i] >= 0) t[p[i]] = 1;
    for (i = n_u = 0; i < n; ++i) if (t[i] == 0 && v[i] >= min_sc) ++n_u;
    if (n_u == 0) { free(a); free(f); free(p); free(t); free(v); return 0; }
    u = (uint64_t*)xmalloc(n_u * 8);
    for (i = n_u = 0; i < n; ++i)
        if (t[i] == 0 && v[i] >= min_sc) {
            j = i;
            while (j >= 0 && f[j] < v[j]) j = p[j];
            if (j < 0) j = i;
            u[n_u++] = (uint64_t)f[j] << 32 | (uint64_t)j;
        }
    lqo_radix_sort_64(u, u + n_u);
    for (i = 0; i < n_u >> 1; ++i) { uint64_t x = u[i]; u[i] = u[n_u - i - 1]; u[n_u - i - 1] = x; }

    /* backtrack, best first (chain.c:108-125) */
    memset(t, 0, n * 4);
    for (i = n_v = k = 0; i < n_u; ++i) {
        int32_t n_v0 = n_v, k0 = k;
        j = (int32_t)u[i];
        do { v[n_v++] = (int32_t)j; t[j] = 1; j = p[j]; } while (j >= 0 && t[j] == 0);
        if (j < 0) {
            if (n_v - n_v0 >= min_cnt) u[k++] = u[i] >> 32 << 32 | (uint64_t)(n_v - n_v0);
        } else if ((int32_t)(u[i] >> 32) - f[j] >= min_sc) {
            if (n_v - n_v0 >= min_cnt) u[k++] = ((u[i] >> 32) - (uint64_t)f[j]) << 32 | (uint64_t)(n_v - n_v0);
        }
        if (k0 == k) n_v = n_v0;
    }
    *n_u_ = n_u = k; *u_ = u;
    free(f); free(p); free(t);

    /* anchors per chain in ascending order (chain.c:130-136) */
    b = (lqo_mm128*)xmalloc((size_t)n_v * sizeof(lqo_mm128));
    for (i = 0, k = 0; i < n_u; ++i) {
        int32_t k0 = k, ni = (int32_t)u[i];
        for (j = 0; j < ni; ++j) b[k++] = a[v[k0 + (ni - j - 1)]];
    }
    free(v);

    /* order chains by first anchor (chain.c:139-155); uses the unstable sort, so restated exactly */
    w = (lqo_mm128*)xmalloc((size_t)n_u * sizeof(lqo_mm128));
    for (i = k = 0; i < n_u; ++i) { w[i].x = b[k].x; w[i].y = (uint64_t)k << 32 | (uint64_t)i; k += (int32_t)u[i]; }
    lqo_radix_sort_128x(w, w + n_u);
    {
        uint64_t *u2 = (uint64_t*)xmalloc(n_u * 8);
        for (i = k = 0; i < n_u; ++i) {
            int32_t jj = (int32_t)w[i].y, nn = (int32_t)u[jj];
            u2[i] = u[jj];
            memcpy(&a[k], &b[w[i].y >> 32], nn * sizeof(lqo_mm128));
            k += nn;
        }
        memcpy(u, u2, n_u * 8);
        memcpy(b, a, k * sizeof(lqo_mm128));
        free(u2);
    }
    free(a); free(w);
    return b;
}

/* ------------------------------------------------------------------ chains -> regions (hit.c:23-88) */

typedef struct {
    int32_t cnt, rid, score0, qs, qe, rs, re, as, rev;
} lqo_reg;

static inline uint64_t mix64(uint64_t key) /* hit.c:40-50 */
{
    key = (~key + (key << 21));
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8));
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4));
    key = key ^ key >> 28;
    key = (key + (key << 31));
    return key;
}

static inline uint32_t wang32(uint32_t key) /* khash.h:400-409 */
{
    key += ~(key << 15); key ^= (key >> 10); key += (key << 3);
    key ^= (key >> 6);   key += ~(key << 11); key ^= (key >> 16);
    return key;
}

static uint32_t x31_str(const char *s) /* khash.h:383-388 */
{
    uint32_t h = (uint32_t)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
    return h;
}

static lqo_reg *gen_regs(uint32_t hash, int qlen, int n_u, const uint64_t *u, const lqo_mm128 *a)
{
    lqo_mm128 *z;
    lqo_reg *r;
    int i, k;
    if (n_u == 0) return 0;
    z = (lqo_mm128*)xmalloc(n_u * sizeof(lqo_mm128));
    for (i = k = 0; i < n_u; ++i) {
        uint32_t h = (uint32_t)mix64((mix64(a[k].x) + mix64(a[k].y)) ^ hash);
        z[i].x = u[i] ^ h;
        z[i].y = (uint64_t)k << 32 | (uint32_t)(int32_t)u[i];
        k += (int32_t)u[i];
    }
    lqo_radix_sort_128x(z, z + n_u);
    for (i = 0; i < n_u >> 1; ++i) { lqo_mm128 t = z[i]; z[i] = z[n_u-1-i]; z[n_u-1-i] = t; }
    r = (lqo_reg*)xcalloc(n_u, sizeof(lqo_reg));
    for (i = 0; i < n_u; ++i) { /* hit.c:23-38 */
        lqo_reg *ri = &r[i];
        int32_t s, span;
        ri->score0 = (int32_t)(z[i].x >> 32);
        ri->cnt = (int32_t)z[i].y;
        ri->as = (int32_t)(z[i].y >> 32);
        s = ri->as; span = (int32_t)(a[s].y >> 32 & 0xff);
        ri->rev = (int32_t)(a[s].x >> 63);
        ri->rid = (int32_t)(a[s].x << 1 >> 33);
        ri->rs = (int32_t)a[s].x + 1 > span ? (int32_t)a[s].x + 1 - span : 0;
        ri->re = (int32_t)a[s + ri->cnt - 1].x + 1;
        if (!ri->rev) {
            ri->qs = (int32_t)a[s].y + 1 - span;
            ri->qe = (int32_t)a[s + ri->cnt - 1].y + 1;
        } else {
            ri->qs = qlen - ((int32_t)a[s + ri->cnt - 1].y + 1);
            ri->qe = qlen - ((int32_t)a[s].y + 1 - span);
        }
    }
    free(z);
    return r;
}

/* ------------------------------------------------------------------ coverage accounting (esterr.c:17-140) */

typedef struct { uint32_t start, end; } lqo_sub;
typedef struct { size_t n, m; lqo_sub *a; } lqo_sub_v;
static void sub_push(lqo_sub_v *v, lqo_sub s)
{
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 16; v->a = (lqo_sub*)xrealloc(v->a, v->m * sizeof(lqo_sub)); }
    v->a[v->n++] = s;
}

static inline int32_t fwd_qpos(int32_t qlen, const lqo_mm128 *s) /* esterr.c:17-24 */
{
    int32_t x = (int32_t)s->y, span = (int32_t)(s->y >> 32 & 0xff);
    if (s->x >> 63) x = qlen - 1 - (x + 1 - span);
    return x;
}

static int find_mini(int qlen, const lqo_mm128 *s, int32_t n, const uint64_t *mini_pos) /* esterr.c:26-38 */
{
    int32_t x = fwd_qpos(qlen, s), lo = 0, hi = n - 1;
    while (lo <= hi) {
        int32_t mid = (int32_t)(((uint64_t)lo + hi) >> 1), y = (int32_t)mini_pos[mid];
        if (y < x) lo = mid + 1; else if (y > x) hi = mid - 1; else return mid;
    }
    return -1;
}

#define LQO_COVT 150 /* minimap2-coverage.h:20 */

static void cnt_match(const lqo_opt *opt, const lqo_index *ix, int qlen, int n_regs, const lqo_reg *regs, const lqo_mm128 *a,
                      int32_t n, const uint64_t *mini_pos, uint16_t *cnt, lqo_sub_v *cv, uint64_t *lambda, uint64_t *lambda2, float *avg_k)
{
    int i;
    uint16_t sc_med = (uint16_t)opt->min_score_med, sc_good = (uint16_t)opt->min_score_good; /* packed into 16 bits each, lqmap.c:841 */
    if (n == 0) return;
    if (*lambda / qlen > LQO_COVT && *avg_k != 0.0) return; /* esterr.c:87-88 */
    /* the frac<1 branch of esterr.c:89-91 needs lambda>COVT*qlen with avg_k still unset: unreachable */
    if (*avg_k == 0.0) {
        uint64_t s = 0;
        for (i = 0; i < n; ++i) s += mini_pos[i] >> 32 & 0xff;
        *avg_k = (float)s / n;
    }
    for (i = 0; i < n_regs; ++i) {
        const lqo_reg *r = &regs[i];
        int32_t st, j, k;
        uint32_t qs, qe, rs, re, rl, h5, h3, flag = 0;
        lqo_sub s;
        if (r->cnt == 0) continue;
        st = find_mini(qlen, r->rev ? &a[r->as + r->cnt - 1] : &a[r->as], n, mini_pos);
        if (st < 0) continue;
        rl = ix->len[r->rid];
        qs = r->qs; qe = r->qe; rs = r->rs; re = r->re;
        h5 = qs < rs ? qs : rs;
        h3 = (uint32_t)qlen - qe < rl - re ? (uint32_t)qlen - qe : rl - re;
        if ((qe - qs) < (qe - qs + h5 + h3) * opt->min_ratio || h5 > (uint32_t)opt->max_overhang || h3 > (uint32_t)opt->max_overhang)
            continue; /* esterr.c:118-119: u32 arithmetic, compared as double; signed max_overhang converts to unsigned */
        *lambda += (qe - qs + 1);
        if (r->score0 >= sc_med) flag |= 2;
        s.start = qs << 3 | flag; s.end = qe << 3 | flag | 1;
        sub_push(cv, s);
        if (r->score0 < sc_good) continue;
        *lambda2 += (qe - qs + 1);
        if (cnt[st] < UINT16_MAX) cnt[st]++;
        for (k = 1, j = st + 1; j < n && k < r->cnt; ++j) {
            int32_t x = fwd_qpos(qlen, r->rev ? &a[r->as + r->cnt - 1 - k] : &a[r->as + k]);
            if (x == (int32_t)mini_pos[j]) { ++k; if (cnt[st] < UINT16_MAX) cnt[j]++; } /* sic: tests cnt[st] (esterr.c:136) */
        }
    }
}

/* lqmap.c:25-100 */
static void filter_redundant(lqo_sub_v *v, lqo_sub_v *cv, uint32_t min_cov)
{
    size_t i, j, nvc;
    uint32_t *vc, med_start = 0, med_cov = 0;
    lqo_sub_v mc = { 0, 0, 0 };
    if (cv->n == 0) { free(cv->a); cv->a = 0; cv->m = 0; return; }
    nvc = cv->n * 2;
    vc = (uint32_t*)xmalloc(nvc * 4);
    for (i = 0; i < cv->n; ++i) { vc[2*i] = cv->a[i].start; vc[2*i+1] = cv->a[i].end; }
    lqo_radix_sort_32(vc, vc + nvc);
    for (j = 0; j < nvc; ++j) {
        uint32_t old = med_cov, e = vc[j];
        if (e & 2) {
            if (e & 1) med_cov -= (e & 4) ? min_cov : 1;
            else       med_cov += (e & 4) ? min_cov : 1;
        }
        if (old < min_cov && med_cov >= min_cov) med_start = e; /* kept encoded (lqmap.c:61) */
        else if (old >= min_cov && med_cov < min_cov) {
            uint32_t mlen = (e >> 3) - med_start;
            if (mlen > 0) {
                lqo_sub m, marker;
                m.start = med_start; m.end = e; sub_push(&mc, m);
                marker.start = med_start | 4; marker.end = e | 4; sub_push(v, marker);
            }
        }
    }
    free(vc);
    for (i = 0; i < cv->n; ++i) {
        int inside = 0;
        if (!(cv->a[i].start & 4))
            for (j = 0; j < mc.n; ++j)
                if (cv->a[i].start >= mc.a[j].start && cv->a[i].end <= mc.a[j].end) inside = 1;
        if (!inside) sub_push(v, cv->a[i]);
    }
    free(cv->a); cv->a = 0; cv->n = cv->m = 0;
    free(mc.a);
}

/* lqutils.c:83-155 */
static void reliable_region(const lqo_sub_v *v, uint32_t min_cov, lqo_sub_v *coords, lqo_sub_v *mcoords)
{
    size_t j, nvc = v->n * 2;
    uint32_t *vc = (uint32_t*)xmalloc(nvc * 4), start = 0, cov = 0, med_start = 0, med_cov = 0;
    for (j = 0; j < v->n; ++j) { vc[2*j] = v->a[j].start; vc[2*j+1] = v->a[j].end; }
    lqo_radix_sort_32(vc, vc + nvc);
    for (j = 0; j < nvc; ++j) {
        uint32_t e = vc[j], oc = cov, om = med_cov, pos = e >> 3;
        if (e & 1) {
            --cov;
            if (e & 2) { if (e & 4) { med_cov -= min_cov; cov -= (min_cov - 1); } else --med_cov; }
        } else {
            ++cov;
            if (e & 2) { if (e & 4) { med_cov += min_cov; cov += (min_cov - 1); } else ++med_cov; }
        }
        if (oc < min_cov && cov >= min_cov) {
            start = pos;
            if (om < min_cov && med_cov >= min_cov) med_start = pos;
        } else if (oc >= min_cov && cov < min_cov) {
            if (pos - start > 0) { lqo_sub c; c.start = start; c.end = pos; sub_push(coords, c); }
            if (om >= min_cov && med_cov < min_cov)
                if (pos - med_start > 0) { lqo_sub c; c.start = med_start; c.end = pos; sub_push(mcoords, c); }
        } else if (om < min_cov && med_cov >= min_cov) {
            med_start = pos;
        } else if (om >= min_cov && med_cov < min_cov) {
            if (pos - med_start > 0) { lqo_sub c; c.start = med_start; c.end = pos; sub_push(mcoords, c); }
        }
    }
    free(vc);
}

/* ------------------------------------------------------------------ quality (lqutils.c:26-80) */

static double q2p_tab[128];
static int q2p_ready = 0;
/* The reference hard-codes 127 literals with 15 decimals (lqutils.c:26-49).  They equal strtod("%.15f" % 10^(-q/10))
 * except for eight entries whose last digit is one higher (checked entry by entry against oracle/_ref in
 * tests/test_oracle_vs_reference.py::test_q2p_table_and_meanq). */
static void q2p_init(void)
{
    static const int up[8] = { 34, 39, 58, 62, 67, 71, 72, 82 };
    int q, i;
    char buf[64];
    if (q2p_ready) return;
    for (q = 0; q < 128; ++q) {
        int n = snprintf(buf, sizeof(buf), "%.15f", pow(10.0, -q / 10.0));
        for (i = 0; i < 8; ++i)
            if (up[i] == q) { int j = n - 1; while (j >= 0 && buf[j] == '9') buf[j--] = '0'; if (j >= 0 && buf[j] != '.') ++buf[j]; }
        q2p_tab[q] = strtod(buf, 0);
    }
    q2p_ready = 1;
}
double lqo_q2p(int q) { q2p_init(); return q2p_tab[q]; }

double lqo_meanQ(const char *qual, int len)
{
    int i; double sum = 0.0;
    q2p_init();
    for (i = 0; i < len; ++i) {
        int q = (int)qual[i] - 33;
        sum += q2p_tab[q < 0 ? 0 : q > 126 ? 126 : q]; /* out-of-table input is outside the parity domain */
    }
    return -10 * log10(sum / len);
}

int lqo_getQV(const char *qual, int threshold, int len)
{
    int i, n = 0, t = threshold + 33;
    for (i = 0; i < len; ++i) if ((int)qual[i] > t) ++n;
    return n;
}

/* ------------------------------------------------------------------ FASTA/FASTQ reader (kseq.h:185-224, bseq.c:56-66) */

typedef struct { gzFile fp; unsigned char *buf; int beg, end, eof; } lqo_stream;
#define LQO_BUFSZ 16384
static int st_getc(lqo_stream *s)
{
    if (s->eof && s->beg >= s->end) return -1;
    if (s->beg >= s->end) {
        s->beg = 0; s->end = gzread(s->fp, s->buf, LQO_BUFSZ);
        if (s->end < LQO_BUFSZ) s->eof = 1;
        if (s->end <= 0) { s->end = 0; return -1; }
    }
    return s->buf[s->beg++];
}
typedef struct { size_t l, m; char *s; } lqo_str;
static void str_putc(lqo_str *t, int c)
{
    if (t->l + 2 > t->m) { t->m = t->m ? t->m * 2 : 256; t->s = (char*)xrealloc(t->s, t->m); }
    t->s[t->l++] = (char)c; t->s[t->l] = 0;
}
/* read up to a delimiter; delim 0 = any isspace, 2 = newline (strip one trailing \r if l>1). returns delimiter or -1 */
static int st_until(lqo_stream *s, int delim, lqo_str *t, int append)
{
    int c;
    if (!append) { t->l = 0; if (t->s) t->s[0] = 0; }
    if (s->eof && s->beg >= s->end) return -2;
    while ((c = st_getc(s)) != -1) {
        if (delim == 2 ? c == '\n' : isspace(c)) break;
        str_putc(t, c);
    }
    if (!t->s) { t->m = 1; t->s = (char*)xcalloc(1, 1); }
    if (delim == 2 && t->l > 1 && t->s[t->l - 1] == '\r') t->s[--t->l] = 0;
    return c;
}

typedef struct { lqo_stream st; lqo_str name, comment, seq, qual; int last_char; } lqo_kseq;

static int kseq_next(lqo_kseq *ks) /* >=0 length, -1 EOF, -2 bad quality */
{
    int c, r;
    lqo_stream *s = &ks->st;
    if (ks->last_char == 0) {
        while ((c = st_getc(s)) != -1 && c != '>' && c != '@');
        if (c == -1) return -1;
        ks->last_char = c;
    }
    ks->comment.l = ks->seq.l = ks->qual.l = 0;
    if ((r = st_until(s, 0, &ks->name, 0)) == -2) return -1;
    c = r;
    if (c != '\n' && c != -1) st_until(s, 2, &ks->comment, 0);
    if (!ks->seq.s) { ks->seq.m = 256; ks->seq.s = (char*)xmalloc(256); ks->seq.s[0] = 0; }
    while ((c = st_getc(s)) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        str_putc(&ks->seq, c);
        st_until(s, 2, &ks->seq, 1);
    }
    if (c == '>' || c == '@') ks->last_char = c;
    if (c != '+') return (int)ks->seq.l;
    while ((c = st_getc(s)) != -1 && c != '\n');
    if (c == -1) return -2;
    while (st_until(s, 2, &ks->qual, 1) != -2 && ks->qual.l < ks->seq.l);
    ks->last_char = 0;
    if (ks->seq.l != ks->qual.l) return -2;
    return (int)ks->seq.l;
}

int lqo_reads_load(const char *fn, lqo_reads *out)
{
    lqo_kseq ks;
    int cap = 0, l;
    memset(out, 0, sizeof(*out));
    memset(&ks, 0, sizeof(ks));
    ks.st.fp = gzopen(fn, "r");
    if (!ks.st.fp) return -1;
    ks.st.buf = (unsigned char*)xmalloc(LQO_BUFSZ);
    while ((l = kseq_next(&ks)) >= 0) {
        if (out->n == cap) {
            cap = cap ? cap * 2 : 1024;
            out->name = (char**)xrealloc(out->name, cap * sizeof(char*));
            out->seq = (char**)xrealloc(out->seq, cap * sizeof(char*));
            out->qual = (char**)xrealloc(out->qual, cap * sizeof(char*));
            out->len = (int*)xrealloc(out->len, cap * sizeof(int));
        }
        out->name[out->n] = strdup(ks.name.s);
        out->seq[out->n] = (char*)xmalloc((size_t)l + 1);
        memcpy(out->seq[out->n], ks.seq.s, (size_t)l); out->seq[out->n][l] = 0;
        if (ks.qual.l) { out->qual[out->n] = (char*)xmalloc((size_t)l + 1); memcpy(out->qual[out->n], ks.qual.s, (size_t)l); out->qual[out->n][l] = 0; }
        else out->qual[out->n] = 0;
        out->len[out->n] = l;
        ++out->n;
    }
    gzclose(ks.st.fp);
    free(ks.st.buf); free(ks.name.s); free(ks.comment.s); free(ks.seq.s); free(ks.qual.s);
    return 0;
}

void lqo_reads_free(lqo_reads *r)
{
    int i;
    for (i = 0; i < r->n; ++i) { free(r->name[i]); free(r->seq[i]); free(r->qual[i]); }
    free(r->name); free(r->seq); free(r->qual); free(r->len);
    memset(r, 0, sizeof(*r));
}

/* ------------------------------------------------------------------ per-query mapping (lqmap.c:207-326) */

typedef struct {
    uint64_t lambda, lambda2;
    float avg_k;
    uint16_t *cnt; uint32_t n_cnt;
    lqo_sub_v ovlp;
} lqo_qacc;

static void map_query(const lqo_opt *opt, const lqo_index *ix, int mid_occ, const char *qname, const char *qseq, int qlen,
                      lqo_qacc *acc, lqo_trace *tr)
{
    lqo_mm128_v mv = { 0, 0, 0 };
    lqo_mm128 *a;
    int64_t n_a;
    int n_mini_pos, n_u = 0;
    uint64_t *mini_pos, *u = 0;
    lqo_reg *regs;
    lqo_sub_v cv = { 0, 0, 0 };
    uint32_t hash;
    if (qlen == 0) return;
    hash = qname ? x31_str(qname) : 0;
    hash ^= wang32((uint32_t)qlen) + wang32((uint32_t)opt->seed);
    hash = wang32(hash);
    lqo_sketch(qseq, qlen, ix->w, ix->k, 0, ix->is_hpc, &mv); /* lqmap.c:131 */
    a = collect_seeds(opt, mid_occ, ix, qname, &mv, qlen, &n_a, &n_mini_pos, &mini_pos);
    if (tr) {
        tr->n_mini = (int)mv.n; tr->n_kept = n_mini_pos; tr->n_seeds = n_a;
        tr->seeds_unsorted = (lqo_mm128*)xmalloc((size_t)n_a * sizeof(lqo_mm128));
        memcpy(tr->seeds_unsorted, a, (size_t)n_a * sizeof(lqo_mm128));
        tr->mini_pos = (uint64_t*)xmalloc((size_t)n_mini_pos * 8);
        memcpy(tr->mini_pos, mini_pos, (size_t)n_mini_pos * 8);
    }
    lqo_radix_sort_128x(a, a + n_a); /* lqmap.c:238 */
    if (tr) {
        tr->seeds_sorted = (lqo_mm128*)xmalloc((size_t)n_a * sizeof(lqo_mm128));
        memcpy(tr->seeds_sorted, a, (size_t)n_a * sizeof(lqo_mm128));
    }
    a = lqo_chain_dp(opt->max_gap, opt->max_gap, opt->bw, opt->max_chain_skip, opt->min_cnt, opt->min_chain_score, n_a, a, &n_u, &u);
    if (tr) {
        int i; int64_t na = 0;
        tr->n_chains = n_u;
        for (i = 0; i < n_u; ++i) na += (int32_t)u[i];
        tr->n_anchors = na;
        tr->u = (uint64_t*)xmalloc((size_t)n_u * 8); if (n_u) memcpy(tr->u, u, (size_t)n_u * 8);
        tr->anchors = (lqo_mm128*)xmalloc((size_t)na * sizeof(lqo_mm128)); if (na) memcpy(tr->anchors, a, (size_t)na * sizeof(lqo_mm128));
    }
    regs = gen_regs(hash, qlen, n_u, u, a);
    if (acc) {
        cnt_match(opt, ix, qlen, n_u, regs, a, n_mini_pos, mini_pos, acc->cnt, &cv, &acc->lambda, &acc->lambda2, &acc->avg_k);
        filter_redundant(&acc->ovlp, &cv, (uint32_t)opt->min_coverage);
    }
    free(cv.a); free(regs); free(mv.a); free(a); free(u); free(mini_pos);
}

void lqo_trace_free(lqo_trace *t)
{
    free(t->seeds_unsorted); free(t->seeds_sorted); free(t->mini_pos); free(t->u); free(t->anchors);
    memset(t, 0, sizeof(*t));
}

int lqo_trace_query(const lqo_opt *opt, const lqo_reads *targets, const lqo_reads *queries, int qi, int mid_occ, lqo_trace *t)
{
    lqo_index ix;
    memset(t, 0, sizeof(*t));
    index_build(&ix, opt, targets->name, targets->seq, targets->len, (uint32_t)targets->n);
    if (mid_occ <= 0) mid_occ = index_mid_occ(&ix, opt->mid_occ_frac);
    t->mid_occ = mid_occ;
    map_query(opt, &ix, mid_occ, queries->name[qi], queries->seq[qi], queries->len[qi], 0, t);
    index_free(&ix);
    return 0;
}

/* ------------------------------------------------------------------ whole program (minimap2-coverage.c:406-617) */

static void print_row(const lqo_opt *opt, FILE *out, const char *name, int len, const char *qual, lqo_qacc *acc)
{
    lqo_sub_v regs = { 0, 0, 0 }, mregs = { 0, 0, 0 };
    uint32_t j, sum = 0, tot = 0;
    int32_t n_match = 0;
    double div, mq;
    for (j = 0; j < acc->n_cnt; ++j) sum += acc->cnt[j];
    sum /= acc->n_cnt; /* parity domain: every query has >=1 minimizer (SIGFPE otherwise, minimap2-coverage.c:558) */
    for (j = 0; j < acc->n_cnt; ++j) if (acc->cnt[j] > sum) ++n_match;
    div = n_match > 0 ? logf((float)acc->n_cnt / n_match) / acc->avg_k : 1.0;
    reliable_region(&acc->ovlp, (uint32_t)opt->min_coverage, &regs, &mregs);
    mq = lqo_meanQ(qual, qual ? len : 0);
    if (regs.n > 0) {
        fprintf(out, "%s\t%d\t%" PRIu64 "\t", name, len, acc->lambda);
        for (j = 0; j < regs.n; ++j) {
            fprintf(out, "%s%d-%d", j ? "," : "", regs.a[j].start, regs.a[j].end);
            tot += regs.a[j].end - regs.a[j].start;
        }
        fputc('\t', out);
        if (mregs.n > 0) for (j = 0; j < mregs.n; ++j) fprintf(out, "%s%d-%d", j ? "," : "", mregs.a[j].start, mregs.a[j].end);
        else fputc('0', out);
        if (opt->filter) fprintf(out, "\t%.3f\t%.3f\t%.3f\t0.0\n", (double)tot / len, mq, div);
        else fprintf(out, "\t%.3f\t%.3f\t%.3f\t%.3f\n", (double)acc->lambda / tot, mq, div, (double)acc->lambda2 / tot);
    } else {
        fprintf(out, "%s\t%d\t%" PRIu64 "\t0\t0\t0.0\t%.3f\t%.3f\t0.0\n", name, len, acc->lambda, mq, div);
    }
    free(regs.a); free(mregs.a);
}

int lqo_run(const lqo_opt *opt, const lqo_reads *targets, const lqo_reads *queries, FILE *out, int *mid_occ_out, int *n_parts_out)
{
    lqo_qacc *acc = (lqo_qacc*)xcalloc(queries->n, sizeof(lqo_qacc));
    int q, mid_occ = 0, n_parts = 0;
    int64_t t0 = 0;
    uint64_t mini = (uint64_t)opt->mini_batch_size < opt->batch_size ? (uint64_t)opt->mini_batch_size : opt->batch_size; /* index.c:316 */

    for (q = 0; q < queries->n; ++q) { /* minimap2-coverage.c:418-427 */
        lqo_mm128_v mv = { 0, 0, 0 };
        if (queries->len[q] > 0) lqo_sketch(queries->seq[q], queries->len[q], opt->w, opt->k, (uint32_t)q, opt->is_hpc, &mv);
        acc[q].n_cnt = (uint32_t)mv.n;
        acc[q].cnt = (uint16_t*)xcalloc(mv.n, 2);
        free(mv.a);
    }
    while (t0 < targets->n) { /* one index part per iteration (minimap2-coverage.c:450; index.c:238-330) */
        lqo_index ix;
        uint64_t sum_len = 0;
        int64_t t1 = t0;
        while (t1 < targets->n && !(sum_len > opt->batch_size)) { /* index.c:244: checked before each mini-batch */
            uint64_t sz = 0;
            while (t1 < targets->n) { /* bseq.c:82-87: a mini-batch ends with the read that makes size >= chunk */
                sz += (uint64_t)targets->len[t1]; sum_len += (uint64_t)targets->len[t1]; ++t1;
                if (sz >= mini) break;
            }
        }
        index_build(&ix, opt, targets->name + t0, targets->seq + t0, targets->len + t0, (uint32_t)(t1 - t0));
        if (mid_occ <= 0) mid_occ = index_mid_occ(&ix, opt->mid_occ_frac); /* map.c:50-51: frozen from the first part */
        for (q = 0; q < queries->n; ++q)
            map_query(opt, &ix, mid_occ, queries->name[q], queries->seq[q], queries->len[q], &acc[q], 0);
        index_free(&ix);
        ++n_parts; t0 = t1;
    }
    if (out)
        for (q = 0; q < queries->n; ++q)
            print_row(opt, out, queries->name[q], queries->len[q], queries->qual[q], &acc[q]);
    for (q = 0; q < queries->n; ++q) { free(acc[q].cnt); free(acc[q].ovlp.a); }
    free(acc);
    if (mid_occ_out) *mid_occ_out = mid_occ;
    if (n_parts_out) *n_parts_out = n_parts;
    return 0;
}

/* ------------------------------------------------------------------ sdust (sdust.c:72-223) */

#define SD_W3 64 /* 4^3 triplet codes */
typedef struct { int start, finish, r, l; } sd_intv;

typedef struct {
    int win[64], w_front, w_n;     /* triplet deque, at most W-2 entries; capacity grows to 64 (kdq bits) */
    sd_intv *P; int nP, mP;        /* perfect intervals, descending start */
    uint64_t *res; int nres, mres;
} sd_state;

static inline int sd_at(const sd_state *s, int i) { return s->win[(s->w_front + i) & 63]; }

static void sd_save(sd_state *s, int start) /* sdust.c:94-108 */
{
    int i;
    sd_intv *p;
    if (s->nP == 0 || s->P[s->nP - 1].start >= start) return;
    p = &s->P[s->nP - 1];
    if (s->nres && p->start <= (int)(uint32_t)s->res[s->nres - 1]) {
        int b = (int)(s->res[s->nres - 1] >> 32), f = (int)(uint32_t)s->res[s->nres - 1];
        s->res[s->nres - 1] = (uint64_t)b << 32 | (uint32_t)(f > p->finish ? f : p->finish);
    } else {
        if (s->nres == s->mres) { s->mres = s->mres ? s->mres * 2 : 16; s->res = (uint64_t*)xrealloc(s->res, s->mres * 8); }
        s->res[s->nres++] = (uint64_t)p->start << 32 | (uint32_t)p->finish;
    }
    for (i = s->nP - 1; i >= 0 && s->P[i].start < start; --i);
    s->nP = i + 1;
}

uint64_t *lqo_sdust(const uint8_t *seq, int l_seq, int T, int W, int *n)
{
    sd_state s;
    int cw[SD_W3], cv[SD_W3], rw = 0, rv = 0, L = 0, i, l = 0, start;
    unsigned t = 0;
    memset(&s, 0, sizeof(s)); memset(cw, 0, sizeof(cw)); memset(cv, 0, sizeof(cv));
    if (l_seq < 0) l_seq = (int)strlen((const char*)seq);
    if (W > 66) W = 66; /* deque capacity here; LongQC never changes W=64 */
    for (i = 0; i <= l_seq; ++i) {
        int b = i < l_seq ? nt4_sdust(seq[i]) : 4;
        if (b < 4) {
            ++l; t = (t << 2 | (unsigned)b) & 63;
            if (l >= 3) {
                int x;
                start = (l - W > 0 ? l - W : 0) + (i + 1 - l);
                sd_save(&s, start);
                /* shift_window (sdust.c:72-92) */
                if (s.w_n >= W - 3 + 1) {
                    x = s.win[s.w_front]; s.w_front = (s.w_front + 1) & 63; --s.w_n;
                    rw -= --cw[x];
                    if (L > s.w_n) { --L; rv -= --cv[x]; }
                }
                s.win[(s.w_front + s.w_n++) & 63] = (int)t;
                ++L;
                rw += cw[t]++; rv += cv[t]++;
                if (cv[t] * 10 > T << 1) {
                    do { x = sd_at(&s, s.w_n - L); rv -= --cv[x]; --L; } while (x != (int)t);
                }
                if (rw * 10 > L * T) { /* find_perfect (sdust.c:110-134) */
                    int c[SD_W3], r = rv, ii, max_r = 0, max_l = 0;
                    memcpy(c, cv, sizeof(c));
                    for (ii = s.w_n - L - 1; ii >= 0; --ii) {
                        int j, tt = sd_at(&s, ii), new_r, new_l;
                        r += c[tt]++;
                        new_r = r; new_l = s.w_n - ii - 1;
                        if (new_r * 10 > T * new_l) {
                            for (j = 0; j < s.nP && s.P[j].start >= ii + start; ++j)
                                if (max_r == 0 || s.P[j].r * max_l > max_r * s.P[j].l) { max_r = s.P[j].r; max_l = s.P[j].l; }
                            if (max_r == 0 || new_r * max_l >= max_r * new_l) {
                                max_r = new_r; max_l = new_l;
                                if (s.nP == s.mP) { s.mP = s.mP ? s.mP * 2 : 16; s.P = (sd_intv*)xrealloc(s.P, s.mP * sizeof(sd_intv)); }
                                memmove(&s.P[j + 1], &s.P[j], (size_t)(s.nP - j) * sizeof(sd_intv));
                                ++s.nP;
                                s.P[j].start = ii + start; s.P[j].finish = s.w_n + 2 + start; s.P[j].r = new_r; s.P[j].l = new_l;
                            }
                        }
                    }
                }
            }
        } else { /* N or end: flush; the deque and counts are NOT cleared (sdust.c:158-162) */
            start = (l - W + 1 > 0 ? l - W + 1 : 0) + (i + 1 - l);
            while (s.nP) sd_save(&s, start++);
            l = 0; t = 0;
        }
    }
    free(s.P);
    *n = s.nres;
    return s.res;
}

int lqo_sdust_run(const lqo_reads *reads, int W, int T, FILE *out)
{
    int i, j, n;
    for (i = 0; i < reads->n; ++i) {
        uint32_t masked = 0;
        uint64_t *r = lqo_sdust((const uint8_t*)reads->seq[i], reads->len[i], T, W, &n); /* the reference passes -1: strlen */
        int ql = reads->qual[i] ? reads->len[i] : 0;
        for (j = 0; j < n; ++j) masked += (uint32_t)((int)r[j] - (int)(r[j] >> 32));
        fprintf(out, "%s\t%d\t%d\t%.3f\t%.3f\t%d\n", reads->name[i], masked, reads->len[i],
                (double)masked / reads->len[i], lqo_meanQ(reads->qual[i], ql), lqo_getQV(reads->qual[i], 7, ql));
        free(r);
    }
    return 0;
}
