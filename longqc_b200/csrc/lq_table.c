/* lq_table.c -- host-side overlap bookkeeping and table formatting (see lq_host.h). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>
#include <inttypes.h>
#include "lq_host.h"

static void *xrealloc(void *p, size_t n) { void *q = realloc(p, n ? n : 1); if (!q) { fprintf(stderr, "[lqcov] out of host memory\n"); abort(); } return q; }

void lqh_sub_push(lqh_sub_v *v, lqh_sub s)
{
    if (v->n == v->m) { v->m = v->m ? v->m << 1 : 32; v->a = (lqh_sub*)xrealloc(v->a, v->m * sizeof(lqh_sub)); }
    v->a[v->n++] = s;
}

void lqh_str_printf(lqh_str *s, const char *fmt, ...)
{
    va_list ap; int need;
    for (;;) {
        va_start(ap, fmt);
        need = vsnprintf(s->s ? s->s + s->l : NULL, s->s ? s->m - s->l : 0, fmt, ap);
        va_end(ap);
        if (s->s && (size_t)need < s->m - s->l) { s->l += (size_t)need; return; }
        s->m = (s->m + (size_t)need + 64) * 2;
        s->s = (char*)xrealloc(s->s, s->m);
    }
}

static int cmp_u32(const void *a, const void *b) { uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return x < y ? -1 : x > y; }

/* endpoints of all intervals, ascending (the reference uses radix_sort_32: a full-key sort, so any sort agrees) */
static uint32_t *sorted_endpoints(const lqh_sub *a, size_t n)
{
    uint32_t *e = (uint32_t*)xrealloc(NULL, 2 * n * sizeof(uint32_t));
    size_t i;
    for (i = 0; i < n; ++i) { e[2 * i] = a[i].start; e[2 * i + 1] = a[i].end; }
    qsort(e, 2 * n, sizeof(uint32_t), cmp_u32);
    return e;
}

/* Endpoint encoding: bit0 = end, bit1 = overlap reached the medium score (-p), bit2 = marker interval
 * standing for min_cov medium overlaps. */
void lqh_filter_redundant(lqh_sub_v *v, const lqh_sub *cv, size_t n_cv, uint32_t min_cov)
{
    uint32_t *e, depth = 0, open_at = 0;
    lqh_sub_v solid = { 0, 0, 0 }; /* stretches covered by >= min_cov medium overlaps in this part */
    size_t i, j;
    if (n_cv == 0) return;
    e = sorted_endpoints(cv, n_cv);
    for (j = 0; j < 2 * n_cv; ++j) {
        const uint32_t x = e[j], before = depth;
        if (x & 2) {
            const uint32_t step = (x & 4) ? min_cov : 1;
            depth = (x & 1) ? depth - step : depth + step;
        }
        if (before < min_cov && depth >= min_cov) open_at = x;             /* stays encoded (lqmap.c:61) */
        else if (before >= min_cov && depth < min_cov) {
            if ((uint32_t)((x >> 3) - open_at) > 0) {
                lqh_sub s, mk;
                s.start = open_at; s.end = x; lqh_sub_push(&solid, s);
                mk.start = open_at | 4; mk.end = x | 4; lqh_sub_push(v, mk);  /* marker into the persistent list */
            }
        }
    }
    free(e);
    for (i = 0; i < n_cv; ++i) { /* keep what is not swallowed by a solid stretch (encoded compare, lqmap.c:81-84) */
        int inside = 0;
        if (!(cv[i].start & 4))
            for (j = 0; j < solid.n; ++j)
                if (cv[i].start >= solid.a[j].start && cv[i].end <= solid.a[j].end) inside = 1;
        if (!inside) lqh_sub_push(v, cv[i]);
    }
    free(solid.a);
}

void lqh_reliable_region(const lqh_sub_v *v, uint32_t min_cov, lqh_sub_v *coords, lqh_sub_v *mcoords)
{
    uint32_t *e = sorted_endpoints(v->a, v->n), cov = 0, med = 0, cov_at = 0, med_at = 0;
    size_t j;
    for (j = 0; j < 2 * v->n; ++j) {
        const uint32_t x = e[j], pos = x >> 3, c0 = cov, m0 = med;
        const uint32_t dc = (x & 2) && (x & 4) ? min_cov : 1;         /* a marker moves total coverage by min_cov (1 + (min_cov-1)) */
        const uint32_t dm = (x & 2) ? ((x & 4) ? min_cov : 1) : 0;
        if (x & 1) { cov -= dc; med -= dm; } else { cov += dc; med += dm; }
        if (c0 < min_cov && cov >= min_cov) {
            cov_at = pos;
            if (m0 < min_cov && med >= min_cov) med_at = pos;
        } else if (c0 >= min_cov && cov < min_cov) {
            if (pos - cov_at > 0) { lqh_sub s; s.start = cov_at; s.end = pos; lqh_sub_push(coords, s); }
            if (m0 >= min_cov && med < min_cov && pos - med_at > 0) { lqh_sub s; s.start = med_at; s.end = pos; lqh_sub_push(mcoords, s); }
        } else if (m0 < min_cov && med >= min_cov) {
            med_at = pos;
        } else if (m0 >= min_cov && med < min_cov) {
            if (pos - med_at > 0) { lqh_sub s; s.start = med_at; s.end = pos; lqh_sub_push(mcoords, s); }
        }
    }
    free(e);
}

/* The reference's Phred table (lqutils.c:26-49) lists 10^(-q/10) with 15 decimals.  Printing pow() with %.15f and
 * reading the text back gives the same doubles except for eight entries whose last printed digit is one higher in
 * the reference (q = 34, 39, 58, 62, 67, 71, 72, 82); tests/test_oracle_vs_reference.py checks all 127 entries
 * against the compiled reference. */
static double lq_q2p_entry(int q)
{
    static const int plus1[8] = { 34, 39, 58, 62, 67, 71, 72, 82 };
    char b[48]; int i, n;
    n = snprintf(b, sizeof b, "%.15f", pow(10.0, -q / 10.0));
    for (i = 0; i < 8; ++i)
        if (plus1[i] == q) { int j = n - 1; while (j >= 0 && b[j] == '9') b[j--] = '0'; if (j >= 0 && b[j] != '.') ++b[j]; }
    return strtod(b, NULL);
}
static double q2p[127];
static int q2p_ok = 0;
static void q2p_build(void) { int q; for (q = 0; q < 127; ++q) q2p[q] = lq_q2p_entry(q); q2p_ok = 1; }
double lqh_q2p(int q) { return lq_q2p_entry(q); }

double lqh_meanQ(const char *qual, int len)
{
    double acc = 0.0; int i;
    if (!q2p_ok) q2p_build();
    for (i = 0; i < len; ++i) { int q = (int)qual[i] - 33; acc += q2p[q < 0 ? 0 : q > 126 ? 126 : q]; }
    return -10 * log10(acc / len);
}

int lqh_getQV(const char *qual, int threshold, int len)
{
    int i, n = 0;
    for (i = 0; i < len; ++i) n += (int)qual[i] > threshold + 33;
    return n;
}

static void put_regions(lqh_str *out, const lqh_sub_v *r)
{
    size_t j;
    for (j = 0; j < r->n; ++j) lqh_str_printf(out, "%s%d-%d", j ? "," : "", r->a[j].start, r->a[j].end);
}

void lqh_format_row(lqh_str *out, const char *name, size_t name_len, int len, int has_qual, double sum_p, uint64_t lambda, uint64_t lambda2,
                    uint32_t n_mini, uint32_t n_match, float avg_k, const lqh_sub_v *ovlp, int min_cov, int filter)
{
    lqh_sub_v reg = { 0, 0, 0 }, mreg = { 0, 0, 0 };
    /* minimap2-coverage.c:563: float log of a float ratio, divided by the float mean k-mer span */
    const double div = n_match > 0 ? logf((float)n_mini / n_match) / avg_k : 1.0;
    const double mq = has_qual ? -10 * log10(sum_p / len) : lqh_meanQ(NULL, 0);   /* lqutils.c:57; FASTA: 0/0 -> -nan */
    uint32_t tot = 0; size_t j;
    lqh_reliable_region(ovlp, (uint32_t)min_cov, &reg, &mreg);
    for (j = 0; j < reg.n; ++j) tot += reg.a[j].end - reg.a[j].start;
    lqh_str_printf(out, "%.*s\t%d\t%" PRIu64 "\t", (int)name_len, name, len, lambda);
    if (reg.n > 0) {
        put_regions(out, &reg);
        lqh_str_printf(out, "\t");
        if (mreg.n > 0) put_regions(out, &mreg); else lqh_str_printf(out, "0");
        if (filter) lqh_str_printf(out, "\t%.3f\t%.3f\t%.3f\t0.0\n", (double)tot / len, mq, div);
        else lqh_str_printf(out, "\t%.3f\t%.3f\t%.3f\t%.3f\n", (double)lambda / tot, mq, div, (double)lambda2 / tot);
    } else lqh_str_printf(out, "0\t0\t0.0\t%.3f\t%.3f\t0.0\n", mq, div);
    free(reg.a); free(mreg.a);
}

void lqh_format_sdust_row(lqh_str *out, const char *name, size_t name_len, uint32_t masked, int len, const char *qual, double sum_p, int n_q7)
{
    /* sum_p: the ordered sum of error probabilities (lqutils.c:54-56), accumulated on the device in read order */
    const double mq = qual ? -10 * log10(sum_p / len) : lqh_meanQ(NULL, 0);
    lqh_str_printf(out, "%.*s\t%d\t%d\t%.3f\t%.3f\t%d\n", (int)name_len, name, masked, len, (double)masked / len, mq, n_q7);
}
