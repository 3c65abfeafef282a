/* lq_host.h -- host-side (plain C) pieces of the path: per-query overlap bookkeeping and the
 * output table.  The data per query is a few hundred intervals, so these stay on the CPU exactly
 * where the reference has them (lqmap.c:25-100, lqutils.c:26-155, minimap2-coverage.c:545-617). */
#ifndef LQ_HOST_H
#define LQ_HOST_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t start, end; } lqh_sub;          /* minimap2-coverage.h:22-25, endpoints encoded pos<<3|flags */
typedef struct { size_t n, m; lqh_sub *a; } lqh_sub_v;
typedef struct { size_t l, m; char *s; } lqh_str;

void lqh_sub_push(lqh_sub_v *v, lqh_sub s);
void lqh_str_printf(lqh_str *s, const char *fmt, ...);

/* lqmap.c:25-100: fold the overlaps `cv` found in one index part into the query's persistent list `v` */
void lqh_filter_redundant(lqh_sub_v *v, const lqh_sub *cv, size_t n_cv, uint32_t min_cov);
/* lqutils.c:83-155 */
void lqh_reliable_region(const lqh_sub_v *v, uint32_t min_cov, lqh_sub_v *coords, lqh_sub_v *mcoords);
/* lqutils.c:51-58, 72-80 */
double lqh_meanQ(const char *qual, int len);
double lqh_q2p(int q);            /* entry q of the reference's Phred->probability table */
int lqh_getQV(const char *qual, int threshold, int len);
/* minimap2-coverage.c:552-604: one table row */
/* has_qual: sum_p is the ordered sum of the read's error probabilities (computed on the device); else the row prints -nan */
void lqh_format_row(lqh_str *out, const char *name, size_t name_len, int len, int has_qual, double sum_p, uint64_t lambda, uint64_t lambda2,
                    uint32_t n_mini, uint32_t n_match, float avg_k, const lqh_sub_v *ovlp, int min_cov, int filter);
/* sdust.c:211-217 */
void lqh_format_sdust_row(lqh_str *out, const char *name, size_t name_len, uint32_t masked, int len, const char *qual, double sum_p, int n_q7);

#ifdef __cplusplus
}
#endif
#endif
