/* lq_oracle.h -- CPU restatement of LongQC's minimap2-coverage / sdust hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under longqc_b200/ may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do.
 *
 * Parity is PINNED: tests/test_oracle_vs_reference.py checks this port byte-for-byte
 * against the compiled, unmodified reference (oracle/_ref, built by oracle/Makefile from
 * /root/reference/minimap2-coverage) both at function level (mm_sketch,
 * radix_sort_128x, mm_chain_dp through _ref/libmm2ref.so) and at table level
 * (committed tests/golden/ tables produced by oracle/make_golden.py).
 *
 * Every function cites the reference file:line it restates.
 */
#ifndef LQ_ORACLE_H
#define LQ_ORACLE_H

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t x, y; } lqo_mm128;                 /* minimap.h:42 */
typedef struct { size_t n, m; lqo_mm128 *a; } lqo_mm128_v;    /* minimap.h:43 */

typedef struct {
    /* indexing (minimap2-coverage.c:252-276, index.c:31-38) */
    int k, w, is_hpc;
    uint64_t batch_size;       /* -I */
    int mini_batch_size;       /* 50,000,000 (index.c:36) */
    /* mapping (minimap2-coverage.c:229-241,302-367; map.c:12-44) */
    int no_self;               /* MM_F_NO_SELF: -X or -Y */
    int ava;                   /* MM_F_AVA: -X only */
    int max_gap, min_cnt, min_chain_score, min_score_med, min_score_good, max_chain_skip, bw;
    float mid_occ_frac;        /* 2e-4f */
    int seed;                  /* 11 */
    /* filtering (minimap2-coverage.c:280-388) */
    int max_overhang, min_ovlp, min_coverage;
    double min_ratio;
    int filter;                /* --filter */
} lqo_opt;

void lqo_opt_default(lqo_opt *o);   /* the values main() ends with when only -Y is given */

/* ---- function-level pieces (pinned against _ref/libmm2ref.so) ---- */
uint64_t lqo_hash64(uint64_t key, uint64_t mask);                                           /* sketch.c:27-37 */
void lqo_sketch(const char *str, int len, int w, int k, uint32_t rid, int is_hpc, lqo_mm128_v *p); /* sketch.c:76-142 */
void lqo_radix_sort_128x(lqo_mm128 *beg, lqo_mm128 *end);                                   /* ksort.h:84-134, misc.c:125-126 */
void lqo_radix_sort_64(uint64_t *beg, uint64_t *end);                                       /* misc.c:128-129 */
void lqo_radix_sort_32(uint32_t *beg, uint32_t *end);                                       /* lqutils.c:9-10 */
/* chain.c:22-157.  a[] (n seeds, sorted) is consumed; returns malloc'ed compacted anchors, *n_u chains in u[]. */
lqo_mm128 *lqo_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int min_cnt, int min_sc,
                        int64_t n, lqo_mm128 *a, int *n_u, uint64_t **u);
double lqo_meanQ(const char *qual, int len);                                                /* lqutils.c:51-58 */
int lqo_getQV(const char *qual, int threshold, int len);                                    /* lqutils.c:72-80 */
/* sdust.c:136-185: masked intervals (start<<32|finish), malloc'ed */
uint64_t *lqo_sdust(const uint8_t *seq, int l_seq, int T, int W, int *n);

/* ---- read sets in memory ---- */
typedef struct {
    int n;
    char **name, **seq, **qual;   /* qual[i] may be NULL (FASTA) */
    int *len;
} lqo_reads;
void lqo_reads_free(lqo_reads *r);
/* kseq.h:185-224 + bseq.c:56-66 semantics; reads the whole file (gz ok). Returns 0 / -1 on open error */
int lqo_reads_load(const char *fn, lqo_reads *out);

/* ---- whole program (minimap2-coverage.c main) ---- */
/* Run all-vs-subsample coverage: targets are split into index parts exactly as index.c:238-330 does
 * (mini-batch rule), queries mapped against each part, table written to `out`.
 * Returns 0 on success.  If mid_occ_out != NULL it receives the frozen mid_occ (map.c:50). */
int lqo_run(const lqo_opt *opt, const lqo_reads *targets, const lqo_reads *queries, FILE *out, int *mid_occ_out, int *n_parts_out);

/* sdust main (sdust.c:187-223) over an in-memory read set */
int lqo_sdust_run(const lqo_reads *reads, int W, int T, FILE *out);

/* ---- intermediate dumps for kernel-level parity (one query vs one part) ---- */
typedef struct {
    int mid_occ;
    int n_mini;            /* all query minimizers */
    int n_kept;            /* mini_pos entries */
    int64_t n_seeds;
    lqo_mm128 *seeds_unsorted, *seeds_sorted;  /* malloc'ed, n_seeds each */
    uint64_t *mini_pos;    /* n_kept */
    int n_chains;
    uint64_t *u;           /* score<<32|cnt per chain, chain order of mm_chain_dp's result */
    lqo_mm128 *anchors;    /* compacted */
    int64_t n_anchors;
} lqo_trace;
void lqo_trace_free(lqo_trace *t);
/* index `targets` as ONE part (rid = order), map query `qi` of `queries`, fill trace. mid_occ<=0 => compute. */
int lqo_trace_query(const lqo_opt *opt, const lqo_reads *targets, const lqo_reads *queries, int qi, int mid_occ, lqo_trace *t);

#ifdef __cplusplus
}
#endif
#endif
