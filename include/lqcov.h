/* lqcov.h -- C ABI of the B200-native minimap2-coverage / sdust path (liblqcov.so).
 *
 * The reference (yfukasawa/LongQC) has no library interface for this path: LongQC drives two
 * executables, `minimap2-coverage` and `sdust`, through lq_exec.py:13-38 / lq_mask.py:17-23, and
 * consumes their stdout.  The drop-in boundary is therefore argv + stdout bytes + exit status
 * (lqcov_main / lqcov_sdust_main below are the two `main`s; longqc_b200/bin/ holds the thin
 * executables).  The remaining entry points expose the same pipeline on HOST BUFFERS so that
 * bindings (ctypes, cgo, JNI ...) and the parity tests can call it without going through files;
 * each names the reference code it replaces.
 *
 * Plain C types only.  Every function returns 0 on success, non-zero on error (message on stderr)
 * unless stated otherwise.  There is NO CPU implementation behind this ABI: lqcov_create() fails if
 * no CUDA device is usable.
 */
#ifndef LQCOV_H
#define LQCOV_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LQCOV_ABI_VERSION 1

/* Options == what main() of the reference holds after argp + defaulting
 * (minimap2-coverage.c:217-388, index.c:31-38, map.c:12-44). */
typedef struct {
    /* indexing */
    int k;                    /* -k, default 12 */
    int w;                    /* -w, default 5 */
    int is_hpc;               /* -H */
    uint64_t batch_size;      /* -I, default 4,000,000,000 bases per index part */
    int mini_batch_size;      /* 50,000,000 (index.c:36): granularity of the part boundary */
    /* mapping */
    int no_self;              /* MM_F_NO_SELF: set by -X and -Y */
    int ava;                  /* MM_F_AVA: -X only */
    int max_gap;              /* -g 10000 */
    int min_cnt;              /* -n 3 */
    int min_chain_score;      /* -m 40 */
    int min_score_med;        /* -p (default = -m) */
    int min_score_good;       /* -q (default = -m) */
    int max_chain_skip;       /* -s 25 */
    int bw;                   /* 500 (map.c:21) */
    float mid_occ_frac;       /* 2e-4f (map.c:16) */
    /* filtering */
    int max_overhang;         /* -a 2000 */
    int min_ovlp;             /* -l 1000 (parsed, unused by the reference too) */
    int min_coverage;         /* -c 3 */
    double min_ratio;         /* -r 0.4 */
    int filter;               /* -f / --filter */
    /* execution */
    int n_threads;            /* -t: host worker threads (parsing) */
    int device;               /* CUDA device ordinal; -1 = current / CUDA_VISIBLE_DEVICES order 0 */
    uint64_t seed_budget;     /* seeds per device batch (0 = default 1000 M, ~45 bytes of HBM each) */
    int verbose;              /* 0..3, stderr progress lines like the reference's mm_verbose */
} lqcov_opt_t;

/* A read set in memory.  seq/qual: one blob, read i occupies [seq_off[i], seq_off[i+1]).
 * names: one blob of n NUL-terminated names back to back is NOT required; name i occupies
 * [name_off[i], name_off[i+1]) without terminator.  qual may be NULL (FASTA).
 * seq_on_device != 0: `seq` is a CUDA device pointer (offsets and names stay on the host). */
typedef struct {
    uint32_t n;
    const char *seq;
    const uint64_t *seq_off;
    const char *qual;
    const char *names;
    const uint64_t *name_off;
    int seq_on_device;
} lqcov_reads_t;

typedef struct lqcov_ctx lqcov_ctx;

/* per-run counters (diagnostics, bench.py) */
typedef struct {
    uint64_t target_bases, query_bases, target_minimizers, query_minimizers;
    uint64_t seeds, groups, chains, overlaps, batches, walk_buckets;
    int32_t mid_occ; int32_t n_parts;
    double t_upload_ms, t_sketch_ms, t_index_ms, t_map_ms, t_post_ms; /* host-clock phase times, informational */
} lqcov_stats_t;

int  lqcov_abi_version(void);
int  lqcov_device_count(void);                               /* CUDA devices visible to this process */
void lqcov_opt_init(lqcov_opt_t *o);                       /* the defaults listed above, with -Y semantics */

/* minimap2-coverage.c:217-621 as a library ---------------------------------------------------- */
lqcov_ctx *lqcov_create(const lqcov_opt_t *o);              /* NULL on error (no usable GPU, bad options) */
void lqcov_destroy(lqcov_ctx *c);
int  lqcov_reset(lqcov_ctx *c);                               /* start a new job in the same context (buffers are kept) */
/* query pre-pass: sketch every query once, allocate the per-query accumulators (minimap2-coverage.c:406-444) */
int  lqcov_set_queries(lqcov_ctx *c, const lqcov_reads_t *queries);
/* one index part: mm_idx_gen (index.c:311-330) + mm_mapopt_update (map.c:46-54, mid_occ frozen from the
 * first part) + lq_map_file for every query (lqmap.c:851-855) */
int  lqcov_add_part(lqcov_ctx *c, const lqcov_reads_t *part);
/* the same in phases, so that several GPUs can share one part (one process per GPU; the collectives between
 * the phases are issued by the caller, see longqc_b200/dist.py and INTEGRATION.md):
 *   lqcov_part_sketch        sketch THIS rank's contiguous shard of the part's reads (rid = rid_base + i), count minimizers
 *   lqcov_part_device_views  device pointers of the count table (u32[4^k]) and of the local records (diagnostics)
 *   lqcov_part_gather_buffers device buffers for the records of ALL ranks, concatenated in rank order (a caller-driven all-gather;
 *                            superseded by lqcov_part_exchange, kept for callers that bring their own collectives)
 *   lqcov_part_finish        offsets + stable sort + mid_occ + name tables; `part` describes the WHOLE part
 *                            (n, seq_off for the lengths, names; seq may be NULL)
 *   lqcov_map_part           map this context's queries against the finished part */
int  lqcov_part_sketch(lqcov_ctx *c, const lqcov_reads_t *shard, uint32_t rid_base);
/* lqcov_part_sketch with the shard arriving in CHUNKS of consecutive reads (the executable's reader threads fill pinned staging
 * buffers while the device copies, packs and sketches the chunk before; replaces the reference's 3-stage kt_pipeline of
 * index.c:238-309 / kthread.c:96-158):
 *   lqcov_part_begin   expect_bases: first sizing of the device arrays (0 = unknown; they grow); returns 1 when the configuration
 *                      has no chunked form (-H): use lqcov_part_sketch
 *   lqcov_stage        n pinned host buffers of `bytes` bytes each, owned by the context
 *   lqcov_part_chunk   queue copy + pack + sketch of a chunk whose bases lie in staging buffer `stage_index` (-1: any host memory); returns at once
 *   lqcov_stage_wait   block until staging buffer `stage_index` may be refilled
 *   lqcov_part_end     all chunks are in: minimizers counted; go on with the collectives / lqcov_part_finish / lqcov_map_part */
int  lqcov_part_begin(lqcov_ctx *c, uint64_t expect_bases, uint32_t rid_base);
int  lqcov_stage(lqcov_ctx *c, int n, size_t bytes, char **bufs);
int  lqcov_part_chunk(lqcov_ctx *c, const lqcov_reads_t *chunk, int stage_index);
int  lqcov_stage_wait(lqcov_ctx *c, int stage_index);
int  lqcov_part_end(lqcov_ctx *c);
int  lqcov_part_device_views(lqcov_ctx *c, void **counts, uint64_t *n_counts, void **key, void **y, uint64_t *n_rec);
int  lqcov_part_gather_buffers(lqcov_ctx *c, uint64_t n_total, void **key, void **y);
int  lqcov_part_finish(lqcov_ctx *c, const lqcov_reads_t *part);
int  lqcov_map_part(lqcov_ctx *c);
/* Several GPUs on one index part (one context per GPU; NCCL over NVLink is loaded on demand, see lq_comm.cu).  The part's reads are
 * owned by the ranks in contiguous, rank-ordered ranges; every rank sketches its range (lqcov_part_sketch / _begin.._end with
 * rid_base = position of its first read in the part), then ALL ranks call lqcov_part_exchange: all-reduce of the minimizer counts,
 * every rank sorts only its own records, the sorted shards are exchanged and placed into the replicated index.  Then
 * lqcov_part_finish (whole-part metadata) and lqcov_map_part (this rank's queries) as on one GPU.
 *   lqcov_comm_unique_id   128 opaque bytes made by one rank and handed to the others by the launcher (torchrun: a broadcast)
 *   lqcov_comm_init_rank   one process per GPU
 *   lqcov_comm_init_all    one process driving n contexts (the drop-in executable); rank = index in ctxs
 *   lqcov_comm_gather_rows table rows of all ranks on rank 0, in rank order (*all malloc'ed on rank 0, NULL elsewhere) */
int  lqcov_comm_unique_id(void *id128);
int  lqcov_comm_init_rank(lqcov_ctx *c, const void *id128, int nranks, int rank);
int  lqcov_comm_init_all(lqcov_ctx **ctxs, int n);
int  lqcov_comm_size(const lqcov_ctx *c);
int  lqcov_comm_rank(const lqcov_ctx *c);
int  lqcov_part_exchange(lqcov_ctx *c);
int  lqcov_comm_gather_rows(lqcov_ctx *c, const char *mine, size_t len, char **all, size_t *all_len);
/* whole target set in memory: cut into parts exactly as index.c:238-330 does (mini-batch rule) and add each */
int  lqcov_add_targets(lqcov_ctx *c, const lqcov_reads_t *targets);
/* `-d FILE` and prebuilt indexes (index.c:390-479: the "MMI\2" image of the reference's index, byte for byte, khash slot order
 * included; longQC.py --db writes one with `-d` and maps against it later):
 *   lqcov_index_dump        append the part just indexed (lqcov_index_part / lqcov_add_part; `part` with its bases in host memory) to `file` (a FILE*)
 *   lqcov_index_peek        1 when `path` is such a file: k, w, -H of its first part (they override the command line's: index.c:524-526)
 *   lqcov_load_part         the next part of an index file becomes the current part (then lqcov_map_part); 1 = loaded, 0 = end of file
 *   lqcov_set_prepass_counts  the row's minimizer count `n` is the COMMAND LINE's k / w (minimap2-coverage.c:418-427) even when an index
 *                           built with other parameters is mapped against: records those counts (call after lqcov_set_queries) */
int  lqcov_index_dump(lqcov_ctx *c, const lqcov_reads_t *part, void *file);
int  lqcov_index_peek(const char *path, int *k, int *w, int *is_hpc);
int  lqcov_load_part(lqcov_ctx *c, void *file);
int  lqcov_set_prepass_counts(lqcov_ctx *c, const lqcov_reads_t *queries, int k, int w, int is_hpc);
/* the stdout table of minimap2-coverage.c:545-617, one row per query; *buf is malloc'ed (lqcov_free) */
int  lqcov_table(lqcov_ctx *c, char **buf, size_t *len);
int  lqcov_get_stats(const lqcov_ctx *c, lqcov_stats_t *s);
void lqcov_free(void *p);

/* sdust.c:187-223 as a library: the stdout table of `sdust <reads>` */
int  lqcov_sdust_table(const lqcov_opt_t *o, const lqcov_reads_t *reads, int W, int T, char **buf, size_t *len);
/* the same, chunk by chunk (the executable: reader threads fill pinned staging buffers while the device works on the chunk before):
 *   lqcov_sdust_begin  stage_bytes > 0: two pinned (sequence, quality) buffer pairs of that size are handed out in stage_seq[2] / stage_qual[2]
 *   lqcov_sdust_chunk  queue copy + kernel of a chunk of whole reads; *rows (malloc'ed) = the rows of the chunk BEFORE it, possibly empty.
 *                      The chunk's host buffers may be overwritten once the next lqcov_sdust_chunk / _end has returned
 *   lqcov_sdust_end    the remaining rows; frees the state */
typedef struct lqcov_sdust lqcov_sdust;
lqcov_sdust *lqcov_sdust_begin(const lqcov_opt_t *o, int W, int T, size_t stage_bytes, char **stage_seq, char **stage_qual);
int  lqcov_sdust_chunk(lqcov_sdust *s, const lqcov_reads_t *reads, char **rows, size_t *rows_len);
int  lqcov_sdust_end(lqcov_sdust *s, char **rows, size_t *rows_len);

/* stage-level entry points (parity tests against the oracle) --------------------------------- */
/* mm_sketch (sketch.c:76-142) of every read; records in (read, position) order.  x[i] = hash<<8|span,
 * y[i] = rid<<32|lastPos<<1|strand, rid = rid_base + read index.  *x,*y malloc'ed. */
int  lqcov_sketch(const lqcov_opt_t *o, const lqcov_reads_t *reads, uint32_t rid_base, uint64_t **x, uint64_t **y, uint64_t *n);
/* seeds of query q against the CURRENT part, before and after the reference-exact sort (lqmap.c:237-238):
 * x/y as the reference lays mm128_t seeds out (y without the tandem flag).  Arrays malloc'ed, n each. */
int  lqcov_debug_seeds(lqcov_ctx *c, uint32_t q, uint64_t **ux, uint64_t **uy, uint64_t **sx, uint64_t **sy, uint64_t *n);
/* test switch for the sketch (w = 5/10, k <= 15): 0 = the defaults (packed-key kernel fed by bulk copies for w = 5, k = 12 / 15,
 * rolling kernel otherwise), 1 = the tiled position-parallel kernel, 2 = the packed-key kernel fed by plain loads,
 * 3 = the rolling kernel everywhere */
void lqcov_debug_sketch_tiled(int on);
/* index one part WITHOUT mapping (used with lqcov_debug_seeds) */
int  lqcov_index_part(lqcov_ctx *c, const lqcov_reads_t *part);

/* per-kernel device timing (CUDA events on the library's stream), launch and transfer counters: bench.py */
void   lqcov_profile_enable(int on);
void   lqcov_profile_reset(void);
size_t lqcov_profile_json(char *buf, size_t cap);            /* returns the length needed (incl. NUL) */

/* host helpers shared with the executables -------------------------------------------------- */
/* FASTA/FASTQ(.gz) reader with kseq.h:185-224 + bseq.c:56-66 semantics.  Reads records until the
 * accumulated sequence length reaches `chunk` (chunk <= 0: whole file).  The returned set lives until
 * the next call or lqcov_reader_close.  Returns 1 = records returned, 0 = end of input, 2 = an empty set (see below), <0 = error.
 * A FASTQ record kseq_read() rejects (truncated or mismatched quality, kseq.h:216-222) is not delivered and ends what the reading
 * loop of the reference would end there: with chunk <= 0 the whole input (one kseq_read loop: minimap2-coverage.c:418, sdust.c:198),
 * with chunk > 0 the batch (bseq.c:76), for index parts the mini-batch -- and the part when that mini-batch is empty (index.c:246-247).
 * Plain files are read by several threads (lq_ingest.c; LQCOV_READER_THREADS), anything else sequentially. */
typedef struct lqcov_reader lqcov_reader;
lqcov_reader *lqcov_reader_open(const char *path);          /* "-" = stdin */
int  lqcov_reader_next(lqcov_reader *r, int64_t chunk, lqcov_reads_t *out);
/* index.c:238-290: one index PART = mini-batches until sum_len > batch_size */
int  lqcov_reader_next_part(lqcov_reader *r, uint64_t batch_size, int mini_batch_size, lqcov_reads_t *out);
void lqcov_reader_close(lqcov_reader *r);

/* the two executables */
int  lqcov_main(int argc, char **argv);                     /* minimap2-coverage.c:199 main */
int  lqcov_sdust_main(int argc, char **argv);               /* sdust.c:187 main */

#ifdef __cplusplus
}
#endif
#endif
