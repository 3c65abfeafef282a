/* lq_ingest.h -- host-side readers shared by lq_fastx.c, lq_ingest.c and the executables (internal; the exported calls are in
 * include/lqcov.h). */
#ifndef LQ_INGEST_H
#define LQ_INGEST_H
#include <stdint.h>
#include "lqcov.h"
#ifdef __cplusplus
extern "C" {
#endif

/* sequential kseq-compatible reader (gz, pipes, stdin): lq_fastx.c */
typedef struct lqs_reader lqs_reader;
lqs_reader *lqs_open(const char *path);
int  lqs_next(lqs_reader *r, int64_t chunk, lqcov_reads_t *out);   /* records until their lengths reach `chunk` (<= 0: all) or a record kseq rejects; 0 = end of input */
int  lqs_broke(const lqs_reader *r);                               /* the last batch ended on a rejected record */
void lqs_close(lqs_reader *r);

/* multi-threaded reader: lq_ingest.c */
typedef struct lqi_reader lqi_reader;
typedef struct {
    uint32_t n;                 /* records of this chunk */
    uint64_t n_bases;
    const uint64_t *seq_off;    /* n+1 offsets into the caller's buffers */
    const char *names; const uint64_t *name_off;   /* n+1; valid until the next call */
    int has_qual;               /* at least one record came with qualities */
    int part_end;               /* the rule in force closes the record set with this chunk */
    int eof;                    /* nothing follows */
    uint64_t need;              /* rc == -2: the next record alone needs this many bytes */
} lqi_chunk;
lqi_reader *lqi_open(const char *path, int n_threads);
void lqi_close(lqi_reader *r);
uint64_t lqi_bases_left_bound(const lqi_reader *r);   /* upper bound on the bases still to come (0 = unknown) */
/* index.c:244,316: a part = mini-batches (reads until their size >= min(mini, batch)) while the part's sum of lengths <= batch */
void lqi_part_rule(lqi_reader *r, uint64_t batch_size, int mini_batch_size);
/* bseq.c:86-87: a batch = reads until their size >= chunk (0: no rule, everything is one set) */
void lqi_batch_rule(lqi_reader *r, uint64_t chunk);
/* The next records, as many as fit `cap` bytes: bases to seq_dst, qualities to qual_dst (NULL: not wanted; records without
 * qualities are zero-filled).  Returns 1 = records delivered, 0 = none left, -2 = the next record does not fit (out->need). */
int lqi_next_chunk(lqi_reader *r, uint64_t cap, char *seq_dst, char *qual_dst, lqi_chunk *out);
#ifdef __cplusplus
}
#endif
#endif
