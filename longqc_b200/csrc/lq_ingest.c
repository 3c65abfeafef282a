/* lq_ingest.c -- multi-threaded FASTA/FASTQ ingest with the record semantics of the reference's kseq.h:185-224.
 *
 * The reference reads its inputs through one kseq stream (bseq.c:68-102 inside a 3-stage kt_pipeline, index.c:238-309).  Here a plain
 * file is mapped into memory and cut into blocks; worker threads SCAN the blocks for records in parallel (offsets only, no copy), a
 * stitcher accepts a block only where the reference's sequential state machine would arrive at exactly the position the block
 * started from (otherwise the block is scanned again from the true state), and the accepted records are COPIED by the same workers
 * straight into the buffer the caller names (a pinned staging buffer of the GPU upload).  Nothing is parsed twice on well-formed
 * input and the result equals kseq_read()'s on every input, well-formed or not.  Target qualities are never copied.
 *
 * Compressed files, pipes and stdin take the sequential reader of lq_fastx.c (lqs_*), one record at a time, behind the same calls.
 *
 * kseq details that are reproduced because they change what a record is:
 *   - a record starts at the next '>' or '@' found ANYWHERE after a FASTQ record, or at the '>' / '@' that ended a FASTA record;
 *   - name = header up to the first isspace() character, the rest of the line is dropped;
 *   - sequence lines are appended until a line starts with '>', '@' or '+'; empty lines are skipped; one trailing '\r' is dropped
 *     per line when the accumulated string is longer than one character (kseq.h:138);
 *   - quality lines are appended until they cover the sequence; a length mismatch or a missing quality ends the whole input
 *     (kseq_read() < 0 stops mm_bseq_read2's loop);
 *   - kseq only knows it is at end of file after a short read of its 16 KB buffer: a line read at the very end of a file whose
 *     length is a multiple of 16384 still applies the '\r' rule once (kseq.h:98-106,138).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <unistd.h>
#include <fcntl.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include "lqcov.h"
#include "lq_ingest.h"

#define LQI_BLOCK_DEFAULT (4u << 20) /* bytes per scan block (LQCOV_READER_BLOCK overrides: tests cut small files into many blocks) */
#define LQI_AHEAD 96                 /* blocks scanned ahead of the stitcher at most */
#define LQI_NOQUAL (~0ULL)

typedef struct {
    uint64_t name_off, seq_off, qual_off;
    uint32_t name_len, len;
    uint32_t simple;                 /* bit 0: the sequence is `len` contiguous bytes at seq_off; bit 1: same for the quality at qual_off */
} lqi_rec;

typedef struct { lqi_rec *a; size_t n, m; } rec_v;

/* the reference's stream state between two records */
typedef struct {
    uint64_t pos;
    int last_char;                   /* header character already consumed (kseq_t.last_char) */
    int zero_read;                   /* kseq's buffer has seen a zero-length read (then it knows it is at end of file) */
    int failed;                      /* kseq_read() returned -2: nothing more is delivered */
    int eof;
} lqi_state;

typedef struct { const uint8_t *b; uint64_t n; } lqi_mem;

static void *xrealloc(void *p, size_t n) { void *q = realloc(p, n ? n : 1); if (!q) { fprintf(stderr, "[lqcov] out of host memory\n"); abort(); } return q; }
static inline void recv_push(rec_v *v, const lqi_rec *r) { if (v->n == v->m) { v->m = v->m ? v->m * 2 : 256; v->a = (lqi_rec*)xrealloc(v->a, v->m * sizeof(lqi_rec)); } v->a[v->n++] = *r; }

static inline int known_eof(const lqi_mem *m, const lqi_state *s) { return (m->n % 16384u) != 0 || s->zero_read; }

/* Accumulated string of kseq (seq or qual) as a length and, in the copy pass, a destination.  Nothing is ever written past the
 * final length: the copy pass of two neighbouring records runs on different threads, so the byte a '\r' rule drops must not
 * touch the next record's first byte even transiently. */
typedef struct { uint8_t *dst; uint64_t l; uint8_t lastc; } acc_t;   /* dst may be NULL (count only) */

/* ks_getuntil2(KS_SEP_LINE, append = 1) with `pre` bytes of the same line already consumed by the caller (the first character of
 * a sequence line, kseq.h:205-206; 0 for quality lines): appends [s->pos - pre, end of line) and applies the '\r' rule (kseq.h:138).
 * Returns -1 where kseq returns -1 (nothing left and kseq knows it: the rule is NOT applied then), else 0. */
static inline int line_append(const lqi_mem *m, lqi_state *s, acc_t *a, uint32_t pre)
{
    const uint64_t p0 = s->pos - pre;
    uint64_t e; int trim = 1, rc = 0;
    if (s->pos >= m->n) {
        if (known_eof(m, s)) { trim = 0; rc = -1; } else s->zero_read = 1;
        e = m->n;
    } else {
        const uint8_t *q = (const uint8_t*)memchr(m->b + s->pos, '\n', (size_t)(m->n - s->pos));
        if (q) { e = (uint64_t)(q - m->b); s->pos = e + 1; }
        else { e = m->n; s->pos = m->n; s->zero_read = 1; }
    }
    {
        uint64_t keep = e - p0;
        uint8_t lastc = keep ? m->b[e - 1] : a->lastc;
        if (trim && a->l + keep > 1 && lastc == '\r') {
            if (keep) --keep; else --a->l;         /* keep == 0: an earlier, still untrimmed byte of this record */
            lastc = 0;                              /* one '\r' per call: the byte before it is not looked at */
        }
        if (keep && a->dst) memcpy(a->dst + a->l, m->b + p0, keep);
        a->l += keep; a->lastc = lastc;
    }
    return rc;
}
static inline int st_getc(const lqi_mem *m, lqi_state *s)
{
    if (s->pos < m->n) return m->b[s->pos++];
    s->zero_read = 1;
    return -1;
}

/* the sequence of a record whose first line starts at s->pos: returns the accumulated length, leaves the stream after the character
 * that ended it (*endc = that character, -1 at end of input); *n_lines = lines with content; *first = position of the first
 * content byte.  With dst != NULL the bytes are written there. */
static uint64_t walk_seq(const lqi_mem *m, lqi_state *s, uint8_t *dst, int *endc, uint32_t *n_lines, uint64_t *first)
{
    acc_t a; int c; uint32_t nl = 0;
    a.dst = dst; a.l = 0; a.lastc = 0;
    while ((c = st_getc(m, s)) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        if (nl == 0 && first) *first = s->pos - 1;
        ++nl;
        line_append(m, s, &a, 1);
    }
    *endc = c;
    if (n_lines) *n_lines = nl;
    return a.l;
}

/* the quality of a record whose first quality line starts at s->pos: lines until they cover `want` bases (kseq.h:220) */
static uint64_t walk_qual(const lqi_mem *m, lqi_state *s, uint8_t *dst, uint64_t want, uint32_t *n_calls)
{
    acc_t a; uint32_t nc = 0;
    a.dst = dst; a.l = 0; a.lastc = 0;
    for (;;) {
        if (line_append(m, s, &a, 0) < 0) break;
        ++nc;
        if (!(a.l < want)) break;
    }
    if (n_calls) *n_calls = nc;
    return a.l;
}

/* kseq_read(): one record from state s.  Returns 1 = record in *r, 0 = end of input, 2 = kseq_read() returned -2 (truncated or
 * mismatched quality): *r is a BREAK marker.  The stream goes on after a break -- what a break ends is decided by the reading loop
 * in force (see take_break()). */
#define LQI_BREAK 0x80000000u
static int scan_one(const lqi_mem *m, lqi_state *s, lqi_rec *r)
{
    int c; uint32_t nl = 0, nc = 0; uint64_t first = 0;
    if (s->failed || s->eof) return 0;
    if (s->last_char == 0) {
        /* jump to the next header character (kseq.h:191-195) */
        const uint8_t *p = m->b + (s->pos < m->n ? s->pos : m->n), *e = m->b + m->n;
        while (p < e && *p != '>' && *p != '@') ++p;
        s->pos = (uint64_t)(p - m->b);
        if (s->pos >= m->n) { s->zero_read = 1; s->eof = 1; return 0; }
        s->last_char = m->b[s->pos++];
    }
    /* name: up to the first whitespace (ks_getuntil(ks, 0, ...) returns -1 only when nothing is left and kseq knows it) */
    if (s->pos >= m->n && known_eof(m, s)) { s->eof = 1; return 0; }
    if (s->pos >= m->n) s->zero_read = 1;
    r->name_off = s->pos;
    {
        const uint8_t *p = m->b + s->pos, *e = m->b + m->n;
        while (p < e && !isspace(*p)) ++p;
        r->name_len = (uint32_t)(p - (m->b + s->pos));
        if (p < e) { c = *p; s->pos = (uint64_t)(p - m->b) + 1; }
        else { c = 0; s->pos = m->n; s->zero_read = 1; }
    }
    if (c != '\n') { acc_t a; a.dst = 0; a.l = 0; a.lastc = 0; line_append(m, s, &a, 0); }   /* comment */
    r->seq_off = s->pos;
    {
        const uint64_t l = walk_seq(m, s, 0, &c, &nl, &first);
        if (l > 0x7fffffffULL) { fprintf(stderr, "[lqcov] a read longer than 2^31 bases\n"); s->failed = 1; return 0; }
        r->len = (uint32_t)l;
    }
    r->simple = 0;
    if (nl == 1) { r->simple |= 1u; r->seq_off = first; }
    if (c == '>' || c == '@') s->last_char = c;   /* at end of input (c == -1) last_char keeps this record's header character, as in kseq */
    r->qual_off = LQI_NOQUAL;
    if (c != '+') return 1;          /* FASTA record */
    while ((c = st_getc(m, s)) != -1 && c != '\n');
    if (c == -1) { r->simple = LQI_BREAK; r->len = 0; return 2; }
    r->qual_off = s->pos;
    {
        const uint64_t ql = walk_qual(m, s, 0, r->len, &nc);
        s->last_char = 0;
        if (ql != r->len) { r->simple = LQI_BREAK; r->len = 0; return 2; }
        if (nc == 1) r->simple |= 2u;
    }
    return 1;
}

/* ---- FASTA records leave last_char set, so "no header search" is part of the state: a FASTA record that ended at end of input
 *      (c == -1) keeps the OLD last_char (non-zero), and the next kseq_read() fails in the name read.  scan_one() handles that:
 *      s->pos >= n with known_eof -> 0; with !known_eof the name read yields an empty name and an empty record ... exactly as kseq. */

/* records from state *s while the record's header lies before `stop`; the header position of the NEXT record is not known for
 * FASTQ records (search pending), so the test is made on the position where that search would start */
static void scan_range(const lqi_mem *m, lqi_state *s, uint64_t stop, rec_v *out)
{
    for (;;) {
        lqi_rec r;
        const uint64_t at = s->last_char ? s->pos - 1 : s->pos;
        if (at >= stop) break;
        if (s->last_char == 0) {     /* where is the next header?  not in this block -> leave the search to the next one */
            const uint8_t *p = m->b + s->pos, *e = m->b + (stop < m->n ? stop : m->n);
            while (p < e && *p != '>' && *p != '@') ++p;
            if (p >= e && stop < m->n) { s->pos = stop; break; }   /* equivalent state: no header character was skipped */
            s->pos = (uint64_t)(p - m->b);
        }
        if (!scan_one(m, s, &r)) break;
        recv_push(out, &r);          /* records and break markers alike */
    }
}

/* ---- speculative start of a block: a position that looks like a record header at the start of a line ---- */
static inline uint64_t line_end(const lqi_mem *m, uint64_t p)   /* position of the '\n' ending the line that contains p, or n */
{
    const uint8_t *q = p < m->n ? (const uint8_t*)memchr(m->b + p, '\n', (size_t)(m->n - p)) : 0;
    return q ? (uint64_t)(q - m->b) : m->n;
}
static int looks_like_fastq_at(const lqi_mem *m, uint64_t p)
{
    uint64_t e0, s, e1, e2, q, e3;
    if (p >= m->n || m->b[p] != '@') return 0;
    e0 = line_end(m, p); if (e0 >= m->n) return 0;
    s = e0 + 1; if (s >= m->n || m->b[s] == '@' || m->b[s] == '+' || m->b[s] == '>' || m->b[s] == '\n') return 0;
    e1 = line_end(m, s); if (e1 >= m->n) return 0;
    if (e1 + 1 >= m->n || m->b[e1 + 1] != '+') return 0;
    e2 = line_end(m, e1 + 1); if (e2 >= m->n) return 0;
    q = e2 + 1; e3 = line_end(m, q);
    if (e3 - q != e1 - s) return 0;
    if (e3 + 1 < m->n && m->b[e3 + 1] != '@') return 0;
    return 1;
}
static uint64_t guess_start(const lqi_mem *m, uint64_t from, uint64_t to, int fasta)
{
    uint64_t p = from;
    if (p > 0 && m->b[p - 1] != '\n') { p = line_end(m, p); if (p >= m->n) return m->n; ++p; }
    while (p < to && p < m->n) {
        if (fasta ? m->b[p] == '>' : looks_like_fastq_at(m, p)) return p;
        p = line_end(m, p);
        if (p >= m->n) break;
        ++p;
    }
    return m->n;
}

/* ------------------------------------------------------------------------------------------------ the reader */

typedef struct {
    int state;                       /* 0 = not started, 1 = being scanned, 2 = done */
    uint64_t guess;                  /* where the speculative scan started (n = no candidate) */
    rec_v recs; lqi_state end;
} lqi_block;

typedef struct { const lqi_rec *recs; const uint64_t *off; uint32_t lo, hi; uint8_t *seq, *qual; } copy_slice;

struct lqi_reader {
    /* memory mode */
    int fd; lqi_mem mem; int fasta; uint64_t block;
    uint64_t n_blocks; lqi_block *blk;
    uint64_t next_scan, next_stitch;  /* block cursors */
    lqi_state st;                    /* true state after the last accepted record */
    rec_v pend; size_t pend_at;      /* accepted records not yet delivered */
    /* pool */
    int n_threads; pthread_t *th; pthread_mutex_t mu; pthread_cond_t cv_work, cv_done; int quit;
    copy_slice *slices; uint32_t n_slices, next_slice, done_slices;
    /* part rule (index.c:244,316; bseq.c:86-87) */
    uint64_t batch_size, mini, mb_size, sum_len; int every_mini; uint32_t mb_n; int halted;
    /* chunk description handed out */
    uint64_t *off; size_t off_m; char *names; size_t names_m; uint64_t *name_off; size_t name_off_m;
    /* stream mode (gz, pipes) */
    lqs_reader *seq_reader; lqcov_reads_t carry; uint32_t carry_at; int carry_valid, carry_broke, stream_eof;
};

static void do_copy_slice(const lqi_reader *r, const copy_slice *s)
{
    for (uint32_t i = s->lo; i < s->hi; ++i) {
        const lqi_rec *rec = &s->recs[i];
        if (rec->len == 0) continue;
        if (rec->simple & 1u) memcpy(s->seq + s->off[i], r->mem.b + rec->seq_off, rec->len);
        else { lqi_state t; int c; memset(&t, 0, sizeof t); t.pos = rec->seq_off; walk_seq(&r->mem, &t, s->seq + s->off[i], &c, 0, 0); }
        if (s->qual) {
            if (rec->qual_off == LQI_NOQUAL) memset(s->qual + s->off[i], 0, rec->len);
            else if (rec->simple & 2u) memcpy(s->qual + s->off[i], r->mem.b + rec->qual_off, rec->len);
            else { lqi_state t; memset(&t, 0, sizeof t); t.pos = rec->qual_off; walk_qual(&r->mem, &t, s->qual + s->off[i], rec->len, 0); }
        }
    }
}

static void scan_block(lqi_reader *r, uint64_t b)
{
    lqi_block *k = &r->blk[b];
    const uint64_t lo = b * r->block, hi = lo + r->block < r->mem.n ? lo + r->block : r->mem.n;
    lqi_state s; memset(&s, 0, sizeof s);
    k->recs.n = 0;
    if (b == 0) { k->guess = 0; s.pos = 0; }
    else {
        k->guess = guess_start(&r->mem, lo, hi, r->fasta);
        s.pos = k->guess;
    }
    if (k->guess < hi || b == 0) scan_range(&r->mem, &s, hi, &k->recs);
    else s.pos = hi;
    k->end = s;
}

static void *worker(void *ud)
{
    lqi_reader *r = (lqi_reader*)ud;
    pthread_mutex_lock(&r->mu);
    for (;;) {
        if (r->quit) break;
        if (r->next_slice < r->n_slices) {
            const copy_slice s = r->slices[r->next_slice++];
            pthread_mutex_unlock(&r->mu);
            do_copy_slice(r, &s);
            pthread_mutex_lock(&r->mu);
            if (++r->done_slices == r->n_slices) pthread_cond_broadcast(&r->cv_done);
            continue;
        }
        if (r->next_scan < r->n_blocks && r->next_scan < r->next_stitch + LQI_AHEAD) {
            const uint64_t b = r->next_scan++;
            r->blk[b].state = 1;
            pthread_mutex_unlock(&r->mu);
            scan_block(r, b);
            pthread_mutex_lock(&r->mu);
            r->blk[b].state = 2;
            pthread_cond_broadcast(&r->cv_done);
            continue;
        }
        pthread_cond_wait(&r->cv_work, &r->mu);
    }
    pthread_mutex_unlock(&r->mu);
    return 0;
}

static int is_plain_regular(const char *path, int *fd_out, uint64_t *size)
{
    struct stat sb; unsigned char magic[2]; int fd;
    if (!path || strcmp(path, "-") == 0) return 0;
    fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { close(fd); return 0; }
    if (sb.st_size >= 2 && pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b) { close(fd); return 0; }
    *fd_out = fd; *size = (uint64_t)sb.st_size;
    return 1;
}

lqi_reader *lqi_open(const char *path, int n_threads)
{
    int fd = -1; uint64_t size = 0;
    const int plain = getenv("LQCOV_SEQUENTIAL_READER") ? 0 : is_plain_regular(path, &fd, &size);
    lqi_reader *r;
    if (plain < 0) return 0;
    r = (lqi_reader*)calloc(1, sizeof(*r));
    r->fd = -1;
    if (!plain) {
        r->seq_reader = lqs_open(path);
        if (!r->seq_reader) { free(r); return 0; }
        return r;
    }
    r->fd = fd; r->mem.n = size;
    if (size) {
        void *p = mmap(0, (size_t)size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) { close(fd); free(r); return 0; }
        madvise(p, (size_t)size, MADV_SEQUENTIAL);
        r->mem.b = (const uint8_t*)p;
    }
    { uint64_t i = 0; while (i < size && isspace(r->mem.b[i])) ++i; r->fasta = i < size && r->mem.b[i] == '>'; }
    r->block = LQI_BLOCK_DEFAULT;
    { const char *e = getenv("LQCOV_READER_BLOCK"); if (e && atol(e) > 0) r->block = (uint64_t)atol(e); }
    r->n_blocks = size ? (size + r->block - 1) / r->block : 0;
    r->blk = (lqi_block*)calloc(r->n_blocks ? r->n_blocks : 1, sizeof(lqi_block));
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    r->n_threads = n_threads;
    pthread_mutex_init(&r->mu, 0); pthread_cond_init(&r->cv_work, 0); pthread_cond_init(&r->cv_done, 0);
    r->th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int i = 0; i < n_threads; ++i) pthread_create(&r->th[i], 0, worker, r);
    return r;
}

void lqi_close(lqi_reader *r)
{
    if (!r) return;
    if (r->seq_reader) lqs_close(r->seq_reader);
    else {
        pthread_mutex_lock(&r->mu); r->quit = 1; pthread_cond_broadcast(&r->cv_work); pthread_mutex_unlock(&r->mu);
        for (int i = 0; i < r->n_threads; ++i) pthread_join(r->th[i], 0);
        free(r->th);
        pthread_mutex_destroy(&r->mu); pthread_cond_destroy(&r->cv_work); pthread_cond_destroy(&r->cv_done);
        for (uint64_t b = 0; b < r->n_blocks; ++b) free(r->blk[b].recs.a);
        free(r->blk); free(r->pend.a); free(r->slices);
        if (r->mem.b) munmap((void*)r->mem.b, (size_t)r->mem.n);
        if (r->fd >= 0) close(r->fd);
    }
    free(r->off); free(r->names); free(r->name_off);
    free(r);
}

/* upper bound on the bases still to come (0 = unknown: compressed or piped input) */
uint64_t lqi_bases_left_bound(const lqi_reader *r)
{
    if (r->seq_reader) return 0;
    { const uint64_t done = r->st.pos < r->mem.n ? r->st.pos : r->mem.n; const uint64_t left = r->mem.n - done; return r->fasta ? left : left / 2 + 1; }
}

void lqi_part_rule(lqi_reader *r, uint64_t batch_size, int mini_batch_size)
{
    r->batch_size = batch_size;
    r->mini = batch_size ? ((uint64_t)mini_batch_size < batch_size ? (uint64_t)mini_batch_size : batch_size) : 0;
    r->mb_size = r->sum_len = 0; r->every_mini = 0; r->mb_n = 0;
}

void lqi_batch_rule(lqi_reader *r, uint64_t chunk)
{
    r->batch_size = chunk; r->mini = chunk; r->mb_size = r->sum_len = 0; r->every_mini = 1; r->mb_n = 0;
}

/* a record was added to the current set: does the rule in force close the set? */
static inline int rule_after_record(lqi_reader *r, uint64_t len)
{
    if (!r->mini) return 0;
    r->mb_size += len; r->sum_len += len; ++r->mb_n;
    if (r->mb_size >= r->mini) {
        r->mb_size = 0; r->mb_n = 0;
        if (r->every_mini || r->sum_len > r->batch_size) { r->sum_len = 0; return 1; }
    }
    return 0;
}
/* kseq_read() < 0 in the middle of the input.  What it ends depends on who is reading:
 *   no rule (one kseq_read loop over the file: minimap2-coverage.c:418, sdust.c:198)   the input ends there
 *   mm_bseq_read batches (bseq.c:76-100)                                                the batch ends, even empty
 *   index parts (index.c:242-247)         the mini-batch ends; an empty mini-batch ends the part (mm_bseq_read returned nothing) */
static inline int take_break(lqi_reader *r)
{
    if (!r->mini) { r->halted = 1; return 1; }
    if (r->every_mini) { r->mb_size = 0; r->mb_n = 0; r->sum_len = 0; return 1; }
    if (r->mb_n > 0) {
        r->mb_size = 0; r->mb_n = 0;
        if (r->sum_len > r->batch_size) { r->sum_len = 0; return 1; }
        return 0;
    }
    r->sum_len = 0;
    return 1;
}

/* accept the next block: append its records to the pending list.  Returns 0 when the input is exhausted. */
static int stitch_next(lqi_reader *r)
{
    lqi_block *k;
    uint64_t b, lo, hi, h;
    if (r->st.failed || r->st.eof || r->next_stitch >= r->n_blocks) return 0;
    b = r->next_stitch;
    pthread_mutex_lock(&r->mu);
    pthread_cond_broadcast(&r->cv_work);
    while (r->blk[b].state != 2) {
        if (r->blk[b].state == 0 && r->next_scan <= b && r->next_scan >= r->next_stitch + LQI_AHEAD) { /* cannot happen: b == next_stitch < next_scan bound */ }
        pthread_cond_wait(&r->cv_done, &r->mu);
    }
    pthread_mutex_unlock(&r->mu);
    k = &r->blk[b];
    lo = b * r->block; hi = lo + r->block < r->mem.n ? lo + r->block : r->mem.n;
    /* where does the reference's state machine find the next header? */
    if (r->st.last_char) h = r->st.pos - 1;
    else {
        const uint8_t *p = r->mem.b + r->st.pos, *e = r->mem.b + r->mem.n;
        /* the search only needs to go as far as this block's end to decide */
        const uint8_t *lim = r->mem.b + hi;
        if (p < lim) { while (p < lim && *p != '>' && *p != '@') ++p; }
        h = p < lim ? (uint64_t)(p - r->mem.b) : (r->st.pos > hi ? r->st.pos : hi);
        (void)e;
    }
    if (b == 0 || (k->guess < hi && h == k->guess && !r->st.last_char) ) {
        /* block 0 starts from the true initial state; a later block is accepted when the pending header search ends exactly at
         * its guess (then "search from st.pos" and "search from guess" are the same state) */
        for (size_t i = 0; i < k->recs.n; ++i) recv_push(&r->pend, &k->recs.a[i]);
        { const int zr = r->st.zero_read; r->st = k->end; r->st.zero_read |= zr; }
    } else if (k->guess < hi && r->st.last_char && h == k->guess) {
        /* FASTA: the header character itself was consumed by the previous record; the block re-read it from its guess */
        for (size_t i = 0; i < k->recs.n; ++i) recv_push(&r->pend, &k->recs.a[i]);
        { const int zr = r->st.zero_read; r->st = k->end; r->st.zero_read |= zr; }
    } else if (h >= hi) {
        /* the previous record reaches past this block, or no header lies inside it: nothing starts here */
        if (!r->st.last_char && r->st.pos < hi) r->st.pos = hi;
    } else {
        /* the guess was wrong (or missing): scan this block from the true state */
        rec_v tmp; tmp.a = 0; tmp.n = tmp.m = 0;
        scan_range(&r->mem, &r->st, hi, &tmp);
        for (size_t i = 0; i < tmp.n; ++i) recv_push(&r->pend, &tmp.a[i]);
        free(tmp.a);
    }
    free(k->recs.a); k->recs.a = 0; k->recs.n = k->recs.m = 0;
    pthread_mutex_lock(&r->mu);
    ++r->next_stitch;
    pthread_cond_broadcast(&r->cv_work);
    pthread_mutex_unlock(&r->mu);
    if (r->next_stitch >= r->n_blocks && !r->st.failed) {
        /* past the last block: whatever the state machine still finds (a record whose header search was left pending) */
        rec_v tmp; tmp.a = 0; tmp.n = tmp.m = 0;
        scan_range(&r->mem, &r->st, r->mem.n + 1, &tmp);
        for (size_t i = 0; i < tmp.n; ++i) recv_push(&r->pend, &tmp.a[i]);
        free(tmp.a);
        r->st.eof = 1;
    }
    return 1;
}

static void ensure_desc(lqi_reader *r, size_t n_rec, size_t name_bytes)
{
    if (n_rec + 2 > r->off_m) { r->off_m = (n_rec + 2) * 2; r->off = (uint64_t*)xrealloc(r->off, r->off_m * 8); }
    if (n_rec + 2 > r->name_off_m) { r->name_off_m = (n_rec + 2) * 2; r->name_off = (uint64_t*)xrealloc(r->name_off, r->name_off_m * 8); }
    if (name_bytes + 1 > r->names_m) { r->names_m = (name_bytes + 1) * 2; r->names = (char*)xrealloc(r->names, r->names_m); }
}

/* stream mode: the sequential reader delivers whole mini-batches; they are re-cut into chunks here */
static int next_chunk_stream(lqi_reader *r, uint64_t cap, char *seq_dst, char *qual_dst, lqi_chunk *out)
{
    uint32_t n = 0; uint64_t nb = 0, nn = 0;
    memset(out, 0, sizeof *out);
    while (!r->halted) {
        if (!r->carry_valid) {
            if (r->stream_eof) break;
            /* one mini-batch at a time: the sequential reader stops at rejected records, whose effect depends on the rule in force */
            const int64_t want = r->mini ? (int64_t)r->mini : 50000000;
            if (lqs_next(r->seq_reader, want, &r->carry) <= 0) { r->stream_eof = 1; break; }
            r->carry_valid = 1; r->carry_at = 0; r->carry_broke = lqs_broke(r->seq_reader);
        }
        while (r->carry_at < r->carry.n) {
            const uint32_t i = r->carry_at;
            const uint64_t L = r->carry.seq_off[i + 1] - r->carry.seq_off[i], nl = r->carry.name_off[i + 1] - r->carry.name_off[i];
            if (nb + L > cap) {
                if (n == 0) { out->need = L; return -2; }
                goto done;
            }
            ensure_desc(r, n + 1, nn + nl);
            memcpy(seq_dst + nb, r->carry.seq + r->carry.seq_off[i], L);
            if (qual_dst) { if (r->carry.qual) memcpy(qual_dst + nb, r->carry.qual + r->carry.seq_off[i], L); else memset(qual_dst + nb, 0, L); }
            if (r->carry.qual) out->has_qual = 1;
            memcpy(r->names + nn, r->carry.names + r->carry.name_off[i], nl);
            r->off[n] = nb; r->name_off[n] = nn; nb += L; nn += nl; ++n; ++r->carry_at;
            if (rule_after_record(r, L)) { out->part_end = 1; goto done; }
        }
        r->carry_valid = 0;
        if (r->carry_broke) { r->carry_broke = 0; if (take_break(r)) { out->part_end = 1; goto done; } }
    }
done:
    if (r->carry_valid && r->carry_at >= r->carry.n && !r->carry_broke) r->carry_valid = 0;
    ensure_desc(r, n + 1, nn);
    r->off[n] = nb; r->name_off[n] = nn;
    out->n = n; out->n_bases = nb; out->seq_off = r->off; out->names = r->names; out->name_off = r->name_off;
    if (r->halted || (!r->carry_valid && r->stream_eof)) { out->eof = 1; if (r->mini && n) out->part_end = 1; }
    return n > 0 ? 1 : 0;
}

int lqi_next_chunk(lqi_reader *r, uint64_t cap, char *seq_dst, char *qual_dst, lqi_chunk *out)
{
    uint32_t n = 0; uint64_t nb = 0, nn = 0; size_t at;
    int part_end = 0;
    if (r->seq_reader) return next_chunk_stream(r, cap, seq_dst, qual_dst, out);
    memset(out, 0, sizeof *out);
    /* select consecutive records: until the buffer is full, the part ends or the input ends */
    at = r->pend_at;
    for (;;) {
        if (r->halted) break;
        if (at >= r->pend.n) {
            /* compact, then accept another block */
            if (r->pend_at > 0) { memmove(r->pend.a, r->pend.a + r->pend_at, (r->pend.n - r->pend_at) * sizeof(lqi_rec)); r->pend.n -= r->pend_at; at -= r->pend_at; r->pend_at = 0; }
            if (!stitch_next(r)) break;
            continue;
        }
        {
            const lqi_rec *rec = &r->pend.a[at];
            if (rec->simple & LQI_BREAK) {          /* a chunk never spans a break: the copy below indexes consecutive records */
                ++at;
                if (take_break(r)) { part_end = 1; break; }
                if (n > 0) break;
                r->pend_at = at;                     /* leading break, nothing selected yet: go on behind it */
                continue;
            }
            if (nb + rec->len > cap) {
                if (n == 0) { out->need = rec->len; r->pend_at = at; return -2; }
                break;
            }
            ensure_desc(r, n + 1, nn + rec->name_len);
            r->off[n] = nb; r->name_off[n] = nn;
            memcpy(r->names + nn, r->mem.b + rec->name_off, rec->name_len);
            if (rec->qual_off != LQI_NOQUAL) out->has_qual = 1;
            nb += rec->len; nn += rec->name_len; ++n; ++at;
            if (rule_after_record(r, rec->len)) { part_end = 1; break; }
        }
    }
    ensure_desc(r, n + 1, nn);
    r->off[n] = nb; r->name_off[n] = nn;
    /* copy the bases (and qualities) of records [pend_at, at) in parallel */
    if (n) {
        const lqi_rec *recs = r->pend.a + r->pend_at;   /* leading breaks were stepped over above: the n records start here */
        uint32_t want = (uint32_t)r->n_threads * 4, ns = 0, lo = 0;
        const uint64_t per = nb / (want ? want : 1) + 1;
        if (!r->slices) r->slices = (copy_slice*)xrealloc(0, (size_t)(64 * 4 + 8) * sizeof(copy_slice));
        while (lo < n) {
            uint32_t hi = lo; uint64_t acc = 0;
            while (hi < n && (acc < per || hi == lo)) { acc += recs[hi].len; ++hi; }
            if (ns == 64 * 4 + 7) hi = n;
            r->slices[ns].recs = recs; r->slices[ns].off = r->off; r->slices[ns].lo = lo; r->slices[ns].hi = hi;
            r->slices[ns].seq = (uint8_t*)seq_dst; r->slices[ns].qual = (uint8_t*)qual_dst;
            ++ns; lo = hi;
        }
        pthread_mutex_lock(&r->mu);
        r->n_slices = ns; r->next_slice = 0; r->done_slices = 0;
        pthread_cond_broadcast(&r->cv_work);
        while (r->next_slice < r->n_slices) {          /* the caller copies too */
            const copy_slice s = r->slices[r->next_slice++];
            pthread_mutex_unlock(&r->mu);
            do_copy_slice(r, &s);
            pthread_mutex_lock(&r->mu);
            ++r->done_slices;
        }
        while (r->done_slices < r->n_slices) pthread_cond_wait(&r->cv_done, &r->mu);
        r->n_slices = 0; r->next_slice = 0;
        pthread_mutex_unlock(&r->mu);
    }
    r->pend_at = at;
    out->n = n; out->n_bases = nb; out->seq_off = r->off; out->names = r->names; out->name_off = r->name_off;
    out->part_end = part_end;
    if (r->halted || (r->pend_at >= r->pend.n && (r->st.eof || r->st.failed || r->next_stitch >= r->n_blocks))) {
        if (!r->halted && !(r->st.eof || r->st.failed)) {   /* the last block is in: is anything left behind it? */
            while (r->pend_at >= r->pend.n && stitch_next(r)) {}
        }
        if (r->halted || r->pend_at >= r->pend.n) {
            out->eof = 1;
            if (r->mini && n) out->part_end = 1;
        }
    }
    return n > 0 ? 1 : 0;
}
