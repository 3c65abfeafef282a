/* lq_fastx.c -- the SEQUENTIAL FASTA/FASTQ(.gz) reader (lqs_*): compressed files, pipes and stdin; plain files go through the
 * multi-threaded reader of lq_ingest.c, which calls this one for everything it cannot map.  Also the public lqcov_reader_* calls of
 * include/lqcov.h (at the end), which sit on top of lq_ingest.c.
 *
 * Record semantics follow the reference's kseq.h:185-224 (what its binaries accept):
 *   - a record starts at the next '>' or '@' found anywhere after the previous record;
 *   - name = header up to the first whitespace; the rest of the header line is a comment;
 *   - sequence lines are concatenated until a line starting with '>', '@' or '+';
 *   - after '+', quality lines are read until they cover the sequence; a length mismatch ends the input;
 *   - a trailing '\r' is dropped from every line.
 * bseq.c:61-63 turns U/u into T/t; the packer's base table maps both to 3, so bytes are kept as read.
 * Part/mini-batch boundaries follow bseq.c:82-87 and index.c:244.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <unistd.h>
#include <zlib.h>
#include "lqcov.h"
#include "lq_ingest.h"

#define RD_BUF (1 << 20)

typedef struct { size_t n, m; char *a; } blob_t;
typedef struct { size_t n, m; uint64_t *a; } offs_t;

struct lqs_reader {
    gzFile fp;
    unsigned char *buf; int beg, end, eof;
    uint64_t total; int zero_read;   /* bytes delivered so far; kseq's 16 KB reader has already seen a zero-length read */
    int last_char;       /* header character already consumed (kseq.h last_char) */
    int broke;           /* the last lqs_next() ended on a truncated/mismatched FASTQ record (kseq_read() == -2), not on its size or on EOF */
    blob_t seq, qual, names;
    offs_t seq_off, name_off;
    int any_qual;
};

static void *xrealloc(void *p, size_t n) { void *q = realloc(p, n ? n : 1); if (!q) { fprintf(stderr, "[lqcov] out of host memory\n"); abort(); } return q; }
static inline void blob_reserve(blob_t *b, size_t extra) { if (b->n + extra > b->m) { b->m = (b->n + extra) * 2 + 4096; b->a = (char*)xrealloc(b->a, b->m); } }
static inline void offs_push(offs_t *o, uint64_t v) { if (o->n == o->m) { o->m = o->m ? o->m * 2 : 1024; o->a = (uint64_t*)xrealloc(o->a, o->m * 8); } o->a[o->n++] = v; }

static inline int rd_fill(lqs_reader *r)
{
    if (r->eof) return 0;
    r->beg = 0; r->end = gzread(r->fp, r->buf, RD_BUF);
    if (r->end < RD_BUF) r->eof = 1;
    if (r->end <= 0) { r->end = 0; return 0; }
    r->total += (uint64_t)r->end;
    return 1;
}
static inline int rd_getc(lqs_reader *r)
{
    if (r->beg >= r->end && !rd_fill(r)) return -1;
    return r->buf[r->beg++];
}
/* append the rest of the current line to b (without the newline); returns -1 at EOF-with-nothing-read, else 0 */
static int rd_line(lqs_reader *r, blob_t *b, size_t line_start)
{
    int got = 0;
    for (;;) {
        unsigned char *p, *q;
        if (r->beg >= r->end && !rd_fill(r)) {
            /* kseq.h:98 returns -1 before the '\r' trim only when its stream already KNOWS it is at end of file: its 16 KB reads have
             * returned a short count, i.e. the file length is not a multiple of 16384, or a zero-length read has happened before */
            if (!got && (r->total % 16384u != 0 || r->zero_read)) return -1;
            r->zero_read = 1;
            break;
        }
        got = 1;
        p = r->buf + r->beg;
        q = (unsigned char*)memchr(p, '\n', (size_t)(r->end - r->beg));
        {
            size_t len = q ? (size_t)(q - p) : (size_t)(r->end - r->beg);
            if (b) { blob_reserve(b, len + 1); memcpy(b->a + b->n, p, len); b->n += len; }
            r->beg += (int)len + (q ? 1 : 0);
        }
        if (q) break;
    }
    /* ks_getuntil2 drops one trailing '\r' when the (accumulated) string is longer than one character (kseq.h:138) */
    if (b && b->n - line_start > 1 && b->a[b->n - 1] == '\r') --b->n;
    return got ? 0 : -1;
}

lqs_reader *lqs_open(const char *path)
{
    lqs_reader *r;
    gzFile f = (path && strcmp(path, "-")) ? gzopen(path, "r") : gzdopen(fileno(stdin), "r");
    if (!f) return NULL;
    gzbuffer(f, 1 << 20);
    r = (lqs_reader*)calloc(1, sizeof(*r));
    r->fp = f; r->buf = (unsigned char*)xrealloc(NULL, RD_BUF);
    return r;
}

void lqs_close(lqs_reader *r)
{
    if (!r) return;
    gzclose(r->fp);
    free(r->buf); free(r->seq.a); free(r->qual.a); free(r->names.a); free(r->seq_off.a); free(r->name_off.a);
    free(r);
}

/* one record appended to the blobs; returns sequence length, -1 at end of input, -2 for a record kseq_read() rejects (the stream goes on) */
static int64_t rd_record(lqs_reader *r)
{
    int c;
    size_t s0, q0;
    if (r->last_char == 0) {
        while ((c = rd_getc(r)) != -1 && c != '>' && c != '@');
        if (c == -1) return -1;
        r->last_char = c;
    }
    /* name: up to the first whitespace */
    if (r->beg >= r->end && !rd_fill(r)) return -1;
    offs_push(&r->name_off, r->names.n);
    while ((c = rd_getc(r)) != -1 && !isspace(c)) { blob_reserve(&r->names, 1); r->names.a[r->names.n++] = (char)c; }
    if (c != '\n' && c != -1) rd_line(r, NULL, 0); /* comment */
    offs_push(&r->seq_off, r->seq.n);
    s0 = r->seq.n;
    while ((c = rd_getc(r)) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        blob_reserve(&r->seq, 1); r->seq.a[r->seq.n++] = (char)c;
        rd_line(r, &r->seq, s0);
    }
    if (c == '>' || c == '@') r->last_char = c;
    /* qualities are kept aligned with the sequence blob: FASTA records get a gap that is never read */
    q0 = s0;
    if (r->qual.n < s0) { blob_reserve(&r->qual, s0 - r->qual.n); memset(r->qual.a + r->qual.n, 0, s0 - r->qual.n); r->qual.n = s0; }
    if (c != '+') {
        blob_reserve(&r->qual, r->seq.n - r->qual.n); memset(r->qual.a + r->qual.n, 0, r->seq.n - r->qual.n); r->qual.n = r->seq.n;
        return (int64_t)(r->seq.n - s0);
    }
    r->any_qual = 1;
    while ((c = rd_getc(r)) != -1 && c != '\n');
    if (c == -1) goto drop;
    while (rd_line(r, &r->qual, q0) == 0 && r->qual.n - q0 < r->seq.n - s0);
    r->last_char = 0;
    if (r->qual.n - q0 != r->seq.n - s0) goto drop;
    return (int64_t)(r->seq.n - s0);
drop: /* kseq_read() returned -2: the record is not delivered */
    r->seq.n = s0; r->qual.n = q0; r->names.n = r->name_off.a[r->name_off.n - 1];
    --r->name_off.n; --r->seq_off.n;
    return -2;
}

static void rd_reset(lqs_reader *r)
{
    r->seq.n = r->qual.n = r->names.n = 0; r->seq_off.n = r->name_off.n = 0; r->any_qual = 0;
}

static void rd_export(lqs_reader *r, lqcov_reads_t *out)
{
    offs_push(&r->seq_off, r->seq.n); offs_push(&r->name_off, r->names.n);
    --r->seq_off.n; --r->name_off.n; /* terminators are stored but not counted */
    out->n = (uint32_t)r->seq_off.n;
    out->seq = r->seq.a; out->seq_off = r->seq_off.a;
    out->qual = r->any_qual ? r->qual.a : NULL;
    out->names = r->names.a; out->name_off = r->name_off.a;
    out->seq_on_device = 0;
}

int lqs_next(lqs_reader *r, int64_t chunk, lqcov_reads_t *out)
{
    int64_t size = 0, l;
    rd_reset(r);
    r->broke = 0;
    while ((l = rd_record(r)) >= 0) {
        size += l;
        if (chunk > 0 && size >= chunk) break; /* bseq.c:86-87: the batch ends with the read that reaches the chunk size */
    }
    if (l == -2) r->broke = 1;                  /* bseq.c:76: kseq_read() < 0 ends the batch, the stream goes on with the next call */
    rd_export(r, out);
    return out->n > 0 ? 1 : (r->broke ? 2 : 0);
}
int lqs_broke(const lqs_reader *r) { return r->broke; }

/* ------------------------------------------------------------------------------------------------ public reader (include/lqcov.h) */

struct lqcov_reader {
    lqi_reader *in;
    char *seq, *qual; size_t cap;              /* blobs of the record set handed out last */
    uint64_t *seq_off, *name_off; size_t off_m;
    char *names; size_t names_m;
};

static int default_threads(void)
{
    const char *e = getenv("LQCOV_READER_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 32) n = 32;
    return (int)n;
}

lqcov_reader *lqcov_reader_open(const char *path)
{
    lqi_reader *in = lqi_open(path, default_threads());
    lqcov_reader *r;
    if (!in) return NULL;
    r = (lqcov_reader*)calloc(1, sizeof(*r));
    r->in = in;
    return r;
}

void lqcov_reader_close(lqcov_reader *r)
{
    if (!r) return;
    lqi_close(r->in);
    free(r->seq); free(r->qual); free(r->seq_off); free(r->name_off); free(r->names);
    free(r);
}

/* chunks of the current record set until the rule set on r->in closes it (or the input ends) */
static int collect(lqcov_reader *r, lqcov_reads_t *out)
{
    size_t used = 0, n = 0, nn = 0; int any_qual = 0, closed = 0, at_eof = 0;
    if (r->cap == 0) { r->cap = (size_t)64 << 20; r->seq = (char*)xrealloc(r->seq, r->cap); r->qual = (char*)xrealloc(r->qual, r->cap); }
    while (!closed) {
        lqi_chunk c;
        const int rc = lqi_next_chunk(r->in, r->cap - used, r->seq + used, r->qual + used, &c);
        if (rc == -2 || (rc > 0 && !c.part_end && !c.eof && r->cap - used - c.n_bases < ((size_t)1 << 20))) {
            const size_t need = rc == -2 ? (size_t)c.need : 0;
            r->cap = (r->cap + need) * 2;
            r->seq = (char*)xrealloc(r->seq, r->cap); r->qual = (char*)xrealloc(r->qual, r->cap);
            if (rc == -2) continue;
        }
        if (rc <= 0 && rc != -2) { at_eof = c.eof; break; }
        if (n + c.n + 2 > r->off_m) { r->off_m = (n + c.n + 2) * 2; r->seq_off = (uint64_t*)xrealloc(r->seq_off, r->off_m * 8); r->name_off = (uint64_t*)xrealloc(r->name_off, r->off_m * 8); }
        if (nn + c.name_off[c.n] + 1 > r->names_m) { r->names_m = (nn + c.name_off[c.n] + 1) * 2; r->names = (char*)xrealloc(r->names, r->names_m); }
        for (uint32_t i = 0; i < c.n; ++i) { r->seq_off[n + i] = used + c.seq_off[i]; r->name_off[n + i] = nn + c.name_off[i]; }
        memcpy(r->names + nn, c.names, c.name_off[c.n]);
        n += c.n; used += c.n_bases; nn += c.name_off[c.n];
        any_qual |= c.has_qual;
        if (c.part_end || c.eof) closed = 1;
        at_eof = c.eof;
    }
    if (r->off_m < n + 2) { r->off_m = n + 2; r->seq_off = (uint64_t*)xrealloc(r->seq_off, r->off_m * 8); r->name_off = (uint64_t*)xrealloc(r->name_off, r->off_m * 8); }
    r->seq_off[n] = used; r->name_off[n] = nn;
    out->n = (uint32_t)n; out->seq = r->seq; out->seq_off = r->seq_off; out->qual = any_qual ? r->qual : NULL;
    out->names = r->names ? r->names : ""; out->name_off = r->name_off; out->seq_on_device = 0;
    return n > 0 ? 1 : (at_eof ? 0 : 2);   /* 2: an empty set, closed by a record kseq rejects; the input goes on */
}

int lqcov_reader_next(lqcov_reader *r, int64_t chunk, lqcov_reads_t *out)
{
    /* bseq.c:86-87: the batch ends with the read that reaches the chunk size */
    lqi_batch_rule(r->in, chunk > 0 ? (uint64_t)chunk : 0);
    return collect(r, out);
}

int lqcov_reader_next_part(lqcov_reader *r, uint64_t batch_size, int mini_batch_size, lqcov_reads_t *out)
{
    /* index.c:238-246,316: mini = min(mini_batch_size, batch_size); keep reading mini-batches while sum_len <= batch_size */
    lqi_part_rule(r->in, batch_size, mini_batch_size);
    return collect(r, out);
}
