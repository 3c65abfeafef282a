/* lq_fastx.c -- FASTA/FASTQ(.gz) reader producing lqcov_reads_t blobs.
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

#define RD_BUF (1 << 20)

typedef struct { size_t n, m; char *a; } blob_t;
typedef struct { size_t n, m; uint64_t *a; } offs_t;

struct lqcov_reader {
    gzFile fp;
    unsigned char *buf; int beg, end, eof;
    uint64_t total; int zero_read;   /* bytes delivered so far; kseq's 16 KB reader has already seen a zero-length read */
    int last_char;       /* header character already consumed (kseq.h last_char) */
    int failed;          /* a truncated/mismatched FASTQ record ends the stream, like kseq_read() < 0 */
    blob_t seq, qual, names;
    offs_t seq_off, name_off;
    int any_qual;
};

static void *xrealloc(void *p, size_t n) { void *q = realloc(p, n ? n : 1); if (!q) { fprintf(stderr, "[lqcov] out of host memory\n"); abort(); } return q; }
static inline void blob_reserve(blob_t *b, size_t extra) { if (b->n + extra > b->m) { b->m = (b->n + extra) * 2 + 4096; b->a = (char*)xrealloc(b->a, b->m); } }
static inline void offs_push(offs_t *o, uint64_t v) { if (o->n == o->m) { o->m = o->m ? o->m * 2 : 1024; o->a = (uint64_t*)xrealloc(o->a, o->m * 8); } o->a[o->n++] = v; }

static inline int rd_fill(lqcov_reader *r)
{
    if (r->eof) return 0;
    r->beg = 0; r->end = gzread(r->fp, r->buf, RD_BUF);
    if (r->end < RD_BUF) r->eof = 1;
    if (r->end <= 0) { r->end = 0; return 0; }
    r->total += (uint64_t)r->end;
    return 1;
}
static inline int rd_getc(lqcov_reader *r)
{
    if (r->beg >= r->end && !rd_fill(r)) return -1;
    return r->buf[r->beg++];
}
/* append the rest of the current line to b (without the newline); returns -1 at EOF-with-nothing-read, else 0 */
static int rd_line(lqcov_reader *r, blob_t *b, size_t line_start)
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

lqcov_reader *lqcov_reader_open(const char *path)
{
    lqcov_reader *r;
    gzFile f = (path && strcmp(path, "-")) ? gzopen(path, "r") : gzdopen(fileno(stdin), "r");
    if (!f) return NULL;
    gzbuffer(f, 1 << 20);
    r = (lqcov_reader*)calloc(1, sizeof(*r));
    r->fp = f; r->buf = (unsigned char*)xrealloc(NULL, RD_BUF);
    return r;
}

void lqcov_reader_close(lqcov_reader *r)
{
    if (!r) return;
    gzclose(r->fp);
    free(r->buf); free(r->seq.a); free(r->qual.a); free(r->names.a); free(r->seq_off.a); free(r->name_off.a);
    free(r);
}

/* one record appended to the blobs; returns sequence length, -1 at end of input */
static int64_t rd_record(lqcov_reader *r)
{
    int c;
    size_t s0, q0;
    if (r->failed) return -1;
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
    if (c == -1) { r->failed = 1; goto drop; }
    while (rd_line(r, &r->qual, q0) == 0 && r->qual.n - q0 < r->seq.n - s0);
    r->last_char = 0;
    if (r->qual.n - q0 != r->seq.n - s0) { r->failed = 1; goto drop; }
    return (int64_t)(r->seq.n - s0);
drop: /* kseq_read() returned -2: the record is not delivered */
    r->seq.n = s0; r->qual.n = q0; r->names.n = r->name_off.a[r->name_off.n - 1];
    --r->name_off.n; --r->seq_off.n;
    return -1;
}

static void rd_reset(lqcov_reader *r)
{
    r->seq.n = r->qual.n = r->names.n = 0; r->seq_off.n = r->name_off.n = 0; r->any_qual = 0;
}

static void rd_export(lqcov_reader *r, lqcov_reads_t *out)
{
    offs_push(&r->seq_off, r->seq.n); offs_push(&r->name_off, r->names.n);
    --r->seq_off.n; --r->name_off.n; /* terminators are stored but not counted */
    out->n = (uint32_t)r->seq_off.n;
    out->seq = r->seq.a; out->seq_off = r->seq_off.a;
    out->qual = r->any_qual ? r->qual.a : NULL;
    out->names = r->names.a; out->name_off = r->name_off.a;
    out->seq_on_device = 0;
}

int lqcov_reader_next(lqcov_reader *r, int64_t chunk, lqcov_reads_t *out)
{
    int64_t size = 0, l;
    rd_reset(r);
    while ((l = rd_record(r)) >= 0) {
        size += l;
        if (chunk > 0 && size >= chunk) break; /* bseq.c:86-87: the batch ends with the read that reaches the chunk size */
    }
    rd_export(r, out);
    return out->n > 0 ? 1 : 0;
}

int lqcov_reader_next_part(lqcov_reader *r, uint64_t batch_size, int mini_batch_size, lqcov_reads_t *out)
{
    /* index.c:238-246,316: mini = min(mini_batch_size, batch_size); keep reading mini-batches while sum_len <= batch_size */
    const uint64_t mini = (uint64_t)mini_batch_size < batch_size ? (uint64_t)mini_batch_size : batch_size;
    uint64_t sum_len = 0;
    rd_reset(r);
    while (!(sum_len > batch_size)) {
        uint64_t size = 0; int64_t l; int got = 0;
        while ((l = rd_record(r)) >= 0) { got = 1; size += (uint64_t)l; sum_len += (uint64_t)l; if (size >= mini) break; }
        if (!got || l < 0) break;
    }
    rd_export(r, out);
    return out->n > 0 ? 1 : 0;
}
