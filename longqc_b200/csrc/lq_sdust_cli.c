/* lq_sdust_cli.c -- main() of the drop-in `sdust` executable (reference sdust.c:187-223):
 *   sdust [-w 64] [-t 20] <in.fx[.gz] | ->     ->  name, masked, len, masked/len, meanQ, #Q>7  per read
 * The file is read once by the multi-threaded reader (lq_ingest.c; one kseq_read loop: the input ends at the first record kseq
 * rejects, sdust.c:198) straight into two pinned staging buffer pairs; the device works on chunk c while the reader fills chunk
 * c+1 and the host threads format the rows of chunk c-1.  Rows leave in file order. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "lqcov.h"
#include "lq_ingest.h"

#define SD_STAGE ((size_t)48 << 20)

int lqcov_sdust_main(int argc, char **argv)
{
    int W = 64, T = 20, c, rc = 0, i = 0;
    lqcov_opt_t o;
    lqi_reader *r;
    lqcov_sdust *s;
    char *sseq[2], *squal[2];
    while ((c = getopt(argc, argv, "w:t:")) >= 0) {
        if (c == 'w') W = atoi(optarg);
        else if (c == 't') T = atoi(optarg);
    }
    if (optind == argc) { fprintf(stderr, "Usage: sdust [-w %d] [-t %d] <in.fa>\n", W, T); return 1; }
    lqcov_opt_init(&o);
    {
        const char *e = getenv("LQCOV_READER_THREADS");
        long nt = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
        if (nt > 8) nt = 8;                                   /* lq_mask.py runs one sdust per input chunk from a process pool */
        r = lqi_open(argv[optind], nt < 1 ? 1 : (int)nt);
    }
    if (!r) { fprintf(stderr, "ERROR: failed to open file '%s'\n", argv[optind]); return 1; }
    lqi_batch_rule(r, 0);
    s = lqcov_sdust_begin(&o, W, T, SD_STAGE, sseq, squal);
    if (!s) { lqi_close(r); return 1; }
    for (;;) {
        lqi_chunk ch; lqcov_reads_t rd; char *rows = 0, *big_s = 0, *big_q = 0; size_t len = 0;
        int got = lqi_next_chunk(r, SD_STAGE, sseq[i & 1], squal[i & 1], &ch);
        memset(&rd, 0, sizeof rd);
        rd.seq = sseq[i & 1]; rd.qual = squal[i & 1];
        if (got == -2) {                                       /* one read longer than a staging buffer: through pageable memory */
            big_s = (char*)malloc(ch.need + 1); big_q = (char*)malloc(ch.need + 1);
            got = lqi_next_chunk(r, ch.need, big_s, big_q, &ch);
            rd.seq = big_s; rd.qual = big_q;
        }
        if (got == 1) {
            rd.n = ch.n; rd.seq_off = ch.seq_off; rd.names = ch.names; rd.name_off = ch.name_off;
            if (!ch.has_qual) rd.qual = 0;
            if (lqcov_sdust_chunk(s, &rd, &rows, &len) != 0) rc = 1;
            else { fwrite(rows, 1, len, stdout); lqcov_free(rows); }
            ++i;
        }
        if (big_s) {                                           /* the pageable copies were staged by the driver before the call returned */
            free(big_s); free(big_q);
        }
        if (rc || got != 1 || ch.eof) break;
    }
    {
        char *rows = 0; size_t len = 0;
        if (lqcov_sdust_end(s, &rows, &len) != 0) rc = 1;
        else fwrite(rows, 1, len, stdout);
        lqcov_free(rows);
    }
    fflush(stdout);
    lqi_close(r);
    return rc;
}
