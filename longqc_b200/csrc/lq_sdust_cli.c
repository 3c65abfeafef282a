/* lq_sdust_cli.c -- main() of the drop-in `sdust` executable (reference sdust.c:187-223):
 *   sdust [-w 64] [-t 20] <in.fx[.gz] | ->     ->  name, masked, len, masked/len, meanQ, #Q>7  per read */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "lqcov.h"

int lqcov_sdust_main(int argc, char **argv)
{
    int W = 64, T = 20, c, rc = 0;
    lqcov_opt_t o;
    lqcov_reader *r;
    lqcov_reads_t batch;
    while ((c = getopt(argc, argv, "w:t:")) >= 0) {
        if (c == 'w') W = atoi(optarg);
        else if (c == 't') T = atoi(optarg);
    }
    if (optind == argc) { fprintf(stderr, "Usage: sdust [-w %d] [-t %d] <in.fa>\n", W, T); return 1; }
    lqcov_opt_init(&o);
    r = lqcov_reader_open(argv[optind]);
    if (!r) { fprintf(stderr, "ERROR: failed to open file '%s'\n", argv[optind]); return 1; }
    while (rc == 0 && lqcov_reader_next(r, 500000000, &batch) > 0) { /* rows stream out batch by batch, in file order */
        char *tab = 0; size_t len = 0;
        if (lqcov_sdust_table(&o, &batch, W, T, &tab, &len) != 0) rc = 1;
        else { fwrite(tab, 1, len, stdout); lqcov_free(tab); }
    }
    lqcov_reader_close(r);
    return rc;
}
