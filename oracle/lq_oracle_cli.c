/* lq_oracle_cli.c -- command-line front end of the CPU restatement (TEST INFRASTRUCTURE ONLY).
 *   lq_oracle_cli cov   [minimap2-coverage flags] <target.fx[.gz]> <query.fx[.gz]>   -> stdout table
 *   lq_oracle_cli sdust <in.fx[.gz]>                                                  -> stdout table
 * Flag letters follow minimap2-coverage.c:166-195 so a reference argv can be replayed unchanged. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "lq_oracle.h"

static uint64_t parse_num(const char *s) /* minimap2-coverage.c:22-31 */
{
    char *p; double x = strtod(s, &p);
    if (*p == 'G' || *p == 'g') x *= 1e9; else if (*p == 'M' || *p == 'm') x *= 1e6; else if (*p == 'K' || *p == 'k') x *= 1e3;
    return (uint64_t)(int64_t)(x + .499);
}

int main(int argc, char **argv)
{
    if (argc >= 3 && strcmp(argv[1], "sdust") == 0) {
        lqo_reads r;
        if (lqo_reads_load(argv[2], &r) != 0) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
        lqo_sdust_run(&r, 64, 20, stdout);
        lqo_reads_free(&r);
        return 0;
    }
    if (argc >= 2 && strcmp(argv[1], "cov") == 0) {
        lqo_opt o; lqo_reads t, q; int c, p_set = 0, q_set = 0, mid, parts, x = 0, y = 0;
        lqo_opt_default(&o);
        optind = 2;
        while ((c = getopt(argc, argv, "Hk:w:I:g:n:m:p:q:s:XYa:l:c:r:u:t:zf")) >= 0) {
            switch (c) {
            case 'H': o.is_hpc = 1; break;
            case 'k': o.k = atoi(optarg); break;
            case 'w': o.w = atoi(optarg); break;
            case 'I': o.batch_size = parse_num(optarg); break;
            case 'g': o.max_gap = atoi(optarg); break;
            case 'n': o.min_cnt = atoi(optarg); break;
            case 'm': o.min_chain_score = atoi(optarg); break;
            case 'p': o.min_score_med = atoi(optarg); p_set = 1; break;
            case 'q': o.min_score_good = atoi(optarg); q_set = 1; break;
            case 's': o.max_chain_skip = atoi(optarg); break;
            case 'X': x = 1; break;
            case 'Y': y = 1; break;
            case 'a': o.max_overhang = atoi(optarg); break;
            case 'l': o.min_ovlp = atoi(optarg); break;
            case 'c': o.min_coverage = atoi(optarg); break;
            case 'r': o.min_ratio = atof(optarg); break;
            case 'f': o.filter = 1; break;
            default: break;
            }
        }
        if (!p_set || o.min_score_med == 0) o.min_score_med = o.min_chain_score;
        if (!q_set || o.min_score_good == 0) o.min_score_good = o.min_chain_score;
        o.no_self = 1; o.ava = x && !y;
        if (argc - optind < 2) { fprintf(stderr, "usage: lq_oracle_cli cov [flags] target query\n"); return 1; }
        if (lqo_reads_load(argv[optind], &t) != 0 || lqo_reads_load(argv[optind + 1], &q) != 0) { fprintf(stderr, "cannot open input\n"); return 1; }
        lqo_run(&o, &t, &q, stdout, &mid, &parts);
        fprintf(stderr, "[oracle] mid_occ=%d parts=%d targets=%d queries=%d\n", mid, parts, t.n, q.n);
        lqo_reads_free(&t); lqo_reads_free(&q);
        return 0;
    }
    fprintf(stderr, "usage: lq_oracle_cli cov|sdust ...\n");
    return 1;
}
