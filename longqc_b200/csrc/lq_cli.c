/* lq_cli.c -- main() of the drop-in `minimap2-coverage` executable.
 *
 * Command line == the reference's (minimap2-coverage.c:63-197): same option letters, long names,
 * argument kinds, "0 / -1 means default" sentinels (:252-388), the same fatal checks, stdout carries
 * only the table, everything else goes to stderr.  Not supported by this build: -d (index dump).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <argp.h>
#include <inttypes.h>
#include <sys/time.h>
#include <sys/resource.h>
#include "lqcov.h"

const char *argp_program_version = "minimap2-coverage 0.3 (lqcov-b200; CLI of LongQC's fork of minimap2 2.6-r639)";
const char *argp_program_bug_address = "<lqcov-b200 maintainers>";

struct cli {
    int h_flag, ava, avs, filter, zflag;
    int k, w, min_cov, n_subset, max_gap, min_cnt, min_score, score_med, score_good, chain_skip, max_ohang, min_ovlp, threads;
    double min_ratio;
    uint64_t batch;
    char *args[2], *dump;
};

static int64_t parse_num(const char *s) /* K/M/G suffixes, minimap2-coverage.c:22-31 */
{
    char *e; double x = strtod(s, &e);
    if (*e == 'G' || *e == 'g') x *= 1e9; else if (*e == 'M' || *e == 'm') x *= 1e6; else if (*e == 'K' || *e == 'k') x *= 1e3;
    return (int64_t)(x + .499);
}

static error_t on_opt(int key, char *arg, struct argp_state *st)
{
    struct cli *a = (struct cli*)st->input;
    switch (key) {
    case 'H': a->h_flag = 1; break;
    case 'k': a->k = atoi(arg); break;
    case 'w': a->w = atoi(arg); break;
    case 'I': a->batch = (uint64_t)parse_num(arg); break;
    case 'd': a->dump = arg; break;
    case 'g': a->max_gap = atoi(arg); break;
    case 'n': a->min_cnt = atoi(arg); break;
    case 'm': a->min_score = atoi(arg); break;
    case 'p': a->score_med = atoi(arg); break;
    case 'q': a->score_good = atoi(arg); break;
    case 's': a->chain_skip = atoi(arg); break;
    case 'X': a->ava = 1; break;
    case 'Y': a->avs = 1; break;
    case 'a': a->max_ohang = atoi(arg); break;
    case 'l': a->min_ovlp = atoi(arg); break;
    case 'c': a->min_cov = atoi(arg); break;
    case 'r': a->min_ratio = atof(arg); break;
    case 'f': a->filter = 1; break;
    case 'z': a->zflag = 1; break;
    case 'u': a->n_subset = atoi(arg); break;
    case 't': a->threads = atoi(arg); break;
    case ARGP_KEY_ARG:
        if (st->arg_num >= 2) argp_usage(st);
        a->args[st->arg_num] = arg;
        break;
    case ARGP_KEY_INIT:
        memset(a, 0, sizeof(*a));
        a->threads = 1; a->min_cov = a->n_subset = -1; a->chain_skip = a->max_ohang = a->min_ovlp = -1;
        break;
    case ARGP_KEY_END:
        if (a->dump == 0 && st->arg_num < 2) argp_usage(st);
        break;
    default: return ARGP_ERR_UNKNOWN;
    }
    return 0;
}

static struct argp_option opts[] = {
    { 0, 0, 0, 0, "Indexing options:", 1 },
    { "homopolymer",       'H', 0,        0, "sketch homopolymer-compressed k-mers" },
    { "k-mer",             'k', "INT",    0, "k-mer size" },
    { "window",            'w', "INT",    0, "minimizer window size" },
    { "index-size",        'I', "STRING", 0, "start a new index part every ~NUM target bases (K/M/G suffix)" },
    { "dump-index",        'd', "FILE",   0, "dump the index to FILE (not supported by this build)" },
    { 0, 0, 0, 0, "Mapping options", 2 },
    { "max-gap-length",    'g', "INT",    0, "maximum distance between chained minimizers" },
    { "min-cnt",           'n', "INT",    0, "minimum number of minimizers on a chain" },
    { "min-score",         'm', "INT",    0, "minimum chaining score" },
    { "min-score-t2",      'p', "INT",    0, "medium chaining-score threshold, >= min-score" },
    { "min-score-t3",      'q', "INT",    0, "good chaining-score threshold, >= min-score-t2" },
    { "max-chain-skip",    's', "INT",    OPTION_HIDDEN, "as in minimap2" },
    { "skip-self-ava",     'X', 0,        0, "skip self and dual mappings (all-vs-all)" },
    { "skip-self",         'Y', 0,        0, "skip self mappings (all-vs-subsample)" },
    { 0, 0, 0, 0, "Filtering options", 3 },
    { "max-overhang",      'a', "INT",    0, "maximum overhang of an accepted overlap" },
    { "min-overlap-len",   'l', "INT",    0, "minimum overlap length" },
    { "min-coverage",      'c', "INT",    0, "coverage needed for a reliable region" },
    { "min-overlap-ratio", 'r', "NUM",    0, "minimum overlap / (overlap + overhang) ratio" },
    { 0, 0, 0, 0, "Misc options", 4 },
    { "num-subset",        'u', "INT",    0, "number of query sequences (informational)" },
    { "threads",           't', "INT",    0, "number of host threads" },
    { "minimizer-cnt",     'z', 0,        0, "accepted for compatibility; no effect on the output" },
    { "filter",            'f', 0,        0, "read filtering mode (spike-in control)" },
    { 0 }
};

static struct argp the_argp = { opts, on_opt, "reference reads",
    "minimap2-coverage computes, for every query read, the overlaps with all target reads and prints a coverage table "
    "(LongQC's overlap/coverage pass); this build runs the pass on an NVIDIA B200.\v " };

static double wall(void) { struct timeval t; gettimeofday(&t, 0); return t.tv_sec + t.tv_usec * 1e-6; }
static double cpu(void) { struct rusage r; getrusage(RUSAGE_SELF, &r); return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec); }

int lqcov_main(int argc, char **argv)
{
    struct cli a;
    lqcov_opt_t o;
    const double t0 = wall();
    argp_parse(&the_argp, argc, argv, 0, 0, &a);
    lqcov_opt_init(&o);
    o.verbose = 3;

    if (a.ava && a.avs) { fprintf(stderr, "Error: -X and -Y are mutually exclusive\n"); return 1; }
    if (!a.ava && !a.avs && !a.dump) { fprintf(stderr, "Error: Choose either -X (all-vs-all) or -Y (all-vs-sub)\n"); return 1; }
    o.no_self = 1; o.ava = a.ava ? 1 : 0;
    if (a.h_flag) o.is_hpc = 1; else fprintf(stderr, "Homopolymer compression is not applied.\n");
    if (a.k == 0) { fprintf(stderr, "Warning: Apply default k=12 instead. \n"); o.k = 12; } else o.k = a.k;
    if (a.w == 0) { fprintf(stderr, "Warning: Apply default w=5 instead. \n"); o.w = 5; } else o.w = a.w;
    if (a.batch == 0) fprintf(stderr, "Warning: Apply default I=4G instead. \n"); else o.batch_size = a.batch;
    if (a.dump) { fprintf(stderr, "ERROR: -d (index dump) is not supported by this build\n"); return 1; }
    if (a.min_cov == -1) { fprintf(stderr, "Warning: Apply default c=3 instead. \n"); o.min_coverage = 3; } else o.min_coverage = a.min_cov;
    if (a.n_subset == -1) { fprintf(stderr, "Warning: -s shouldn't be zero. Apply default s=100000 instead.\n"); a.n_subset = 100000; }
    if (a.max_gap == 0) { fprintf(stderr, "Warning: Apply default g=10000 instead.\n"); o.max_gap = 10000; } else o.max_gap = a.max_gap;
    if (a.min_cnt == 0) { fprintf(stderr, "Warning: Apply default n=3 instead.\n"); o.min_cnt = 3; } else o.min_cnt = a.min_cnt;
    if (a.min_score == 0) { fprintf(stderr, "Warning: Apply default m=40 instead.\n"); o.min_chain_score = 40; } else o.min_chain_score = a.min_score;
    if (a.score_med == 0) { fprintf(stderr, "Warning: Apply default p=m instead.\n"); a.score_med = o.min_chain_score; }
    if (a.score_good == 0) { fprintf(stderr, "Warning: Apply default q=m instead.\n"); a.score_good = o.min_chain_score; }
    if (a.score_med < o.min_chain_score) { fprintf(stderr, "Error: -p must be larger than or equal to -m.\n"); return 1; }
    if (a.score_good < o.min_chain_score || a.score_good < a.score_med) { fprintf(stderr, "Error: -q must be larger than or equal to -m and -p.\n"); return 1; }
    o.min_score_med = a.score_med; o.min_score_good = a.score_good;
    if (a.chain_skip == -1) { fprintf(stderr, "Warning: Apply default s=25 instead.\n"); o.max_chain_skip = 25; } else o.max_chain_skip = a.chain_skip;
    if (a.max_ohang == -1) { fprintf(stderr, "Warning: Apply default a=2000 instead.\n"); o.max_overhang = 2000; } else o.max_overhang = a.max_ohang;
    if (a.min_ovlp == -1) { fprintf(stderr, "Warning: Apply default l=1000 instead.\n"); o.min_ovlp = 1000; } else o.min_ovlp = a.min_ovlp;
    if (a.min_ratio == 0.0) { fprintf(stderr, "Warning: Apply default r=0.4 instead.\n"); o.min_ratio = 0.4; } else o.min_ratio = a.min_ratio;
    o.filter = a.filter; o.n_threads = a.threads > 1 ? a.threads : 1;

    fprintf(stderr, "=== Parameters are listed below === \n");
    fprintf(stderr, "Inputs are target: %s, query: %s\n", a.args[0], a.args[1]);
    fprintf(stderr, "kmer %d, window %d, index loading size %" PRIu64 "\n", o.k, o.w, o.batch_size);
    fprintf(stderr, "min-score %d, min-score-med %d, min-score-good %d, max-gap %d, min-cnt %d\n", o.min_chain_score, o.min_score_med, o.min_score_good, o.max_gap, o.min_cnt);
    fprintf(stderr, "Homo-polymer compression: %d, Filtering: %d, minimizer-count: %d\n", a.h_flag, a.filter, a.zflag);
    fprintf(stderr, "max-overhang %d, min-overlaplen %d, min-overapratio %.2f\n", o.max_overhang, o.min_ovlp, o.min_ratio);
    fprintf(stderr, "num of threads %d, num of query seqs %d\n===\n", o.n_threads, a.n_subset);

    lqcov_reader *tr = lqcov_reader_open(a.args[0]);
    if (!tr) { fprintf(stderr, "ERROR: failed to open file '%s'\n", a.args[0]); return 1; }
    lqcov_reader *qr = lqcov_reader_open(a.args[1]);
    if (!qr) { fprintf(stderr, "ERROR: failed to open file '%s'\n", a.args[1]); lqcov_reader_close(tr); return 1; }
    lqcov_ctx *c = lqcov_create(&o);
    if (!c) { lqcov_reader_close(tr); lqcov_reader_close(qr); return 1; }
    lqcov_reads_t q, part;
    int rc = 0;
    lqcov_reader_next(qr, 0, &q);
    if (lqcov_set_queries(c, &q) != 0) rc = 1;
    lqcov_reader_close(qr);
    fprintf(stderr, "[M::%s::%.3f*%.2f] loaded %u sequence(s).\n", __func__, wall() - t0, cpu() / (wall() - t0), q.n);
    while (rc == 0 && lqcov_reader_next_part(tr, o.batch_size, o.mini_batch_size, &part) > 0)
        if (lqcov_add_part(c, &part) != 0) rc = 1;
    lqcov_reader_close(tr);
    if (rc == 0) {
        char *tab = 0; size_t len = 0;
        if (lqcov_table(c, &tab, &len) != 0) rc = 1;
        else { fwrite(tab, 1, len, stdout); fflush(stdout); lqcov_free(tab); }
    }
    lqcov_destroy(c);
    fprintf(stderr, "\n[M::%s] Real time: %.3f sec; CPU: %.3f sec\n", __func__, wall() - t0, cpu());
    return rc;
}
