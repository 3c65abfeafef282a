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
#include <pthread.h>
#include "lqcov.h"
#include "lq_ingest.h"

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

static double wall(void);
static double g_t0;
#define TL(what) fprintf(stderr, "[T::%.3f] %s\n", wall() - g_t0, (what))
static double wall(void) { struct timeval t; gettimeofday(&t, 0); return t.tv_sec + t.tv_usec * 1e-6; }
static double cpu(void) { struct rusage r; getrusage(RUSAGE_SELF, &r); return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec); }

/* ---- the part loop: index parts arrive in chunks of consecutive reads (lq_ingest.c), the device copies, packs and sketches chunk i
 *      while the reader threads fill chunk i+1 (replaces kt_pipeline's read -> sketch -> dispatch stages, index.c:238-309) ---- */
#define CLI_NSTAGE 4
#define CLI_STAGE_BYTES ((size_t)16 << 20)

struct init_job { const lqcov_opt_t *o; lqcov_ctx *c; char *stage[CLI_NSTAGE]; int rc; };
static void *init_thread(void *p)       /* the CUDA context: hundreds of ms, during which the reader threads already parse */
{
    struct init_job *j = (struct init_job*)p;
    j->c = lqcov_create(j->o);
    TL("CUDA context + stream created");
    return 0;
}
static void *stage_thread(void *p)      /* page-locking the staging buffers runs beside the query sketch */
{
    struct init_job *j = (struct init_job*)p;
    j->rc = lqcov_stage(j->c, CLI_NSTAGE, CLI_STAGE_BYTES, j->stage);
    TL("pinned staging buffers allocated");
    return 0;
}

static int reader_threads(int t_opt)
{
    const char *e = getenv("LQCOV_READER_THREADS");
    int n = e ? atoi(e) : t_opt;
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    return n;
}

typedef struct { uint64_t *seq_off, *name_off; char *names; size_t n, m, nn, nm; } part_meta;
static void meta_add(part_meta *pm, const lqi_chunk *ch)
{
    if (pm->n + ch->n + 2 > pm->m) { pm->m = (pm->n + ch->n + 2) * 2; pm->seq_off = (uint64_t*)realloc(pm->seq_off, pm->m * 8); pm->name_off = (uint64_t*)realloc(pm->name_off, pm->m * 8); }
    if (pm->nn + ch->name_off[ch->n] + 1 > pm->nm) { pm->nm = (pm->nn + ch->name_off[ch->n] + 1) * 2; pm->names = (char*)realloc(pm->names, pm->nm); }
    if (pm->n == 0) { pm->seq_off[0] = 0; pm->name_off[0] = 0; }
    for (uint32_t i = 0; i < ch->n; ++i) {
        pm->seq_off[pm->n + i + 1] = pm->seq_off[pm->n] + ch->seq_off[i + 1];
        pm->name_off[pm->n + i + 1] = pm->nn + ch->name_off[i + 1];
    }
    memcpy(pm->names + pm->nn, ch->names, ch->name_off[ch->n]);
    pm->n += ch->n; pm->nn += ch->name_off[ch->n];
}

static int run_parts(lqcov_ctx *c, lqi_reader *tr, const lqcov_opt_t *o, char **stage, double t0)
{
    part_meta pm; memset(&pm, 0, sizeof pm);
    int rc = 0, eof = 0;
    lqi_part_rule(tr, o->batch_size, o->mini_batch_size);
    while (rc == 0 && !eof) {
        uint64_t expect = lqi_bases_left_bound(tr);           /* first sizing of the device arrays; they grow if the part turns out larger */
        if (expect > o->batch_size + 2 * (uint64_t)o->mini_batch_size) expect = o->batch_size + 2 * (uint64_t)o->mini_batch_size;
        const int streamed = lqcov_part_begin(c, expect, 0);
        TL("part begun (device arrays sized)");  /* 1: no chunked form for this configuration (-H): whole part at once */
        char *whole = 0; size_t whole_cap = 0, whole_n = 0;
        int i = 0, part_end = 0;
        if (streamed < 0) { rc = 1; break; }
        pm.n = 0; pm.nn = 0;
        while (rc == 0 && !part_end && !eof) {
            lqi_chunk ch; int r;
            char *dst; size_t cap;
            if (streamed == 0) { lqcov_stage_wait(c, i % CLI_NSTAGE); dst = stage[i % CLI_NSTAGE]; cap = CLI_STAGE_BYTES; }
            else {
                if (whole_cap - whole_n < CLI_STAGE_BYTES) { whole_cap = whole_cap ? whole_cap * 2 : 4 * CLI_STAGE_BYTES; whole = (char*)realloc(whole, whole_cap); }
                dst = whole + whole_n; cap = whole_cap - whole_n;
            }
            r = lqi_next_chunk(tr, cap, dst, 0, &ch);
            if (r == -2) {                                    /* a read longer than a staging buffer */
                if (streamed != 0) { whole_cap = whole_n + ch.need + CLI_STAGE_BYTES; whole = (char*)realloc(whole, whole_cap); continue; }
                /* through pageable memory: cudaMemcpyAsync returns once a pageable source has been staged, so `big` can be freed at once */
                char *big = (char*)malloc(ch.need + 1);
                r = lqi_next_chunk(tr, ch.need, big, 0, &ch);
                if (r == 1) {
                    lqcov_reads_t cr; memset(&cr, 0, sizeof cr);
                    cr.n = ch.n; cr.seq = big; cr.seq_off = ch.seq_off; cr.names = ch.names; cr.name_off = ch.name_off;
                    meta_add(&pm, &ch);
                    if (lqcov_part_chunk(c, &cr, -1) != 0) rc = 1;
                }
                free(big);
                part_end = ch.part_end; eof = ch.eof;
                continue;
            }
            part_end = ch.part_end; eof = ch.eof;
            if (r <= 0) continue;
            meta_add(&pm, &ch);
            if (streamed == 0) {
                lqcov_reads_t cr; memset(&cr, 0, sizeof cr);
                cr.n = ch.n; cr.seq = dst; cr.seq_off = ch.seq_off; cr.names = ch.names; cr.name_off = ch.name_off;
                if (lqcov_part_chunk(c, &cr, i % CLI_NSTAGE) != 0) rc = 1;
                ++i;
            } else whole_n += ch.n_bases;
        }
        if (rc == 0 && pm.n > 0) {
            lqcov_reads_t part; memset(&part, 0, sizeof part);
            part.n = (uint32_t)pm.n; part.seq_off = pm.seq_off; part.names = pm.names; part.name_off = pm.name_off;
            if (streamed == 0) {
                TL("part parsed, all chunks queued");
                if (lqcov_part_end(c) != 0) rc = 1;
                TL("part sketched + counted");
                if (rc == 0 && lqcov_part_finish(c, &part) != 0) rc = 1;
                fprintf(stderr, "[M::%s::%.3f*%.2f] indexed %u target sequence(s)\n", __func__, wall() - t0, cpu() / (wall() - t0), part.n);
                if (rc == 0 && lqcov_map_part(c) != 0) rc = 1;
            } else {
                part.seq = whole;
                if (lqcov_add_part(c, &part) != 0) rc = 1;
            }
            fprintf(stderr, "[M::%s::%.3f*%.2f] mapped against part of %u sequence(s)\n", __func__, wall() - t0, cpu() / (wall() - t0), part.n);
        } else if (rc == 0 && streamed == 0) {
            if (lqcov_part_end(c) != 0) rc = 1;               /* an empty part (a rejected record closed it): nothing to map against */
        }
        free(whole);
    }
    free(pm.seq_off); free(pm.name_off); free(pm.names);
    return rc;
}

int lqcov_main(int argc, char **argv)
{
    struct cli a;
    lqcov_opt_t o;
    const double t0 = wall();
    g_t0 = t0;
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

    /* The CUDA context (a few hundred ms) is created on a second thread while the reader threads already parse the inputs. */
    struct init_job ij; pthread_t ith;
    memset(&ij, 0, sizeof ij); ij.o = &o;
    pthread_create(&ith, 0, init_thread, &ij);
    const int n_rd = reader_threads(o.n_threads);
    lqi_reader *tr = lqi_open(a.args[0], n_rd);
    lqcov_reader *qr = tr ? lqcov_reader_open(a.args[1]) : 0;
    lqcov_reads_t q; memset(&q, 0, sizeof q);
    if (qr) lqcov_reader_next(qr, 0, &q);                     /* one kseq_read loop over the query file (minimap2-coverage.c:418) */
    TL("queries parsed");
    pthread_join(ith, 0);
    lqcov_ctx *c = ij.c;
    if (!tr) fprintf(stderr, "ERROR: failed to open file '%s'\n", a.args[0]);
    else if (!qr) fprintf(stderr, "ERROR: failed to open file '%s'\n", a.args[1]);
    if (!tr || !qr || !c) { if (tr) lqi_close(tr); if (qr) lqcov_reader_close(qr); if (c) lqcov_destroy(c); return 1; }
    int rc = 0;
    pthread_create(&ith, 0, stage_thread, &ij);
    if (lqcov_set_queries(c, &q) != 0) rc = 1;
    pthread_join(ith, 0);
    if (ij.rc != 0) rc = 1;
    TL("queries sketched");
    fprintf(stderr, "[M::%s::%.3f*%.2f] loaded %u sequence(s).\n", __func__, wall() - t0, cpu() / (wall() - t0), q.n);
    lqcov_reader_close(qr);
    if (rc == 0) rc = run_parts(c, tr, &o, ij.stage, t0);
    lqi_close(tr);
    if (rc == 0) {
        char *tab = 0; size_t len = 0;
        if (lqcov_table(c, &tab, &len) != 0) rc = 1;
        else { fwrite(tab, 1, len, stdout); fflush(stdout); lqcov_free(tab); }
    }
    TL("table written");
    if (!getenv("LQCOV_FAST_EXIT")) lqcov_destroy(c);
    fprintf(stderr, "\n[M::%s] Real time: %.3f sec; CPU: %.3f sec\n", __func__, wall() - t0, cpu());
    return rc;
}
