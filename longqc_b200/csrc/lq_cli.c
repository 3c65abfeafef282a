/* lq_cli.c -- main() of the drop-in `minimap2-coverage` executable.
 *
 * Command line == the reference's (minimap2-coverage.c:63-197): same option letters, long names,
 * argument kinds, "0 / -1 means default" sentinels (:252-388), the same fatal checks, stdout carries
 * only the table, everything else goes to stderr.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <argp.h>
#include <inttypes.h>
#include <sys/time.h>
#include <sys/resource.h>
#include <pthread.h>
#include <unistd.h>
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
    { "dump-index",        'd', "FILE",   0, "dump the index to FILE (the reference's MMI format)" },
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

/* ---- the part loop: index parts arrive in chunks of consecutive reads (lq_ingest.c); the device copies, packs and sketches chunk i
 *      while the reader threads fill chunk i+1 (replaces kt_pipeline's read -> sketch -> dispatch stages, index.c:238-309).
 *      With several visible GPUs the executable drives them all from this one process (one context and one host thread per GPU,
 *      NCCL communicator inside the library): a part's reads go to the GPUs in contiguous ranges, the queries are split, every GPU
 *      maps its share against the replicated index and the rows are concatenated in GPU order. ---- */
#define CLI_NSTAGE 4
#define CLI_STAGE_BYTES ((size_t)16 << 20)
#define CLI_MAX_DEV 16

typedef struct {
    int dev, n_dev, rc;
    lqcov_opt_t o; lqcov_ctx *c; char *stage[CLI_NSTAGE];
    lqcov_reads_t q;                     /* this GPU's share of the queries */
    const lqcov_reads_t *part;           /* whole-part metadata for the finish step */
    int chunks, begun;
    char *tab; size_t tab_len;
    double t0;
} cli_dev;

static void *dev_create(void *p)         /* the CUDA context: hundreds of ms, during which the reader threads already parse */
{
    cli_dev *d = (cli_dev*)p;
    d->c = lqcov_create(&d->o);
    d->rc = d->c ? 0 : 1;
    TL("CUDA context + stream created");
    return 0;
}
static void *dev_stage(void *p)          /* page-locking the staging buffers runs beside the query sketch */
{
    cli_dev *d = (cli_dev*)p;
    if (lqcov_stage(d->c, CLI_NSTAGE, CLI_STAGE_BYTES, d->stage) != 0) d->rc = 1;
    return 0;
}
static void *dev_queries(void *p)
{
    cli_dev *d = (cli_dev*)p;
    if (lqcov_set_queries(d->c, &d->q) != 0) d->rc = 1;
    return 0;
}
static void *dev_finish_part(void *p)    /* everything behind the last chunk: count, exchange, replicated index, mapping */
{
    cli_dev *d = (cli_dev*)p;
    if (lqcov_part_end(d->c) != 0) { d->rc = 1; return 0; }
    if (d->dev == 0) TL("part sketched + counted");
    if (d->n_dev > 1 && lqcov_part_exchange(d->c) != 0) { d->rc = 1; return 0; }
    if (d->part->n == 0) return 0;       /* an empty part (a rejected record closed it): nothing to map against */
    if (lqcov_part_finish(d->c, d->part) != 0) { d->rc = 1; return 0; }
    if (d->dev == 0) fprintf(stderr, "[M::%s::%.3f*%.2f] indexed %u target sequence(s)\n", __func__, wall() - d->t0, cpu() / (wall() - d->t0), d->part->n);
    if (lqcov_map_part(d->c) != 0) d->rc = 1;
    return 0;
}
static void *dev_table(void *p)
{
    cli_dev *d = (cli_dev*)p;
    if (lqcov_table(d->c, &d->tab, &d->tab_len) != 0) d->rc = 1;
    return 0;
}
/* fn on every device, one thread each; returns non-zero when any failed */
static int on_all(cli_dev *dv, int n, void *(*fn)(void*))
{
    pthread_t th[CLI_MAX_DEV]; int rc = 0;
    if (n == 1) { fn(&dv[0]); return dv[0].rc; }
    for (int i = 0; i < n; ++i) pthread_create(&th[i], 0, fn, &dv[i]);
    for (int i = 0; i < n; ++i) { pthread_join(th[i], 0); rc |= dv[i].rc; }
    return rc;
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

/* configurations without a chunked form (-H: the spike-in run, a 4 kb target): one GPU, whole parts */
static int run_parts_whole(cli_dev *d, lqi_reader *tr, const lqcov_opt_t *o, double t0)
{
    part_meta pm; memset(&pm, 0, sizeof pm);
    int rc = 0, eof = 0;
    char *whole = 0; size_t whole_cap = 0;
    lqi_part_rule(tr, o->batch_size, o->mini_batch_size);
    while (rc == 0 && !eof) {
        size_t whole_n = 0; int part_end = 0;
        pm.n = 0; pm.nn = 0;
        while (!part_end && !eof) {
            lqi_chunk ch; int r;
            if (whole_cap - whole_n < CLI_STAGE_BYTES) { whole_cap = whole_cap ? whole_cap * 2 : 4 * CLI_STAGE_BYTES; whole = (char*)realloc(whole, whole_cap); }
            r = lqi_next_chunk(tr, whole_cap - whole_n, whole + whole_n, 0, &ch);
            if (r == -2) { whole_cap = whole_n + ch.need + CLI_STAGE_BYTES; whole = (char*)realloc(whole, whole_cap); continue; }
            part_end = ch.part_end; eof = ch.eof;
            if (r <= 0) continue;
            meta_add(&pm, &ch);
            whole_n += ch.n_bases;
        }
        if (pm.n > 0) {
            lqcov_reads_t part; memset(&part, 0, sizeof part);
            part.n = (uint32_t)pm.n; part.seq = whole; part.seq_off = pm.seq_off; part.names = pm.names; part.name_off = pm.name_off;
            if (lqcov_add_part(d->c, &part) != 0) rc = 1;
            fprintf(stderr, "[M::%s::%.3f*%.2f] mapped against part of %u sequence(s)\n", __func__, wall() - t0, cpu() / (wall() - t0), part.n);
        }
    }
    free(whole); free(pm.seq_off); free(pm.name_off); free(pm.names);
    return rc;
}

static int run_parts(cli_dev *dv, int n_dev, lqi_reader *tr, const lqcov_opt_t *o, double t0)
{
    part_meta pm; memset(&pm, 0, sizeof pm);
    int rc = 0, eof = 0;
    lqi_part_rule(tr, o->batch_size, o->mini_batch_size);
    while (rc == 0 && !eof) {
        uint64_t expect = lqi_bases_left_bound(tr);           /* first sizing of the device arrays; they grow if the part turns out larger */
        int cur = 0, part_end = 0; uint64_t pb = 0;
        lqcov_reads_t part;
        if (expect > o->batch_size + 2 * (uint64_t)o->mini_batch_size) expect = o->batch_size + 2 * (uint64_t)o->mini_batch_size;
        pm.n = 0; pm.nn = 0;
        for (int i = 0; i < n_dev; ++i) { dv[i].begun = 0; dv[i].chunks = 0; }
        while (rc == 0 && !part_end && !eof) {
            lqi_chunk ch; int r; cli_dev *d;
            lqcov_reads_t cr;
            /* GPU `cur` owns the reads until the part's bases pass its share of the expected total */
            while (cur < n_dev - 1 && expect && pb >= expect / (uint64_t)n_dev * (uint64_t)(cur + 1)) ++cur;
            d = &dv[cur];
            if (!d->begun) { if (lqcov_part_begin(d->c, expect / (uint64_t)n_dev + (expect >> 4), (uint32_t)pm.n) != 0) { rc = 1; break; } d->begun = 1; if (cur == 0) TL("part begun (device arrays sized)"); }
            lqcov_stage_wait(d->c, d->chunks % CLI_NSTAGE);
            r = lqi_next_chunk(tr, CLI_STAGE_BYTES, d->stage[d->chunks % CLI_NSTAGE], 0, &ch);
            memset(&cr, 0, sizeof cr);
            if (r == -2) {                                    /* a read longer than a staging buffer: through pageable memory */
                char *big = (char*)malloc(ch.need + 1);       /* (cudaMemcpyAsync returns once a pageable source has been staged: freed at once) */
                r = lqi_next_chunk(tr, ch.need, big, 0, &ch);
                if (r == 1) {
                    cr.n = ch.n; cr.seq = big; cr.seq_off = ch.seq_off; cr.names = ch.names; cr.name_off = ch.name_off;
                    meta_add(&pm, &ch); pb += ch.n_bases;
                    if (lqcov_part_chunk(d->c, &cr, -1) != 0) rc = 1;
                }
                free(big);
                part_end = ch.part_end; eof = ch.eof;
                continue;
            }
            part_end = ch.part_end; eof = ch.eof;
            if (r <= 0) continue;
            meta_add(&pm, &ch); pb += ch.n_bases;
            cr.n = ch.n; cr.seq = d->stage[d->chunks % CLI_NSTAGE]; cr.seq_off = ch.seq_off; cr.names = ch.names; cr.name_off = ch.name_off;
            if (lqcov_part_chunk(d->c, &cr, d->chunks % CLI_NSTAGE) != 0) rc = 1;
            ++d->chunks;
        }
        if (rc != 0) break;
        TL("part parsed, all chunks queued");
        for (int i = 0; i < n_dev; ++i)                        /* GPUs that got nothing of this part still take part in the exchange */
            if (!dv[i].begun) { if (lqcov_part_begin(dv[i].c, 0, (uint32_t)pm.n) != 0) rc = 1; dv[i].begun = 1; }
        memset(&part, 0, sizeof part);
        if (pm.n == 0) { pm.m = pm.m ? pm.m : 4; if (!pm.seq_off) { pm.seq_off = (uint64_t*)calloc(pm.m, 8); pm.name_off = (uint64_t*)calloc(pm.m, 8); pm.names = (char*)calloc(4, 1); pm.nm = 4; } pm.seq_off[0] = pm.name_off[0] = 0; }
        part.n = (uint32_t)pm.n; part.seq_off = pm.seq_off; part.names = pm.names; part.name_off = pm.name_off;
        for (int i = 0; i < n_dev; ++i) { dv[i].part = &part; dv[i].t0 = t0; }
        if (rc == 0 && on_all(dv, n_dev, dev_finish_part) != 0) rc = 1;
        if (pm.n > 0) fprintf(stderr, "[M::%s::%.3f*%.2f] mapped against part of %u sequence(s)\n", __func__, wall() - t0, cpu() / (wall() - t0), part.n);
    }
    free(pm.seq_off); free(pm.name_off); free(pm.names);
    return rc;
}

/* GPUs this run drives: all visible ones (LQCOV_GPUS limits), one for jobs that do not shard (-H; k > 15, whose index addresses are
 * numbered per GPU: lq_widx.cu) */
static int pick_devices(const lqcov_opt_t *o)
{
    int n = lqcov_device_count();
    const char *e = getenv("LQCOV_GPUS");
    if (e && atoi(e) > 0 && atoi(e) < n) n = atoi(e);
    if (n > CLI_MAX_DEV) n = CLI_MAX_DEV;
    if (o->is_hpc || o->k > 15 || n < 1) n = 1;
    return n;
}

/* `-d FILE` (write every part's index as the reference's "MMI\2" image, index.c:390-426) and index files as the target argument
 * (longQC.py --db): one GPU, a part at a time in host memory -- the dump needs the bases for the packed-sequence block. */
static int run_dump_or_indexfile(cli_dev *d, const char *target, int is_idx, const char *dump_path, lqi_reader *tr, const lqcov_opt_t *o, int have_q, double t0)
{
    int rc = 0;
    if (is_idx) {
        FILE *fp = fopen(target, "rb");
        int r;
        if (!fp) { fprintf(stderr, "ERROR: failed to open file '%s'\n", target); return 1; }
        while ((r = lqcov_load_part(d->c, fp)) == 1) {
            fprintf(stderr, "[M::%s::%.3f*%.2f] loaded/built the index for target sequence(s)\n", __func__, wall() - t0, cpu() / (wall() - t0));
            if (have_q && lqcov_map_part(d->c) != 0) { rc = 1; break; }
        }
        if (r < 0) rc = 1;
        fclose(fp);
        return rc;
    }
    FILE *fd = dump_path ? fopen(dump_path, "wb") : 0;
    if (dump_path && !fd) { fprintf(stderr, "ERROR: failed to open file '%s' for writing\n", dump_path); return 1; }
    part_meta pm; memset(&pm, 0, sizeof pm);
    int eof = 0;
    char *whole = 0; size_t whole_cap = 0;
    lqi_part_rule(tr, o->batch_size, o->mini_batch_size);
    while (rc == 0 && !eof) {
        size_t whole_n = 0; int part_end = 0;
        pm.n = 0; pm.nn = 0;
        while (!part_end && !eof) {
            lqi_chunk ch; int r;
            if (whole_cap - whole_n < CLI_STAGE_BYTES) { whole_cap = whole_cap ? whole_cap * 2 : 4 * CLI_STAGE_BYTES; whole = (char*)realloc(whole, whole_cap); }
            r = lqi_next_chunk(tr, whole_cap - whole_n, whole + whole_n, 0, &ch);
            if (r == -2) { whole_cap = whole_n + ch.need + CLI_STAGE_BYTES; whole = (char*)realloc(whole, whole_cap); continue; }
            part_end = ch.part_end; eof = ch.eof;
            if (r <= 0) continue;
            meta_add(&pm, &ch);
            whole_n += ch.n_bases;
        }
        if (pm.n > 0) {
            lqcov_reads_t part; memset(&part, 0, sizeof part);
            part.n = (uint32_t)pm.n; part.seq = whole; part.seq_off = pm.seq_off; part.names = pm.names; part.name_off = pm.name_off;
            if (lqcov_index_part(d->c, &part) != 0) rc = 1;
            fprintf(stderr, "[M::%s::%.3f*%.2f] loaded/built the index for %u target sequence(s)\n", __func__, wall() - t0, cpu() / (wall() - t0), part.n);
            if (rc == 0 && fd && lqcov_index_dump(d->c, &part, fd) != 0) rc = 1;
            if (rc == 0 && have_q && lqcov_map_part(d->c) != 0) rc = 1;
        }
    }
    if (fd) fclose(fd);
    free(whole); free(pm.seq_off); free(pm.name_off); free(pm.names);
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
    /* stdout carries the table and nothing else (lq_coverage.py parses it): whatever a library prints there (NCCL's version banner
     * under NCCL_DEBUG, for one) is sent to stderr instead, and the table goes to the saved descriptor at the end */
    fflush(stdout);
    const int out_fd = dup(STDOUT_FILENO);
    dup2(STDERR_FILENO, STDOUT_FILENO);

    if (a.ava && a.avs) { fprintf(stderr, "Error: -X and -Y are mutually exclusive\n"); return 1; }
    if (!a.ava && !a.avs && !a.dump) { fprintf(stderr, "Error: Choose either -X (all-vs-all) or -Y (all-vs-sub)\n"); return 1; }
    o.no_self = 1; o.ava = a.ava ? 1 : 0;
    if (a.h_flag) o.is_hpc = 1; else fprintf(stderr, "Homopolymer compression is not applied.\n");
    if (a.k == 0) { fprintf(stderr, "Warning: Apply default k=12 instead. \n"); o.k = 12; } else o.k = a.k;
    if (a.w == 0) { fprintf(stderr, "Warning: Apply default w=5 instead. \n"); o.w = 5; } else o.w = a.w;
    if (a.batch == 0) fprintf(stderr, "Warning: Apply default I=4G instead. \n"); else o.batch_size = a.batch;
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

    /* an index file as the target (index.c:481-498): its k / w / -H replace the command line's for the mapping, not for the rows' `n` */
    int idx_k = 0, idx_w = 0, idx_hpc = 0;
    const int is_idx = lqcov_index_peek(a.args[0], &idx_k, &idx_w, &idx_hpc);
    const int cl_k = o.k, cl_w = o.w, cl_hpc = o.is_hpc, have_q = a.args[1] != 0;
    if (is_idx < 0) { fprintf(stderr, "ERROR: failed to open file '%s'\n", a.args[0]); return 1; }
    if (is_idx == 1) {
        if (idx_k != o.k || idx_w != o.w || idx_hpc != o.is_hpc)
            fprintf(stderr, "[WARNING]\033[1;31m Indexing parameters (-k, -w or -H) overridden by parameters used in the prebuilt index.\033[0m\n");
        o.k = idx_k; o.w = idx_w; o.is_hpc = idx_hpc;
    }
    /* The CUDA contexts (hundreds of ms each) are created on their own threads while the reader threads already parse the inputs. */
    cli_dev dv[CLI_MAX_DEV]; pthread_t cth[CLI_MAX_DEV];
    const int n_dev = (is_idx == 1 || a.dump) ? 1 : pick_devices(&o);
    memset(dv, 0, sizeof dv);
    for (int i = 0; i < n_dev; ++i) { dv[i].dev = i; dv[i].n_dev = n_dev; dv[i].o = o; dv[i].o.device = n_dev > 1 ? i : -1; pthread_create(&cth[i], 0, dev_create, &dv[i]); }
    const int n_rd = reader_threads(o.n_threads);
    lqi_reader *tr = is_idx == 1 ? 0 : lqi_open(a.args[0], n_rd);
    lqcov_reader *qr = (tr || is_idx == 1) && have_q ? lqcov_reader_open(a.args[1]) : 0;
    lqcov_reads_t q; memset(&q, 0, sizeof q);
    static const uint64_t zero_off[1] = { 0 };
    q.seq_off = zero_off; q.name_off = zero_off; q.seq = ""; q.names = "";
    if (qr) lqcov_reader_next(qr, 0, &q);                     /* one kseq_read loop over the query file (minimap2-coverage.c:418) */
    TL("queries parsed");
    int rc = 0;
    for (int i = 0; i < n_dev; ++i) { pthread_join(cth[i], 0); rc |= dv[i].rc; }
    if (!tr && is_idx != 1) fprintf(stderr, "ERROR: failed to open file '%s'\n", a.args[0]);
    else if (!qr && have_q) fprintf(stderr, "ERROR: failed to open file '%s'\n", a.args[1]);
    if ((!tr && is_idx != 1) || (!qr && have_q) || rc) { if (tr) lqi_close(tr); if (qr) lqcov_reader_close(qr); for (int i = 0; i < n_dev; ++i) if (dv[i].c) lqcov_destroy(dv[i].c); return 1; }
    if (n_dev > 1) {
        lqcov_ctx *cs[CLI_MAX_DEV];
        for (int i = 0; i < n_dev; ++i) cs[i] = dv[i].c;
        if (lqcov_comm_init_all(cs, n_dev) != 0) rc = 1;
        fprintf(stderr, "[M::%s] %d GPUs: targets sharded for sketching, minimizer counts all-reduced, index replicated, queries split\n", __func__, n_dev);
    }
    /* staging buffers are page-locked beside the query sketch; GPU i takes queries [nq*i/n, nq*(i+1)/n) */
    for (int i = 0; i < n_dev; ++i) pthread_create(&cth[i], 0, dev_stage, &dv[i]);
    for (int i = 0; i < n_dev; ++i) {
        const uint32_t lo = (uint32_t)((uint64_t)q.n * i / n_dev), hi = (uint32_t)((uint64_t)q.n * (i + 1) / n_dev);
        dv[i].q = q; dv[i].q.n = hi - lo; dv[i].q.seq_off = q.seq_off + lo; dv[i].q.name_off = q.name_off + lo;
    }
    { cli_dev qd[CLI_MAX_DEV]; for (int i = 0; i < n_dev; ++i) { qd[i] = dv[i]; qd[i].rc = 0; } if (rc == 0 && on_all(qd, n_dev, dev_queries) != 0) rc = 1; }
    for (int i = 0; i < n_dev; ++i) { pthread_join(cth[i], 0); rc |= dv[i].rc; }
    TL("queries sketched, staging buffers allocated");
    fprintf(stderr, "[M::%s::%.3f*%.2f] loaded %u sequence(s).\n", __func__, wall() - t0, cpu() / (wall() - t0), q.n);
    if (rc == 0 && is_idx == 1 && have_q && (cl_k != o.k || cl_w != o.w || cl_hpc != o.is_hpc) && lqcov_set_prepass_counts(dv[0].c, &q, cl_k, cl_w, cl_hpc) != 0) rc = 1;
    if (rc == 0) {
        if (is_idx == 1 || a.dump) rc = run_dump_or_indexfile(&dv[0], a.args[0], is_idx == 1, a.dump, tr, &o, have_q, t0);
        else rc = lqcov_part_begin(dv[0].c, 0, 0) == 1 ? run_parts_whole(&dv[0], tr, &o, t0) : run_parts(dv, n_dev, tr, &o, t0);
    }
    if (tr) lqi_close(tr);
    if (!have_q) {                       /* -d without a query file: the index is written, nothing is mapped (minimap2-coverage.c:460-468) */
        for (int i = 0; i < n_dev; ++i) lqcov_destroy(dv[i].c);
        close(out_fd);
        fprintf(stderr, "\n[M::%s] Real time: %.3f sec; CPU: %.3f sec\n", __func__, wall() - t0, cpu());
        return rc;
    }
    if (rc == 0 && on_all(dv, n_dev, dev_table) != 0) rc = 1;
    if (rc == 0) {
        FILE *out = fdopen(out_fd, "w");
        for (int i = 0; i < n_dev; ++i) { fwrite(dv[i].tab, 1, dv[i].tab_len, out); lqcov_free(dv[i].tab); }
        fclose(out);
    }
    if (qr) lqcov_reader_close(qr);
    TL("table written");
    { const char *fe = getenv("LQCOV_FAST_EXIT"); if (!fe || fe[0] == '0') for (int i = 0; i < n_dev; ++i) lqcov_destroy(dv[i].c); }   /* set (by lq_main_cov.c): the process is about to _exit */
    fprintf(stderr, "\n[M::%s] Real time: %.3f sec; CPU: %.3f sec\n", __func__, wall() - t0, cpu());
    return rc;
}
