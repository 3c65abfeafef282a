/* lq_api.cu -- the C ABI of include/lqcov.h: context, part loop, host bookkeeping.
 *
 * Control flow mirrors main() of the reference (minimap2-coverage.c:199-621):
 *   lqcov_set_queries  ~ :406-444   query pre-pass + accumulators
 *   lqcov_add_part     ~ :450-458   mm_idx_reader_read + mm_mapopt_update + lq_map_file
 *   lqcov_table        ~ :545-617   final pass
 * All compute is on the device (lq_sketch.cu, lq_index.cu, lq_map.cu); the host keeps the per-query
 * interval lists and formats the table (lq_table.c).
 */
#include <string>
#include <vector>
#include <unordered_map>
#include <algorithm>
#include <chrono>
#include <thread>
#include <future>
#include <string.h>
#include "lq_cuda.cuh"
#include "lq_device.h"
#include "lq_index.h"
#include "lq_map.h"
#include "lq_host.h"
#include "lqcov.h"
#include "lq_prof.h"

int lq_qualsum_run(const uint8_t *d_qual, const uint64_t *d_off, uint32_t n_reads, double *d_sum, cudaStream_t st);

/* host worker threads for the per-query bookkeeping (interval folding, table rows): LQCOV_HOST_THREADS, else the cores we may use, <= 16 */
static unsigned host_threads()
{
    static unsigned n = 0;
    if (n == 0) {
        const char *e = getenv("LQCOV_HOST_THREADS");
        n = e ? (unsigned)atoi(e) : std::thread::hardware_concurrency();
        if (n < 1) n = 1;
        if (n > 16) n = 16;
    }
    return n;
}
/* fn(lo, hi, chunk index) over [0, n) cut into contiguous chunks, one thread each */
template <class F> static void parallel_chunks(uint32_t n, unsigned n_chunks, F fn)
{
    if (n_chunks > n) n_chunks = n ? n : 1;
    if (n_chunks <= 1) { fn(0u, n, 0u); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < n_chunks; ++t) {
        const uint32_t lo = (uint32_t)((uint64_t)n * t / n_chunks), hi = (uint32_t)((uint64_t)n * (t + 1) / n_chunks);
        th.emplace_back([=]() { fn(lo, hi, t); });
    }
    for (size_t t = 0; t < th.size(); ++t) th[t].join();
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

#include "lq_ctx.h"
#include "lq_comm.h"
#include "lq_mmi.h"

extern "C" int lqcov_abi_version(void) { return LQCOV_ABI_VERSION; }
extern "C" int lqcov_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
/* contexts created with an explicit device may be driven from any host thread: every entry point selects the device first */
#define LQ_USE_DEV(c) do { if ((c)->opt.device >= 0 && cudaSetDevice((c)->opt.device) != cudaSuccess) { fprintf(stderr, "[lqcov] ERROR: cannot select CUDA device %d\n", (c)->opt.device); return -1; } } while (0)

extern "C" void lqcov_opt_init(lqcov_opt_t *o)
{
    memset(o, 0, sizeof(*o));
    o->k = 12; o->w = 5; o->is_hpc = 0; o->batch_size = 4000000000ULL; o->mini_batch_size = 50000000;
    o->no_self = 1; o->ava = 0;
    o->max_gap = 10000; o->min_cnt = 3; o->min_chain_score = 40; o->min_score_med = 40; o->min_score_good = 40;
    o->max_chain_skip = 25; o->bw = 500; o->mid_occ_frac = 2e-4f;
    o->max_overhang = 2000; o->min_ovlp = 1000; o->min_coverage = 3; o->min_ratio = 0.4; o->filter = 0;
    o->n_threads = 1; o->device = -1; o->seed_budget = 0; o->verbose = 0;
}

extern "C" void lqcov_free(void *p) { free(p); }

extern "C" lqcov_ctx *lqcov_create(const lqcov_opt_t *o)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        fprintf(stderr, "[lqcov] ERROR: no usable CUDA device (%s). This library has no CPU path.\n", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return 0;
    }
    if (o->device >= 0) { if (cudaSetDevice(o->device) != cudaSuccess) { fprintf(stderr, "[lqcov] ERROR: cannot select CUDA device %d\n", o->device); return 0; } }
    if (o->w < 1 || o->w > LQ_MAX_W) { fprintf(stderr, "[lqcov] ERROR: -w %d outside 1..%d supported by the GPU path\n", o->w, LQ_MAX_W); return 0; }
    if (o->k < 1 || o->k > LQ_MAX_K) { fprintf(stderr, "[lqcov] ERROR: -k %d outside 1..%d\n", o->k, LQ_MAX_K); return 0; }   /* minimap2-coverage.c:171 */
    lqcov_ctx *c = new lqcov_ctx();
    c->opt = *o; c->stage_bytes = 0; c->comm = 0; c->placed = false; c->nq = 0; c->q_has_qual = false; c->part_ready = false; c->use_full = false; c->mid_occ = 0;
    memset(&c->stats, 0, sizeof(c->stats));
    if (cudaStreamCreate(&c->st) != cudaSuccess) { fprintf(stderr, "[lqcov] ERROR: cudaStreamCreate failed\n"); delete c; return 0; }
    return c;
}

/* forget everything learnt from previous parts (mid_occ, accumulators); buffers stay allocated */
extern "C" int lqcov_reset(lqcov_ctx *c)
{
    LQ_USE_DEV(c);
    c->mid_occ = 0; c->part_ready = false; c->use_full = false;
    memset(&c->stats, 0, sizeof(c->stats));
    for (size_t i = 0; i < c->ovlp.size(); ++i) { free(c->ovlp[i].a); c->ovlp[i].a = 0; c->ovlp[i].n = c->ovlp[i].m = 0; }
    return 0;
}

extern "C" void lqcov_destroy(lqcov_ctx *c)
{
    if (!c) return;
    if (c->opt.device >= 0) cudaSetDevice(c->opt.device);
    cudaStreamSynchronize(c->st);
    lq_comm_release(c);
    for (size_t i = 0; i < c->ovlp.size(); ++i) free(c->ovlp[i].a);
    c->stream.release(); for (size_t i = 0; i < c->stage.size(); ++i) cudaFreeHost(c->stage[i]);
    c->qd.release(); c->ix.release(); c->sc.release(); c->treads.release(); c->tmins.release(); c->full.release(); c->ws.release(); c->qual_dev.release(); c->qsum_dev.release();
    cudaStreamDestroy(c->st);
    delete c;
}

static std::string name_of(const lqcov_reads_t *r, uint32_t i) { return std::string(r->names + r->name_off[i], (size_t)(r->name_off[i + 1] - r->name_off[i])); }

static inline uint64_t name_hash(const char *s, size_t n) { uint64_t h = 1469598103934665603ULL; for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 1099511628211ULL; } return h ^ (h >> 29); }

static void qnames_build(lqcov_ctx *c)
{
    uint32_t cap = 16; while (cap < 4 * c->nq + 16) cap <<= 1;
    c->qn_mask = cap - 1; c->qn_slot.assign(cap, -1); c->qn_next.assign(c->nq, -1);
    for (uint32_t q = c->nq; q-- > 0; ) { /* reverse, so that chains list the queries in ascending order */
        const std::string &nm = c->qname[q];
        uint32_t h = (uint32_t)name_hash(nm.data(), nm.size()) & c->qn_mask;
        for (;; h = (h + 1) & c->qn_mask) {
            const int32_t s = c->qn_slot[h];
            if (s < 0) { c->qn_slot[h] = (int32_t)q; break; }
            if (c->qname[s] == nm) { c->qn_next[q] = s; c->qn_slot[h] = (int32_t)q; break; }
        }
    }
}
/* first query with this name (or -1); further ones through qn_next */
static inline int32_t qnames_find(const lqcov_ctx *c, const char *s, size_t n)
{
    for (uint32_t h = (uint32_t)name_hash(s, n) & c->qn_mask;; h = (h + 1) & c->qn_mask) {
        const int32_t q = c->qn_slot[h];
        if (q < 0) return -1;
        if (c->qname[q].size() == n && memcmp(c->qname[q].data(), s, n) == 0) return q;
    }
}

extern "C" int lqcov_set_queries(lqcov_ctx *c, const lqcov_reads_t *q)
{
    LQ_USE_DEV(c);
    const double t0 = now_ms();
    c->nq = q->n; c->n_prepass.clear();
    c->qname.resize(q->n); c->qlen.resize(q->n);
    for (uint32_t i = 0; i < q->n; ++i) {
        c->qname[i] = name_of(q, i);
        c->qlen[i] = (int)(q->seq_off[i + 1] - q->seq_off[i]);
    }
    qnames_build(c);
    c->q_has_qual = q->qual != 0;
    c->qsum_p.assign(q->n, 0.0);
    for (size_t i = 0; i < c->ovlp.size(); ++i) free(c->ovlp[i].a);
    c->ovlp.assign(q->n, lqh_sub_v());
    for (uint32_t i = 0; i < q->n; ++i) { c->ovlp[i].n = c->ovlp[i].m = 0; c->ovlp[i].a = 0; }
    c->avg_k.assign(q->n, 0.f);
    LqQueryDev *qd = &c->qd;
    qd->nq = q->n;
    LQ_TRY(lq_reads_upload(&qd->reads, (const uint8_t*)q->seq, q->seq_off, q->n, q->seq_on_device, 0, c->st));
    LQ_TRY(lq_sketch_run(&qd->reads, c->opt.w, c->opt.k, c->opt.is_hpc, 0, &qd->mins, c->ws, c->st));
    qd->n_min = qd->mins.n;
    LQ_TRY(lq_read_first(&qd->mins, 0, q->n, qd->first, c->st));
    if (qd->mins.wide) {   /* k > 15: equal keys <=> equal ids of a table over the query set's own keys; per part the ids are the part's (lqcov_map_part) */
        LqWideTable qt; int kb = 1;
        int rc = lq_wide_build(&qt, qd->mins.key64.as<uint64_t>(), qd->n_min, qd->mins.key.as<uint32_t>(), c->st);
        while (kb < 32 && ((uint64_t)qt.n_ids >> kb)) ++kb;
        if (rc == 0) rc = lq_map_flag_dups(qd, kb, c->ws, c->st);
        qt.release();
        if (rc != 0) return -1;
    } else LQ_TRY(lq_map_flag_dups(qd, 2 * c->opt.k, c->ws, c->st));
    if (q->qual && q->n) { /* meanQ of the table rows (lqutils.c:51-58): the additions must happen in read order, one thread per read */
        const uint64_t nbq = q->seq_off[q->n] - q->seq_off[0];
        LQ_TRY(c->qual_dev.ensure(nbq + 16)); LQ_TRY(c->qsum_dev.ensure(((size_t)q->n + 1) * 8));
        LQ_CUDA_OK(cudaMemcpyAsync(c->qual_dev.p, q->qual + q->seq_off[0], nbq, cudaMemcpyHostToDevice, c->st)); lq_prof_h2d(nbq);
        /* qd->reads.off holds the absolute offsets; the copy starts at seq_off[0] */
        LQ_TRY(lq_qualsum_run(c->qual_dev.as<uint8_t>() - q->seq_off[0], qd->reads.off.as<uint64_t>(), q->n, c->qsum_dev.as<double>(), c->st));
        LQ_CUDA_OK(cudaMemcpyAsync(c->qsum_p.data(), c->qsum_dev.p, (size_t)q->n * 8, cudaMemcpyDeviceToHost, c->st)); lq_prof_d2h((uint64_t)q->n * 8);
    }
    c->qfirst.resize((size_t)q->n + 1);
    LQ_CUDA_OK(cudaMemcpyAsync(c->qfirst.data(), qd->first.p, ((size_t)q->n + 1) * 8, cudaMemcpyDeviceToHost, c->st));
    LQ_TRY(qd->lambda.ensure(((size_t)q->n + 1) * 8)); LQ_TRY(qd->lambda2.ensure(((size_t)q->n + 1) * 8));
    LQ_TRY(qd->mcnt.ensure((qd->n_min + 1) * 4));
    LQ_CUDA_OK(cudaMemsetAsync(qd->lambda.p, 0, ((size_t)q->n + 1) * 8, c->st));
    LQ_CUDA_OK(cudaMemsetAsync(qd->lambda2.p, 0, ((size_t)q->n + 1) * 8, c->st));
    LQ_CUDA_OK(cudaMemsetAsync(qd->mcnt.p, 0, (qd->n_min + 1) * 4, c->st));
    LQ_CUDA_OK(cudaStreamSynchronize(c->st));
    for (uint32_t i = 0; i < q->n; ++i)
        if (c->qfirst[i + 1] - c->qfirst[i] >= (1u << 24)) { fprintf(stderr, "[lqcov] ERROR: query %u has >= 2^24 minimizers\n", i); return -1; }
    c->stats.query_bases = qd->reads.n_bases; c->stats.query_minimizers = qd->n_min;
    c->stats.t_sketch_ms += now_ms() - t0;
    if (c->opt.verbose >= 3) fprintf(stderr, "[M::lqcov] loaded %u query sequence(s). Total m_cnt: %llu\n", q->n, (unsigned long long)qd->n_min);
    return 0;
}

/* ---- one index part in three phases, so that several GPUs can share it:
 *   lqcov_part_sketch   every rank: pack + sketch ITS shard of the part's reads (rid = rid_base + i) and count minimizers
 *   (collective)        all-reduce of the count table, all-gather of the (key, y) records in rank order  [longqc_b200/dist.py]
 *   lqcov_part_finish   every rank: offsets, stable sort by key, mid_occ, name tables  -> replicated index
 * lqcov_index_part() is the single-GPU composition. */
extern "C" int lqcov_part_sketch(lqcov_ctx *c, const lqcov_reads_t *shard, uint32_t rid_base)
{
    LQ_USE_DEV(c);
    double t0 = now_ms();
    LqIndexDev *ix = &c->ix;
    c->part_ready = false; c->use_full = false; c->placed = false;
    int piped = 1;
    if (!shard->seq_on_device) {   /* host bases: copy, pack and sketch overlap chunk by chunk where the rolling kernel applies */
        piped = lq_upload_sketch_pipelined(&c->treads, (const uint8_t*)shard->seq, shard->seq_off, shard->n, c->opt.w, c->opt.k, c->opt.is_hpc, rid_base, &ix->rec, c->ws, c->st);
        if (piped < 0) return -1;
    }
    if (piped == 1) {
        LQ_TRY(lq_reads_upload(&c->treads, (const uint8_t*)shard->seq, shard->seq_off, shard->n, shard->seq_on_device, 0, c->st));
        LQ_CUDA_OK(cudaStreamSynchronize(c->st));
        c->stats.t_upload_ms += now_ms() - t0; t0 = now_ms();
        LQ_TRY(lq_sketch_run(&c->treads, c->opt.w, c->opt.k, c->opt.is_hpc, rid_base, &ix->rec, c->ws, c->st)); /* rid restarts at 0 in every part (index.c:287) */
    }
    LQ_CUDA_OK(cudaStreamSynchronize(c->st));
    c->stats.t_sketch_ms += now_ms() - t0; t0 = now_ms();
    uint64_t n_addr = 0;
    if (ix->rec.wide) {    /* k > 15: dense addresses for the part's distinct keys, one more for "not in this part" */
        LQ_TRY(lq_wide_build(&ix->wide, ix->rec.key64.as<uint64_t>(), ix->rec.n, ix->rec.key.as<uint32_t>(), c->st));
        n_addr = (uint64_t)ix->wide.n_ids + 1;
    }
    LQ_TRY(lq_index_alloc(ix, c->opt.k, n_addr, c->st));
    LQ_TRY(lq_index_count(ix, &ix->rec, c->st));
    LQ_CUDA_OK(cudaStreamSynchronize(c->st));
    c->stats.t_index_ms += now_ms() - t0;
    c->stats.target_bases += c->treads.n_bases;
    return 0;
}

/* ---- the same, with the shard arriving in chunks from pinned staging buffers while the caller's reader threads parse the file ---- */
extern "C" int lqcov_part_begin(lqcov_ctx *c, uint64_t expect_bases, uint32_t rid_base)
{
    LQ_USE_DEV(c);
    if (!lq_stream_ok(c->opt.w, c->opt.k, c->opt.is_hpc)) return 1;   /* not applicable: hand the whole part to lqcov_part_sketch */
    c->t_part0 = now_ms();
    c->part_ready = false; c->use_full = false; c->placed = false;
    LQ_TRY(lq_stream_begin(&c->stream, &c->treads, &c->ix.rec, c->opt.w, c->opt.k, rid_base, expect_bases, c->st));
    return 0;
}

extern "C" int lqcov_stage(lqcov_ctx *c, int n, size_t bytes, char **bufs)
{
    LQ_USE_DEV(c);
    if ((int)c->stage.size() != n || c->stage_bytes != bytes) {
        for (size_t i = 0; i < c->stage.size(); ++i) cudaFreeHost(c->stage[i]);
        c->stage.assign((size_t)n, (char*)0); c->stage_ev.assign((size_t)n, (cudaEvent_t)0); c->stage_bytes = bytes;
        for (int i = 0; i < n; ++i) LQ_CUDA_OK(cudaHostAlloc((void**)&c->stage[i], bytes, cudaHostAllocDefault));
    }
    for (int i = 0; i < n; ++i) { bufs[i] = c->stage[i]; c->stage_ev[i] = 0; }
    return 0;
}

extern "C" int lqcov_part_chunk(lqcov_ctx *c, const lqcov_reads_t *chunk, int stage_index)
{
    LQ_USE_DEV(c);
    cudaEvent_t ev = 0;
    LQ_TRY(lq_stream_push(&c->stream, (const uint8_t*)chunk->seq, chunk->seq_off, chunk->n, &ev));
    if (stage_index >= 0 && stage_index < (int)c->stage_ev.size()) c->stage_ev[stage_index] = ev;
    return 0;
}

extern "C" int lqcov_stage_wait(lqcov_ctx *c, int stage_index)
{
    LQ_USE_DEV(c);
    if (stage_index >= 0 && stage_index < (int)c->stage_ev.size() && c->stage_ev[stage_index]) {
        LQ_CUDA_OK(cudaEventSynchronize(c->stage_ev[stage_index]));   /* the event may have been re-recorded by a later chunk: waiting longer is harmless */
        c->stage_ev[stage_index] = 0;
    }
    return 0;
}

extern "C" int lqcov_part_end(lqcov_ctx *c)
{
    LQ_USE_DEV(c);
    LqIndexDev *ix = &c->ix;
    LQ_TRY(lq_stream_end(&c->stream, c->ws));
    c->stats.t_sketch_ms += now_ms() - c->t_part0;
    const double t0 = now_ms();
    LQ_TRY(lq_index_alloc(ix, c->opt.k, 0, c->st));
    LQ_TRY(lq_index_count(ix, &ix->rec, c->st));
    LQ_CUDA_OK(cudaStreamSynchronize(c->st));
    c->stats.t_index_ms += now_ms() - t0;
    c->stats.target_bases += c->treads.n_bases;
    return 0;
}

extern "C" int lqcov_part_device_views(lqcov_ctx *c, void **counts, uint64_t *n_counts, void **key, void **y, uint64_t *n_rec)
{
    *counts = c->ix.counts.p; *n_counts = c->ix.n_keyspace;
    *key = c->ix.rec.key.p; *y = c->ix.rec.y.p; *n_rec = c->ix.rec.n;
    return 0;
}

extern "C" int lqcov_part_gather_buffers(lqcov_ctx *c, uint64_t n_total, void **key, void **y)
{
    if (c->opt.is_hpc) { fprintf(stderr, "[lqcov] ERROR: HPC sketches are not supported on the multi-GPU path\n"); return -1; }
    LQ_TRY(c->full.key.ensure((size_t)(n_total + 1) * 4)); LQ_TRY(c->full.y.ensure((size_t)(n_total + 1) * 8));
    c->full.n = n_total; c->full.has_span = 0; c->use_full = true;
    *key = c->full.key.p; *y = c->full.y.p;
    return 0;
}

/* name tables for the self-diagonal / dual-mapping skips (lqmap.c:180-189): host-only work */
static void build_name_tables(lqcov_ctx *c, const lqcov_reads_t *part)
{
    const uint32_t nq = c->nq;
    /* (query, target) pairs with equal names, then CSR by query with ascending target ids */
    std::vector<std::pair<uint32_t, uint32_t> > hits;
    if (c->opt.no_self || c->opt.ava)
        for (uint32_t t = 0; t < part->n; ++t)
            for (int32_t q = qnames_find(c, part->names + part->name_off[t], (size_t)(part->name_off[t + 1] - part->name_off[t])); q >= 0; q = c->qn_next[q])
                hits.push_back(std::make_pair((uint32_t)q, t));
    c->self_off.assign((size_t)nq + 1, 0); c->self_list.assign(hits.size(), 0);
    for (size_t i = 0; i < hits.size(); ++i) ++c->self_off[hits[i].first + 1];
    for (uint32_t q = 0; q < nq; ++q) c->self_off[q + 1] += c->self_off[q];
    {
        std::vector<uint32_t> fill(c->self_off.begin(), c->self_off.end() - 1);
        for (size_t i = 0; i < hits.size(); ++i) c->self_list[fill[hits[i].first]++] = hits[i].second; /* t ascending within a query */
    }
    c->qrank.clear(); c->trank.clear();
    if (c->opt.ava) { /* strcmp order of names == rank among the sorted distinct names of queries and targets */
        std::vector<std::string> all; all.reserve((size_t)nq + part->n);
        for (uint32_t q = 0; q < nq; ++q) all.push_back(c->qname[q]);
        for (uint32_t t = 0; t < part->n; ++t) all.push_back(name_of(part, t));
        std::vector<std::string> uniq(all);
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        c->qrank.resize(nq); c->trank.resize(part->n);
        for (uint32_t q = 0; q < nq; ++q) c->qrank[q] = (uint32_t)(std::lower_bound(uniq.begin(), uniq.end(), all[q]) - uniq.begin());
        for (uint32_t t = 0; t < part->n; ++t) c->trank[t] = (uint32_t)(std::lower_bound(uniq.begin(), uniq.end(), all[nq + t]) - uniq.begin());
    }
}

static int part_finish_device(lqcov_ctx *c, const lqcov_reads_t *part)
{
    LqIndexDev *ix = &c->ix;
    if (c->use_full) { std::swap(ix->rec.key, c->full.key); std::swap(ix->rec.y, c->full.y); ix->rec.n = c->full.n; ix->rec.has_span = 0; c->use_full = false; }
    if (c->placed) c->placed = false;                      /* lqcov_part_exchange left offsets and positions of the replicated index in place */
    else LQ_TRY(lq_index_finish(ix, &ix->rec, c->ws, c->st));
    ix->n_seq = part->n;
    LQ_TRY(ix->tlen.ensure(((size_t)part->n + 1) * 4));
    {
        std::vector<uint32_t> tl(part->n);
        for (uint32_t i = 0; i < part->n; ++i) tl[i] = (uint32_t)(part->seq_off[i + 1] - part->seq_off[i]);
        if (part->n) LQ_CUDA_OK(cudaMemcpyAsync(ix->tlen.p, tl.data(), (size_t)part->n * 4, cudaMemcpyHostToDevice, c->st));
        LQ_CUDA_OK(cudaStreamSynchronize(c->st));
        lq_prof_h2d((uint64_t)part->n * 4);
    }
    if (c->mid_occ <= 0) { /* map.c:50-51: only while still unset, i.e. from the first part */
        uint64_t nd = 0;
        LQ_TRY(lq_index_mid_occ(ix, c->opt.mid_occ_frac, &c->mid_occ, &nd, c->ws, c->st));
        if (c->opt.verbose >= 3) fprintf(stderr, "[M::lqcov] mid_occ = %d (distinct minimizers: %llu)\n", c->mid_occ, (unsigned long long)nd);
    }
    LQ_CUDA_OK(cudaStreamSynchronize(c->st));
    return 0;
}

extern "C" int lqcov_part_finish(lqcov_ctx *c, const lqcov_reads_t *part)
{
    LQ_USE_DEV(c);
    const double t0 = now_ms();
    /* the name tables are host-only work: built on a second thread while the device sorts the records */
    std::future<void> names = std::async(std::launch::async, build_name_tables, c, part);
    const int rc = part_finish_device(c, part);
    c->stats.t_index_ms += now_ms() - t0;
    const double t1 = now_ms();
    names.get();
    c->stats.t_post_ms += now_ms() - t1;
    if (rc != 0) return rc;
    c->part_ready = true;
    c->stats.target_minimizers += c->ix.n_rec;
    c->stats.n_parts += 1; c->stats.mid_occ = c->mid_occ;
    if (c->opt.verbose >= 3) fprintf(stderr, "[M::lqcov] loaded/built the index for %u target sequence(s)\n", part->n);
    return 0;
}

extern "C" int lqcov_index_part(lqcov_ctx *c, const lqcov_reads_t *part)
{
    LQ_TRY(lqcov_part_sketch(c, part, 0));
    return lqcov_part_finish(c, part);
}

static void map_opt_of(const lqcov_opt_t *o, LqMapOpt *m)
{
    m->no_self = o->no_self; m->ava = o->ava;
    m->max_dist = o->max_gap; m->bw = o->bw; m->max_skip = o->max_chain_skip; m->min_cnt = o->min_cnt; m->min_sc = o->min_chain_score;
    m->min_sc_med = o->min_score_med; m->min_sc_good = o->min_score_good;
    m->max_overhang = o->max_overhang; m->min_ratio = o->min_ratio; m->covt = 150;
}


extern "C" int lqcov_map_part(lqcov_ctx *c)
{
    LQ_USE_DEV(c);
    if (!c->part_ready) { fprintf(stderr, "[lqcov] ERROR: lqcov_map_part without an index part\n"); return -1; }
    if (c->nq == 0) return 0;
    double t0 = now_ms();
    LqMapOpt mo; map_opt_of(&c->opt, &mo);
    std::vector<LqOvl> ovl; std::vector<LqQStat> hs; LqMapStats ms; memset(&ms, 0, sizeof(ms));
    const uint64_t cap = c->opt.seed_budget ? c->opt.seed_budget : 1000000000ULL;
    if (c->qd.mins.wide)   /* k > 15: the query minimizers take the addresses this part gave their keys */
        LQ_TRY(lq_wide_translate(&c->ix.wide, c->qd.mins.key64.as<uint64_t>(), c->qd.n_min, c->qd.mins.key.as<uint32_t>(), c->st));
    LQ_TRY(lq_map_part(&c->qd, &c->ix, &mo, c->mid_occ, c->self_off.data(), c->self_list.data(), c->qrank.data(), c->trank.data(), cap, &c->sc, &ovl, &hs, &ms, c->st));
    lq_prof_collect();
    c->stats.t_map_ms += now_ms() - t0; t0 = now_ms();
    /* esterr.c:93-97: the mean k-mer span is fixed by the first part in which the query keeps a minimizer */
    for (uint32_t q = 0; q < c->nq; ++q)
        if (hs[q].n_kept > 0 && !hs[q].gate_closed && c->avg_k[q] == 0.f) c->avg_k[q] = (float)(uint64_t)hs[q].sum_span_kept / (int32_t)hs[q].n_kept;
    /* lqmap.c:287: fold this part's overlaps into every query's persistent interval list */
    {
        std::vector<uint32_t> off((size_t)c->nq + 1, 0);
        for (size_t i = 0; i < ovl.size(); ++i) ++off[ovl[i].q + 1];
        for (uint32_t q = 0; q < c->nq; ++q) off[q + 1] += off[q];
        std::vector<lqh_sub> cv(ovl.size());
        std::vector<uint32_t> at(off.begin(), off.end() - 1);
        for (size_t i = 0; i < ovl.size(); ++i) { lqh_sub s; s.start = ovl[i].start; s.end = ovl[i].end; cv[at[ovl[i].q]++] = s; }
        const uint32_t min_cov = (uint32_t)c->opt.min_coverage;
        parallel_chunks(c->nq, host_threads(), [&](uint32_t lo, uint32_t hi, unsigned) {   /* queries are independent */
            for (uint32_t q = lo; q < hi; ++q)
                if (off[q + 1] > off[q]) lqh_filter_redundant(&c->ovlp[q], cv.data() + off[q], off[q + 1] - off[q], min_cov);
        });
    }
    c->stats.seeds += ms.n_seeds; c->stats.groups += ms.n_groups; c->stats.chains += ms.n_chains; c->stats.overlaps += ms.n_ovl;
    c->stats.batches += ms.n_batches; c->stats.walk_buckets += ms.n_walk_buckets;
    c->stats.t_post_ms += now_ms() - t0;
    if (c->opt.verbose >= 3) fprintf(stderr, "[M::lqcov] mapped %u sequences (%llu seeds, %llu chains, %llu overlaps)\n", c->nq,
                                    (unsigned long long)ms.n_seeds, (unsigned long long)ms.n_chains, (unsigned long long)ms.n_ovl);
    return 0;
}

/* ---- `-d FILE` and index files as the target argument (index.c:390-479, lq_mmi.cpp) ---- */
extern "C" int lqcov_index_dump(lqcov_ctx *c, const lqcov_reads_t *part, void *file)
{
    LQ_USE_DEV(c);
    if (!c->part_ready || !part->seq || part->seq_on_device) { fprintf(stderr, "[lqcov] ERROR: lqcov_index_dump needs the part just indexed, with its bases in host memory\n"); return -1; }
    LqIndexDev *ix = &c->ix;
    if (ix->rec.wide) { fprintf(stderr, "[lqcov] ERROR: -d with k > 15 is not supported by this build (index files: k <= 15)\n"); return -1; }
    std::vector<uint32_t> counts((size_t)ix->n_keyspace); std::vector<uint64_t> offs((size_t)ix->n_keyspace + 1), pos((size_t)ix->n_rec + 1);
    LQ_CUDA_OK(cudaMemcpyAsync(counts.data(), ix->counts.p, counts.size() * 4, cudaMemcpyDeviceToHost, c->st));
    LQ_CUDA_OK(cudaMemcpyAsync(offs.data(), ix->offs.p, offs.size() * 8, cudaMemcpyDeviceToHost, c->st));
    if (ix->n_rec) LQ_CUDA_OK(cudaMemcpyAsync(pos.data(), ix->rec.y.p, (size_t)ix->n_rec * 8, cudaMemcpyDeviceToHost, c->st));
    LQ_CUDA_OK(cudaStreamSynchronize(c->st));
    return lq_mmi_dump_part((FILE*)file, c->opt.w, c->opt.k, c->opt.is_hpc, part, counts.data(), offs.data(), pos.data());
}

extern "C" int lqcov_index_peek(const char *path, int *k, int *w, int *is_hpc)
{
    const int is = lq_mmi_is_index(path);
    if (is <= 0) return is;
    FILE *fp = fopen(path, "rb");
    uint32_t x[6];
    if (!fp || fread(x, 4, 6, fp) != 6) { if (fp) fclose(fp); return -1; }
    fclose(fp);
    *w = (int)x[1]; *k = (int)x[2]; *is_hpc = (int)(x[5] & 1u);
    return 1;
}

/* the next part of an index file becomes the current part: 1 = loaded, 0 = end of file */
extern "C" int lqcov_load_part(lqcov_ctx *c, void *file)
{
    LQ_USE_DEV(c);
    LqMmiPart mp;
    const int rc = lq_mmi_load_part((FILE*)file, &mp);
    if (rc <= 0) { if (rc < 0) fprintf(stderr, "[lqcov] ERROR: damaged index file\n"); return rc; }
    if (mp.k != c->opt.k || mp.w != c->opt.w || (int)(mp.flag & 1u) != (c->opt.is_hpc ? 1 : 0)) {
        fprintf(stderr, "[lqcov] ERROR: the index part was built with -k %d -w %d%s, the context with -k %d -w %d%s (lqcov_index_peek tells before lqcov_create)\n",
                mp.k, mp.w, mp.flag & 1 ? " -H" : "", c->opt.k, c->opt.w, c->opt.is_hpc ? " -H" : "");
        return -1;
    }
    const double t0 = now_ms();
    LqIndexDev *ix = &c->ix;
    c->part_ready = false; c->use_full = false; c->placed = false;
    const uint64_t n = mp.key.size();
    LQ_TRY(ix->rec.key.ensure((size_t)(n + 1) * 4)); LQ_TRY(ix->rec.y.ensure((size_t)(n + 1) * 8));
    ix->rec.n = n; ix->rec.has_span = 0; ix->rec.wide = 0;
    if (n) {
        LQ_CUDA_OK(cudaMemcpyAsync(ix->rec.key.p, mp.key.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
        LQ_CUDA_OK(cudaMemcpyAsync(ix->rec.y.p, mp.y.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->st));
    }
    LQ_TRY(lq_index_alloc(ix, c->opt.k, 0, c->st));
    LQ_TRY(lq_index_count(ix, &ix->rec, c->st));
    LQ_CUDA_OK(cudaStreamSynchronize(c->st));
    c->stats.t_index_ms += now_ms() - t0;
    lqcov_reads_t part; memset(&part, 0, sizeof part);
    part.n = mp.n_seq; part.seq_off = mp.seq_off.data(); part.names = mp.names.data(); part.name_off = mp.name_off.data();
    c->stats.target_bases += mp.seq_off[mp.n_seq];
    LQ_TRY(lqcov_part_finish(c, &part));
    return 1;
}

/* minimap2-coverage.c:418-427 sketches the queries with the COMMAND LINE's k / w only to size the per-minimizer counters; mapping
 * uses the index's own k / w (index.c:468-470).  When an index file built with other parameters is mapped against, the row's
 * `n` (minimap2-coverage.c:552-563) is still the command line's: this call records those counts. */
extern "C" int lqcov_set_prepass_counts(lqcov_ctx *c, const lqcov_reads_t *q, int k, int w, int is_hpc)
{
    LQ_USE_DEV(c);
    if (q->n != c->nq) return -1;
    LqReadsDev rd; LqMinimizers m; LqDevBuf first; int rc = -1;
    std::vector<uint64_t> hf((size_t)q->n + 1);
    if (lq_reads_upload(&rd, (const uint8_t*)q->seq, q->seq_off, q->n, q->seq_on_device, 0, c->st) == 0 &&
        lq_sketch_run(&rd, w, k, is_hpc, 0, &m, c->ws, c->st) == 0 && lq_read_first(&m, 0, q->n, first, c->st) == 0 &&
        cudaMemcpyAsync(hf.data(), first.p, ((size_t)q->n + 1) * 8, cudaMemcpyDeviceToHost, c->st) == cudaSuccess &&
        cudaStreamSynchronize(c->st) == cudaSuccess) {
        c->n_prepass.resize(q->n);
        for (uint32_t i = 0; i < q->n; ++i) c->n_prepass[i] = (uint32_t)(hf[i + 1] - hf[i]);
        rc = 0;
    }
    rd.release(); m.release(); first.release();
    return rc;
}

extern "C" int lqcov_add_part(lqcov_ctx *c, const lqcov_reads_t *part)
{
    LQ_TRY(lqcov_index_part(c, part));
    return lqcov_map_part(c);
}

extern "C" int lqcov_add_targets(lqcov_ctx *c, const lqcov_reads_t *t)
{
    /* index.c:238-246,316 + bseq.c:82-87: a part is a run of mini-batches, opened while sum_len <= batch_size */
    const uint64_t mini = (uint64_t)c->opt.mini_batch_size < c->opt.batch_size ? (uint64_t)c->opt.mini_batch_size : c->opt.batch_size;
    uint32_t t0 = 0;
    while (t0 < t->n) {
        uint64_t sum_len = 0; uint32_t t1 = t0;
        while (t1 < t->n && !(sum_len > c->opt.batch_size)) {
            uint64_t sz = 0;
            while (t1 < t->n) { const uint64_t L = t->seq_off[t1 + 1] - t->seq_off[t1]; sz += L; sum_len += L; ++t1; if (sz >= mini) break; }
        }
        lqcov_reads_t part = *t;
        part.n = t1 - t0; part.seq_off = t->seq_off + t0; part.name_off = t->name_off + t0;
        LQ_TRY(lqcov_add_part(c, &part));
        t0 = t1;
    }
    return 0;
}

extern "C" int lqcov_table(lqcov_ctx *c, char **buf, size_t *len)
{
    LQ_USE_DEV(c);
    const double t0 = now_ms();
    std::vector<uint32_t> n_match; std::vector<uint64_t> lam(c->nq), lam2(c->nq);
    const bool pre = c->n_prepass.size() == c->nq && c->nq > 0;
    LQ_TRY(lq_map_nmatch(&c->qd, pre ? c->n_prepass.data() : (const uint32_t*)0, &n_match, c->st));
    if (c->nq) {
        LQ_CUDA_OK(cudaMemcpyAsync(lam.data(), c->qd.lambda.p, (size_t)c->nq * 8, cudaMemcpyDeviceToHost, c->st));
        LQ_CUDA_OK(cudaMemcpyAsync(lam2.data(), c->qd.lambda2.p, (size_t)c->nq * 8, cudaMemcpyDeviceToHost, c->st));
        LQ_CUDA_OK(cudaStreamSynchronize(c->st));
    }
    lq_prof_d2h((uint64_t)c->nq * 16); lq_prof_collect();
    for (uint32_t q = 0; q < c->nq; ++q)
        if (pre ? c->n_prepass[q] == 0 : c->qfirst[q + 1] == c->qfirst[q]) { /* the reference divides by zero here (minimap2-coverage.c:558, SIGFPE) */
            fprintf(stderr, "[lqcov] ERROR: query '%s' yields no minimizer; the reference binary crashes on such input\n", c->qname[q].c_str());
            return -1;
        }
    /* rows are independent: contiguous query ranges are formatted by the host threads and concatenated in order */
    (void)lqh_meanQ(NULL, 0);   /* builds the Phred table once, before the threads read it */
    const unsigned nt = host_threads();
    std::vector<lqh_str> piece(nt);
    for (unsigned t = 0; t < nt; ++t) { piece[t].l = piece[t].m = 0; piece[t].s = 0; }
    parallel_chunks(c->nq, nt, [&](uint32_t lo, uint32_t hi, unsigned t) {
        for (uint32_t q = lo; q < hi; ++q)
            lqh_format_row(&piece[t], c->qname[q].data(), c->qname[q].size(), c->qlen[q], c->q_has_qual, c->qsum_p[q], lam[q], lam2[q],
                           pre ? c->n_prepass[q] : (uint32_t)(c->qfirst[q + 1] - c->qfirst[q]), n_match[q], c->avg_k[q], &c->ovlp[q], c->opt.min_coverage, c->opt.filter);
    });
    lqh_str out; out.l = 0; out.m = 1; out.s = 0;
    for (unsigned t = 0; t < nt; ++t) out.m += piece[t].l;
    out.s = (char*)malloc(out.m);
    if (!out.s) { fprintf(stderr, "[lqcov] out of host memory\n"); return -1; }
    for (unsigned t = 0; t < nt; ++t) { if (piece[t].l) memcpy(out.s + out.l, piece[t].s, piece[t].l); out.l += piece[t].l; free(piece[t].s); }
    out.s[out.l] = 0;
    *buf = out.s; *len = out.l;
    c->stats.t_post_ms += now_ms() - t0;
    return 0;
}

extern "C" int lqcov_get_stats(const lqcov_ctx *c, lqcov_stats_t *s) { *s = c->stats; return 0; }

extern "C" int lqcov_sketch(const lqcov_opt_t *o, const lqcov_reads_t *reads, uint32_t rid_base, uint64_t **x, uint64_t **y, uint64_t *n)
{
    LqReadsDev rd; LqMinimizers m; LqDevBuf ws; cudaStream_t st = 0;
    int rc = -1;
    *x = *y = 0; *n = 0;
    if (o->device >= 0 && cudaSetDevice(o->device) != cudaSuccess) { fprintf(stderr, "[lqcov] ERROR: cannot select CUDA device %d\n", o->device); return -1; }
    if (lq_reads_upload(&rd, (const uint8_t*)reads->seq, reads->seq_off, reads->n, reads->seq_on_device, 0, st) == 0 &&
        lq_sketch_run(&rd, o->w, o->k, o->is_hpc, rid_base, &m, ws, st) == 0) {
        std::vector<uint32_t> key(m.wide ? 0 : m.n); std::vector<uint64_t> key64(m.wide ? m.n : 0); std::vector<uint8_t> sp(m.n);
        uint64_t *hy = (uint64_t*)malloc((m.n + 1) * 8), *hx = (uint64_t*)malloc((m.n + 1) * 8);
        bool ok = true;
        if (m.n) {
            ok = (m.wide ? cudaMemcpy(key64.data(), m.key64.p, m.n * 8, cudaMemcpyDeviceToHost) : cudaMemcpy(key.data(), m.key.p, m.n * 4, cudaMemcpyDeviceToHost)) == cudaSuccess &&
                 cudaMemcpy(hy, m.y.p, m.n * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
            if (ok && m.has_span) ok = cudaMemcpy(sp.data(), m.span.p, m.n, cudaMemcpyDeviceToHost) == cudaSuccess;
        }
        if (ok && cudaDeviceSynchronize() == cudaSuccess) {
            for (uint64_t i = 0; i < m.n; ++i) hx[i] = (m.wide ? key64[i] : (uint64_t)key[i]) << 8 | (uint64_t)(m.has_span ? sp[i] : o->k);
            *x = hx; *y = hy; *n = m.n; rc = 0;
        } else { fprintf(stderr, "[lqcov] CUDA error in lqcov_sketch: %s\n", cudaGetErrorString(cudaGetLastError())); free(hx); free(hy); }
    }
    rd.release(); m.release(); ws.release();
    return rc;
}

extern "C" int lqcov_debug_seeds(lqcov_ctx *c, uint32_t q, uint64_t **ux, uint64_t **uy, uint64_t **sx, uint64_t **sy, uint64_t *n)
{
    LQ_USE_DEV(c);
    if (!c->part_ready || q >= c->nq) return -1;
    LqMapOpt mo; map_opt_of(&c->opt, &mo);
    std::vector<lq_mm128> u, s;
    LQ_TRY(lq_map_debug_sorted_seeds(&c->qd, &c->ix, &mo, c->mid_occ, q, c->self_off.data(), c->self_list.data(), c->qrank.data(), c->trank.data(), &c->sc, &u, &s, c->st));
    *n = u.size();
    *ux = (uint64_t*)malloc((u.size() + 1) * 8); *uy = (uint64_t*)malloc((u.size() + 1) * 8);
    *sx = (uint64_t*)malloc((u.size() + 1) * 8); *sy = (uint64_t*)malloc((u.size() + 1) * 8);
    for (size_t i = 0; i < u.size(); ++i) { (*ux)[i] = u[i].x; (*uy)[i] = u[i].y; (*sx)[i] = s[i].x; (*sy)[i] = s[i].y; }
    return 0;
}
