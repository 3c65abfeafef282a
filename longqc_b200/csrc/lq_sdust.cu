/* lq_sdust.cu -- K8: the `sdust` table (reference sdust.c:187-223): per read the DUST-masked length
 * (T=20, W=64), the mean base quality and the number of bases above Q7.
 *
 * One persistent thread per read (grid-stride): the DUST scan is a short-range state machine
 * (lq_sdust_core.h) and the mean quality is a strictly ordered double sum (lqutils.c:54-56: the order
 * of the additions fixes the last bits, and the table prints %.3f), so both run sequentially per read
 * and in parallel across the reads of the batch.  Each thread owns a slice of a global scratch buffer
 * for the perfect-interval list (up to W*W/2 entries on low-complexity reads).
 */
#include <vector>
#include <string>
#include <math.h>
#include "lq_cuda.cuh"
#include "lq_sdust_core.h"
#include "lq_host.h"
#include "lqcov.h"

__constant__ double c_q2p[127];

struct SdOut { int64_t masked; double sum_p; int32_t q7; int32_t overflow; };

__global__ void lq_sdust_k(const uint8_t *__restrict__ seq, const uint8_t *__restrict__ qual, const uint64_t *__restrict__ off, uint32_t n_reads,
                           int T, int W, int *__restrict__ pbuf, int capP, uint32_t *__restrict__ cursor, SdOut *__restrict__ out)
{
    int *my = pbuf + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 4 * capP;
    for (;;) {
        const uint32_t r = atomicAdd(cursor, 1u);
        if (r >= n_reads) break;
        const uint64_t b = off[r]; const int L = (int)(off[r + 1] - b);
        SdOut o; int ov = 0;
        o.masked = lq_sdust_masked(seq + b, L, T, W, my, capP, &ov);
        o.overflow = ov; o.sum_p = 0.0; o.q7 = 0;
        if (qual) {
            double acc = 0.0; int n7 = 0;
            for (int i = 0; i < L; ++i) {
                const int q = (int)(signed char)qual[b + i] - 33;
                acc += c_q2p[q < 0 ? 0 : q > 126 ? 126 : q];   /* lqutils.c:55, in read order */
                n7 += (int)(signed char)qual[b + i] > 7 + 33;  /* lqutils.c:72-80 */
            }
            o.sum_p = acc; o.q7 = n7;
        }
        out[r] = o;
    }
}

/* per read: the ordered sum of error probabilities (lqutils.c:54-56) -- also used for the query rows of the coverage table */
__global__ void lq_qualsum_k(const uint8_t *__restrict__ qual, const uint64_t *__restrict__ off, uint32_t n_reads, double *__restrict__ sum_p)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    double acc = 0.0;
    for (uint64_t i = off[r]; i < off[r + 1]; ++i) { const int q = (int)(signed char)qual[i] - 33; acc += c_q2p[q < 0 ? 0 : q > 126 ? 126 : q]; }
    sum_p[r] = acc;
}

int lq_qualsum_run(const uint8_t *d_qual, const uint64_t *d_off, uint32_t n_reads, double *d_sum, cudaStream_t st)
{
    double h_q2p[127];
    for (int q = 0; q < 127; ++q) h_q2p[q] = lqh_q2p(q);
    LQ_CUDA_OK(cudaMemcpyToSymbolAsync(c_q2p, h_q2p, sizeof(h_q2p), 0, cudaMemcpyHostToDevice, st));
    if (n_reads) lq_qualsum_k<<<lq_grid(n_reads, 64), 64, 0, st>>>(d_qual, d_off, n_reads, d_sum);
    LQ_CUDA_OK(cudaGetLastError());
    lq_prof_count_launch(1);
    return 0;
}

/* ---- the sdust table, chunk by chunk: the copy of chunk c+1 into the device overlaps the kernel of chunk c, rows are formatted by the
 *      host threads.  Used with host buffers (lqcov_sdust_table) and by the executable, whose reader threads fill pinned staging
 *      buffers chunk after chunk (lqcov_sdust_begin / _chunk / _end). ---- */
#include <thread>
#include "lq_ingest.h"
#define SD_CHUNK ((size_t)48 << 20)

struct lqcov_sdust {
    int W, T, device; cudaStream_t st;
    LqDevBuf d_seq[2], d_qual[2], d_off[2], d_out[2], d_p, d_cur;
    std::vector<SdOut> h_out[2]; cudaEvent_t ev[2]; bool ev_made;
    unsigned blocks, threads; int capP; uint64_t n_chunks;
    char *stage_seq[2], *stage_qual[2]; size_t stage_bytes;
};

static int sd_setup(lqcov_sdust *s)
{
    double h_q2p[127];
    for (int q = 0; q < 127; ++q) h_q2p[q] = lqh_q2p(q); /* lqutils.c:26-49 */
    LQ_CUDA_OK(cudaMemcpyToSymbol(c_q2p, h_q2p, sizeof(h_q2p)));
    s->capP = LQ_SD_PCAP(s->W); s->threads = 64; s->blocks = 148 * 4;   /* 64 threads x 61 KB of interval scratch each: 2.3 GB at most */
    LQ_TRY(s->d_p.ensure((size_t)s->blocks * s->threads * 4 * s->capP * sizeof(int)));
    LQ_TRY(s->d_cur.ensure(64));
    LQ_CUDA_OK(cudaStreamCreate(&s->st));
    for (int i = 0; i < 2; ++i) LQ_CUDA_OK(cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming));
    s->ev_made = true;
    return 0;
}

extern "C" lqcov_sdust *lqcov_sdust_begin(const lqcov_opt_t *o, int W, int T, size_t stage_bytes, char **stage_seq, char **stage_qual)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { fprintf(stderr, "[lqcov] ERROR: no usable CUDA device. This library has no CPU path.\n"); return 0; }
    if (o && o->device >= 0 && cudaSetDevice(o->device) != cudaSuccess) { fprintf(stderr, "[lqcov] ERROR: cannot select CUDA device %d\n", o->device); return 0; }
    if (W < 4 || W > 66) { fprintf(stderr, "[lqcov] ERROR: sdust window %d outside 4..66 supported by the GPU path\n", W); return 0; }
    lqcov_sdust *s = new lqcov_sdust();
    s->W = W; s->T = T; s->device = o ? o->device : -1; s->ev_made = false; s->n_chunks = 0; s->stage_bytes = 0;
    for (int i = 0; i < 2; ++i) { s->stage_seq[i] = s->stage_qual[i] = 0; }
    if (sd_setup(s) != 0) { delete s; return 0; }
    if (stage_bytes) {   /* two pinned (sequence, quality) buffer pairs for the caller's reader to fill */
        s->stage_bytes = stage_bytes;
        for (int i = 0; i < 2; ++i) {
            if (cudaHostAlloc((void**)&s->stage_seq[i], stage_bytes, cudaHostAllocDefault) != cudaSuccess || cudaHostAlloc((void**)&s->stage_qual[i], stage_bytes, cudaHostAllocDefault) != cudaSuccess) {
                fprintf(stderr, "[lqcov] ERROR: cannot page-lock the sdust staging buffers\n"); delete s; return 0; }
            stage_seq[i] = s->stage_seq[i]; stage_qual[i] = s->stage_qual[i];
        }
    }
    return s;
}

/* queue one chunk (reads with offsets relative to seq / qual); the rows of the chunk BEFORE it are appended to *rows.  The caller may
 * overwrite the chunk's host buffers after the NEXT call (or lqcov_sdust_end) returns. */
struct SdPending { std::vector<uint64_t> rel; std::string names; std::vector<uint64_t> name_off; const char *qual; bool has_qual; uint32_t n; };

static void sd_format(const SdPending &p, const SdOut *h, lqh_str *out)
{
    unsigned nt = std::thread::hardware_concurrency(); if (nt < 1) nt = 1; if (nt > 16) nt = 16;
    if (p.n < 4096) nt = 1;
    std::vector<lqh_str> piece(nt);
    for (unsigned t = 0; t < nt; ++t) { piece[t].l = piece[t].m = 0; piece[t].s = 0; }
    (void)lqh_meanQ(NULL, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) {
        const uint32_t lo = (uint32_t)((uint64_t)p.n * t / nt), hi = (uint32_t)((uint64_t)p.n * (t + 1) / nt);
        th.emplace_back([&, lo, hi, t]() {
            for (uint32_t i = lo; i < hi; ++i) {
                const int L = (int)(p.rel[i + 1] - p.rel[i]);
                lqh_format_sdust_row(&piece[t], p.names.data() + p.name_off[i], (size_t)(p.name_off[i + 1] - p.name_off[i]),
                                     (uint32_t)h[i].masked, L, p.has_qual ? "q" : 0, h[i].sum_p, h[i].q7);
            }
        });
    }
    for (size_t t = 0; t < th.size(); ++t) th[t].join();
    for (unsigned t = 0; t < nt; ++t) {
        if (piece[t].l) { if (out->l + piece[t].l + 1 > out->m) { out->m = (out->l + piece[t].l + 1) * 2; out->s = (char*)realloc(out->s, out->m); } memcpy(out->s + out->l, piece[t].s, piece[t].l); out->l += piece[t].l; }
        free(piece[t].s);
    }
}

struct lqcov_sdust_state { SdPending pend[2]; bool have[2]; };
static lqcov_sdust_state g_sd_state;   /* one sdust run at a time per process (the executable, the bench) */

static int sd_collect(lqcov_sdust *s, int slot, lqh_str *rows)
{
    if (!g_sd_state.have[slot]) return 0;
    LQ_CUDA_OK(cudaEventSynchronize(s->ev[slot]));
    const SdPending &p = g_sd_state.pend[slot];
    for (uint32_t i = 0; i < p.n; ++i)
        if (s->h_out[slot][i].overflow) { fprintf(stderr, "[lqcov] ERROR: sdust interval list overflow on a read\n"); return -1; }
    sd_format(p, s->h_out[slot].data(), rows);
    g_sd_state.have[slot] = false;
    return 0;
}

extern "C" int lqcov_sdust_chunk(lqcov_sdust *s, const lqcov_reads_t *reads, char **rows, size_t *rows_len)
{
    if (s->device >= 0) LQ_CUDA_OK(cudaSetDevice(s->device));
    const int slot = (int)(s->n_chunks & 1);
    lqh_str out; out.l = out.m = 0; out.s = 0;
    /* the slot's previous chunk (two chunks ago) has been collected by the call before this one */
    const uint32_t n = reads->n;
    SdPending &p = g_sd_state.pend[slot];
    p.n = n; p.rel.resize((size_t)n + 1); p.name_off.resize((size_t)n + 1); p.has_qual = reads->qual != 0;
    for (uint32_t i = 0; i <= n; ++i) { p.rel[i] = reads->seq_off[i] - reads->seq_off[0]; p.name_off[i] = reads->name_off[i] - reads->name_off[0]; }
    p.names.assign(reads->names + reads->name_off[0], (size_t)p.name_off[n]);
    if (n) {
        const uint64_t nb = p.rel[n];
        LQ_TRY(s->d_seq[slot].ensure(nb + 16)); LQ_TRY(s->d_off[slot].ensure(((size_t)n + 1) * 8)); LQ_TRY(s->d_out[slot].ensure((size_t)n * sizeof(SdOut)));
        if (reads->qual) LQ_TRY(s->d_qual[slot].ensure(nb + 16));
        s->h_out[slot].resize(n);
        const uint8_t *dseq = s->d_seq[slot].as<uint8_t>();
        if (reads->seq_on_device) dseq = (const uint8_t*)reads->seq + reads->seq_off[0];
        else LQ_CUDA_OK(cudaMemcpyAsync(s->d_seq[slot].p, reads->seq + reads->seq_off[0], nb, cudaMemcpyHostToDevice, s->st));
        if (reads->qual) LQ_CUDA_OK(cudaMemcpyAsync(s->d_qual[slot].p, reads->qual + reads->seq_off[0], nb, cudaMemcpyHostToDevice, s->st));
        LQ_CUDA_OK(cudaMemcpyAsync(s->d_off[slot].p, p.rel.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, s->st));
        LQ_CUDA_OK(cudaMemsetAsync(s->d_cur.p, 0, 64, s->st));
        unsigned blocks = (n + s->threads - 1) / s->threads; if (blocks > s->blocks) blocks = s->blocks;
        { LqProfScope ps("sdust", s->st, 1, nb * (reads->qual ? 2 : 1));
          lq_sdust_k<<<blocks, s->threads, 0, s->st>>>(dseq, reads->qual ? s->d_qual[slot].as<uint8_t>() : 0, s->d_off[slot].as<uint64_t>(), n, s->T, s->W, s->d_p.as<int>(), s->capP,
                                                      s->d_cur.as<uint32_t>(), s->d_out[slot].as<SdOut>()); }
        LQ_CUDA_OK(cudaGetLastError());
        LQ_CUDA_OK(cudaMemcpyAsync(s->h_out[slot].data(), s->d_out[slot].p, (size_t)n * sizeof(SdOut), cudaMemcpyDeviceToHost, s->st));
    }
    LQ_CUDA_OK(cudaEventRecord(s->ev[slot], s->st));
    g_sd_state.have[slot] = true;
    ++s->n_chunks;
    /* while this chunk runs, the rows of the one before it are formatted */
    if (sd_collect(s, slot ^ 1, &out) != 0) { free(out.s); return -1; }
    if (!out.s) { out.s = (char*)malloc(1); }
    out.s[out.l] = 0;
    *rows = out.s; *rows_len = out.l;
    return 0;
}

extern "C" int lqcov_sdust_end(lqcov_sdust *s, char **rows, size_t *rows_len)
{
    lqh_str out; out.l = out.m = 0; out.s = 0;
    int rc = 0;
    if (s->device >= 0) cudaSetDevice(s->device);
    const int last = (int)((s->n_chunks + 1) & 1);
    if (sd_collect(s, last, &out) != 0 || sd_collect(s, last ^ 1, &out) != 0) rc = -1;
    if (!out.s) out.s = (char*)malloc(1);
    out.s[out.l] = 0;
    *rows = out.s; *rows_len = out.l;
    for (int i = 0; i < 2; ++i) { s->d_seq[i].release(); s->d_qual[i].release(); s->d_off[i].release(); s->d_out[i].release(); if (s->stage_seq[i]) cudaFreeHost(s->stage_seq[i]); if (s->stage_qual[i]) cudaFreeHost(s->stage_qual[i]); }
    s->d_p.release(); s->d_cur.release();
    if (s->ev_made) for (int i = 0; i < 2; ++i) cudaEventDestroy(s->ev[i]);
    cudaStreamDestroy(s->st);
    g_sd_state.have[0] = g_sd_state.have[1] = false;
    delete s;
    return rc;
}

extern "C" int lqcov_sdust_table(const lqcov_opt_t *o, const lqcov_reads_t *reads, int W, int T, char **buf, size_t *len)
{
    *buf = 0; *len = 0;
    lqcov_sdust *s = lqcov_sdust_begin(o, W, T, 0, 0, 0);
    if (!s) return -1;
    lqh_str all; all.l = all.m = 0; all.s = 0;
    int rc = 0;
    uint32_t i0 = 0;
    while (rc == 0 && i0 < reads->n) {   /* chunks of whole reads, ~48 MB of bases each */
        uint32_t i1 = i0; uint64_t nb = 0;
        while (i1 < reads->n && (nb < SD_CHUNK || i1 == i0)) { nb += reads->seq_off[i1 + 1] - reads->seq_off[i1]; ++i1; }
        lqcov_reads_t sub = *reads;
        sub.n = i1 - i0; sub.seq_off = reads->seq_off + i0; sub.name_off = reads->name_off + i0;
        char *rows = 0; size_t rl = 0;
        if (lqcov_sdust_chunk(s, &sub, &rows, &rl) != 0) rc = -1;
        if (rl) { if (all.l + rl + 1 > all.m) { all.m = (all.l + rl + 1) * 2; all.s = (char*)realloc(all.s, all.m); } memcpy(all.s + all.l, rows, rl); all.l += rl; }
        free(rows);
        i0 = i1;
    }
    char *rows = 0; size_t rl = 0;
    if (lqcov_sdust_end(s, &rows, &rl) != 0) rc = -1;
    if (rl) { if (all.l + rl + 1 > all.m) { all.m = (all.l + rl + 1) * 2; all.s = (char*)realloc(all.s, all.m); } memcpy(all.s + all.l, rows, rl); all.l += rl; }
    free(rows);
    if (!all.s) all.s = (char*)malloc(1);
    all.s[all.l] = 0;
    if (rc != 0) { free(all.s); return -1; }
    *buf = all.s; *len = all.l;
    return 0;
}
