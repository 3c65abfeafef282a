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

/* ---- the segment form (lq_sdust_core.h): a thread per SDS_SEG steps of a read.  It cold-starts far enough before
 *      its segment, keeps the interval list in local memory and leaves the merged runs of the intervals given up during its steps;
 *      lq_sdust_merge_k folds the runs of a read's segments in order (the reference's merge is order-dependent once an ambiguous base
 *      has left stale triplets in the deque: it is NOT the union of the intervals).
 *      A repeat of a few dozen bases already holds hundreds of live intervals (every suffix of the window scores above the one before
 *      it, sdust.c:112-121), up to (W-2)^2: a segment whose list outgrows SDS_PCAP, or which leaves more than SDS_RUNS runs, is put on
 *      a list and scanned again by lq_sdust_redo_k with the full-size list in a global scratch slice and room for SDS_BIGRUNS runs. ---- */
#define SDS_SEG 512
#define SDS_PCAP 96
#define SDS_RUNS 4
#define SDS_BIGRUNS 128  /* runs are disjoint, >= 4 bases long (two equal triplets before r > 0) and lie within SEG + W + 2 bases: at most 116 */
struct SdSegOut { int n, redo; lq_sd_run r[SDS_RUNS]; };   /* redo: 1 + index into the redo list, 0 = runs are here */

struct SdSegAt { uint32_t r; int lo, hi, L; uint64_t b; };
__device__ __forceinline__ SdSegAt sd_seg_at(uint32_t g, const uint64_t *__restrict__ off, const uint32_t *__restrict__ seg0, uint32_t n_reads)
{
    uint32_t lo = 0, hi = n_reads;
    while (hi - lo > 1) { const uint32_t mid = lo + ((hi - lo) >> 1); if (seg0[mid] <= g) lo = mid; else hi = mid; }
    SdSegAt a; a.r = lo; a.b = off[lo]; a.L = (int)(off[lo + 1] - a.b);
    a.lo = (int)((g - seg0[lo]) * SDS_SEG); a.hi = a.lo + SDS_SEG > a.L ? 0x7fffffff : a.lo + SDS_SEG;   /* the last segment holds the end-of-read flush */
    return a;
}

__global__ void __launch_bounds__(128) lq_sdust_seg_k(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ off, const uint32_t *__restrict__ seg0, uint32_t n_reads,
                                                      uint32_t n_segs, int T, int W, SdSegOut *__restrict__ out, uint32_t *__restrict__ redo, uint32_t *__restrict__ n_redo)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_segs) return;
    const SdSegAt a = sd_seg_at(g, off, seg0, n_reads);
    int pbuf[4 * SDS_PCAP];
    SdSegOut o; int ov = 0;
    o.n = lq_sdust_segment(seq + a.b, a.L, T, W, a.lo, a.hi, pbuf, SDS_PCAP, o.r, SDS_RUNS, &ov);
    o.redo = 0;
    if (ov || o.n > SDS_RUNS) { const uint32_t at = atomicAdd(n_redo, 1u); redo[at] = g; o.redo = (int)at + 1; o.n = 0; }
    out[g] = o;
}

/* the segments on the redo list, a thread each, interval list of (W-2)^2+2 entries in global memory; big[at] takes the runs */
__global__ void lq_sdust_redo_k(const uint8_t *__restrict__ seq, const uint64_t *__restrict__ off, const uint32_t *__restrict__ seg0, uint32_t n_reads, int T, int W,
                                const uint32_t *__restrict__ redo, const uint32_t *__restrict__ n_redo, int *__restrict__ pbuf, int capP, uint32_t *__restrict__ cursor,
                                lq_sd_run *__restrict__ big, int *__restrict__ big_n, int *__restrict__ fail)
{
    int *my = pbuf + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 4 * capP;
    const uint32_t n = *n_redo;
    for (;;) {
        const uint32_t at = atomicAdd(cursor, 1u);
        if (at >= n) break;
        const SdSegAt a = sd_seg_at(redo[at], off, seg0, n_reads);
        int ov = 0;
        const int nr = lq_sdust_segment(seq + a.b, a.L, T, W, a.lo, a.hi, my, capP, big + (size_t)at * SDS_BIGRUNS, SDS_BIGRUNS, &ov);
        big_n[at] = nr;
        if (ov || nr > SDS_BIGRUNS) *fail = 1;
    }
}

__global__ void lq_sdust_merge_k(uint32_t n_reads, const uint32_t *__restrict__ seg0, const SdSegOut *__restrict__ segs, const lq_sd_run *__restrict__ big,
                                 const int *__restrict__ big_n, const int *__restrict__ fail, SdOut *__restrict__ out)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    lq_sd_merge_sink m; m.init();
    for (uint32_t g = seg0[r]; g < seg0[r + 1]; ++g) {
        const int redo = segs[g].redo;
        if (redo) { const lq_sd_run *q = big + (size_t)(redo - 1) * SDS_BIGRUNS; const int n = min(big_n[redo - 1], SDS_BIGRUNS); for (int i = 0; i < n; ++i) m.add(q[i].s, q[i].f); }
        else { const int n = segs[g].n; for (int i = 0; i < n; ++i) m.add(segs[g].r[i].s, segs[g].r[i].f); }
    }
    out[r].masked = m.total(); out[r].overflow = *fail;
}

/* mean quality and #Q>7 of every read: strictly ordered double sum (lqutils.c:54-56), one thread per read */
__global__ void lq_sdust_qual_k(const uint8_t *__restrict__ qual, const uint64_t *__restrict__ off, uint32_t n_reads, SdOut *__restrict__ out)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    double acc = 0.0; int n7 = 0;
    if (qual) {
        for (uint64_t i = off[r]; i < off[r + 1]; ++i) {
            const int q = (int)(signed char)qual[i] - 33;
            acc += c_q2p[q < 0 ? 0 : q > 126 ? 126 : q];   /* lqutils.c:55, in read order */
            n7 += (int)(signed char)qual[i] > 7 + 33;      /* lqutils.c:72-80 */
        }
    }
    out[r].sum_p = acc; out[r].q7 = n7;
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
    LqDevBuf d_seq[2], d_qual[2], d_off[2], d_out[2], d_seg0[2], d_segs[2], d_redo[2], d_big[2], d_bign[2], d_p, d_cur;
    SdOut *h_out[2]; size_t h_out_cap[2];   /* pinned: results come back without stalling the caller */
    cudaEvent_t ev[2]; bool ev_made;
    unsigned blocks, threads; int capP; uint64_t n_chunks;
    char *stage_seq[2], *stage_qual[2]; size_t stage_bytes;
};

static int sd_setup(lqcov_sdust *s)
{
    double h_q2p[127];
    for (int q = 0; q < 127; ++q) h_q2p[q] = lqh_q2p(q); /* lqutils.c:26-49 */
    LQ_CUDA_OK(cudaMemcpyToSymbol(c_q2p, h_q2p, sizeof(h_q2p)));
    s->capP = LQ_SD_PCAP(s->W); s->threads = 64; s->blocks = 148 * 8;   /* the redo kernel: 64 threads x 61 KB of interval scratch each (4.7 GB) */
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
    s->W = W; s->T = T; s->device = o ? o->device : -1; s->ev_made = false; s->n_chunks = 0; s->stage_bytes = 0; s->h_out[0] = s->h_out[1] = 0; s->h_out_cap[0] = s->h_out_cap[1] = 0;
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
struct SdPending { std::vector<uint64_t> rel; std::vector<uint32_t> seg0; std::string names; std::vector<uint64_t> name_off; const char *qual; bool has_qual; uint32_t n; const uint8_t *dseq; };

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
    sd_format(p, s->h_out[slot], rows);
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
        if (s->h_out_cap[slot] < n) { if (s->h_out[slot]) cudaFreeHost(s->h_out[slot]); s->h_out[slot] = 0; s->h_out_cap[slot] = (size_t)n + n / 4 + 1024;
                                      LQ_CUDA_OK(cudaHostAlloc((void**)&s->h_out[slot], s->h_out_cap[slot] * sizeof(SdOut), cudaHostAllocDefault)); }
        const uint8_t *dseq = s->d_seq[slot].as<uint8_t>();
        if (reads->seq_on_device) dseq = (const uint8_t*)reads->seq + reads->seq_off[0];
        else LQ_CUDA_OK(cudaMemcpyAsync(s->d_seq[slot].p, reads->seq + reads->seq_off[0], nb, cudaMemcpyHostToDevice, s->st));
        if (reads->qual) LQ_CUDA_OK(cudaMemcpyAsync(s->d_qual[slot].p, reads->qual + reads->seq_off[0], nb, cudaMemcpyHostToDevice, s->st));
        p.dseq = dseq;
        LQ_CUDA_OK(cudaMemcpyAsync(s->d_off[slot].p, p.rel.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, s->st));
        /* segments: floor(L / SDS_SEG) + 1 per read (the last one holds the end-of-read flush) */
        p.seg0.resize((size_t)n + 1);
        uint64_t ns = 0;
        for (uint32_t i = 0; i < n; ++i) { p.seg0[i] = (uint32_t)ns; ns += (p.rel[i + 1] - p.rel[i]) / SDS_SEG + 1; }
        p.seg0[n] = (uint32_t)ns;
        if (ns >= 0xffffffffULL) { fprintf(stderr, "[lqcov] ERROR: sdust chunk too large\n"); return -1; }
        LQ_TRY(s->d_seg0[slot].ensure(((size_t)n + 1) * 4)); LQ_TRY(s->d_segs[slot].ensure((size_t)ns * sizeof(SdSegOut))); LQ_TRY(s->d_redo[slot].ensure(((size_t)ns + 4) * 4));
        LQ_TRY(s->d_big[slot].ensure((size_t)ns * SDS_BIGRUNS * sizeof(lq_sd_run))); LQ_TRY(s->d_bign[slot].ensure((size_t)ns * 4));
        LQ_CUDA_OK(cudaMemcpyAsync(s->d_seg0[slot].p, p.seg0.data(), ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, s->st));
        uint32_t *ctr = s->d_redo[slot].as<uint32_t>() + ns;   /* [0] segments to redo, [1] cursor of the redo kernel, [2] failure flag */
        LQ_CUDA_OK(cudaMemsetAsync(ctr, 0, 16, s->st));
        { LqProfScope ps("sdust", s->st, 4, nb * (reads->qual ? 2 : 1));
          lq_sdust_seg_k<<<lq_grid(ns, 128), 128, 0, s->st>>>(dseq, s->d_off[slot].as<uint64_t>(), s->d_seg0[slot].as<uint32_t>(), n, (uint32_t)ns, s->T, s->W, s->d_segs[slot].as<SdSegOut>(),
                                                           s->d_redo[slot].as<uint32_t>(), ctr);
          lq_sdust_redo_k<<<s->blocks, s->threads, 0, s->st>>>(dseq, s->d_off[slot].as<uint64_t>(), s->d_seg0[slot].as<uint32_t>(), n, s->T, s->W, s->d_redo[slot].as<uint32_t>(), ctr,
                                                            s->d_p.as<int>(), s->capP, ctr + 1, s->d_big[slot].as<lq_sd_run>(), s->d_bign[slot].as<int>(), (int*)(ctr + 2));
          lq_sdust_merge_k<<<lq_grid(n, 128), 128, 0, s->st>>>(n, s->d_seg0[slot].as<uint32_t>(), s->d_segs[slot].as<SdSegOut>(), s->d_big[slot].as<lq_sd_run>(), s->d_bign[slot].as<int>(),
                                                           (const int*)(ctr + 2), s->d_out[slot].as<SdOut>());
          lq_sdust_qual_k<<<lq_grid(n, 64), 64, 0, s->st>>>(reads->qual ? s->d_qual[slot].as<uint8_t>() : 0, s->d_off[slot].as<uint64_t>(), n, s->d_out[slot].as<SdOut>()); }
        LQ_CUDA_OK(cudaGetLastError());
        LQ_CUDA_OK(cudaMemcpyAsync(s->h_out[slot], s->d_out[slot].p, (size_t)n * sizeof(SdOut), cudaMemcpyDeviceToHost, s->st));
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
    for (int i = 0; i < 2; ++i) { s->d_seq[i].release(); s->d_qual[i].release(); s->d_off[i].release(); s->d_out[i].release(); s->d_seg0[i].release(); s->d_segs[i].release(); s->d_redo[i].release(); s->d_big[i].release(); s->d_bign[i].release(); if (s->stage_seq[i]) cudaFreeHost(s->stage_seq[i]); if (s->stage_qual[i]) cudaFreeHost(s->stage_qual[i]); if (s->h_out[i]) cudaFreeHost(s->h_out[i]); }
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
