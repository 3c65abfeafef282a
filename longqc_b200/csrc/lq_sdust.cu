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

extern "C" int lqcov_sdust_table(const lqcov_opt_t *o, const lqcov_reads_t *reads, int W, int T, char **buf, size_t *len)
{
    *buf = 0; *len = 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { fprintf(stderr, "[lqcov] ERROR: no usable CUDA device. This library has no CPU path.\n"); return -1; }
    if (o && o->device >= 0 && cudaSetDevice(o->device) != cudaSuccess) { fprintf(stderr, "[lqcov] ERROR: cannot select CUDA device %d\n", o->device); return -1; }
    if (W < 4 || W > 66) { fprintf(stderr, "[lqcov] ERROR: sdust window %d outside 4..66 supported by the GPU path\n", W); return -1; }
    const uint32_t n = reads->n;
    lqh_str out; out.l = out.m = 0; out.s = 0;
    if (n) {
        double h_q2p[127];
        for (int q = 0; q < 127; ++q) h_q2p[q] = lqh_q2p(q); /* lqutils.c:26-49 */
        LQ_CUDA_OK(cudaMemcpyToSymbol(c_q2p, h_q2p, sizeof(h_q2p)));
        const uint64_t nb = reads->seq_off[n] - reads->seq_off[0];
        LqDevBuf d_seq, d_qual, d_off, d_p, d_out, d_cur;
        const uint8_t *dseq;
        std::vector<uint64_t> rel(n + 1);
        for (uint32_t i = 0; i <= n; ++i) rel[i] = reads->seq_off[i] - reads->seq_off[0];
        if (reads->seq_on_device) dseq = (const uint8_t*)reads->seq + reads->seq_off[0];
        else { LQ_TRY(d_seq.ensure(nb + 16)); LQ_CUDA_OK(cudaMemcpy(d_seq.p, reads->seq + reads->seq_off[0], nb, cudaMemcpyHostToDevice)); dseq = d_seq.as<uint8_t>(); }
        if (reads->qual) { LQ_TRY(d_qual.ensure(nb + 16)); LQ_CUDA_OK(cudaMemcpy(d_qual.p, reads->qual + reads->seq_off[0], nb, cudaMemcpyHostToDevice)); }
        LQ_TRY(d_off.ensure(((size_t)n + 1) * 8)); LQ_CUDA_OK(cudaMemcpy(d_off.p, rel.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice));
        const int capP = LQ_SD_PCAP(W);
        const unsigned threads = 64;
        unsigned blocks = (n + threads - 1) / threads; if (blocks > 148 * 4) blocks = 148 * 4;   /* 64 threads x 61 KB of interval scratch each: 2.3 GB at most */
        LQ_TRY(d_p.ensure((size_t)blocks * threads * 4 * capP * sizeof(int)));
        LQ_TRY(d_out.ensure((size_t)n * sizeof(SdOut))); LQ_TRY(d_cur.ensure(64));
        LQ_CUDA_OK(cudaMemset(d_cur.p, 0, 64));
        lq_sdust_k<<<blocks, threads>>>(dseq, reads->qual ? d_qual.as<uint8_t>() : 0, d_off.as<uint64_t>(), n, T, W, d_p.as<int>(), capP, d_cur.as<uint32_t>(), d_out.as<SdOut>());
        LQ_CUDA_OK(cudaGetLastError());
        std::vector<SdOut> h(n);
        LQ_CUDA_OK(cudaMemcpy(h.data(), d_out.p, (size_t)n * sizeof(SdOut), cudaMemcpyDeviceToHost));
        d_seq.release(); d_qual.release(); d_off.release(); d_p.release(); d_out.release(); d_cur.release();
        for (uint32_t i = 0; i < n; ++i) {
            if (h[i].overflow) { fprintf(stderr, "[lqcov] ERROR: sdust interval list overflow on read %u\n", i); free(out.s); return -1; }
            const int L = (int)(rel[i + 1] - rel[i]);
            lqh_format_sdust_row(&out, reads->names + reads->name_off[i], (size_t)(reads->name_off[i + 1] - reads->name_off[i]),
                                 (uint32_t)h[i].masked, L, reads->qual ? reads->qual + reads->seq_off[i] : 0, h[i].sum_p, h[i].q7);
        }
    }
    if (!out.s) { out.s = (char*)malloc(1); out.s[0] = 0; }
    *buf = out.s; *len = out.l;
    return 0;
}
