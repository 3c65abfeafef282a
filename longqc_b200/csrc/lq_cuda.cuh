/* lq_cuda.cuh -- small CUDA utilities shared by the kernels: error checks, growable device
 * buffers, warp/block scans and a device-wide exclusive scan (all hand-written). */
#ifndef LQ_CUDA_CUH
#define LQ_CUDA_CUH

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "lq_common.h"
#include "lq_prof.h"

#define LQ_CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "[lqcov] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
    return -1; } } while (0)
#define LQ_CUDA_OK_V(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "[lqcov] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
    abort(); } } while (0)
#define LQ_TRY(call) do { int r_ = (call); if (r_ != 0) return r_; } while (0)

/* growable device buffer */
struct LqDevBuf {
    void *p; size_t cap;
    LqDevBuf() : p(0), cap(0) {}
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) { cudaFree(p); p = 0; cap = 0; }
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { fprintf(stderr, "[lqcov] cudaMalloc(%zu) failed: %s\n", want, cudaGetErrorString(e)); p = 0; return -1; }
        cap = want; return 0;
    }
    /* grow, keeping the first `keep` bytes (work queued on `st` that touches the old block is waited for) */
    int ensure_keep(size_t bytes, size_t keep, cudaStream_t st) {
        if (bytes <= cap) return 0;
        if (!p || keep == 0) return ensure(bytes);
        void *q = 0; const size_t want = bytes + (bytes >> 1) + 256;
        cudaError_t e = cudaMalloc(&q, want);
        if (e != cudaSuccess) { fprintf(stderr, "[lqcov] cudaMalloc(%zu) failed: %s\n", want, cudaGetErrorString(e)); return -1; }
        if (cudaMemcpyAsync(q, p, keep < cap ? keep : cap, cudaMemcpyDeviceToDevice, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { cudaFree(q); return -1; }
        cudaFree(p); p = q; cap = want; return 0;
    }
    void release() { if (p) cudaFree(p); p = 0; cap = 0; }
    template <class T> T *as() const { return (T*)p; }
};

static inline unsigned lq_grid(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

/* ---- warp / block primitives ---- */
__device__ __forceinline__ uint32_t lq_lane() { return threadIdx.x & 31; }

template <class T> __device__ __forceinline__ T lq_warp_incl_scan(T v)
{
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) { T t = __shfl_up_sync(0xffffffffu, v, d); if (lq_lane() >= (uint32_t)d) v += t; }
    return v;
}
template <class T> __device__ __forceinline__ T lq_warp_sum(T v)
{
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

/* exclusive scan over the block (blockDim.x multiple of 32, <= 1024); returns the exclusive prefix,
 * *total = block sum.  `sm` must hold 33 T's. */
template <class T> __device__ __forceinline__ T lq_block_excl_scan(T v, T *sm, T *total)
{
    const uint32_t lane = lq_lane(), wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T inc = lq_warp_incl_scan(v);
    if (lane == 31) sm[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        T s = lane < nw ? sm[lane] : (T)0;
        T si = lq_warp_incl_scan(s);
        sm[lane] = si - s;
        if (lane == 31) sm[32] = si;
    }
    __syncthreads();
    T r = sm[wid] + inc - v;
    *total = sm[32];
    __syncthreads();
    return r;
}

/* ---- device-wide exclusive scan (three-phase, recursive on the block sums) ---- */
#define LQ_SCAN_ITEMS 8
#define LQ_SCAN_BLOCK 256
#define LQ_SCAN_TILE (LQ_SCAN_ITEMS * LQ_SCAN_BLOCK)

template <class TI, class TO>
__global__ void lq_scan_reduce_k(const TI *in, size_t n, TO *bsum)
{
    __shared__ TO sm[33];
    size_t base = (size_t)blockIdx.x * LQ_SCAN_TILE;
    TO s = 0;
    #pragma unroll
    for (int j = 0; j < LQ_SCAN_ITEMS; ++j) { size_t i = base + (size_t)j * LQ_SCAN_BLOCK + threadIdx.x; if (i < n) s += (TO)in[i]; }
    TO tot; lq_block_excl_scan(s, sm, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

template <class TI, class TO>
__global__ void lq_scan_down_k(const TI *in, size_t n, const TO *bbase, TO *out)
{
    __shared__ TO sm[33];
    size_t base = (size_t)blockIdx.x * LQ_SCAN_TILE + (size_t)threadIdx.x * LQ_SCAN_ITEMS;
    TO v[LQ_SCAN_ITEMS], s = 0;
    #pragma unroll
    for (int j = 0; j < LQ_SCAN_ITEMS; ++j) { size_t i = base + j; v[j] = i < n ? (TO)in[i] : (TO)0; s += v[j]; }
    TO tot; TO ex = lq_block_excl_scan(s, sm, &tot) + bbase[blockIdx.x];
    #pragma unroll
    for (int j = 0; j < LQ_SCAN_ITEMS; ++j) { size_t i = base + j; if (i < n) out[i] = ex; ex += v[j]; }
}

/* out[i] = sum_{j<i} in[j]; out may alias in only if sizeof(TI)==sizeof(TO).  `out` gets n entries
 * (+1: out[n] = total when `with_total`).  ws: scratch buffer (grown on demand). */
template <class TI, class TO>
static int lq_exclusive_scan(const TI *in, TO *out, size_t n, int with_total, LqDevBuf &ws, cudaStream_t st, size_t ws_off = 0)
{
    if (n == 0) { if (with_total) LQ_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(TO), st)); return 0; }
    size_t nb = (n + LQ_SCAN_TILE - 1) / LQ_SCAN_TILE;
    /* workspace layout: level sums one after another; compute total need up-front on the first call */
    if (ws_off == 0) {
        size_t need = 0, m = nb;
        for (;;) { need += (m + 1) * sizeof(TO); if (m <= 1) break; m = (m + LQ_SCAN_TILE - 1) / LQ_SCAN_TILE; }
        need += 64;
        if (ws.ensure(need + 16) != 0) return -1;
    }
    TO *bsum = (TO*)((char*)ws.p + ws_off);
    lq_scan_reduce_k<TI, TO><<<(unsigned)nb, LQ_SCAN_BLOCK, 0, st>>>(in, n, bsum);
    lq_prof_count_launch(2);
    /* exclusive scan of the block sums, in place, total at bsum[nb] */
    if (nb == 1) {
        /* bsum[1] = bsum[0]; bsum[0] = 0 */
        LQ_CUDA_OK(cudaMemcpyAsync(bsum + 1, bsum, sizeof(TO), cudaMemcpyDeviceToDevice, st));
        LQ_CUDA_OK(cudaMemsetAsync(bsum, 0, sizeof(TO), st));
    } else {
        LQ_TRY((lq_exclusive_scan<TO, TO>(bsum, bsum, nb, 1, ws, st, ws_off + (nb + 1) * sizeof(TO))));
    }
    lq_scan_down_k<TI, TO><<<(unsigned)nb, LQ_SCAN_BLOCK, 0, st>>>(in, n, bsum, out);
    if (with_total) LQ_CUDA_OK(cudaMemcpyAsync(out + n, bsum + nb, sizeof(TO), cudaMemcpyDeviceToDevice, st));
    LQ_CUDA_OK(cudaGetLastError());
    return 0;
}

#endif
