/* lq_widx.h -- open-address table from wide minimizer keys (16 <= k <= 28) to dense ids (see lq_widx.cu) */
#ifndef LQ_WIDX_H
#define LQ_WIDX_H
#include "lq_cuda.cuh"

struct LqWideTable {
    LqDevBuf slots, vals, ctr;   /* u64[cap] keys (all ones = empty), u32[cap] ids, ctr[0] = number of ids */
    uint64_t mask; uint32_t n_ids;
    LqWideTable() : mask(0), n_ids(0) {}
    void release() { slots.release(); vals.release(); ctr.release(); }
};

/* all keys inserted, d_ids[i] = id of d_key64[i]; t->n_ids = number of distinct keys (synchronises the stream) */
int lq_wide_build(LqWideTable *t, const uint64_t *d_key64, uint64_t n, uint32_t *d_ids, cudaStream_t st);
/* look-ups only: absent keys get id n_ids */
int lq_wide_translate(const LqWideTable *t, const uint64_t *d_key64, uint64_t n, uint32_t *d_ids, cudaStream_t st);
#endif
