/* lq_index.h -- minimizer index of one target part on the device (see lq_index.cu). */
#ifndef LQ_INDEX_H
#define LQ_INDEX_H
#include "lq_device.h"
#include "lq_widx.h"

struct LqIndexDev {
    int k;
    uint64_t n_keyspace;   /* 4^k; with wide keys (k > 15) the number of distinct keys of the part */
    int key_bits;          /* significant bits of an index address: 2k, or those of n_keyspace with wide keys */
    LqWideTable wide;      /* k > 15: 2k-bit key -> address (lq_widx.cu) */
    uint64_t n_rec;        /* minimizers in the part */
    uint32_t n_seq;        /* target reads in the part */
    LqDevBuf counts;       /* u32[4^k]   occurrences per minimizer key (the all-reduced table) */
    LqDevBuf offs;         /* u64[4^k+1] exclusive scan of counts */
    LqMinimizers rec;      /* rec.y = positions, stable-sorted by key (rec.span likewise in HPC mode) */
    LqDevBuf tlen;         /* u32[n_seq] target read lengths (overhang filter, esterr.c:113) */
    LqDevBuf tmp_key, tmp_y, tmp_sp, hist;
    LqIndexDev() : k(0), n_keyspace(0), key_bits(0), n_rec(0), n_seq(0) {}
    void release() { counts.release(); offs.release(); rec.release(); tlen.release(); tmp_key.release(); tmp_y.release(); tmp_sp.release(); hist.release(); wide.release(); }
};

/* n_ids: 0 = direct addressing by the key (k <= 15); else the address space lq_wide_build() has just numbered */
int lq_index_alloc(LqIndexDev *ix, int k, uint64_t n_ids, cudaStream_t st);
int lq_index_count(LqIndexDev *ix, const LqMinimizers *m, cudaStream_t st);
int lq_index_finish(LqIndexDev *ix, LqMinimizers *m, LqDevBuf &ws, cudaStream_t st);
int lq_index_mid_occ(const LqIndexDev *ix, float frac, int32_t *mid_occ, uint64_t *n_distinct, LqDevBuf &ws, cudaStream_t st);
int lq_sort_by_key(LqMinimizers *m, int key_bits, LqDevBuf &tmp_key, LqDevBuf &tmp_y, LqDevBuf &tmp_sp, LqDevBuf &hist, LqDevBuf &ws, cudaStream_t st, int keep_keys);
#endif
