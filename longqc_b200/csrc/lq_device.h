/* lq_device.h -- device-side data structures shared by the .cu translation units (C++). */
#ifndef LQ_DEVICE_H
#define LQ_DEVICE_H

#include <vector>
#include "lq_cuda.cuh"

/* a read set packed in HBM (see lq_common.h for the slot layout) */
struct LqReadsDev {
    uint32_t n_reads; uint64_t n_bases, n_slots;
    LqDevBuf ascii, off, len, slot0, slot_read, b2, nm;
    std::vector<uint32_t> h_len; std::vector<uint64_t> h_slot0;
    LqReadsDev() : n_reads(0), n_bases(0), n_slots(0) {}
    void release() { ascii.release(); off.release(); len.release(); slot0.release(); slot_read.release(); b2.release(); nm.release(); }
};

/* minimizers of a read set, ordered by (read, position): key = 2k-bit hash (x>>8 of the reference record),
 * y = rid<<32 | lastPos<<1 | strand (minimap.h:42 / sketch.c:70-72); span only in HPC mode (else == k) */
struct LqMinimizers {
    uint64_t n; int has_span;
    LqDevBuf key, y, span, blk;
    LqMinimizers() : n(0), has_span(0) {}
    void release() { key.release(); y.release(); span.release(); blk.release(); }
};

int lq_reads_upload(LqReadsDev *d, const uint8_t *h_seq, const uint64_t *h_off, uint32_t n_reads, int seq_on_device, int sdust_tbl, cudaStream_t st);
int lq_sketch_run(const LqReadsDev *rd, int w, int k, int is_hpc, uint32_t rid_base, LqMinimizers *out, LqDevBuf &ws, cudaStream_t st);
/* host buffers only: copy, pack and sketch in overlapped chunks; returns 1 when not applicable (use lq_reads_upload + lq_sketch_run) */
int lq_upload_sketch_pipelined(LqReadsDev *d, const uint8_t *h_seq, const uint64_t *h_off, uint32_t n_reads, int w, int k, int is_hpc, uint32_t rid_base,
                               LqMinimizers *out, LqDevBuf &ws, cudaStream_t st);
int lq_read_first(const LqMinimizers *m, uint32_t rid_base, uint32_t n_reads, LqDevBuf &first, cudaStream_t st);

#endif
