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
    int wide;              /* k > 15: key64 holds the 2k-bit hashes, key the dense ids lq_widx.cu gives them */
    LqDevBuf key, y, span, blk, key64;
    LqMinimizers() : n(0), has_span(0), wide(0) {}
    void release() { key.release(); y.release(); span.release(); blk.release(); key64.release(); }
};

int lq_reads_upload(LqReadsDev *d, const uint8_t *h_seq, const uint64_t *h_off, uint32_t n_reads, int seq_on_device, int sdust_tbl, cudaStream_t st);
int lq_sketch_run(const LqReadsDev *rd, int w, int k, int is_hpc, uint32_t rid_base, LqMinimizers *out, LqDevBuf &ws, cudaStream_t st);
/* host buffers only: copy, pack and sketch in overlapped chunks; returns 1 when not applicable (use lq_reads_upload + lq_sketch_run) */
int lq_upload_sketch_pipelined(LqReadsDev *d, const uint8_t *h_seq, const uint64_t *h_off, uint32_t n_reads, int w, int k, int is_hpc, uint32_t rid_base,
                               LqMinimizers *out, LqDevBuf &ws, cudaStream_t st);
int lq_read_first(const LqMinimizers *m, uint32_t rid_base, uint32_t n_reads, LqDevBuf &first, cudaStream_t st);

/* One index part (or shard of it) arriving in CHUNKS of whole reads from host staging buffers: every chunk is copied on a second
 * stream while the chunk before it is packed and sketched; the records of all chunks form one contiguous, ordered array, exactly
 * what lq_reads_upload + lq_sketch_run give for the concatenated reads.  Rolling-kernel configurations only (lq_stream_ok). */
#define LQ_STREAM_RING 2
struct LqPartStream {
    LqReadsDev *d; LqMinimizers *out; int w, k; uint32_t rid_base; cudaStream_t st, st_copy;
    LqDevBuf asc[LQ_STREAM_RING]; cudaEvent_t ev_copied[LQ_STREAM_RING], ev_packed[LQ_STREAM_RING]; bool ev_made;
    LqDevBuf state; size_t state_used; LqDevBuf totals; uint32_t n_chunks;
    std::vector<size_t> chunk_state_off; std::vector<uint32_t> chunk_tiles;
    uint64_t cap_rec; bool open;
    void *h_meta[LQ_STREAM_RING]; size_t h_meta_cap[LQ_STREAM_RING];   /* pinned: per-read arrays of a chunk on their way to the device */
    LqPartStream() : d(0), out(0), w(0), k(0), rid_base(0), st(0), st_copy(0), ev_made(false), state_used(0), n_chunks(0), cap_rec(0), open(false) { for (int i = 0; i < LQ_STREAM_RING; ++i) { h_meta[i] = 0; h_meta_cap[i] = 0; } }
    void release();
};
bool lq_stream_ok(int w, int k, int is_hpc);
int lq_stream_begin(LqPartStream *s, LqReadsDev *d, LqMinimizers *out, int w, int k, uint32_t rid_base, uint64_t expect_bases, cudaStream_t st);
/* h_seq: the chunk's bases in host memory (pinned for a truly asynchronous copy), read i at [h_off[i], h_off[i+1]).  Returns once the
 * copy and the kernels are QUEUED; *copied (if not NULL) is an event that fires when h_seq may be overwritten. */
int lq_stream_push(LqPartStream *s, const uint8_t *h_seq, const uint64_t *h_off, uint32_t n_reads, cudaEvent_t *copied);
int lq_stream_end(LqPartStream *s, LqDevBuf &ws);

#endif
