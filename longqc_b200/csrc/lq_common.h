/* lq_common.h -- shared definitions of the B200 minimap2-coverage path (host + device). */
#ifndef LQ_COMMON_H
#define LQ_COMMON_H

#include <stdint.h>
#include <stddef.h>

#ifdef __CUDACC__
#define LQ_HD __host__ __device__ __forceinline__
#define LQ_D __device__ __forceinline__
#else
#define LQ_HD inline
#define LQ_D inline
#endif

/* Packed read store in HBM ------------------------------------------------------------
 * Reads are laid out in SLOTS of 128 bases; a read owns ceil(len/128) consecutive slots, so
 * every read starts on a 32-byte boundary of the 2-bit plane and a 16-byte boundary of the
 * ambiguity plane, and "global base index" g = slot*128 + offset is contiguous inside a read.
 *   b2[g>>4]  bits 2*(g&15)..+1 : base code 0..3 (A,C,G,T/U); 0 where ambiguous / padding
 *   nm[g>>5]  bit  (g&31)       : 1 = ambiguous base (reference code 4) or padding past the read end
 */
#define LQ_SLOT 128
#define LQ_SLOT_W2 8   /* u32 words of the 2-bit plane per slot */
#define LQ_SLOT_WN 4   /* u32 words of the N plane per slot */

#define LQ_MAX_W 32    /* window sizes the GPU path accepts (LongQC uses 5 and 10) */
#define LQ_MAX_K_DIRECT 15 /* direct-address minimizer table: 4^k counters (k=15 -> 4 GiB) */
#define LQ_MAX_K 28        /* as the reference (minimap2-coverage.c:171); 16..28 through the open-address table of lq_widx.cu */

#define LQ_U64MAX 0xffffffffffffffffULL

typedef struct { uint64_t x, y; } lq_mm128; /* same meaning as minimap.h:42 mm128_t */

#endif
