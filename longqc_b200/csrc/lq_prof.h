/* lq_prof.h -- optional per-kernel device timing (CUDA events on the launching stream) and transfer /
 * launch counters.  Off by default; bench.py switches it on to report the dominant kernel's measured
 * duration next to its algorithmic bytes (roofline) and the launch count of the timed region. */
#ifndef LQ_PROF_H
#define LQ_PROF_H
#include <cuda_runtime.h>
#include <stdint.h>

struct LqKernelStat { const char *name; double ms; uint64_t launches; uint64_t bytes; /* algorithmic bytes moved */ };

void lq_prof_enable(int on);
int  lq_prof_on();
/* bracket a group of launches that make up one named kernel (name must be a string literal) */
void lq_prof_begin(const char *name, cudaStream_t st);
void lq_prof_end(cudaStream_t st, uint64_t launches, uint64_t algorithmic_bytes);
void lq_prof_add_bytes(const char *name, uint64_t bytes);   /* bytes only known after the launch (device-side counters) */
void lq_prof_count_launch(uint64_t n);           /* launches outside begin/end brackets */
void lq_prof_h2d(uint64_t bytes);
void lq_prof_d2h(uint64_t bytes);
void lq_prof_collect();                           /* after a stream sync: fold finished event pairs into the table */

struct LqProfScope {
    cudaStream_t st; uint64_t launches, bytes;
    LqProfScope(const char *name, cudaStream_t s, uint64_t l, uint64_t b) : st(s), launches(l), bytes(b) { lq_prof_begin(name, s); }
    ~LqProfScope() { lq_prof_end(st, launches, bytes); }
};
#endif
