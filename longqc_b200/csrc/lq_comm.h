/* lq_comm.h -- internal side of the multi-GPU layer (lq_comm.cu); the exported calls are in include/lqcov.h */
#ifndef LQ_COMM_H
#define LQ_COMM_H
struct lqcov_ctx;
void lq_comm_release(lqcov_ctx *c);
#endif
