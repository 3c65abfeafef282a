/* lq_ctx.h -- the context behind include/lqcov.h (private to the library's translation units). */
#ifndef LQ_CTX_H
#define LQ_CTX_H
#include <string>
#include <vector>
#include "lq_cuda.cuh"
#include "lq_device.h"
#include "lq_index.h"
#include "lq_map.h"
#include "lq_host.h"
#include "lqcov.h"

struct LqComm;

struct lqcov_ctx {
    lqcov_opt_t opt;
    cudaStream_t st;
    /* queries, host side */
    uint32_t nq;
    std::vector<std::string> qname;
    std::vector<int> qlen;
    std::vector<double> qsum_p; bool q_has_qual;   /* ordered error-probability sums of the query qualities (device) */
    /* query names -> query indices: open-address table over (pointer, length) keys, chained for duplicate names */
    std::vector<int32_t> qn_slot, qn_next; uint32_t qn_mask;
    std::vector<uint64_t> qfirst;
    std::vector<lqh_sub_v> ovlp;        /* ovlp_coords (minimap2-coverage.c:438-444) */
    std::vector<float> avg_k;           /* avg_ks */
    /* device */
    LqQueryDev qd; LqIndexDev ix; LqMapScratch sc; LqReadsDev treads; LqMinimizers tmins, full; LqDevBuf ws, qual_dev, qsum_dev; bool use_full;
    LqPartStream stream; std::vector<char*> stage; size_t stage_bytes; std::vector<cudaEvent_t> stage_ev; double t_part0;
    /* current part */
    std::vector<uint32_t> self_off, self_list, qrank, trank;
    bool part_ready;
    int32_t mid_occ;
    lqcov_stats_t stats;
    std::vector<uint32_t> n_prepass;     /* per query: minimizers under the command line's k / w when an index file with other parameters is mapped against */
    LqComm *comm;                        /* multi-GPU: NULL for a single context (lq_comm.cu) */
    bool placed;                         /* the part's records were merged and placed by lqcov_part_exchange: lqcov_part_finish must not sort them again */
};

#endif
