/* lq_mmi.h -- the reference's on-disk index format (lq_mmi.cpp) */
#ifndef LQ_MMI_H
#define LQ_MMI_H
#include <stdio.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "lqcov.h"

struct LqMmiPart {
    int w, k, b; uint32_t n_seq, flag;
    std::string names; std::vector<uint64_t> name_off, seq_off;   /* n_seq + 1 each; seq_off = cumulative lengths */
    std::vector<uint32_t> key; std::vector<uint64_t> y;           /* every minimizer record; one key's positions ascend in y */
};
int lq_mmi_dump_part(FILE *fp, int w, int k, int is_hpc, const lqcov_reads_t *part, const uint32_t *counts, const uint64_t *offs, const uint64_t *pos);
int lq_mmi_is_index(const char *path);
int lq_mmi_load_part(FILE *fp, LqMmiPart *out);
#endif
