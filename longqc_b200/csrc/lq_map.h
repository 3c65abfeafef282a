/* lq_map.h -- device-side mapping of a query set against one index part (see lq_map.cu). */
#ifndef LQ_MAP_H
#define LQ_MAP_H
#include "lq_device.h"
#include "lq_index.h"

struct LqMapOpt {
    int no_self, ava;                 /* MM_F_NO_SELF, MM_F_AVA (minimap2-coverage.c:236-241) */
    int max_dist, bw, max_skip, min_cnt, min_sc;   /* mm_chain_dp arguments (lqmap.c:252) */
    int min_sc_med, min_sc_good;      /* -p / -q (lqmap.c:841) */
    int max_overhang; double min_ratio;   /* lq_fltopt_t (minimap2-coverage.c:369-388) */
    int covt;                         /* COVT 150 (minimap2-coverage.h:20) */
};

/* one accepted overlap of esterr.c:120-126: (query, start<<3|flags, end<<3|flags|1) */
struct LqOvl { uint32_t q, start, end; };

/* queries: persistent across index parts */
struct LqQueryDev {
    uint32_t nq; uint64_t n_min;
    LqReadsDev reads;
    LqMinimizers mins;                /* key, y (rid = query index), span */
    LqDevBuf first;                   /* u64[nq+1] minimizer range per query */
    LqDevBuf lambda, lambda2;         /* u64[nq]  esterr.c:120,128 */
    LqDevBuf qtied;                   /* u8[nq]: the query has such a minimizer (its seeds are never pre-filtered) */
    LqDevBuf dup;                     /* u8[n_min]: another minimizer of the same query has the same (key, strand) => its seeds tie (lq_afsort_core.h) */
    LqDevBuf mcnt;                    /* u32[n_min] per-minimizer match counters, indexed first[q] + rank among KEPT minimizers (esterr.c:130-137) */
    /* per part */
    LqDevBuf keep, neff, krank, soff; /* u32[n_min], u32[n_min], u32[n_min+1], u64[n_min+1] */
    LqDevBuf qstat;                   /* per query: LqQStat */
    LqDevBuf fmask; uint32_t fmask_stride; /* pre-filter survivor bits: fmask_stride u32 words per query minimizer */
    LqDevBuf self_off, self_list, qrank, trank; /* self-hit tables (u32) */
    LqMinimizers dup_tmp; LqDevBuf dup_tk, dup_ty, dup_ts, dup_hist, nmatch_buf; /* reusable scratch */
    LqQueryDev() : nq(0), n_min(0), fmask_stride(0) {}
    void release();
};

/* n_seeds: seeds the reference would sort (collect_seed_hits); n_sorted: seeds this part actually writes and sorts (after the pre-filter) */
struct LqQStat { uint32_t n_kept; uint32_t sum_span_kept; uint64_t n_seeds; uint64_t sum_span_seeds; uint32_t gate_closed; float avg_span; uint64_t n_sorted; };

struct LqMapScratch {
    LqDevBuf arena1, arena2, bkt, grp, misc, ovl, ws, wst, wph;   /* wst: region starts of the buckets being walked (lq_af_walk3_k); wph: their phase rows */
    void release() { arena1.release(); arena2.release(); bkt.release(); grp.release(); misc.release(); ovl.release(); ws.release(); wst.release(); wph.release(); }
};

struct LqMapStats { uint64_t n_seeds, n_groups, n_chains, n_ovl, n_batches, n_walk_buckets, n_seeds_all; };

/* Map every query against the part `ix`.  h_self_off/h_self_list: per query, target rids of this part
 * with the same name (CSR); h_qrank/h_trank: name ranks (only used with ava).  Appends accepted overlaps
 * to `ovl_out` (host vector), fills h_stat[nq]. */
int lq_map_part(LqQueryDev *qd, const LqIndexDev *ix, const LqMapOpt *opt, int mid_occ,
                const uint32_t *h_self_off, const uint32_t *h_self_list, const uint32_t *h_qrank, const uint32_t *h_trank,
                uint64_t seed_cap, LqMapScratch *sc, std::vector<LqOvl> *ovl_out, std::vector<LqQStat> *h_stat,
                LqMapStats *stats, cudaStream_t st);

/* final per-query reduction of the match counters (minimap2-coverage.c:552-562): n_match[q] */
/* h_npre (may be NULL): per query the minimizer count of the command line's k / w -- the `n` of minimap2-coverage.c:552-563 when it differs
 * from the mapping sketch (index files built with other parameters) */
int lq_map_nmatch(LqQueryDev *qd, const uint32_t *h_npre, std::vector<uint32_t> *n_match, cudaStream_t st);

/* once per query set: flag the minimizers whose (key, strand) occurs more than once in their query */
int lq_map_flag_dups(LqQueryDev *qd, int key_bits, LqDevBuf &ws, cudaStream_t st);

/* test hooks: run seeding+sort for queries [q0,q1) and return the sorted seeds (x, y as the reference lays them out) */
int lq_map_debug_sorted_seeds(LqQueryDev *qd, const LqIndexDev *ix, const LqMapOpt *opt, int mid_occ, uint32_t q,
                              const uint32_t *h_self_off, const uint32_t *h_self_list, const uint32_t *h_qrank, const uint32_t *h_trank,
                              LqMapScratch *sc, std::vector<lq_mm128> *unsorted, std::vector<lq_mm128> *sorted, cudaStream_t st);
#endif
