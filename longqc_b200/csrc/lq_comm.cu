/* lq_comm.cu -- several GPUs share one index part: NCCL over NVLink, under the C ABI.
 *
 * One context per GPU (one process per GPU under torchrun, or one process driving all visible GPUs: the drop-in executable).  The
 * part's reads are owned by the ranks in contiguous, rank-ordered ranges, so rid == position in the part and "ascending y inside a
 * key" (index.c:188) == rank-major order.  Per part, lqcov_part_exchange():
 *   1. ALL-REDUCE (sum) of the per-minimizer count table: the global counts decide mid_occ (index.c:123-144) and the
 *      high-frequency filter (lqmap.c:159,166) -- the collective the method cannot do without;
 *   2. every rank sorts ONLY ITS OWN records by key (stable: 1/N of the work);
 *   3. the per-rank count tables are all-gathered: with their exclusive scans a record's place in the replicated index is
 *        offs[key] + (occurrences of key on lower ranks) + (its rank inside the key's run of its own shard)
 *      -- disjoint slots, no further sorting;
 *   4. the key-sorted position arrays are exchanged (grouped send/recv: an all-gather with uneven shards) and one kernel copies every
 *      (key, rank) run to its slot of the replicated index.
 * NCCL is loaded with dlopen (no link-time dependency: a single-GPU installation needs no NCCL at all); inside a process that has
 * already loaded a libnccl.so.2 (PyTorch) the same object is used.
 */
#include <dlfcn.h>
#include <string.h>
#include <vector>
#include <nccl.h>
#include "lq_ctx.h"
#include "lq_comm.h"

struct LqNccl {
    void *dl;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    const char *(*GetErrorString)(ncclResult_t);
};
static LqNccl g_nccl;

static int nccl_load()
{
    if (g_nccl.dl) return 0;
    const char *names[] = { getenv("LQCOV_NCCL_LIB"), "libnccl.so.2", "libnccl.so", 0 };
    void *dl = 0;
    for (int i = 0; i < 4 && !dl; ++i) if (names[i]) dl = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!dl) { fprintf(stderr, "[lqcov] ERROR: cannot load NCCL (libnccl.so.2): %s\n", dlerror()); return -1; }
#define LQ_SYM(field, sym) do { *(void**)&g_nccl.field = dlsym(dl, sym); if (!g_nccl.field) { fprintf(stderr, "[lqcov] ERROR: %s not found in NCCL\n", sym); return -1; } } while (0)
    LQ_SYM(GetUniqueId, "ncclGetUniqueId"); LQ_SYM(CommInitRank, "ncclCommInitRank"); LQ_SYM(CommInitAll, "ncclCommInitAll"); LQ_SYM(CommDestroy, "ncclCommDestroy");
    LQ_SYM(AllReduce, "ncclAllReduce"); LQ_SYM(AllGather, "ncclAllGather"); LQ_SYM(Send, "ncclSend"); LQ_SYM(Recv, "ncclRecv");
    LQ_SYM(GroupStart, "ncclGroupStart"); LQ_SYM(GroupEnd, "ncclGroupEnd"); LQ_SYM(GetErrorString, "ncclGetErrorString");
#undef LQ_SYM
    g_nccl.dl = dl;
    return 0;
}
#define LQ_NCCL_OK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) { \
    fprintf(stderr, "[lqcov] NCCL error at %s:%d: %s\n", __FILE__, __LINE__, g_nccl.GetErrorString(r_)); return -1; } } while (0)

struct LqComm {
    ncclComm_t comm; int n, rank;
    LqDevBuf all_counts, lo, gather, sizes;
    LqComm() : comm(0), n(1), rank(0) {}
};

extern "C" int lqcov_comm_unique_id(void *id128)
{
    LQ_TRY(nccl_load());
    ncclUniqueId id;
    LQ_NCCL_OK(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return 0;
}

extern "C" int lqcov_comm_init_rank(lqcov_ctx *c, const void *id128, int nranks, int rank)
{
    LQ_TRY(nccl_load());
    if (c->comm) { fprintf(stderr, "[lqcov] ERROR: the context already has a communicator\n"); return -1; }
    if (c->opt.device >= 0) LQ_CUDA_OK(cudaSetDevice(c->opt.device));
    ncclUniqueId id; memcpy(&id, id128, 128);
    LqComm *m = new LqComm(); m->n = nranks; m->rank = rank;
    if (g_nccl.CommInitRank(&m->comm, nranks, id, rank) != ncclSuccess) { fprintf(stderr, "[lqcov] ERROR: ncclCommInitRank failed\n"); delete m; return -1; }
    c->comm = m;
    return 0;
}

/* one process, n contexts (created on n different devices): rank = position in `ctxs` */
extern "C" int lqcov_comm_init_all(lqcov_ctx **ctxs, int n)
{
    LQ_TRY(nccl_load());
    std::vector<int> devs(n); std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; ++i) { if (ctxs[i]->comm) return -1; devs[i] = ctxs[i]->opt.device >= 0 ? ctxs[i]->opt.device : 0; }
    LQ_NCCL_OK(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) { LqComm *m = new LqComm(); m->n = n; m->rank = i; m->comm = comms[i]; ctxs[i]->comm = m; }
    return 0;
}

void lq_comm_release(lqcov_ctx *c)
{
    if (!c->comm) return;
    LqComm *m = c->comm;
    m->all_counts.release(); m->lo.release(); m->gather.release(); m->sizes.release();
    if (m->comm && g_nccl.dl) g_nccl.CommDestroy(m->comm);
    delete m; c->comm = 0;
}

extern "C" int lqcov_comm_size(const lqcov_ctx *c) { return c->comm ? c->comm->n : 1; }
extern "C" int lqcov_comm_rank(const lqcov_ctx *c) { return c->comm ? c->comm->rank : 0; }

/* ---- placement: every (key, rank) run of the gathered, key-sorted shards goes to its slot of the replicated index ----
 * A warp takes 32 consecutive keys at a time (lane = key: the count / offset tables of all ranks are read coalesced), then walks
 * those keys one after the other with the whole warp (lane = element of the key's final run). */
#define PL_MAX_RANKS 16
struct PlaceArgs {
    int n; uint64_t nkeys;
    const uint32_t *all_counts, *lo;     /* [n][nkeys] */
    uint64_t shard_base[PL_MAX_RANKS];    /* first element of rank r's shard in `gather` */
    const uint64_t *gather, *goffs; uint64_t *pos;
};
__global__ void __launch_bounds__(256) lq_place_k(PlaceArgs a)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t nwarp = (uint64_t)gridDim.x * (blockDim.x >> 5), w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint64_t k0 = w * 32; k0 < a.nkeys; k0 += nwarp * 32) {
        const uint64_t k = k0 + lane;
        uint32_t c[PL_MAX_RANKS], l[PL_MAX_RANKS], tot = 0;
        #pragma unroll
        for (int r = 0; r < PL_MAX_RANKS; ++r) {
            c[r] = 0; l[r] = 0;
            if (r < a.n && k < a.nkeys) { c[r] = a.all_counts[(uint64_t)r * a.nkeys + k]; l[r] = a.lo[(uint64_t)r * a.nkeys + k]; }
            tot += c[r];
        }
        const uint64_t g = k < a.nkeys ? a.goffs[k] : 0;
        uint32_t live = __ballot_sync(0xffffffffu, tot != 0);
        while (live) {
            const int kk = __ffs(live) - 1; live &= live - 1;
            const uint64_t gk = __shfl_sync(0xffffffffu, g, kk);
            uint32_t done = 0;
            #pragma unroll
            for (int r = 0; r < PL_MAX_RANKS; ++r) {
                if (r >= a.n) break;
                const uint32_t cr = __shfl_sync(0xffffffffu, c[r], kk), lr = __shfl_sync(0xffffffffu, l[r], kk);
                const uint64_t *src = a.gather + a.shard_base[r] + lr;
                for (uint32_t j = lane; j < cr; j += 32) a.pos[gk + done + j] = src[j];
                done += cr;
            }
        }
    }
}

/* all ranks call this between lqcov_part_sketch / lqcov_part_end and lqcov_part_finish */
extern "C" int lqcov_part_exchange(lqcov_ctx *c)
{
    LqComm *m = c->comm;
    if (!m || m->n == 1) return 0;
    if (c->opt.is_hpc) { fprintf(stderr, "[lqcov] ERROR: -H (spike-in run, a 4 kb target) is a single-GPU job: do not share it between GPUs\n"); return -1; }
    if (c->ix.rec.wide) { fprintf(stderr, "[lqcov] ERROR: k > 15 is a single-GPU job in this build (the index addresses of wide keys are numbered per GPU)\n"); return -1; }
    if (m->n > PL_MAX_RANKS) { fprintf(stderr, "[lqcov] ERROR: more than %d ranks\n", PL_MAX_RANKS); return -1; }
    LqIndexDev *ix = &c->ix; cudaStream_t st = c->st;
    const int n = m->n, me = m->rank; const uint64_t nkeys = ix->n_keyspace;
    if (c->opt.device >= 0) LQ_CUDA_OK(cudaSetDevice(c->opt.device));
    /* shard sizes */
    LQ_TRY(m->sizes.ensure((size_t)(n + 1) * 8));
    std::vector<uint64_t> sz(n, 0);
    const uint64_t my_n = ix->rec.n;
    LQ_CUDA_OK(cudaMemcpyAsync(m->sizes.as<uint64_t>() + me, &my_n, 8, cudaMemcpyHostToDevice, st));
    LQ_NCCL_OK(g_nccl.AllGather(m->sizes.as<uint64_t>() + me, m->sizes.p, 1, ncclUint64, m->comm, st));
    /* the per-rank count tables (rank prefixes) and their sum (THE all-reduce: global occurrence counts) */
    LQ_TRY(m->all_counts.ensure((size_t)n * nkeys * 4 + 64)); LQ_TRY(m->lo.ensure((size_t)n * nkeys * 4 + 64));
    {
        LqProfScope ps("comm_counts", st, 0, (uint64_t)(n + 2) * nkeys * 4);
        LQ_NCCL_OK(g_nccl.AllGather(ix->counts.p, m->all_counts.p, (size_t)nkeys, ncclUint32, m->comm, st));
        LQ_NCCL_OK(g_nccl.AllReduce(ix->counts.p, ix->counts.p, (size_t)nkeys, ncclUint32, ncclSum, m->comm, st));
    }
    LQ_CUDA_OK(cudaMemcpyAsync(sz.data(), m->sizes.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    /* this rank's records, stable-sorted by key */
    ix->rec.has_span = 0;
    LQ_TRY(lq_sort_by_key(&ix->rec, 2 * ix->k, ix->tmp_key, ix->tmp_y, ix->tmp_sp, ix->hist, c->ws, st, 0));
    LQ_CUDA_OK(cudaStreamSynchronize(st));
    std::vector<uint64_t> base(n + 1, 0);
    for (int r = 0; r < n; ++r) base[r + 1] = base[r] + sz[r];
    const uint64_t n_total = base[n];
    /* exchange of the sorted shards */
    LQ_TRY(m->gather.ensure((size_t)(n_total + 1) * 8));
    {
        LqProfScope ps("comm_records", st, 0, n_total * 8);
        LQ_NCCL_OK(g_nccl.GroupStart());
        for (int r = 0; r < n; ++r) {
            if (r == me) continue;
            if (my_n) LQ_NCCL_OK(g_nccl.Send(ix->rec.y.p, (size_t)my_n, ncclUint64, r, m->comm, st));
            if (sz[r]) LQ_NCCL_OK(g_nccl.Recv(m->gather.as<uint64_t>() + base[r], (size_t)sz[r], ncclUint64, r, m->comm, st));
        }
        LQ_NCCL_OK(g_nccl.GroupEnd());
        if (my_n) LQ_CUDA_OK(cudaMemcpyAsync(m->gather.as<uint64_t>() + base[me], ix->rec.y.p, (size_t)my_n * 8, cudaMemcpyDeviceToDevice, st));
    }
    /* offsets: of the replicated index (global counts) and inside every rank's shard */
    { LqProfScope ps("offs_scan", st, 0, nkeys * 16);
      LQ_TRY((lq_exclusive_scan<uint32_t, uint64_t>(ix->counts.as<uint32_t>(), ix->offs.as<uint64_t>(), (size_t)nkeys, 1, c->ws, st))); }
    for (int r = 0; r < n; ++r)
        LQ_TRY((lq_exclusive_scan<uint32_t, uint32_t>(m->all_counts.as<uint32_t>() + (size_t)r * nkeys, m->lo.as<uint32_t>() + (size_t)r * nkeys, (size_t)nkeys, 0, c->ws, st)));
    /* placement into the final position array (the sort's temporary is free again) */
    LQ_TRY(ix->tmp_y.ensure((size_t)(n_total + 1) * 8));
    {
        PlaceArgs a; a.n = n; a.nkeys = nkeys; a.all_counts = m->all_counts.as<uint32_t>(); a.lo = m->lo.as<uint32_t>();
        for (int r = 0; r < PL_MAX_RANKS; ++r) a.shard_base[r] = r < n ? base[r] : 0;
        a.gather = m->gather.as<uint64_t>(); a.goffs = ix->offs.as<uint64_t>(); a.pos = ix->tmp_y.as<uint64_t>();
        LqProfScope ps("index_place", st, 1, n_total * 16 + (uint64_t)n * nkeys * 8);
        lq_place_k<<<148 * 8, 256, 0, st>>>(a);
        LQ_CUDA_OK(cudaGetLastError());
    }
    std::swap(ix->rec.y, ix->tmp_y);
    ix->rec.n = n_total; ix->n_rec = n_total;
    c->placed = true;
    LQ_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

/* rows of all ranks on rank 0, in rank order (the queries are split in rank order): a gather with uneven sizes.  `mine`/`len`: this
 * rank's rows; on rank 0 *all is malloc'ed (lqcov_free) */
extern "C" int lqcov_comm_gather_rows(lqcov_ctx *c, const char *mine, size_t len, char **all, size_t *all_len)
{
    LqComm *m = c->comm;
    *all = 0; *all_len = 0;
    if (!m || m->n == 1) { *all = (char*)malloc(len + 1); memcpy(*all, mine, len); (*all)[len] = 0; *all_len = len; return 0; }
    cudaStream_t st = c->st; const int n = m->n, me = m->rank;
    if (c->opt.device >= 0) LQ_CUDA_OK(cudaSetDevice(c->opt.device));
    LQ_TRY(m->sizes.ensure((size_t)(n + 1) * 8));
    const uint64_t my = len; std::vector<uint64_t> sz(n);
    LQ_CUDA_OK(cudaMemcpyAsync(m->sizes.as<uint64_t>() + me, &my, 8, cudaMemcpyHostToDevice, st));
    LQ_NCCL_OK(g_nccl.AllGather(m->sizes.as<uint64_t>() + me, m->sizes.p, 1, ncclUint64, m->comm, st));
    LQ_CUDA_OK(cudaMemcpyAsync(sz.data(), m->sizes.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    LQ_CUDA_OK(cudaStreamSynchronize(st));
    uint64_t tot = 0; std::vector<uint64_t> base(n + 1, 0);
    for (int r = 0; r < n; ++r) { base[r + 1] = base[r] + sz[r]; } tot = base[n];
    LqDevBuf buf; LQ_TRY(buf.ensure((size_t)(me == 0 ? tot : len) + 16));
    if (len) LQ_CUDA_OK(cudaMemcpyAsync((char*)buf.p + (me == 0 ? base[0] : 0), mine, len, cudaMemcpyHostToDevice, st));
    LQ_NCCL_OK(g_nccl.GroupStart());
    if (me == 0) { for (int r = 1; r < n; ++r) if (sz[r]) LQ_NCCL_OK(g_nccl.Recv((char*)buf.p + base[r], (size_t)sz[r], ncclChar, r, m->comm, st)); }
    else if (len) LQ_NCCL_OK(g_nccl.Send(buf.p, len, ncclChar, 0, m->comm, st));
    LQ_NCCL_OK(g_nccl.GroupEnd());
    if (me == 0) {
        *all = (char*)malloc((size_t)tot + 1);
        if (tot) LQ_CUDA_OK(cudaMemcpyAsync(*all, buf.p, (size_t)tot, cudaMemcpyDeviceToHost, st));
        LQ_CUDA_OK(cudaStreamSynchronize(st));
        (*all)[tot] = 0; *all_len = (size_t)tot;
    } else LQ_CUDA_OK(cudaStreamSynchronize(st));
    buf.release();
    return 0;
}
