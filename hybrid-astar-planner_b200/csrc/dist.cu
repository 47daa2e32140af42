// Multi-GPU host layer behind the C ABI: the communicator (NCCL over NVLink, resolved at run time; an
// in-process transport for several ranks in one process), NavMap::setOccupancyMap tiled by row bands with every stage
// banded (N1, N3 at the original resolution; N5-N9 at the down-sampled one, voronoi.cu), and the query-sharded
// planner call with its result gather. The reference has no counterpart: it is a single-threaded ROS node.
#include <dlfcn.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>

#include "dist.cuh"

// ---- NCCL, resolved at run time -------------------------------------------------------------------------------------
// Only the handful of entry points used here, with the prototypes of nccl.h 2.27 / 2.28 (the ABI of these calls has been
// stable since 2.7). ncclUniqueId is 128 opaque bytes, ncclDataType_t ncclChar = 0.
namespace
{
struct NcclId
{
    char internal[128];
};
struct NcclApi
{
    void *lib = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

bool nccl_load()
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.ok)
        return true;
    if (!g_nccl.lib)
    {
        // a process that already carries NCCL (torch) hands back that very copy for the soname
        const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int k = 0; names[k] && !g_nccl.lib; ++k)
            g_nccl.lib = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
    }
    if (!g_nccl.lib)
    {
        mpros::set_error(std::string("NCCL not found: ") + (dlerror() ? dlerror() : "libnccl.so.2"));
        return false;
    }
#define MPROS_NCCL_SYM(field, name)                                                    \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name)); \
    if (!g_nccl.field)                                                                 \
    {                                                                                  \
        mpros::set_error(std::string("NCCL symbol missing: ") + name);                 \
        return false;                                                                  \
    }
    MPROS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    MPROS_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    MPROS_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    MPROS_NCCL_SYM(Broadcast, "ncclBroadcast")
    MPROS_NCCL_SYM(AllGather, "ncclAllGather")
    MPROS_NCCL_SYM(GroupStart, "ncclGroupStart")
    MPROS_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    MPROS_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef MPROS_NCCL_SYM
    g_nccl.ok = true;
    return true;
}

int nccl_fail(int r, const char *what)
{
    mpros::set_error(std::string("NCCL error in ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    return MPROS_ERR_CUDA;
}
#define MPROS_NCCL(expr)                    \
    do                                      \
    {                                       \
        int r_ = (expr);                    \
        if (r_ != 0)                        \
            return nccl_fail(r_, #expr);    \
    } while (0)
} // namespace

// ---- ranks inside one process -----------------------------------------------------------------------------------------
struct LocalGroup
{
    std::mutex mu;
    std::condition_variable cv;
    int world = 0, waiting = 0, refs = 0;
    uint64_t generation = 0;
    void *ptr[64] = {};
    int share_dev = 0; // device of rank 0's block (mpros_plan_share_create)
    bool broken = false; // a rank gave up waiting: every later barrier fails at once instead of hanging
    // false when a peer rank did not arrive within two minutes (it failed before reaching the same call)
    bool barrier()
    {
        std::unique_lock<std::mutex> lk(mu);
        if (broken)
            return false;
        const uint64_t gen = generation;
        if (++waiting == world)
        {
            waiting = 0;
            ++generation;
            cv.notify_all();
            return true;
        }
        if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return generation != gen || broken; }) || broken)
        {
            broken = true;
            cv.notify_all();
            return false;
        }
        return true;
    }
};
static int local_barrier(LocalGroup *g)
{
    if (g->barrier())
        return MPROS_OK;
    mpros::set_error("local communicator: a peer rank did not reach the exchange (it failed earlier, or the ranks make different calls)");
    return MPROS_ERR_INVALID;
}

namespace mpros
{
int comm_allgatherv(Ctx *c, mpros_comm *cm, void *buf, const size_t *off, const size_t *cnt)
{
    if (!cm || cm->world == 1)
        return MPROS_OK;
    char *b = static_cast<char *>(buf);
    cm->bytes_sent += cnt[cm->rank];
    cm->exchanges++;
    if (cm->kind == 0)
    {
        // one fused group of broadcasts, part r rooted at rank r: in place, ragged extents allowed
        MPROS_NCCL(g_nccl.GroupStart());
        for (int r = 0; r < cm->world; ++r)
            if (cnt[r])
                MPROS_NCCL(g_nccl.Broadcast(b + off[r], b + off[r], cnt[r], 0 /* ncclChar */, r, cm->nccl, c->stream));
        MPROS_NCCL(g_nccl.GroupEnd());
        return MPROS_OK;
    }
    LocalGroup *g = cm->grp;
    MPROS_CUDA(cudaStreamSynchronize(c->stream)); // my part is final
    g->ptr[cm->rank] = buf;
    int rc = local_barrier(g);
    if (rc)
        return rc;
    for (int r = 0; r < cm->world; ++r)
        if (r != cm->rank && cnt[r])
            MPROS_CUDA(cudaMemcpyAsync(b + off[r], static_cast<char *>(g->ptr[r]) + off[r], cnt[r], cudaMemcpyDefault, c->stream));
    MPROS_CUDA(cudaStreamSynchronize(c->stream));
    return local_barrier(g); // nobody overwrites its part before every rank has read it
}

int comm_allgather(Ctx *c, mpros_comm *cm, void *buf, size_t bytes)
{
    if (!cm || cm->world == 1)
        return MPROS_OK;
    if (cm->kind == 0)
    {
        cm->bytes_sent += bytes;
        cm->exchanges++;
        char *b = static_cast<char *>(buf);
        MPROS_NCCL(g_nccl.AllGather(b + (size_t)cm->rank * bytes, b, bytes, 0 /* ncclChar */, cm->nccl, c->stream));
        return MPROS_OK;
    }
    size_t off[64], cnt[64];
    for (int r = 0; r < cm->world; ++r)
        off[r] = (size_t)r * bytes, cnt[r] = bytes;
    return comm_allgatherv(c, cm, buf, off, cnt);
}

int comm_allgather_rows(Ctx *c, mpros_comm *cm, void *layer, size_t row_bytes, int total_rows, int rows_per_rank)
{
    if (!cm || cm->world == 1)
        return MPROS_OK;
    size_t off[64], cnt[64];
    for (int r = 0; r < cm->world; ++r)
    {
        int lo, hi;
        band_of(total_rows, rows_per_rank, r, &lo, &hi);
        off[r] = (size_t)lo * row_bytes;
        cnt[r] = (size_t)(hi - lo) * row_bytes;
    }
    return comm_allgatherv(c, cm, layer, off, cnt);
}
} // namespace mpros

using namespace mpros;

extern "C" int mpros_comm_unique_id(void *id128)
{
    if (!id128)
    {
        set_error("mpros_comm_unique_id: NULL buffer");
        return MPROS_ERR_INVALID;
    }
    if (!nccl_load())
        return MPROS_ERR_CUDA;
    NcclId id;
    MPROS_NCCL(g_nccl.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return MPROS_OK;
}

static bool comm_shape_ok(int world, int rank, const char *who)
{
    if (world < 1 || world > 64 || rank < 0 || rank >= world)
    {
        set_error(std::string(who) + ": world must be 1..64 and 0 <= rank < world");
        return false;
    }
    return true;
}

extern "C" int mpros_comm_create_nccl(mpros_ctx *c, int world, int rank, const void *id128, mpros_comm **out)
{
    if (!c || !id128 || !out || !comm_shape_ok(world, rank, "mpros_comm_create_nccl"))
    {
        if (c && id128 && out)
            return MPROS_ERR_INVALID;
        set_error("mpros_comm_create_nccl: invalid argument");
        return MPROS_ERR_INVALID;
    }
    if (!nccl_load())
        return MPROS_ERR_CUDA;
    MPROS_CUDA(cudaSetDevice(c->device));
    NcclId id;
    std::memcpy(&id, id128, sizeof(id));
    void *comm = nullptr;
    MPROS_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    mpros_comm *cm = new mpros_comm;
    cm->world = world;
    cm->rank = rank;
    cm->kind = 0;
    cm->nccl = comm;
    cm->own_nccl = true;
    *out = cm;
    return MPROS_OK;
}

extern "C" int mpros_comm_adopt_nccl(int world, int rank, void *nccl_comm, mpros_comm **out)
{
    if (!nccl_comm || !out || !comm_shape_ok(world, rank, "mpros_comm_adopt_nccl"))
    {
        if (nccl_comm && out)
            return MPROS_ERR_INVALID;
        set_error("mpros_comm_adopt_nccl: invalid argument");
        return MPROS_ERR_INVALID;
    }
    if (!nccl_load())
        return MPROS_ERR_CUDA;
    mpros_comm *cm = new mpros_comm;
    cm->world = world;
    cm->rank = rank;
    cm->kind = 0;
    cm->nccl = nccl_comm;
    cm->own_nccl = false;
    *out = cm;
    return MPROS_OK;
}

extern "C" int mpros_comm_create_local(int world, mpros_comm **out)
{
    if (!out || world < 1 || world > 64)
    {
        set_error("mpros_comm_create_local: world must be 1..64");
        return MPROS_ERR_INVALID;
    }
    LocalGroup *g = new LocalGroup;
    g->world = world;
    g->refs = world;
    for (int r = 0; r < world; ++r)
    {
        mpros_comm *cm = new mpros_comm;
        cm->world = world;
        cm->rank = r;
        cm->kind = 1;
        cm->grp = g;
        out[r] = cm;
    }
    return MPROS_OK;
}

extern "C" int mpros_comm_destroy(mpros_comm *cm)
{
    if (!cm)
        return MPROS_OK;
    int rc = MPROS_OK;
    if (cm->kind == 0 && cm->own_nccl && cm->nccl && g_nccl.ok)
    {
        int r = g_nccl.CommDestroy(cm->nccl);
        if (r != 0)
            rc = nccl_fail(r, "ncclCommDestroy");
    }
    if (cm->kind == 1 && cm->grp)
    {
        bool last;
        {
            std::lock_guard<std::mutex> lk(cm->grp->mu);
            last = --cm->grp->refs == 0;
        }
        if (last)
            delete cm->grp;
    }
    delete cm;
    return rc;
}

extern "C" int mpros_comm_info(const mpros_comm *cm, int *world, int *rank, uint64_t *bytes_sent, uint64_t *exchanges)
{
    if (!cm)
    {
        set_error("mpros_comm_info: NULL communicator");
        return MPROS_ERR_INVALID;
    }
    if (world)
        *world = cm->world;
    if (rank)
        *rank = cm->rank;
    if (bytes_sent)
        *bytes_sent = cm->bytes_sent;
    if (exchanges)
        *exchanges = cm->exchanges;
    return MPROS_OK;
}

// ---- query sharding ------------------------------------------------------------------------------------------------------
// Contiguous block of query indices of `rank` (sizes differ by at most one).
static void shard_of(size_t n, int rank, int world, size_t *lo, size_t *hi)
{
    const size_t base = n / world, rem = n % world;
    *lo = (size_t)rank * base + std::min<size_t>(rank, rem);
    *hi = *lo + base + ((size_t)rank < rem ? 1 : 0);
}

extern "C" int mpros_plan_shard_range(const mpros_comm *cm, size_t nq, size_t *lo, size_t *hi)
{
    if (!lo || !hi)
    {
        set_error("mpros_plan_shard_range: NULL output");
        return MPROS_ERR_INVALID;
    }
    shard_of(nq, cm ? cm->rank : 0, cm ? cm->world : 1, lo, hi);
    return MPROS_OK;
}

extern "C" int mpros_plan_batch_sharded(mpros_ctx *c, mpros_comm *cm, const double *starts3, const double *goals3, size_t nq,
                                        int32_t *status, double *cost, int32_t *path_len, double *paths, int path_cap, int64_t *stats,
                                        int gather_paths)
{
    if (!c || !starts3 || !goals3 || !status || !cost || !path_len || nq == 0 || nq > (size_t)1 << 30 || path_cap < 0 ||
        (path_cap > 0 && !paths))
    {
        set_error("mpros_plan_batch_sharded: invalid argument");
        return MPROS_ERR_INVALID;
    }
    const int world = cm ? cm->world : 1, rank = cm ? cm->rank : 0;
    size_t lo, hi;
    shard_of(nq, rank, world, &lo, &hi);
    MPROS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    // device arrays for ALL queries: the shard is planned into its own slice, the gather fills the rest in place
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_in = up(nq * 24), b_st = up(nq * 4), b_co = up(nq * 8), b_pl = up(nq * 4), b_stats = up(nq * 48),
                 b_paths = up(nq * (size_t)path_cap * 40);
    const size_t need = 2 * b_in + b_st + b_co + b_pl + b_stats + b_paths;
    if (need > c->dist_bytes)
    {
        if (c->dist_buf)
            MPROS_CUDA(cudaFree(c->dist_buf));
        c->dist_buf = nullptr;
        c->dist_bytes = 0;
        MPROS_CUDA(cudaMalloc(&c->dist_buf, need));
        c->dist_bytes = need;
    }
    char *p = static_cast<char *>(c->dist_buf);
    double *d_s = (double *)p;
    double *d_g = (double *)(p += b_in);
    int32_t *d_st = (int32_t *)(p += b_in);
    double *d_co = (double *)(p += b_st);
    int32_t *d_pl = (int32_t *)(p += b_co);
    int64_t *d_stats = (int64_t *)(p += b_pl);
    double *d_paths = (double *)(p += b_stats);
    const size_t n_loc = hi - lo;
    if (n_loc)
    {
        MPROS_CUDA(cudaMemcpyAsync(d_s + 3 * lo, starts3 + 3 * lo, n_loc * 24, cudaMemcpyHostToDevice, st));
        MPROS_CUDA(cudaMemcpyAsync(d_g + 3 * lo, goals3 + 3 * lo, n_loc * 24, cudaMemcpyHostToDevice, st));
        int rc = mpros_plan_batch_dev(c, d_s + 3 * lo, d_g + 3 * lo, n_loc, d_st + lo, d_co + lo, d_pl + lo,
                                      path_cap ? d_paths + lo * (size_t)path_cap * 5 : nullptr, path_cap, d_stats + 6 * lo);
        if (rc)
            return rc;
    }
    // result gather: the only exchange of the query-sharded planner (fixed-size records; the paths only when asked for)
    if (world > 1)
    {
        size_t off[64], cnt[64];
        auto gather = [&](void *buf, size_t rec) -> int {
            for (int r = 0; r < world; ++r)
            {
                size_t a, b;
                shard_of(nq, r, world, &a, &b);
                off[r] = a * rec;
                cnt[r] = (b - a) * rec;
            }
            return comm_allgatherv(c, cm, buf, off, cnt);
        };
        int rc = gather(d_st, 4);
        if (!rc)
            rc = gather(d_co, 8);
        if (!rc)
            rc = gather(d_pl, 4);
        if (!rc)
            rc = gather(d_stats, 48);
        if (!rc && gather_paths && path_cap)
            rc = gather(d_paths, (size_t)path_cap * 40);
        if (rc)
            return rc;
    }
    MPROS_CUDA(cudaMemcpyAsync(status, d_st, nq * 4, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaMemcpyAsync(cost, d_co, nq * 8, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaMemcpyAsync(path_len, d_pl, nq * 4, cudaMemcpyDeviceToHost, st));
    if (stats)
        MPROS_CUDA(cudaMemcpyAsync(stats, d_stats, nq * 48, cudaMemcpyDeviceToHost, st));
    if (path_cap)
    {
        if (gather_paths || world == 1)
            MPROS_CUDA(cudaMemcpyAsync(paths, d_paths, nq * (size_t)path_cap * 40, cudaMemcpyDeviceToHost, st));
        else if (n_loc)
            MPROS_CUDA(cudaMemcpyAsync(paths + lo * (size_t)path_cap * 5, d_paths + lo * (size_t)path_cap * 5, n_loc * (size_t)path_cap * 40,
                                       cudaMemcpyDeviceToHost, st));
    }
    MPROS_CUDA(cudaStreamSynchronize(st));
    return MPROS_OK;
}

// ---- dynamic sharding over peer memory -------------------------------------------------------------------------------------
// mpros_plan_batch_sharded gives every rank a fixed block of the queries; a rank whose block holds more of the long searches
// (a query that runs into the iteration limit is ~50 times the mean) finishes last while the others idle. Here ALL ranks
// draw query tickets from ONE counter in rank 0's memory (system-scope atomicAdd through the NVLink mapping) and store each
// query's results - status, cost, path length, statistics and the path itself - straight into rank 0's result arrays with
// ordinary stores: the planner kernel does the load balancing and the result gather itself, no collective moves data. Two
// stream-ordered barriers (a 4-byte exchange each) fence the launch: counter cleared before anybody draws, every kernel
// finished before anybody reads.
struct mpros_plan_share
{
    mpros_comm *cm = nullptr;
    void *home = nullptr;      // rank 0's block as this rank addresses it
    bool opened_ipc = false;   // other process: cudaIpcOpenMemHandle; same process / rank 0: the pointer itself
    size_t max_queries = 0;
    int path_cap = 0;
    size_t off_status = 0, off_cost = 0, off_len = 0, off_stats = 0, off_paths = 0, bytes = 0;
    void *d_queries = nullptr; // this rank's copy of all starts / goals
    void *d_sync = nullptr;    // world x 4 bytes for the barriers
};

static int comm_barrier(mpros_ctx *c, mpros_comm *cm, void *d_sync) { return comm_allgather(c, cm, d_sync, 4); }

extern "C" int mpros_plan_share_create(mpros_ctx *c, mpros_comm *cm, size_t max_queries, int path_cap, mpros_plan_share **out)
{
    if (!c || !cm || !out || max_queries == 0 || max_queries > ((size_t)1 << 28) || path_cap < 0)
    {
        set_error("mpros_plan_share_create: invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    mpros_plan_share *sh = new mpros_plan_share;
    sh->cm = cm;
    sh->max_queries = max_queries;
    sh->path_cap = path_cap;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t o = 256; // [0, 256): the ticket counter
    sh->off_status = o, o += up(max_queries * 4);
    sh->off_cost = o, o += up(max_queries * 8);
    sh->off_len = o, o += up(max_queries * 4);
    sh->off_stats = o, o += up(max_queries * 48);
    sh->off_paths = o, o += up(max_queries * (size_t)path_cap * 40);
    sh->bytes = o;
    MPROS_CUDA(cudaMalloc(&sh->d_queries, max_queries * 48));
    MPROS_CUDA(cudaMalloc(&sh->d_sync, (size_t)cm->world * 4 + 64));
    MPROS_CUDA(cudaMemsetAsync(sh->d_sync, 0, (size_t)cm->world * 4 + 64, c->stream));
    // rank 0 owns the block; the others map it
    void *mine = nullptr;
    if (cm->rank == 0)
    {
        MPROS_CUDA(cudaMalloc(&mine, sh->bytes));
        MPROS_CUDA(cudaMemsetAsync(mine, 0, sh->bytes, c->stream));
    }
    if (cm->kind == 1)
    {
        // ranks of one process: the pointer itself (peer access between different devices is enabled on demand)
        LocalGroup *g = cm->grp;
        MPROS_CUDA(cudaStreamSynchronize(c->stream));
        if (cm->rank == 0)
        {
            g->ptr[0] = mine;
            g->share_dev = c->device;
        }
        int rc = local_barrier(g);
        if (rc)
            return rc;
        sh->home = g->ptr[0];
        if (cm->rank != 0 && g->share_dev != c->device)
        {
            cudaError_t e = cudaDeviceEnablePeerAccess(g->share_dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            cudaGetLastError();
        }
        rc = local_barrier(g);
        if (rc)
            return rc;
    }
    else
    {
        // ranks in different processes: the CUDA IPC handle of the block travels through the communicator
        cudaIpcMemHandle_t h;
        std::memset(&h, 0, sizeof(h));
        if (cm->rank == 0)
            MPROS_CUDA(cudaIpcGetMemHandle(&h, mine));
        void *d_h = nullptr;
        MPROS_CUDA(cudaMalloc(&d_h, sizeof(h)));
        MPROS_CUDA(cudaMemcpyAsync(d_h, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
        size_t off[64] = {}, cnt[64] = {};
        cnt[0] = sizeof(h); // a broadcast from rank 0
        int rc = comm_allgatherv(c, cm, d_h, off, cnt);
        if (rc)
            return rc;
        MPROS_CUDA(cudaMemcpyAsync(&h, d_h, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        MPROS_CUDA(cudaStreamSynchronize(c->stream));
        MPROS_CUDA(cudaFree(d_h));
        if (cm->rank == 0)
            sh->home = mine;
        else
        {
            MPROS_CUDA(cudaIpcOpenMemHandle(&sh->home, h, cudaIpcMemLazyEnablePeerAccess));
            sh->opened_ipc = true;
        }
    }
    *out = sh;
    return MPROS_OK;
}

extern "C" int mpros_plan_share_destroy(mpros_ctx *c, mpros_plan_share *sh)
{
    if (!sh)
        return MPROS_OK;
    if (c)
    {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        // nobody unmaps or frees while a peer may still be storing: one more barrier
        comm_barrier(c, sh->cm, sh->d_sync);
        cudaStreamSynchronize(c->stream);
    }
    if (sh->opened_ipc)
        cudaIpcCloseMemHandle(sh->home);
    else if (sh->cm->rank == 0 && sh->home)
        cudaFree(sh->home);
    if (sh->d_queries)
        cudaFree(sh->d_queries);
    if (sh->d_sync)
        cudaFree(sh->d_sync);
    delete sh;
    return MPROS_OK;
}

extern "C" int mpros_plan_batch_shared(mpros_ctx *c, mpros_plan_share *sh, const double *starts3, const double *goals3, size_t nq,
                                       int32_t *status, double *cost, int32_t *path_len, double *paths, int path_cap, int64_t *stats)
{
    if (!c || !sh || !starts3 || !goals3 || nq == 0 || nq > sh->max_queries || path_cap != sh->path_cap)
    {
        set_error("mpros_plan_batch_shared: invalid argument (nq <= max_queries and the path_cap of mpros_plan_share_create)");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    mpros_comm *cm = sh->cm;
    char *home = static_cast<char *>(sh->home);
    double *d_s = static_cast<double *>(sh->d_queries), *d_g = d_s + 3 * sh->max_queries;
    MPROS_CUDA(cudaMemcpyAsync(d_s, starts3, nq * 24, cudaMemcpyHostToDevice, st));
    MPROS_CUDA(cudaMemcpyAsync(d_g, goals3, nq * 24, cudaMemcpyHostToDevice, st));
    if (cm->rank == 0)
        MPROS_CUDA(cudaMemsetAsync(home, 0, 256, st)); // the ticket counter
    int rc = comm_barrier(c, cm, sh->d_sync); // the counter is clear before anybody draws
    if (rc)
        return rc;
    rc = plan_batch_shared_launch(c, d_s, d_g, (int)nq, reinterpret_cast<unsigned int *>(home), reinterpret_cast<int32_t *>(home + sh->off_status),
                                  reinterpret_cast<double *>(home + sh->off_cost), reinterpret_cast<int32_t *>(home + sh->off_len),
                                  path_cap ? reinterpret_cast<double *>(home + sh->off_paths) : nullptr, path_cap,
                                  reinterpret_cast<int64_t *>(home + sh->off_stats));
    if (rc)
        return rc;
    rc = comm_barrier(c, cm, sh->d_sync); // every rank's kernel has finished: rank 0's arrays are complete
    if (rc)
        return rc;
    // any rank that passed result arrays reads them (rank 0 from its own memory, the others through the mapping)
    if (status)
        MPROS_CUDA(cudaMemcpyAsync(status, home + sh->off_status, nq * 4, cudaMemcpyDeviceToHost, st));
    if (cost)
        MPROS_CUDA(cudaMemcpyAsync(cost, home + sh->off_cost, nq * 8, cudaMemcpyDeviceToHost, st));
    if (path_len)
        MPROS_CUDA(cudaMemcpyAsync(path_len, home + sh->off_len, nq * 4, cudaMemcpyDeviceToHost, st));
    if (stats)
        MPROS_CUDA(cudaMemcpyAsync(stats, home + sh->off_stats, nq * 48, cudaMemcpyDeviceToHost, st));
    if (paths && path_cap)
        MPROS_CUDA(cudaMemcpyAsync(paths, home + sh->off_paths, nq * (size_t)path_cap * 40, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaStreamSynchronize(st));
    // the next call clears the counter only after everybody has read
    rc = comm_barrier(c, cm, sh->d_sync);
    if (rc)
        return rc;
    MPROS_CUDA(cudaStreamSynchronize(st));
    return MPROS_OK;
}
