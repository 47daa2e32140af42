// Batched Reeds-Shepp solve / shot (R1-R4), batched node expansion (E1-E5) and the whole-query planner
// (HybridAstar::initialize + Plan + raw setPath, hybrid_a_star.cpp:52-118, 682-811) for MANY independent queries.
//
// Execution model: persistent kernel, one WARP per query, warps pull query indices from a global counter. A* is
// sequential per query (one popped node per step, strict reference order); the parallelism is across queries
// (thousands of resident warps) and inside a step across the primitives' footprint probes, the 48 RS candidates
// and the 4 symmetry images of every OMPL-distance formula. All per-query state (node pool, open heap, closed-set
// hash, on-demand goal wavefront) lives in a per-warp HBM workspace that is reused without clearing (generation tags).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "planner_device.cuh"

namespace mpros
{

struct WarpWorkspaceLayout
{
    size_t node_off, hkey_off, hid_off, table_off, blocked_off, f0_off, f1_off, dist_off, rowtag_off, traj_off, hdr_off;
    size_t stride;
    int node_cap, table_slots, bfs_rows, bfs_wwords;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static WarpWorkspaceLayout make_layout(int node_cap, int H, int W)
{
    WarpWorkspaceLayout L;
    L.node_cap = node_cap;
    int slots = 1;
    while (slots < 2 * node_cap + 16)
        slots <<= 1;
    L.table_slots = slots;
    const int span = 2 * (kHugeInt - 1) + 1;
    L.bfs_rows = std::min(H, span);
    int wpr = (W + 31) / 32;
    L.bfs_wwords = std::min(wpr, span / 32 + 2);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align_up(o + bytes, 256);
        return at;
    };
    L.node_off = take((size_t)node_cap * sizeof(PlanNode));
    L.hkey_off = take((size_t)node_cap * 8);
    L.hid_off = take((size_t)node_cap * 4);
    L.table_off = take((size_t)slots * sizeof(ClosedSlot));
    size_t bw = (size_t)L.bfs_rows * L.bfs_wwords;
    L.blocked_off = take(bw * 4);
    L.f0_off = take(bw * 4);
    L.f1_off = take(bw * 4);
    L.dist_off = take(bw * 32 * 2);
    L.rowtag_off = take((size_t)L.bfs_rows * 4);
    L.traj_off = take((size_t)kTrajCap * 5 * 8);
    L.hdr_off = take(512); // ResumeHeader: state of a query handed from the squad pass to the team pass (plan_team.cuh)
    L.stride = o;
    return L;
}

struct PlanBatchArgs
{
    const double *starts, *goals; // nq x 3
    int nq;
    // outputs
    int32_t *status;   // nq
    double *cost;      // nq  (goal_node_->gvalue_)
    int32_t *path_len; // nq
    double *paths;     // nq x path_cap x 5 (may be null when path_cap == 0)
    int path_cap;
    int64_t *stats; // nq x 6 {iterations, primitives, rs_shots, collision_checks, goal_map_cells, nodes_inserted}
    // work distribution
    const int32_t *todo; // indices of the queries to run in this pass (null = all)
    int n_todo;
    unsigned int *counter;
    int shared_counter; // the ticket counter (and the result arrays) live in ANOTHER GPU's memory, mapped over NVLink: tickets
                        // are drawn with system-scope atomics (mpros_plan_batch_shared, dist.cu)
    const unsigned int *n_todo_dev; // when set: the length of `todo` is read from the device (hand-off list of the solo kernel)
    WorkspacePool pool; // a team takes one workspace for the lifetime of its CTA (plan_team.cuh)
    WarpWorkspaceLayout L;
    // solo pass (plan_solo.cuh): one warp per query with a small workspace; queries that outgrow it go to the hand-off list
    WorkspacePool solo_pool;
    WarpWorkspaceLayout Ls;
    // grown workspaces of the squad pass (plan_squad.cuh: squad_grow): a query that outgrows the small one moves into one of these
    WorkspacePool mid_pool; // n == 0: none
    WarpWorkspaceLayout Lm;
    unsigned int *solo_counter;
    int32_t *handoff;         // nq entries (null: overflowed queries report ST_OVERFLOW)
    int32_t *handoff_ws;      // nq entries: solo workspace that holds the query's state (-1: none, the team pass starts it afresh)
    unsigned int *handoff_n;
    const int32_t *todo_ws;   // team pass: handoff_ws of the list it runs (null: every query starts afresh)
    // throughput policy of the squad pass: hand-overs that another squad may adopt (plan_squad.cuh); mig_n == null: off
    int32_t *mig, *mig_ws;
    unsigned int *mig_n, *mig_taken; // entries [0, *mig_taken) have been adopted, [*mig_taken, *mig_n) are the team pass's
    int mig_cap;
    int migrate;                     // squad pass only: adoption on (the team pass reads the list whenever mig_n is set)
    int age_limit;                   // a query older than this many iterations goes to the team pass for good
    int solo_spare0, solo_spare_n; // workspaces [solo_spare0, solo_spare0 + solo_spare_n) of the solo pool start out unowned: a
                                   // warp that hands its workspace over with a query takes one of them
    const uint32_t *unsafe_tw; // pre-classification layer of the footprint test (map_device.cuh: footprint_preclass)
    int center_probe;
    RsConfig rs;
    int optimal;
    // optional expansion trace (mpros_plan_search_tree, single query): pose of every expanded node in expansion order
    double *trace;         // trace_cap x 3 {x, y, yaw}
    long long trace_cap;
    long long *trace_n;    // number of expansions (may exceed trace_cap: then the trace is truncated)
};

// Phase profile of the team planner (developer build only: make PROFILE=1). Cycle counts are taken by thread 0 of the
// compute group and lane 0 of the heap warp and summed over all queries.
#ifdef MPROS_PROFILE
__device__ unsigned long long g_prof[32];
struct Prof
{
    long long t0;
    unsigned long long *acc; // 24 accumulators in shared memory (TeamShared::prof); each slot has ONE writing thread
    static __device__ __forceinline__ long long now()
    {
        long long t;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
        return t;
    }
    __device__ __forceinline__ void start(unsigned long long *a)
    {
        acc = a;
        t0 = now();
    }
    __device__ __forceinline__ void restart() { t0 = now(); }
    __device__ __forceinline__ void mark(int k)
    {
        long long t1 = now();
        acc[k] += (unsigned long long)(t1 - t0);
        t0 = t1;
    }
    __device__ __forceinline__ void count(int k) { acc[k] += 1; }
};
#else
struct Prof
{
    __device__ __forceinline__ void start(unsigned long long *) {}
    __device__ __forceinline__ void restart() {}
    __device__ __forceinline__ void mark(int) {}
    __device__ __forceinline__ void count(int) {}
};
#endif

// status codes of mpros_plan_batch (include/mpros_gpu.h)
enum
{
    ST_FOUND = 1,
    ST_LIMIT = 0,
    ST_OPEN_EMPTY = -1,
    ST_NOT_INIT = -2,
    ST_OVERFLOW = -3
};

// heuristic of one node computed by a QUAD of lanes (hybrid_a_star.cpp:248-264); h1 is the goal-map value of the cell
__device__ __noinline__ double node_heuristic_quad(const MapView &m, const PlannerView &p, int h1_cells, double x, double y,
                                                   double yaw, double gx, double gy, double gyaw)
{
    double h1 = (double)h1_cells * m.res; // getGoalCost: cost_ * map_resolution_
    double h2 = ompl_rs_distance_quad(p.min_r, x, y, yaw, gx, gy, gyaw);
    return fmax(h1, h2) * (1 + 1e-3);
}

} // namespace mpros

#include "plan_team.cuh"
#include "plan_solo.cuh"
#include "plan_squad.cuh"

namespace mpros
{

// ---- standalone batches (R1-R4, E1-E5) ---------------------------------------------------------------------------------
struct RsBatchArgs
{
    const double *starts5; // n x 5 {x, y, yaw, dir, steer}
    const double *goals3;  // n x 3
    int n;
    double *segs;      // n x 15 {len, steer, dir} x 5
    int32_t *nseg;     // n
    double *cost;      // n
    double *traj;      // n x traj_cap x 5 (optional)
    int32_t *npts;     // n
    uint8_t *verdicts; // n (optional): pathInCollision of the sampled trajectory (R4)
    int traj_cap;
    char *ws;          // per-warp scratch: kTrajCap x 5 doubles
    RsConfig rs;
    const uint32_t *unsafe_tw = nullptr; // pre-classification layer of the footprint test (may be NULL)
    int center_probe = 0;
};

// Trajectory longer than one warp's 32 poses (rare inside the 10 m analytic range): the one-lane GenTraj into the warp's
// scratch, then the poses 32 at a time. Out of line: it must not sit in the hot loop's instruction stream.
__device__ __noinline__ bool rs_shot_long(const MapView &m, const PlannerView &p, const RsBatchArgs &a, const RsWord &w, double x0, double y0,
                                          double yaw0, int q, double *scratch, int &np_out)
{
    const int lane = threadIdx.x & 31;
    double *txyz = scratch, *tsd = scratch + 3 * kTrajCap;
    int np = 0;
    if (lane == 0)
    {
        RsPath path;
        path.n = w.n;
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (k < w.n)
                rs_word_seg(w, k, path.s[k].len, path.s[k].steer, path.s[k].dir);
        np = rs_gen_traj(a.rs, x0, y0, yaw0, path, txyz, tsd, kTrajCap);
    }
    np = __shfl_sync(kFull, np, 0);
    __syncwarp();
    np_out = np;
    const int stored = min(np, kTrajCap);
    if (a.traj)
    {
        double *dst = a.traj + (size_t)q * a.traj_cap * 5;
        for (int k = lane; k < stored && k < a.traj_cap; k += 32)
        {
            dst[5 * k] = txyz[3 * k], dst[5 * k + 1] = txyz[3 * k + 1], dst[5 * k + 2] = txyz[3 * k + 2];
            dst[5 * k + 3] = tsd[2 * k], dst[5 * k + 4] = tsd[2 * k + 1];
        }
    }
    bool hit = np > kTrajCap; // longer than the scratch: report blocked rather than judge a truncated path
    if (a.verdicts)
        for (int k0 = 0; k0 < stored && !hit; k0 += 32)
        {
            const int k = k0 + lane;
            int cls = 0;
            if (k < stored)
                cls = footprint_preclass(m, p, a.unsafe_tw, a.center_probe, txyz[3 * k], txyz[3 * k + 1], txyz[3 * k + 2]);
            if (__any_sync(kFull, cls == 1))
            {
                hit = true;
                break;
            }
            unsigned walk = __ballot_sync(kFull, cls == 2);
            double s = 0, c = 1;
            if (cls == 2)
                dm::sincos_(txyz[3 * k + 2], &s, &c);
            while (walk && !hit)
            {
                const int j = __ffs(walk) - 1;
                walk &= walk - 1;
                const double ts = __shfl_sync(kFull, s, j), tc = __shfl_sync(kFull, c, j);
                hit = warp_footprint_collision(m, p, txyz[3 * (k0 + j)], txyz[3 * (k0 + j) + 1], ts, tc);
            }
        }
    __syncwarp();
    return hit;
}

// R1-R4 for a batch of shots, in two kernels with the winning candidate (three doubles and the word number: 40 bytes)
// handed over in HBM:
//   rs_solve_kernel  FindOptimalPath + PathLength: a warp takes EIGHT shots at a time, lane = (shot, symmetry image), and
//                    walks the twelve words together (rs_find_optimal_quad: every lane in the same closed form, candidates in
//                    registers).
//   rs_traj_kernel   GenTraj + pathInCollision: one shot per warp step, one lane per pose, the poses never leave registers
//                    (rs_gen_traj_regs).
// History: one kernel, one shot per warp, the 48 candidates as per-thread segment arrays in local memory and the trajectory in
// a per-warp HBM scratch: 4 700 warp-instructions per shot with eight divergent formula bodies on four lanes each, 3.4 GB of
// DRAM traffic per 2^20 shots against 0.45 GB algorithmic, 85 M shots/s. Eight shots per warp with everything in registers:
// 2 440 warp-instructions, 0.29 GB, 232 M shots/s, but still instruction-fetch bound (sm__icc hit rate 67 %,
// stall_no_instruction 8.5 per issue): solve and trajectory code together do not fit the instruction cache while the 32
// resident warps of an SM are spread over both. Two kernels keep each SM inside one half of the code at a time.
struct RsWordRec
{
    double t, u, v;
    int f, j, n, pad;
};
#ifndef MPROS_RS_MINB
#define MPROS_RS_MINB 8 // 64 registers: 32 resident warps per SM
#endif
__global__ void __launch_bounds__(128, MPROS_RS_MINB) rs_solve_kernel(const __grid_constant__ RsBatchArgs a, RsWordRec *__restrict__ words)
{
    const int lane = threadIdx.x & 31, quad = lane >> 2, j = lane & 3;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int ngroups = (a.n + 7) >> 3;
    for (int g = warp_global; g < ngroups; g += nwarps)
    {
        const int q = 8 * g + quad;
        const bool live = q < a.n;
        const double *s5 = a.starts5 + 5 * (size_t)(live ? q : a.n - 1), *g3 = a.goals3 + 3 * (size_t)(live ? q : a.n - 1);
        double cost;
        const RsWord best = rs_find_optimal_quad(a.rs, s5[0], s5[1], s5[2], (int)s5[3], s5[4], g3[0], g3[1], g3[2], cost);
        if (live)
        {
            if (j == 0)
            {
                a.cost[q] = cost;
                a.nseg[q] = best.n;
                RsWordRec r;
                r.t = best.t, r.u = best.u, r.v = best.v, r.f = best.f, r.j = best.j, r.n = best.n, r.pad = 0;
                words[q] = r;
            }
            for (int k = j; k < 5; k += 4) // lane j of the quad stores segments j and j + 4
            {
                double len = 0;
                int steer = 0, dir = 0;
                if (k < best.n)
                    rs_word_seg(best, k, len, steer, dir);
                double *o = a.segs + 15 * (size_t)q + 3 * k;
                o[0] = len, o[1] = (double)steer, o[2] = (double)dir;
            }
        }
    }
}

__global__ void __launch_bounds__(128, MPROS_RS_MINB) rs_traj_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p,
                                                                      const __grid_constant__ RsBatchArgs a, const RsWordRec *__restrict__ words)
{
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    double *scratch = reinterpret_cast<double *>(a.ws) + (size_t)warp_global * kTrajCap * 5;
    for (int q = warp_global; q < a.n; q += nwarps)
    {
        const RsWordRec r = words[q]; // the same address on every lane: one broadcast transaction
        RsWord w;
        w.t = r.t, w.u = r.u, w.v = r.v, w.f = r.f, w.j = r.j, w.n = r.n;
        const double *s5 = a.starts5 + 5 * (size_t)q;
        const double x0 = s5[0], y0 = s5[1], yaw0 = s5[2];
        double px, py, pyaw;
        int pst, pdr;
        int np = rs_gen_traj_regs(a.rs, x0, y0, yaw0, w, px, py, pyaw, pst, pdr);
        bool hit = false;
        if (np > 32)
            hit = rs_shot_long(m, p, a, w, x0, y0, yaw0, q, scratch, np);
        else
        {
            const bool pose = lane < np;
            if (a.traj && pose && lane < a.traj_cap)
            {
                double *o = a.traj + ((size_t)q * a.traj_cap + lane) * 5;
                o[0] = px, o[1] = py, o[2] = pyaw, o[3] = (double)pst * a.rs.steer_angle, o[4] = (double)pdr;
            }
            if (a.verdicts)
            {
                // pathInCollision is an OR over the poses: every lane decides its pose by the pose's own cell where that is
                // possible (map_device.cuh: footprint_preclass); the undecided poses get their sin/cos in parallel and the
                // warp-wide probe walk one after the other
                const int cls = pose ? footprint_preclass(m, p, a.unsafe_tw, a.center_probe, px, py, pyaw) : 0;
                hit = __any_sync(kFull, cls == 1);
                unsigned walk = hit ? 0u : __ballot_sync(kFull, cls == 2);
                if (walk)
                {
                    double sn = 0, cs = 1;
                    if (cls == 2)
                        dm::sincos_(pyaw, &sn, &cs);
                    while (walk && !hit)
                    {
                        const int t = __ffs(walk) - 1;
                        walk &= walk - 1;
                        hit = warp_footprint_collision(m, p, __shfl_sync(kFull, px, t), __shfl_sync(kFull, py, t), __shfl_sync(kFull, sn, t),
                                                       __shfl_sync(kFull, cs, t));
                    }
                }
            }
        }
        if (lane == 0)
        {
            a.npts[q] = np;
            if (a.verdicts)
                a.verdicts[q] = hit ? 1 : 0;
        }
    }
}

// both kernels of a shot batch on `st`; d_words: n records of scratch
static int rs_launch(mpros_ctx *c, const RsBatchArgs &a, RsWordRec *d_words, cudaStream_t st)
{
    const int wpb = 4;
    const size_t cap = (size_t)c->n_sm * MPROS_RS_MINB;
    const int b_solve = (int)std::min<size_t>((((size_t)a.n + 7) / 8 + wpb - 1) / wpb, cap);
    const int b_traj = (int)std::min<size_t>(((size_t)a.n + wpb - 1) / wpb, cap);
    rs_solve_kernel<<<b_solve, 128, 0, st>>>(a, d_words);
    MPROS_LAUNCH_CHECK(c);
    rs_traj_kernel<<<b_traj, 128, 0, st>>>(c->view(), c->pv, a, d_words);
    MPROS_LAUNCH_CHECK(c);
    return MPROS_OK;
}

// OMPL-equivalent RS distance for n pose pairs (R5): one quad of lanes per pair
__global__ void __launch_bounds__(128) rs_distance_kernel(double rho, const double *__restrict__ a3, const double *__restrict__ b3, int n,
                                                          double *__restrict__ out)
{
    const int quad_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int nquads = (gridDim.x * blockDim.x) >> 2;
    const int rounds = (n + nquads - 1) / nquads;
    for (int r = 0; r < rounds; ++r)
    {
        int k = r * nquads + quad_global;
        bool live = k < n;
        int kk = live ? k : 0;
        double d = ompl_rs_distance_quad(rho, a3[3 * kk], a3[3 * kk + 1], a3[3 * kk + 2], b3[3 * kk], b3[3 * kk + 1], b3[3 * kk + 2]);
        if (live && (threadIdx.x & 3) == 0)
            out[k] = d;
    }
}

// RS::PathFunction::path1..12 (rs_curve.cpp:353-599) for n normalised goals: one thread per goal
__global__ void __launch_bounds__(128) rs_word_kernel(int word, const double *__restrict__ xyphi, int n, double *__restrict__ segs,
                                                      int32_t *__restrict__ nseg)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    {
        RsPath p;
        rs_formula(word, xyphi[3 * k], xyphi[3 * k + 1], xyphi[3 * k + 2], p);
        nseg[k] = p.n;
        for (int t = 0; t < 5; ++t)
        {
            const bool have = t < p.n;
            segs[15 * k + 3 * t] = have ? p.s[t].len : 0.0;
            segs[15 * k + 3 * t + 1] = have ? (double)p.s[t].steer : 0.0;
            segs[15 * k + 3 * t + 2] = have ? (double)p.s[t].dir : 0.0;
        }
    }
}

// E1-E5 for a batch of parents against the context's goal map (mpros_map_update_goal) with a FRESH closed set per
// parent (only the start/goal seeds), i.e. what HybridAstar::expandNode does right after initialize().
struct ExpandArgs
{
    const double *parents6; // n x 6 {x, y, yaw, steer, dir, g}
    int n;
    double gx, gy, gyaw;
    double *prims;   // n x P x 7 {x, y, yaw, steer, dir, collision, cell-valid}
    double *childs;  // n x P x 10 {x, y, yaw, steer, dir, g, h, f, i, j} (inserted children, in order)
    int32_t *nchild; // n
    unsigned long long start_idx, goal_idx;
    double start_g;
    int have_seeds;
    const uint32_t *unsafe_tw; // pre-classification layer of the footprint test (may be NULL)
    int center_probe;
    // compact output (mpros_expand_batch_compact): only the inserted children, appended as 80-byte records
    mpros_child_record *records = nullptr;
    unsigned long long record_cap = 0;
    unsigned long long *record_count = nullptr;
};

#ifndef MPROS_EXPAND_MINB
#define MPROS_EXPAND_MINB 8 // 64 registers, small spills: 32 resident warps hide the FP64 dependency chains (134 -> 160 M expansions/s)
#endif
__global__ void __launch_bounds__(128, MPROS_EXPAND_MINB) expand_batch_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p, ExpandArgs a)
{
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int P = 2 * p.n_steer;
    for (int q = warp_global; q < a.n; q += nwarps)
    {
        const double *pr = a.parents6 + 6 * (size_t)q;
        const double cx = pr[0], cy = pr[1], cyaw = pr[2], psteer = pr[3], pdir = pr[4], pg = pr[5];
        double s_yaw, c_yaw;
        dm::sincos_(cyaw, &s_yaw, &c_yaw);
        double nx = 0, ny = 0, nyaw = 0, steer = 0, direc = 1, ns = 0, nc = 1;
        if (lane < P)
        {
            primitive_end_pose(p, lane, cx, cy, cyaw, s_yaw, c_yaw, nx, ny, nyaw, steer, direc);
            dm::sincos_(nyaw, &ns, &nc);
        }
        // most end poses are decided by their own cell (surely free / sure hit, map_device.cuh: footprint_preclass), one lane
        // per child; only the undecided ones get the warp-wide probe walk
        const int cls = lane < P ? footprint_preclass(m, p, a.unsafe_tw, a.center_probe, nx, ny, nyaw) : 0;
        unsigned free_mask = __ballot_sync(kFull, lane < P && cls == 0);
        unsigned walk_mask = __ballot_sync(kFull, lane < P && cls == 2);
        while (walk_mask)
        {
            const int c = __ffs(walk_mask) - 1;
            walk_mask &= walk_mask - 1;
            double tx = __shfl_sync(kFull, nx, c), ty = __shfl_sync(kFull, ny, c);
            double ts = __shfl_sync(kFull, ns, c), tc = __shfl_sync(kFull, nc, c);
            if (!warp_footprint_collision(m, p, tx, ty, ts, tc))
                free_mask |= 1u << c;
        }
        bool alive = lane < P && ((free_mask >> lane) & 1u);
        int ci = -1, cj = -1;
        bool cell_ok = true;
        if (alive && !pos_to_index_ds(m, nx, ny, ci, cj))
        {
            alive = false;
            cell_ok = false;
        }
        int h1 = alive ? (int)m.goal[(size_t)ci * m.W + cj] : 0;
        double g = 0;
        if (alive)
            g = node_cost(p, pg, pdir, psteer, (double)__ldg(m.voronoi + (size_t)ci * m.W + cj), steer, direc);
        double hval = 0;
        for (int base = 0; base < P; base += 8)
        {
            int src = base + (lane >> 2);
            double qx = __shfl_sync(kFull, nx, src), qy = __shfl_sync(kFull, ny, src), qyaw = __shfl_sync(kFull, nyaw, src);
            int qh1 = __shfl_sync(kFull, h1, src);
            bool qalive = (__shfl_sync(kFull, (int)alive, src) != 0) && src < P;
            double hq = 0;
            if (__any_sync(kFull, qalive))
            {
                // h = max(h1, h2) * 1.001: when the CSC upper bound of h2 stays below h1 for every child of the group the
                // other nine Reeds-Shepp formulas cannot change the result (exact; the planner kernel skips the same way)
                const double h1v = (double)qh1 * m.res;
                const double ub = ompl_csc_upper_bound_quad(p.min_r, qalive ? qx : a.gx, qalive ? qy : a.gy, qalive ? qyaw : a.gyaw, a.gx, a.gy, a.gyaw);
                if (__all_sync(kFull, !qalive || ub <= h1v))
                    hq = h1v * (1 + 1e-3);
                else
                    hq = node_heuristic_quad(m, p, qh1, qalive ? qx : a.gx, qalive ? qy : a.gy, qalive ? qyaw : a.gyaw, a.gx, a.gy, a.gyaw);
            }
            int owner_quad = lane - base;
            double back = __shfl_sync(kFull, hq, (owner_quad >= 0 && owner_quad < 8) ? owner_quad * 4 : 0);
            if (owner_quad >= 0 && owner_quad < 8)
                hval = back;
        }
        if (lane < P && a.prims)
        {
            double *o = a.prims + ((size_t)q * P + lane) * 7;
            o[0] = nx, o[1] = ny, o[2] = nyaw, o[3] = steer, o[4] = direc, o[5] = ((free_mask >> lane) & 1u) ? 0.0 : 1.0;
            o[6] = cell_ok ? 1.0 : 0.0;
        }
        // prune against the seeded closed set (start g, goal 10000) and against earlier siblings, in primitive order
        unsigned long long idx = alive ? cal_index(m, p, ci, cj, yaw_to_idx(p, nyaw), direc) : ~0ull;
        unsigned alive_mask = __ballot_sync(kFull, alive);
        int count = 0, my_order = -1;
        for (int c = 0; c < P; ++c)
        {
            if (!((alive_mask >> c) & 1u))
                continue;
            unsigned long long tidx = __shfl_sync(kFull, idx, c);
            double tg = __shfl_sync(kFull, g, c);
            // best g already stored for this index by the seeds or by an inserted earlier sibling
            double old_cost = __longlong_as_double(0x7ff0000000000000ll);
            if (a.have_seeds)
            {
                if (tidx == a.start_idx)
                    old_cost = a.start_g;
                if (tidx == a.goal_idx)
                    old_cost = 10000.0;
            }
            // sequential semantics: walk the earlier siblings in order
            for (int e = 0; e < c; ++e)
            {
                if (!((alive_mask >> e) & 1u))
                    continue;
                unsigned long long eidx = __shfl_sync(kFull, idx, e);
                double eg = __shfl_sync(kFull, g, e);
                // sibling e was inserted iff it beat what was stored before it; replay that decision
                if (eidx == tidx && eg < old_cost)
                    old_cost = eg;
            }
            bool ins = old_cost > tg;
            if (ins)
            {
                if (lane == c)
                    my_order = count;
                if (lane == c && a.childs)
                {
                    double *o = a.childs + ((size_t)q * P + count) * 10;
                    o[0] = nx, o[1] = ny, o[2] = nyaw, o[3] = steer, o[4] = direc, o[5] = g, o[6] = hval, o[7] = g + hval;
                    o[8] = (double)ci, o[9] = (double)cj;
                }
                ++count;
            }
        }
        if (a.records)
        {
            // one reservation per parent; the children of a parent stay together and in insertion order
            unsigned long long at = 0;
            if (lane == 0 && count)
                at = atomicAdd(a.record_count, (unsigned long long)count);
            at = __shfl_sync(kFull, at, 0);
            if (my_order >= 0 && at + my_order < a.record_cap)
            {
                mpros_child_record r;
                r.parent = (uint32_t)q, r.prim = (uint16_t)lane, r.dir = (int16_t)(direc > 0 ? 1 : -1), r.i = ci, r.j = cj;
                r.x = nx, r.y = ny, r.yaw = nyaw, r.steer = steer, r.g = g, r.h = hval, r.f = g + hval;
                r.order = (uint32_t)my_order, r.reserved = 0;
                a.records[at + my_order] = r;
            }
        }
        if (lane == 0 && a.nchild)
            a.nchild[q] = count;
        __syncwarp();
    }
}

} // namespace mpros

// =====================================================================================================================
using namespace mpros;

static RsConfig rs_config_from_planner(const Ctx *c)
{
    // hybrid_a_star.cpp:142-143,169-170: setPenalty(GEAR, BACKWARD, STEER_ANGLE, STEER_CHANGE) lands positionally in
    // RSCurve's (gear, backward, steer_change, steer_angle) slots
    RsConfig r;
    r.radius = c->pv.min_r;
    r.steer_angle = c->pp.max_steering_angle;
    r.step = c->pp.step_size;
    r.pen_gear = c->pp.gear_penalty;
    r.pen_back = c->pp.reverse_penalty;
    r.pen_steer_change = c->pp.steering_penalty;
    r.pen_steer_angle = c->pp.steering_change_penalty;
    return r;
}

static int check_planner_ready(mpros_ctx *c, const char *who)
{
    if (!c)
    {
        set_error(std::string(who) + ": NULL context");
        return MPROS_ERR_INVALID;
    }
    if (!c->get_map || !c->planner_configured)
    {
        set_error(std::string(who) + ": needs a map (mpros_map_set_occupancy) and mpros_planner_config first");
        return MPROS_ERR_INVALID;
    }
    return MPROS_OK;
}

extern "C" int mpros_rs_default_config(mpros_ctx *c, mpros_rs_params *out)
{
    int rc = check_planner_ready(c, "mpros_rs_default_config");
    if (rc)
        return rc;
    RsConfig r = rs_config_from_planner(c);
    out->radius = r.radius;
    out->steering_angle = r.steer_angle;
    out->step_size = r.step;
    out->pen_gear = r.pen_gear;
    out->pen_backward = r.pen_back;
    out->pen_steer_change = r.pen_steer_change;
    out->pen_steer_angle = r.pen_steer_angle;
    return MPROS_OK;
}

static int rs_batch_impl(mpros_ctx *c, const mpros_rs_params *rp, const double *starts5, const double *goals3, size_t n,
                         double *segs, int32_t *nseg, double *cost, double *traj, int traj_cap, int32_t *npts, uint8_t *verdicts,
                         bool with_collision)
{
    if (n == 0)
        return MPROS_OK;
    if (!starts5 || !goals3 || !segs || !nseg || !cost || !npts || (with_collision && !verdicts) || traj_cap < 0 ||
        (traj_cap > 0 && !traj) || n > (size_t)INT32_MAX)
    {
        set_error("mpros_rs_*_batch: NULL / invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    RsBatchArgs a;
    if (rp)
    {
        a.rs.radius = rp->radius;
        a.rs.steer_angle = rp->steering_angle;
        a.rs.step = rp->step_size;
        a.rs.pen_gear = rp->pen_gear;
        a.rs.pen_back = rp->pen_backward;
        a.rs.pen_steer_change = rp->pen_steer_change;
        a.rs.pen_steer_angle = rp->pen_steer_angle;
    }
    else
        a.rs = rs_config_from_planner(c);
    // per-warp scratch of the (rare) trajectories longer than 32 poses; MPROS_RS_MINB CTAs of 4 warps are resident per SM
    size_t nwarps = (size_t)c->n_sm * MPROS_RS_MINB * 4;
    size_t in_bytes = n * 8 * 8, seg_bytes = n * 15 * 8, traj_bytes = n * (size_t)traj_cap * 5 * 8;
    size_t ws_bytes = nwarps * kTrajCap * 5 * 8, word_bytes = n * sizeof(RsWordRec);
    size_t total = in_bytes + seg_bytes + n * 4 + n * 8 + traj_bytes + n * 4 + n + ws_bytes + word_bytes + 4096;
    int rc = ensure_stage(c, total);
    if (rc)
        return rc;
    char *base = static_cast<char *>(c->stage);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        char *p = base + o;
        o = (o + bytes + 255) / 256 * 256;
        return p;
    };
    double *d_s = (double *)take(n * 5 * 8), *d_g = (double *)take(n * 3 * 8), *d_seg = (double *)take(seg_bytes);
    int32_t *d_nseg = (int32_t *)take(n * 4);
    double *d_cost = (double *)take(n * 8);
    double *d_traj = traj_cap ? (double *)take(traj_bytes) : nullptr;
    int32_t *d_npts = (int32_t *)take(n * 4);
    uint8_t *d_ver = (uint8_t *)take(n);
    char *d_ws = take(ws_bytes);
    RsWordRec *d_words = (RsWordRec *)take(word_bytes);
    if (o > c->stage_bytes)
    {
        rc = ensure_stage(c, o + 1024);
        set_error("internal: staging layout");
        return MPROS_ERR_INVALID;
    }
    cudaStream_t st = c->stream;
    MPROS_CUDA(cudaMemcpyAsync(d_s, starts5, n * 5 * 8, cudaMemcpyHostToDevice, st));
    MPROS_CUDA(cudaMemcpyAsync(d_g, goals3, n * 3 * 8, cudaMemcpyHostToDevice, st));
    a.starts5 = d_s;
    a.goals3 = d_g;
    a.n = (int)n;
    a.segs = d_seg;
    a.nseg = d_nseg;
    a.cost = d_cost;
    a.traj = d_traj;
    a.npts = d_npts;
    a.verdicts = with_collision ? d_ver : nullptr;
    a.traj_cap = traj_cap;
    a.ws = d_ws;
    if (with_collision)
    {
        rc = map_build_unsafe(c);
        if (rc)
            return rc;
        a.unsafe_tw = c->unsafe_tw;
        a.center_probe = c->unsafe_center_probe;
    }
    rc = rs_launch(c, a, d_words, st);
    if (rc)
        return rc;
    MPROS_CUDA(cudaMemcpyAsync(segs, d_seg, seg_bytes, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaMemcpyAsync(nseg, d_nseg, n * 4, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaMemcpyAsync(cost, d_cost, n * 8, cudaMemcpyDeviceToHost, st));
    if (traj_cap)
        MPROS_CUDA(cudaMemcpyAsync(traj, d_traj, traj_bytes, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaMemcpyAsync(npts, d_npts, n * 4, cudaMemcpyDeviceToHost, st));
    if (with_collision)
        MPROS_CUDA(cudaMemcpyAsync(verdicts, d_ver, n, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaStreamSynchronize(st));
    return MPROS_OK;
}

extern "C" int mpros_rs_solve_batch(mpros_ctx *c, const mpros_rs_params *rp, const double *starts5, const double *goals3, size_t n,
                                    double *segs, int32_t *nseg, double *cost, double *traj, int traj_cap, int32_t *npts)
{
    int rc = check_planner_ready(c, "mpros_rs_solve_batch");
    if (rc)
        return rc;
    return rs_batch_impl(c, rp, starts5, goals3, n, segs, nseg, cost, traj, traj_cap, npts, nullptr, false);
}

extern "C" int mpros_rs_shot_batch(mpros_ctx *c, const double *starts5, const double *goals3, size_t n, double *segs, int32_t *nseg,
                                   double *cost, double *traj, int traj_cap, int32_t *npts, uint8_t *verdicts)
{
    int rc = check_planner_ready(c, "mpros_rs_shot_batch");
    if (rc)
        return rc;
    return rs_batch_impl(c, nullptr, starts5, goals3, n, segs, nseg, cost, traj, traj_cap, npts, verdicts, true);
}

// genRSPath's shot with every buffer in device memory, asynchronous on the context's stream; no trajectory output
extern "C" int mpros_rs_shot_batch_dev(mpros_ctx *c, const double *d_starts5, const double *d_goals3, size_t n, double *d_segs,
                                       int32_t *d_nseg, double *d_cost, int32_t *d_npts, uint8_t *d_verdicts)
{
    int rc = check_planner_ready(c, "mpros_rs_shot_batch_dev");
    if (rc)
        return rc;
    if (n == 0)
        return MPROS_OK;
    if (!d_starts5 || !d_goals3 || !d_segs || !d_nseg || !d_cost || !d_npts || !d_verdicts || n > (size_t)INT32_MAX)
    {
        set_error("mpros_rs_shot_batch_dev: NULL / invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    RsBatchArgs a;
    a.rs = rs_config_from_planner(c);
    const size_t ws_bytes = ((size_t)c->n_sm * MPROS_RS_MINB * 4 * kTrajCap * 5 * 8 + 255) & ~(size_t)255, word_bytes = n * sizeof(RsWordRec);
    rc = map_build_unsafe(c); // before the scratch is handed out: the builder uses it for its temporaries
    if (rc)
        return rc;
    rc = ensure_scratch(c, ws_bytes + word_bytes + 256);
    if (rc)
        return rc;
    a.starts5 = d_starts5;
    a.goals3 = d_goals3;
    a.n = (int)n;
    a.segs = d_segs;
    a.nseg = d_nseg;
    a.cost = d_cost;
    a.traj = nullptr;
    a.npts = d_npts;
    a.verdicts = d_verdicts;
    a.traj_cap = 0;
    a.unsafe_tw = c->unsafe_tw;
    a.center_probe = c->unsafe_center_probe;
    a.ws = static_cast<char *>(c->scratch);
    return rs_launch(c, a, reinterpret_cast<RsWordRec *>(static_cast<char *>(c->scratch) + ws_bytes), c->stream);
}

extern "C" int mpros_rs_distance_batch(mpros_ctx *c, double rho, const double *a3, const double *b3, size_t n, double *out)
{
    if (!c)
    {
        set_error("mpros_rs_distance_batch: NULL context");
        return MPROS_ERR_INVALID;
    }
    if (n == 0)
        return MPROS_OK;
    if (!a3 || !b3 || !out || !(rho > 0) || n > (size_t)INT32_MAX)
    {
        set_error("mpros_rs_distance_batch: NULL / invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    int rc = ensure_stage(c, n * 7 * 8 + 1024);
    if (rc)
        return rc;
    double *d_a = static_cast<double *>(c->stage), *d_b = d_a + 3 * n, *d_o = d_b + 3 * n;
    cudaStream_t st = c->stream;
    MPROS_CUDA(cudaMemcpyAsync(d_a, a3, n * 24, cudaMemcpyHostToDevice, st));
    MPROS_CUDA(cudaMemcpyAsync(d_b, b3, n * 24, cudaMemcpyHostToDevice, st));
    int blocks = (int)std::min<size_t>((n * 4 + 127) / 128, (size_t)c->n_sm * 16);
    rs_distance_kernel<<<blocks, 128, 0, st>>>(rho, d_a, d_b, (int)n, d_o);
    MPROS_LAUNCH_CHECK(c);
    MPROS_CUDA(cudaMemcpyAsync(out, d_o, n * 8, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaStreamSynchronize(st));
    return MPROS_OK;
}

extern "C" int mpros_rs_word_batch(mpros_ctx *c, int word, const double *xyphi, size_t n, double *segs, int32_t *nseg)
{
    if (!c || word < 1 || word > 12 || (n && (!xyphi || !segs || !nseg)) || n > (size_t)INT32_MAX)
    {
        set_error("mpros_rs_word_batch: NULL / invalid argument (word is 1..12)");
        return MPROS_ERR_INVALID;
    }
    if (n == 0)
        return MPROS_OK;
    MPROS_CUDA(cudaSetDevice(c->device));
    const size_t in_b = align_up(n * 24, 256), seg_b = align_up(n * 120, 256);
    int rc = ensure_stage(c, in_b + seg_b + n * 4 + 256);
    if (rc)
        return rc;
    char *base = static_cast<char *>(c->stage);
    double *d_in = (double *)base, *d_seg = (double *)(base + in_b);
    int32_t *d_n = (int32_t *)(base + in_b + seg_b);
    cudaStream_t st = c->stream;
    MPROS_CUDA(cudaMemcpyAsync(d_in, xyphi, n * 24, cudaMemcpyHostToDevice, st));
    int blocks = (int)std::min<size_t>((n + 127) / 128, (size_t)c->n_sm * 16);
    rs_word_kernel<<<blocks, 128, 0, st>>>(word - 1, d_in, (int)n, d_seg, d_n);
    MPROS_LAUNCH_CHECK(c);
    MPROS_CUDA(cudaMemcpyAsync(segs, d_seg, n * 120, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaMemcpyAsync(nseg, d_n, n * 4, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaStreamSynchronize(st));
    return MPROS_OK;
}

// expandNode for n parents of one query; device buffers, asynchronous on the context's stream
static int expand_launch(mpros_ctx *c, const char *who, const double *start3, const double *goal3, const double *d_parents6, size_t n,
                         double *d_prims, double *d_childs, int32_t *d_nchild, mpros_child_record *d_records = nullptr, size_t record_cap = 0,
                         unsigned long long *d_record_count = nullptr)
{
    int rc = check_planner_ready(c, who);
    if (rc)
        return rc;
    if (!c->get_goal)
    {
        set_error(std::string(who) + ": no goal map (mpros_map_update_goal first)"); // hybrid_a_star.cpp:96-100
        return MPROS_ERR_INVALID;
    }
    if (n == 0)
        return MPROS_OK;
    const bool compact = d_records != nullptr;
    if (!goal3 || !d_parents6 || (!compact && (!d_prims || !d_childs || !d_nchild)) || (compact && !d_record_count) || n > (size_t)INT32_MAX)
    {
        set_error(std::string(who) + ": NULL / invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    const int P = 2 * c->pv.n_steer;
    ExpandArgs a;
    a.n = (int)n;
    a.gx = goal3[0], a.gy = goal3[1], a.gyaw = goal3[2];
    a.have_seeds = 0;
    a.start_idx = a.goal_idx = ~0ull;
    a.start_g = 0.0;
    // closed-set seeds of initialize() (hybrid_a_star.cpp:107-112), computed like the device does
    auto cell_index = [&](const double *pose, unsigned long long &idx) {
        double dx = pose[0] - c->ox, dy = pose[1] - c->oy;
        double qj = (dx * c->ocos + dy * c->osin) / c->res, qi = (dx * -c->osin + dy * c->ocos) / c->res;
        if (!(qi > -1.0 && qi < (double)c->H && qj > -1.0 && qj < (double)c->W))
            return false;
        double yaw = pose[2], yaw_pos = yaw < 0 ? yaw + 2 * M_PI : yaw;
        int iyaw = (int)(yaw_pos / c->pv.yaw_resol);
        idx = ((unsigned long long)((long long)(int)qi * c->W + (int)qj) * c->pv.num_yaw + (unsigned long long)iyaw) * 2ull + 0;
        return true;
    };
    if (start3)
    {
        unsigned long long si, gi;
        if (cell_index(start3, si) && cell_index(goal3, gi))
        {
            a.have_seeds = 1;
            a.start_idx = si;
            a.goal_idx = gi;
            if (si == gi)
                a.start_idx = ~0ull; // the goal seed overwrites a shared cell
        }
    }
    cudaStream_t st = c->stream;
    rc = map_build_unsafe(c);
    if (rc)
        return rc;
    a.unsafe_tw = c->unsafe_tw;
    a.center_probe = c->unsafe_center_probe;
    if (d_childs)
        MPROS_CUDA(cudaMemsetAsync(d_childs, 0, n * P * 10 * 8, st));
    if (compact)
        MPROS_CUDA(cudaMemsetAsync(d_record_count, 0, 8, st));
    a.records = d_records;
    a.record_cap = record_cap;
    a.record_count = d_record_count;
    a.parents6 = d_parents6;
    a.prims = d_prims;
    a.childs = d_childs;
    a.nchild = d_nchild;
    // one warp per parent, 4 warps per CTA; grid = a multiple of the SM count (16 resident CTAs each), grid-stride beyond
    int blocks = (int)std::min<size_t>((n + 3) / 4, (size_t)c->n_sm * 16);
    expand_batch_kernel<<<blocks, 128, 0, st>>>(c->view(), c->pv, a);
    MPROS_LAUNCH_CHECK(c);
    return MPROS_OK;
}

extern "C" int mpros_expand_batch_dev(mpros_ctx *c, const double *start3, const double *goal3, const double *d_parents6, size_t n,
                                      double *d_prims, double *d_childs, int32_t *d_nchild)
{
    return expand_launch(c, "mpros_expand_batch_dev", start3, goal3, d_parents6, n, d_prims, d_childs, d_nchild);
}

extern "C" int mpros_expand_batch_compact_dev(mpros_ctx *c, const double *start3, const double *goal3, const double *d_parents6, size_t n,
                                              mpros_child_record *d_records, size_t cap_records, unsigned long long *d_n_records,
                                              int32_t *d_nchild)
{
    return expand_launch(c, "mpros_expand_batch_compact_dev", start3, goal3, d_parents6, n, nullptr, nullptr, d_nchild, d_records, cap_records,
                         d_n_records);
}

extern "C" int mpros_expand_batch_compact(mpros_ctx *c, const double *start3, const double *goal3, const double *parents6, size_t n,
                                          mpros_child_record *records, size_t cap_records, size_t *n_records, int32_t *nchild)
{
    int rc = check_planner_ready(c, "mpros_expand_batch_compact");
    if (rc)
        return rc;
    if (!n_records)
    {
        set_error("mpros_expand_batch_compact: NULL n_records");
        return MPROS_ERR_INVALID;
    }
    *n_records = 0;
    if (n == 0)
        return MPROS_OK;
    if (!goal3 || !parents6 || (cap_records && !records) || n > (size_t)INT32_MAX)
    {
        set_error("mpros_expand_batch_compact: NULL / invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    const int P = 2 * c->pv.n_steer;
    const size_t in_b = n * 6 * 8, cap_dev = std::min(cap_records, n * (size_t)P), rec_b = cap_dev * sizeof(mpros_child_record);
    rc = ensure_stage(c, in_b + rec_b + n * 4 + 2048);
    if (rc)
        return rc;
    char *base = static_cast<char *>(c->stage);
    double *d_in = (double *)base;
    mpros_child_record *d_rec = (mpros_child_record *)(base + align_up(in_b, 256));
    int32_t *d_nc = (int32_t *)((char *)d_rec + align_up(rec_b + 16, 256));
    unsigned long long *d_cnt = (unsigned long long *)((char *)d_nc + align_up(n * 4, 256));
    cudaStream_t st = c->stream;
    MPROS_CUDA(cudaMemcpyAsync(d_in, parents6, in_b, cudaMemcpyHostToDevice, st));
    // a NULL record buffer with capacity 0 only counts
    rc = expand_launch(c, "mpros_expand_batch_compact", start3, goal3, d_in, n, nullptr, nullptr, d_nc, d_rec, cap_dev, d_cnt);
    if (rc)
        return rc;
    unsigned long long cnt = 0;
    MPROS_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, st));
    if (nchild)
        MPROS_CUDA(cudaMemcpyAsync(nchild, d_nc, n * 4, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaStreamSynchronize(st));
    *n_records = (size_t)cnt;
    // only the records that exist travel back: ~2 children x 80 B per expansion instead of 816 B of fixed-size arrays
    const size_t have = std::min<size_t>((size_t)cnt, cap_dev);
    if (have)
        MPROS_CUDA(cudaMemcpy(records, d_rec, have * sizeof(mpros_child_record), cudaMemcpyDeviceToHost));
    if (cnt > cap_records)
    {
        set_error("mpros_expand_batch_compact: record buffer too small (n_records holds the required size)");
        return MPROS_ERR_CAPACITY;
    }
    return MPROS_OK;
}

extern "C" int mpros_expand_batch(mpros_ctx *c, const double *start3, const double *goal3, const double *parents6, size_t n,
                                  double *prims, double *childs, int32_t *nchild)
{
    int rc = check_planner_ready(c, "mpros_expand_batch");
    if (rc)
        return rc;
    if (n == 0)
        return MPROS_OK;
    if (!goal3 || !parents6 || !prims || !childs || !nchild || n > (size_t)INT32_MAX)
    {
        set_error("mpros_expand_batch: NULL / invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    const int P = 2 * c->pv.n_steer;
    size_t in_b = n * 6 * 8, pr_b = n * P * 7 * 8, ch_b = n * P * 10 * 8;
    rc = ensure_stage(c, in_b + pr_b + ch_b + n * 4 + 2048);
    if (rc)
        return rc;
    char *base = static_cast<char *>(c->stage);
    double *d_in = (double *)base;
    double *d_pr = (double *)(base + align_up(in_b, 256));
    double *d_ch = (double *)((char *)d_pr + align_up(pr_b, 256));
    int32_t *d_nc = (int32_t *)((char *)d_ch + align_up(ch_b, 256));
    cudaStream_t st = c->stream;
    MPROS_CUDA(cudaMemcpyAsync(d_in, parents6, in_b, cudaMemcpyHostToDevice, st));
    rc = expand_launch(c, "mpros_expand_batch", start3, goal3, d_in, n, d_pr, d_ch, d_nc);
    if (rc)
        return rc;
    MPROS_CUDA(cudaMemcpyAsync(prims, d_pr, pr_b, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaMemcpyAsync(childs, d_ch, ch_b, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaMemcpyAsync(nchild, d_nc, n * 4, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaStreamSynchronize(st));
    return MPROS_OK;
}

// ---- whole-query planner -------------------------------------------------------------------------------------------------
namespace mpros
{
#ifndef MPROS_TEAM_WARPS
#define MPROS_TEAM_WARPS 8
#endif
#ifndef MPROS_TEAM_MIN_BLOCKS
#define MPROS_TEAM_MIN_BLOCKS 2
#endif
constexpr int kTeamWarps = MPROS_TEAM_WARPS;
constexpr int kTeamMinBlocks = MPROS_TEAM_MIN_BLOCKS;

static_assert(sizeof(WarpWorkspaceLayout) <= sizeof(PlanPool::last_layout), "PlanPool::last_layout too small");

// every batch in flight on any lane of the context is complete (the pool is about to be re-allocated or cleared)
static int plan_quiesce(mpros_ctx *c)
{
    MPROS_CUDA(cudaStreamSynchronize(c->stream));
    for (int l = 1; l < MPROS_PLAN_LANES; ++l)
        if (c->lane_stream[l])
            MPROS_CUDA(cudaStreamSynchronize(c->lane_stream[l]));
    return MPROS_OK;
}

#ifndef MPROS_SOLO_OFF
// Solo pass (plan_solo.cuh): every query of the batch on its own warp with a small workspace; the queries that outgrow it are
// appended to the lane's hand-off list, which the team pass behind it on the same stream reads from the device.
static int plan_solo_pass(mpros_ctx *c, int lane, cudaStream_t stream, PlanBatchArgs &a, int nq, long long max_nodes)
{
    PlanScratch &S = c->plan_scratch[lane][0];
    PlanPool &PL = c->solo_pool;
    const WarpWorkspaceLayout L = make_layout(c->opt.solo_nodes, c->H, c->W);
    // Two forms of the pass: plan_squad_kernel (W queries per CTA, heuristics evaluated by the whole CTA; needs 2 * n_steer <=
    // kSquadMaxReq) and plan_solo_kernel (independent warps; any configuration).
    const bool squad_ok = !c->opt.plan_nosquad && 2 * c->pv.n_steer <= kSquadMaxReq;
    if (!c->solo_resident)
    {
        int r = 0, rs = 0;
        MPROS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, plan_solo_kernel<MPROS_SOLO_WARPS, MPROS_SOLO_MINB>, 32 * MPROS_SOLO_WARPS, 0));
        MPROS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rs, plan_squad_kernel<MPROS_SQUAD_WARPS, MPROS_SQUAD_MINB>, 32 * MPROS_SQUAD_WARPS, 0));
        c->solo_resident = std::max(1, r);
        c->squad_resident = std::max(1, rs);
    }
    // one workspace per warp the device can keep resident, however many batches are in flight (as for the team pool)
    const int warps = squad_ok ? c->n_sm * c->squad_resident * MPROS_SQUAD_WARPS : c->n_sm * c->solo_resident * MPROS_SOLO_WARPS;
    const bool layout_change = std::memcmp(PL.last_layout, &L, sizeof(L)) != 0;
    // the squad form hands long queries over WITH their workspace: that many spare workspaces on top (WorkspacePool::custody)
    const int spare = squad_ok ? std::min(nq, std::max(256, warps * 3 / 8)) : 0;
    const int wanted = (int)std::max<long long>(layout_change ? 0 : PL.n, std::min<long long>(warps, (long long)nq * 2) + spare);
    if (layout_change || PL.n < wanted)
    {
        size_t free_b = 0, total_b = 0;
        MPROS_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t budget = (free_b + PL.bytes) / 2; // half of the free memory stays with the caller (and the team pool)
        const int n = (int)std::min<long long>(wanted, std::max<long long>(1, (long long)(budget / L.stride)));
        if (layout_change || n > PL.n)
        {
            int rc = plan_quiesce(c);
            if (rc)
                return rc;
            const size_t need = (size_t)n * L.stride;
            if (need > PL.bytes)
            {
                if (PL.base)
                    MPROS_CUDA(cudaFree(PL.base));
                PL.base = nullptr;
                PL.bytes = 0;
                MPROS_CUDA(cudaMalloc(&PL.base, need));
                PL.bytes = need;
            }
            if (n > PL.meta_cap)
            {
                if (PL.owner)
                    MPROS_CUDA(cudaFree(PL.owner));
                if (PL.gen)
                    MPROS_CUDA(cudaFree(PL.gen));
                PL.owner = nullptr, PL.gen = nullptr, PL.meta_cap = 0;
                MPROS_CUDA(cudaMalloc(&PL.owner, (size_t)n * 4 + 64)); // + the custody counter behind the flags
                MPROS_CUDA(cudaMalloc(&PL.gen, (size_t)n * 4));
                PL.meta_cap = n;
            }
            MPROS_CUDA(cudaMemsetAsync(PL.base, 0, PL.bytes, stream));
            MPROS_CUDA(cudaMemsetAsync(PL.owner, 0, (size_t)PL.meta_cap * 4 + 64, stream));
            MPROS_CUDA(cudaMemsetAsync(PL.gen, 0, (size_t)PL.meta_cap * 4, stream));
            MPROS_CUDA(cudaStreamSynchronize(stream));
            PL.n = n;
            PL.stride = L.stride;
            PL.issued = 0;
            std::memcpy(PL.last_layout, &L, sizeof(L));
        }
    }
    if (PL.issued + (unsigned long long)nq + 2ull >= (1ull << 30)) // generation tags stay below 2^30
    {
        int rc = plan_quiesce(c);
        if (rc)
            return rc;
        MPROS_CUDA(cudaMemsetAsync(PL.base, 0, PL.bytes, stream));
        MPROS_CUDA(cudaMemsetAsync(PL.gen, 0, (size_t)PL.meta_cap * 4, stream));
        MPROS_CUDA(cudaStreamSynchronize(stream));
        PL.issued = 0;
    }
    PL.issued += (unsigned long long)nq + 1ull;
    // grown workspaces (squad_grow): eight per SM, four times the nodes; left out when memory is short
    PlanPool &PM = c->mid_pool;
    WarpWorkspaceLayout Lmid = make_layout((int)std::min<long long>((long long)c->opt.solo_nodes * 4, max_nodes), c->H, c->W);
    if (squad_ok && c->opt.plan_grow)
    {
        const int want_mid = std::min(nq, 8 * c->n_sm);
        const bool mid_change = std::memcmp(PM.last_layout, &Lmid, sizeof(Lmid)) != 0;
        if (mid_change || PM.n < want_mid)
        {
            size_t free_b = 0, total_b = 0;
            MPROS_CUDA(cudaMemGetInfo(&free_b, &total_b));
            // the team pool (allocated after this one) must still fit into half of what is left
            const WarpWorkspaceLayout big = make_layout((int)max_nodes, c->H, c->W);
            const size_t team_need = c->plan_pool[0].bytes ? 0 : big.stride * (size_t)(kTeamMinBlocks * c->n_sm);
            const size_t avail = free_b + PM.bytes > 2 * team_need ? (free_b + PM.bytes - 2 * team_need) / 2 : 0;
            const int n = (int)std::min<long long>(want_mid, (long long)(avail / Lmid.stride));
            if (n >= 1 && (mid_change || n > PM.n))
            {
                int rc = plan_quiesce(c);
                if (rc)
                    return rc;
                const size_t need = (size_t)n * Lmid.stride;
                if (need > PM.bytes)
                {
                    if (PM.base)
                        MPROS_CUDA(cudaFree(PM.base));
                    PM.base = nullptr;
                    PM.bytes = 0;
                    MPROS_CUDA(cudaMalloc(&PM.base, need));
                    PM.bytes = need;
                }
                if (n > PM.meta_cap)
                {
                    if (PM.owner)
                        MPROS_CUDA(cudaFree(PM.owner));
                    if (PM.gen)
                        MPROS_CUDA(cudaFree(PM.gen));
                    PM.owner = nullptr, PM.gen = nullptr, PM.meta_cap = 0;
                    MPROS_CUDA(cudaMalloc(&PM.owner, (size_t)n * 4));
                    MPROS_CUDA(cudaMalloc(&PM.gen, (size_t)n * 4));
                    PM.meta_cap = n;
                }
                MPROS_CUDA(cudaMemsetAsync(PM.base, 0, PM.bytes, stream));
                MPROS_CUDA(cudaMemsetAsync(PM.owner, 0, (size_t)PM.meta_cap * 4, stream));
                MPROS_CUDA(cudaMemsetAsync(PM.gen, 0, (size_t)PM.meta_cap * 4, stream));
                MPROS_CUDA(cudaStreamSynchronize(stream));
                PM.n = n;
                PM.stride = Lmid.stride;
                PM.issued = 0;
                std::memcpy(PM.last_layout, &Lmid, sizeof(Lmid));
            }
            else if (mid_change)
                PM.n = 0; // no room: the squad pass works without grown workspaces
        }
        if (PM.n > 0)
        {
            if (PM.issued + (unsigned long long)nq + 2ull >= (1ull << 30))
            {
                int rc = plan_quiesce(c);
                if (rc)
                    return rc;
                MPROS_CUDA(cudaMemsetAsync(PM.base, 0, PM.bytes, stream));
                MPROS_CUDA(cudaMemsetAsync(PM.gen, 0, (size_t)PM.meta_cap * 4, stream));
                MPROS_CUDA(cudaStreamSynchronize(stream));
                PM.issued = 0;
            }
            PM.issued += (unsigned long long)nq + 1ull;
        }
    }
    if (!S.counter)
        MPROS_CUDA(cudaMalloc(&S.counter, 256));
    if ((size_t)nq > S.handoff_cap)
    {
        // the list may still be read by the lane's previous batch: the lane is idle by contract (submit after wait), and the
        // stream order covers the synchronous entry points
        MPROS_CUDA(cudaStreamSynchronize(stream));
        if (S.handoff)
            MPROS_CUDA(cudaFree(S.handoff));
        S.handoff = nullptr;
        S.handoff_cap = 0;
        MPROS_CUDA(cudaMalloc(&S.handoff, ((size_t)nq + 256) * 8)); // query indices, then the workspaces that travel with them
        S.handoff_cap = (size_t)nq + 256;
    }
    MPROS_CUDA(cudaMemsetAsync(S.counter, 0, 256, stream));
    // pre-classification layer of the footprint test; built on the context's own stream when the map or the footprint changed
    {
        const bool was_valid = c->unsafe_valid;
        int rc = map_build_unsafe(c);
        if (rc)
            return rc;
        if (!was_valid && stream != c->stream)
            MPROS_CUDA(cudaStreamSynchronize(c->stream));
    }
    a.unsafe_tw = c->unsafe_tw;
    a.center_probe = c->unsafe_center_probe;
    a.Ls = L;
    a.solo_pool.base = PL.base;
    a.solo_pool.owner = PL.owner;
    a.solo_pool.gen = PL.gen;
    a.solo_pool.n = PL.n;
    a.solo_counter = S.counter + 16;
    a.handoff = S.handoff;
    a.handoff_n = S.counter + 32;
    // a squad CTA must own a workspace for each of its warps (a warp waiting for one would stall its CTA at the round barrier)
    const bool squad = squad_ok && PL.n >= MPROS_SQUAD_WARPS;
    // workspaces beyond one per resident warp may be in custody of the team pass at any time (a memory-limited pool: none)
    const int custody_max = squad ? std::max(0, PL.n - std::min(warps, PL.n)) : 0;
    a.solo_pool.custody = PL.owner + PL.meta_cap;
    a.solo_pool.custody_max = custody_max;
    a.handoff_ws = (squad && custody_max > 0) ? S.handoff + S.handoff_cap : nullptr;
    a.todo_ws = a.handoff_ws;
    // Policy. Alone on the device, a batch is done when its longest queries are: the squad pass hands what is still running
    // to the team pass as soon as its CTAs thin out (a team steps a query in ~7 us, a squad in ~50). With other batches in
    // flight the device never idles and cost per iteration is what counts: the long queries stay in squad form (idle warps of
    // well-filled squads adopt them from the migration list) and only the oldest go to the team pass.
    a.migrate = 0;
    a.mig_n = nullptr;
    a.age_limit = c->opt.plan_age_limit;
    if (a.handoff_ws && c->opt.plan_policy != 1)
    {
        bool busy = c->opt.plan_policy == 2;
        for (int l = 0; l < MPROS_PLAN_LANES && !busy; ++l)
        {
            cudaStream_t st = l == 0 ? c->stream : c->lane_stream[l];
            if (l != lane && st && cudaStreamQuery(st) == cudaErrorNotReady)
                busy = true;
        }
        (void)cudaGetLastError(); // cudaErrorNotReady is not an error
        if (busy)
        {
            const size_t cap = (size_t)nq * 4 + 16384;
            if (cap > S.mig_cap)
            {
                MPROS_CUDA(cudaStreamSynchronize(stream));
                if (S.mig)
                    MPROS_CUDA(cudaFree(S.mig));
                S.mig = nullptr;
                S.mig_cap = 0;
                MPROS_CUDA(cudaMalloc(&S.mig, cap * 8));
                S.mig_cap = cap;
            }
            a.mig = S.mig;
            a.mig_ws = S.mig + S.mig_cap;
            a.mig_cap = (int)S.mig_cap;
            a.mig_n = S.counter + 48;
            a.mig_taken = S.counter + 40;
            a.migrate = 1;
            MPROS_CUDA(cudaMemsetAsync(a.mig_ws, 0xfe, S.mig_cap * 4, stream)); // kNotReady
        }
    }
    a.solo_spare0 = std::min(warps, PL.n);
    a.solo_spare_n = std::max(1, custody_max);
    const bool grow = squad && c->opt.plan_grow && PM.n > 0 && a.handoff_ws != nullptr && std::memcmp(PM.last_layout, &Lmid, sizeof(Lmid)) == 0;
    a.Lm = Lmid;
    a.mid_pool.base = PM.base;
    a.mid_pool.owner = PM.owner;
    a.mid_pool.gen = PM.gen;
    a.mid_pool.n = grow ? PM.n : 0;
    a.mid_pool.custody = nullptr;
    a.mid_pool.custody_max = 0;
    const int wpc = squad ? MPROS_SQUAD_WARPS : MPROS_SOLO_WARPS;
    const int ctas = squad ? std::max(1, std::min(std::min(warps, PL.n) / wpc, (nq + wpc - 1) / wpc))
                           : std::max(1, std::min((std::min(warps, PL.n) + wpc - 1) / wpc, (nq + wpc - 1) / wpc));
    const bool dbg = c->opt.debug;
    cudaEvent_t e0, e1;
    if (dbg)
    {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, stream);
    }
    if (squad)
        plan_squad_kernel<MPROS_SQUAD_WARPS, MPROS_SQUAD_MINB><<<ctas, 32 * MPROS_SQUAD_WARPS, 0, stream>>>(c->view(), c->pv, a);
    else
        plan_solo_kernel<MPROS_SOLO_WARPS, MPROS_SOLO_MINB><<<ctas, 32 * MPROS_SOLO_WARPS, 0, stream>>>(c->view(), c->pv, a);
    MPROS_LAUNCH_CHECK(c);
    if (dbg)
    {
        cudaEventRecord(e1, stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        unsigned int nh = 0, nt = 0, nm = 0;
        cudaMemcpy(&nh, S.counter + 32, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&nt, S.counter + 40, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&nm, S.counter + 48, 4, cudaMemcpyDeviceToHost);
#ifdef MPROS_PROFILE
        if (squad)
        {
            unsigned long long h[32];
            cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
            unsigned long long z[32] = {0};
            cudaMemcpyToSymbol(g_prof, z, sizeof(z));
            const double wr = (double)std::max<unsigned long long>(1, h[8]);
            const char *names[8] = {"own steps", "wait (1)", "round A", "wait (2)", "skip + wait (3)", "round B", "wait (4)", "inserts"};
            std::fprintf(stderr, "[mpros]   wait (w1) %9.1f, walk pass %9.1f, wait (w2) %9.1f\n", h[24] / wr, h[25] / wr, h[26] / wr);
            std::fprintf(stderr, "[mpros] squad profile: %.0f warp-rounds; per warp-round: expansions %.3f, requests %.3f, round-B requests %.3f, round B present %.3f, waiting for the goal map %.3f, idle/done %.3f, init %.3f, shots %.4f; cycles per warp-round:\n",
                         wr, h[12] / wr, h[10] / wr, h[11] / wr, h[9] / wr, h[13] / wr, h[14] / wr, h[16] / wr, h[15] / wr);
            for (int k = 0; k < 8; ++k)
                std::fprintf(stderr, "[mpros]   %-16s %9.1f\n", names[k], h[k] / wr);
            {
                unsigned long long hh[64];
                cudaMemcpyFromSymbol(hh, g_prof_hist, sizeof(hh));
                unsigned long long zz[64] = {0};
                cudaMemcpyToSymbol(g_prof_hist, zz, sizeof(zz));
                const char *cause[7] = {"plain expansion", "stale pops", ">= 2 probe walks", "wavefront steps", "shot", "query start", "no work"};
                std::fprintf(stderr, "[mpros]   own step by cause, share of warp-rounds with < 16 k / < 32 k / < 64 k / < 128 k / more cycles:\n");
                for (int k = 0; k < 7; ++k)
                    std::fprintf(stderr, "[mpros]     %-18s %7.4f %7.4f %7.4f %7.4f %7.4f\n", cause[k], hh[5 * k] / wr, hh[5 * k + 1] / wr, hh[5 * k + 2] / wr,
                                 hh[5 * k + 3] / wr, hh[5 * k + 4] / wr);
            }
            const char *sub[6] = {"pop (+ stale)", "node, primitives", "pre-classification", "probe walks", "closed set, g", "goal-map values"};
            const double ex = (double)std::max<unsigned long long>(1, h[12]);
            std::fprintf(stderr, "[mpros]   inside the own step, cycles per EXPANSION:\n");
            for (int k = 0; k < 6; ++k)
                std::fprintf(stderr, "[mpros]     %-18s %9.1f\n", sub[k], h[17 + k] / ex);
        }
#endif
        std::fprintf(stderr, "[mpros] %s pass: %d queries, node_cap %d, %d CTAs x %d warps (%d resident per SM), %.1f MB/warp, %d grown workspaces of %.1f MB, %.3f ms, %u handed to the team pass, %u migrated, %u of them adopted (%s policy)\n",
                     squad ? "squad" : "solo", nq, L.node_cap, ctas, wpc, squad ? c->squad_resident : c->solo_resident, L.stride / 1048576.0,
                     a.mid_pool.n, Lmid.stride / 1048576.0, ms, nh, nm, nt, a.migrate ? "throughput" : "latency");
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return MPROS_OK;
}
#endif

static int plan_pass(mpros_ctx *c, int lane, cudaStream_t stream, PlanBatchArgs &a, int pass, int node_cap, int n_run, const int32_t *d_todo,
                     int max_warps, const unsigned int *n_todo_dev = nullptr)
{
    PlanScratch &S = c->plan_scratch[lane][pass]; // ticket counter / re-run list of this lane
    PlanPool &PL = c->plan_pool[pass];            // workspaces: owned by the context, shared by its lanes
    WarpWorkspaceLayout L = make_layout(node_cap, c->H, c->W);
    // resident teams: a multiple of the SM count. Device queries are cached in the context: a batch submission should cost
    // the host microseconds (cudaGetDeviceProperties alone takes milliseconds).
    const int n_sm = c->n_sm;
    // team = one CTA (default) or a cluster of two CTAs on the two SMs of a TPC (MPROS_PLAN_MODE=cluster; plan_team.cuh).
    // Measured on the 16 384-query synthetic batch: CTA teams 1.57 s (296 teams, 6.8-7.5 us per iteration alone, ~11 us with
    // two teams per SM), CTA pairs 2.24 s (148 teams, 8.2-8.6 us per iteration alone AND loaded: no interference between
    // co-resident teams any more, but the two cluster barriers cost ~4 k cycles per iteration). Results are identical.
    const bool cluster = c->opt.plan_cluster;
    const bool duo = !cluster && c->opt.plan_duo;
    int &resident = c->plan_resident[cluster ? 1 : duo ? 2 : 0];
    if (!resident)
    {
        if (duo)
        {
            MPROS_CUDA(cudaFuncSetAttribute(plan_duo_kernel<kTeamWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)duo_smem_bytes<kTeamWarps>()));
            resident = 2; // one CTA of two teams per SM
        }
        else if (cluster)
            MPROS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, plan_cluster_kernel<kTeamWarps, kTeamMinBlocks>, kTeamWarps * 32, 0));
        else
            MPROS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, plan_team_kernel<kTeamWarps, kTeamMinBlocks>, kTeamWarps * 32, 0));
        resident = std::max(1, resident);
    }
    // The pool holds one workspace per team the DEVICE can keep resident (the register file decides: two CTAs per SM), however
    // many batches are in flight: the CTAs of a later batch's kernel take the workspaces - and the SM slots - an earlier one
    // frees. (Round 1 gave every lane its own set: 4 x 33 GB; now 33 GB serve any number of lanes.)
    int teams = (cluster ? n_sm / 2 : n_sm) * resident;
    if (duo)
        teams = (teams + 1) & ~1;
    const bool layout_change = std::memcmp(PL.last_layout, &L, sizeof(L)) != 0;
    // ... but never more than the batches seen so far could use: a context that only ever plans single queries (the node's
    // own use) keeps one ~110 MB workspace, not 33 GB. A caller that pipelines on the extra lanes gets room for two batches.
    const int wanted = (int)std::min<long long>(teams, std::max<long long>(layout_change ? 0 : PL.n, (long long)n_run * (lane == 0 ? 1 : 2)));
    if (layout_change || PL.n < wanted)
    {
        // a different layout re-interprets the same bytes (stale tags could alias), a bigger pool moves it: nothing may be
        // in flight on any lane, and the pool starts out cleared
        size_t free_b = 0, total_b = 0;
        MPROS_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t budget = (free_b + PL.bytes) / 2;
        int n = (int)std::min<long long>(wanted, std::max<long long>(1, (long long)(budget / L.stride)));
        if (duo)
            n = std::max(2, (n + 1) & ~1);
        if (layout_change || n > PL.n)
        {
            int rc = plan_quiesce(c);
            if (rc)
                return rc;
            const size_t need = (size_t)n * L.stride;
            if (need > PL.bytes)
            {
                if (PL.base)
                    MPROS_CUDA(cudaFree(PL.base));
                PL.base = nullptr;
                PL.bytes = 0;
                MPROS_CUDA(cudaMalloc(&PL.base, need));
                PL.bytes = need;
            }
            if (n > PL.meta_cap)
            {
                if (PL.owner)
                    MPROS_CUDA(cudaFree(PL.owner));
                if (PL.gen)
                    MPROS_CUDA(cudaFree(PL.gen));
                PL.owner = nullptr, PL.gen = nullptr, PL.meta_cap = 0;
                MPROS_CUDA(cudaMalloc(&PL.owner, (size_t)n * 4));
                MPROS_CUDA(cudaMalloc(&PL.gen, (size_t)n * 4));
                PL.meta_cap = n;
            }
            MPROS_CUDA(cudaMemsetAsync(PL.base, 0, PL.bytes, stream));
            MPROS_CUDA(cudaMemsetAsync(PL.owner, 0, (size_t)PL.meta_cap * 4, stream));
            MPROS_CUDA(cudaMemsetAsync(PL.gen, 0, (size_t)PL.meta_cap * 4, stream));
            MPROS_CUDA(cudaStreamSynchronize(stream)); // the other lanes' streams are not ordered behind this one
            PL.n = n;
            PL.stride = L.stride;
            PL.issued = 0;
            std::memcpy(PL.last_layout, &L, sizeof(L));
        }
    }
    // generation tags must stay below 2^30 and unique per (workspace, query); a workspace sees at most every query issued
    if (PL.issued + (unsigned long long)n_run + 2ull >= (1ull << 30))
    {
        int rc = plan_quiesce(c);
        if (rc)
            return rc;
        MPROS_CUDA(cudaMemsetAsync(PL.base, 0, PL.bytes, stream));
        MPROS_CUDA(cudaMemsetAsync(PL.gen, 0, (size_t)PL.meta_cap * 4, stream));
        MPROS_CUDA(cudaStreamSynchronize(stream));
        PL.issued = 0;
    }
    PL.issued += (unsigned long long)n_run + 1ull;
    // one team per query at a time; never more CTAs than workspaces (a CTA without one would only wait for one)
    int blocks = std::min(std::min(teams, PL.n), n_run);
    if (max_warps > 0)
        blocks = std::min(blocks, std::max(1, max_warps));
    if (duo)
        blocks = (blocks + 1) & ~1; // a CTA always owns two workspaces
    if (!S.counter)
        MPROS_CUDA(cudaMalloc(&S.counter, 256));
    if (!a.shared_counter)
    {
        MPROS_CUDA(cudaMemsetAsync(S.counter, 0, 4, stream));
        a.counter = S.counter;
    }
    a.L = L;
    a.pool.base = PL.base;
    a.pool.owner = PL.owner;
    a.pool.gen = PL.gen;
    a.pool.n = PL.n;
    a.todo = d_todo;
    a.n_todo = n_run;
    a.n_todo_dev = n_todo_dev;
    const bool dbg = c->opt.debug;
    cudaEvent_t e0, e1;
    if (dbg)
    {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, stream);
    }
    if (duo)
        plan_duo_kernel<kTeamWarps><<<(blocks + 1) / 2, kTeamWarps * 64, duo_smem_bytes<kTeamWarps>(), stream>>>(c->view(), c->pv, a);
    else if (cluster)
        plan_cluster_kernel<kTeamWarps, kTeamMinBlocks><<<2 * blocks, kTeamWarps * 32, 0, stream>>>(c->view(), c->pv, a);
    else
        plan_team_kernel<kTeamWarps, kTeamMinBlocks><<<blocks, kTeamWarps * 32, 0, stream>>>(c->view(), c->pv, a);
    MPROS_LAUNCH_CHECK(c);
    if (dbg)
    {
        cudaEventRecord(e1, stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
#ifdef MPROS_PROFILE
        {
            unsigned long long h[24];
            cudaMemcpyFromSymbol(h, g_prof, sizeof(h));
            unsigned long long z[24] = {0};
            cudaMemcpyToSymbol(g_prof, z, sizeof(z));
            const char *names[12] = {"wait_pop", "shot", "expand", "h1_peek", "bfs_steps", "h2_roundA", "h2_roundB", "h2_tail", "insert", "wait_heap",
                                     "heap_pop", "heap_restore"};
            double its = (double)h[12];
            std::fprintf(stderr, "[mpros] profile: %.0f iterations, %.3f (bfs steps + stale roots dropped)/it, %.3f (roundB + sequential heap pushes)/it, %.4f shots/it; cycles per iteration:\n", its,
                         h[13] / its, h[14] / its, h[15] / its);
            for (int k = 0; k < 12; ++k)
                std::fprintf(stderr, "[mpros]   %-12s %9.1f\n", names[k], h[k] / its);
            const char *names2[8] = {"col:pose+sincos", "col:footprint", "closed:A loads", "closed:B-D", "closed:insert", "heap:remove root", "heap:pushes", "duo:wait for partner"};
            for (int k = 16; k < 24; ++k)
                std::fprintf(stderr, "[mpros]   %-18s %9.1f\n", names2[k - 16], h[k] / its);
        }
#endif
        std::fprintf(stderr, "[mpros] plan pass %d: %d queries, node_cap %d, %d teams (%s) x %d warps, %.1f MB/team, %.3f ms\n", pass, n_run,
                     node_cap, blocks, cluster ? "CTA pairs" : duo ? "two per CTA" : "CTAs", kTeamWarps, L.stride / 1048576.0, ms);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return MPROS_OK;
}
} // namespace mpros

static int plan_batch_dev_impl(mpros_ctx *c, int lane, cudaStream_t stream, const double *d_starts, const double *d_goals, int nq,
                               int32_t *d_status, double *d_cost, int32_t *d_path_len, double *d_paths, int path_cap, int64_t *d_stats,
                               double *d_trace = nullptr, long long trace_cap = 0, long long *d_trace_n = nullptr)
{
    PlanBatchArgs a;
    std::memset(&a, 0, sizeof(a));
    a.trace = d_trace;
    a.trace_cap = trace_cap;
    a.trace_n = d_trace_n;
    a.starts = d_starts;
    a.goals = d_goals;
    a.nq = nq;
    a.status = d_status;
    a.cost = d_cost;
    a.path_len = d_path_len;
    a.paths = d_paths;
    a.path_cap = path_cap;
    a.stats = d_stats;
    a.rs = rs_config_from_planner(c);
    a.optimal = c->pp.optimal_path;
    PlanScratch &S = c->plan_scratch[lane][0];
    // pass 1: every query with a small node pool; later passes: only the queries that overflowed, bigger pools
    const long long max_nodes = (long long)(c->pp.max_iteration + 2) * (2 * c->pv.n_steer) + 16;
    // node pools: when every resident team can own a pool for the iteration limit (B200: 296 teams x ~100 MB) each query
    // runs exactly once; otherwise small pools first and the queries that overflowed are re-run with bigger ones
    long long caps[3] = {16384, 131072, max_nodes};
    {
        WarpWorkspaceLayout big = make_layout((int)max_nodes, c->H, c->W);
        const size_t n_ws = std::min<size_t>((size_t)kTeamMinBlocks * c->n_sm, (size_t)nq * (lane == 0 ? 1 : 2)); // what plan_pass will ask for
        const size_t full = big.stride * n_ws;
        if (c->plan_pool[0].bytes >= full && std::memcmp(c->plan_pool[0].last_layout, &big, sizeof(big)) == 0)
            caps[0] = max_nodes; // the context already owns full-size workspaces for this batch
        else
        {
            size_t free_b = 0, total_b = 0;
            MPROS_CUDA(cudaMemGetInfo(&free_b, &total_b));
            size_t have = free_b;
            for (int k = 0; k < kPlanPasses; ++k)
                have += c->plan_pool[k].bytes;
            if (full < have / 2) // half of the memory stays free for the caller
                caps[0] = max_nodes;
        }
    }
#ifndef MPROS_SOLO_OFF
    // Default for big batches: every query starts on its own warp with a small workspace (solo pass); the team pass behind it runs only the
    // queries that outgrew it, with full-size workspaces, and finds them in a device-side list (no host round trip, so the
    // asynchronous entry points stay asynchronous). The expansion trace is a team-kernel feature.
    if (c->opt.plan_solo && !c->opt.plan_cluster && !c->opt.plan_duo && !d_trace && caps[0] >= max_nodes && c->opt.solo_nodes < max_nodes &&
        nq >= c->opt.solo_min_queries)
    {
        int rc = plan_solo_pass(c, lane, stream, a, nq, max_nodes);
        if (rc)
            return rc;
        return plan_pass(c, lane, stream, a, 0, (int)max_nodes, nq, a.handoff, 0, a.handoff_n);
    }
#endif
    int rc = plan_pass(c, lane, stream, a, 0, (int)std::min<long long>(caps[0], max_nodes), nq, nullptr, 0);
    if (rc)
        return rc;
    if (caps[0] >= max_nodes)
        return MPROS_OK;
    std::vector<int32_t> h_status(nq);
    for (int pass = 1; pass < 3; ++pass)
    {
        MPROS_CUDA(cudaMemcpyAsync(h_status.data(), d_status, (size_t)nq * 4, cudaMemcpyDeviceToHost, stream));
        MPROS_CUDA(cudaStreamSynchronize(stream));
        std::vector<int32_t> todo;
        for (int q = 0; q < nq; ++q)
            if (h_status[q] == ST_OVERFLOW)
                todo.push_back(q);
        if (todo.empty())
            break;
        if (todo.size() > S.todo_cap)
        {
            if (S.todo)
                MPROS_CUDA(cudaFree(S.todo));
            S.todo = nullptr;
            MPROS_CUDA(cudaMalloc(&S.todo, todo.size() * 4 + 1024));
            S.todo_cap = todo.size() + 256;
        }
        MPROS_CUDA(cudaMemcpyAsync(S.todo, todo.data(), todo.size() * 4, cudaMemcpyHostToDevice, stream));
        // `todo` is pageable: the copy above has been staged before the call returns, the vector may go away
        long long cap = std::min(caps[pass], max_nodes);
        rc = plan_pass(c, lane, stream, a, pass, (int)cap, (int)todo.size(), S.todo, 0);
        if (rc)
            return rc;
        if (cap >= max_nodes)
            break;
    }
    return MPROS_OK;
}

// One launch of the planner whose ticket counter and result arrays are given by the caller and may live in a peer GPU's
// memory (dist.cu: mpros_plan_batch_shared). Always with full-size node pools: a query that overflowed a small pool could
// not be re-run by the rank that happened to draw it without another exchange.
namespace mpros
{
int plan_batch_shared_launch(mpros_ctx *c, const double *d_starts, const double *d_goals, int nq, unsigned int *counter, int32_t *status,
                             double *cost, int32_t *path_len, double *paths, int path_cap, int64_t *stats)
{
    PlanBatchArgs a;
    std::memset(&a, 0, sizeof(a));
    a.starts = d_starts;
    a.goals = d_goals;
    a.nq = nq;
    a.status = status;
    a.cost = cost;
    a.path_len = path_len;
    a.paths = paths;
    a.path_cap = path_cap;
    a.stats = stats;
    a.rs = rs_config_from_planner(c);
    a.optimal = c->pp.optimal_path;
    a.counter = counter;
    a.shared_counter = 1;
    const long long max_nodes = (long long)(c->pp.max_iteration + 2) * (2 * c->pv.n_steer) + 16;
    // pass 0 with pools for the iteration limit (as many as fit into half of the free memory): every query runs exactly once
    return plan_pass(c, 0, c->stream, a, 0, (int)max_nodes, nq, nullptr, 0);
}
} // namespace mpros

namespace mpros
{
// HybridAstar::getSearchTree (hybrid_a_star.cpp:1114-1164): every tree_ entry (expanded node x primitive) becomes
// leaf_num short segments {next_x, next_y, curr_x, curr_y}. One thread per entry; the per-primitive leaf displacement is a
// host-computed constant (PlannerView::leaf_*), the rotation into the map frame runs here.
__global__ void __launch_bounds__(256) search_tree_kernel(const __grid_constant__ PlannerView p, const double *__restrict__ trace,
                                                          long long n_exp, double *__restrict__ tree4)
{
    const int P = 2 * p.n_steer;
    const long long total = n_exp * P;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
    {
        const long long k = e / P;
        const int c = (int)(e % P);
        double x = trace[3 * k], y = trace[3 * k + 1], yaw = trace[3 * k + 2];
        const double dx = p.leaf_dx[c], dy = p.leaf_dy[c], dyaw = p.leaf_dyaw[c];
        const bool straight = p.steer_set[c >> 1] == 0;
        double *o = tree4 + (size_t)e * p.leaf_num * 4;
        for (int i = 0; i < p.leaf_num; ++i)
        {
            double sn, cs;
            sincos(yaw, &sn, &cs);
            const double nx = dx * cs - dy * sn + x, ny = dx * sn + dy * cs + y;
            o[4 * i] = nx, o[4 * i + 1] = ny, o[4 * i + 2] = x, o[4 * i + 3] = y;
            x = nx, y = ny;
            if (!straight)
                yaw = regular_yaw(yaw + dyaw);
        }
    }
}
} // namespace mpros

extern "C" int mpros_plan_search_tree(mpros_ctx *c, const double *start3, const double *goal3, int32_t *status, double *cost, double *tree4,
                                      size_t cap_segments, size_t *n_segments)
{
    int rc = check_planner_ready(c, "mpros_plan_search_tree");
    if (rc)
        return rc;
    if (!start3 || !goal3 || !status || !cost || !n_segments)
    {
        set_error("mpros_plan_search_tree: NULL argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int P = 2 * c->pv.n_steer, leaf = c->pv.leaf_num;
    const long long max_exp = (long long)c->pp.max_iteration + 2;
    // device scratch: poses, results, trace
    double *d_sg = nullptr, *d_cost = nullptr, *d_trace = nullptr;
    int32_t *d_status = nullptr, *d_len = nullptr;
    long long *d_n = nullptr;
    const size_t trace_bytes = (size_t)max_exp * 24;
    MPROS_CUDA(cudaMalloc(&d_sg, 48 + 8 + 8 + 8 + 8));
    MPROS_CUDA(cudaMalloc(&d_trace, trace_bytes));
    d_cost = d_sg + 6;
    d_status = reinterpret_cast<int32_t *>(d_sg + 7);
    d_len = d_status + 1;
    d_n = reinterpret_cast<long long *>(d_sg + 8);
    auto cleanup = [&]() {
        cudaFree(d_sg);
        cudaFree(d_trace);
    };
    double sg[6] = {start3[0], start3[1], start3[2], goal3[0], goal3[1], goal3[2]};
    cudaMemcpyAsync(d_sg, sg, sizeof(sg), cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(d_n, 0, 8, st);
    rc = plan_batch_dev_impl(c, 0, st, d_sg, d_sg + 3, 1, d_status, d_cost, d_len, nullptr, 0, nullptr, d_trace, max_exp, d_n);
    if (rc)
    {
        cleanup();
        return rc;
    }
    long long n_exp = 0;
    cudaMemcpyAsync(&n_exp, d_n, 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(status, d_status, 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(cost, d_cost, 8, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess)
    {
        cleanup();
        return cuda_fail(cudaGetLastError(), "mpros_plan_search_tree", __FILE__, __LINE__);
    }
    if (n_exp > max_exp)
        n_exp = max_exp;
    const size_t n_seg = (size_t)n_exp * P * leaf;
    *n_segments = n_seg;
    if (!tree4 || n_seg > cap_segments)
    {
        cleanup();
        if (!tree4 && cap_segments == 0)
            return MPROS_OK; // size query
        set_error("mpros_plan_search_tree: output buffer too small (*n_segments holds the required size)");
        return MPROS_ERR_CAPACITY;
    }
    if (n_seg)
    {
        double *d_tree = nullptr;
        if (cudaMalloc(&d_tree, n_seg * 32) != cudaSuccess)
        {
            cleanup();
            return cuda_fail(cudaGetLastError(), "cudaMalloc(search tree)", __FILE__, __LINE__);
        }
        size_t blocks = ((size_t)n_exp * P + 255) / 256;
        if (blocks > (size_t)c->n_sm * 8)
            blocks = (size_t)c->n_sm * 8;
        search_tree_kernel<<<(unsigned)blocks, 256, 0, st>>>(c->pv, d_trace, n_exp, d_tree);
        c->launches++;
        cudaMemcpyAsync(tree4, d_tree, n_seg * 32, cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        cudaFree(d_tree);
        if (e != cudaSuccess)
        {
            cleanup();
            return cuda_fail(e, "search_tree_kernel", __FILE__, __LINE__);
        }
    }
    cleanup();
    return MPROS_OK;
}

namespace mpros
{
// Lanes: up to MPROS_PLAN_LANES query batches in flight per context. Lane 0 is the context's own stream; the others get a
// stream, a staging area and planner workspaces of their own on first use, so the few long queries at the end of one batch
// (the iteration-limit tail: a dependent chain of up to max_iteration expansions) overlap the next batch instead of
// leaving the SMs idle. The map layers are shared and read-only while batches are in flight.
static int lane_stream(mpros_ctx *c, int lane, cudaStream_t *out)
{
    if (lane < 0 || lane >= MPROS_PLAN_LANES)
    {
        set_error("plan lane out of range (0 .. MPROS_PLAN_LANES-1)");
        return MPROS_ERR_INVALID;
    }
    if (lane == 0)
    {
        *out = c->stream;
        return MPROS_OK;
    }
    if (!c->lane_stream[lane])
    {
        MPROS_CUDA(cudaStreamCreateWithFlags(&c->lane_stream[lane], cudaStreamNonBlocking));
        MPROS_CUDA(cudaEventCreateWithFlags(&c->lane_event[lane], cudaEventDisableTiming));
    }
    // everything issued on the context's stream so far (map layers, planner configuration) precedes the batch
    MPROS_CUDA(cudaEventRecord(c->lane_event[lane], c->stream));
    MPROS_CUDA(cudaStreamWaitEvent(c->lane_stream[lane], c->lane_event[lane], 0));
    *out = c->lane_stream[lane];
    return MPROS_OK;
}

static int check_plan_args(mpros_ctx *c, const char *who, const void *s, const void *g, size_t nq, const void *status, const void *cost,
                           const void *len, const void *paths, int path_cap)
{
    if (!s || !g || !status || !cost || !len || path_cap < 0 || (path_cap > 0 && !paths) || nq > (size_t)(1 << 28))
    {
        set_error(std::string(who) + ": NULL / invalid argument");
        return MPROS_ERR_INVALID;
    }
    if ((unsigned long long)c->H * c->W * c->pv.num_yaw * 2ull >= (1ull << 34))
    {
        set_error("mpros_plan_batch: H*W*yaw*2 exceeds the closed-set key range (2^34)");
        return MPROS_ERR_INVALID;
    }
    return MPROS_OK;
}
} // namespace mpros

extern "C" void *mpros_plan_lane_stream(mpros_ctx *c, int lane)
{
    if (!c || lane < 0 || lane >= MPROS_PLAN_LANES)
        return nullptr;
    return lane == 0 ? (void *)c->stream : (void *)c->lane_stream[lane];
}

extern "C" int mpros_plan_batch_wait(mpros_ctx *c, int lane)
{
    if (!c || lane >= MPROS_PLAN_LANES)
    {
        set_error("mpros_plan_batch_wait: invalid argument");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    for (int l = 0; l < MPROS_PLAN_LANES; ++l)
    {
        if (lane >= 0 && l != lane)
            continue;
        cudaStream_t st = l == 0 ? c->stream : c->lane_stream[l];
        if (l == 0 || st)
            MPROS_CUDA(cudaStreamSynchronize(st));
    }
    return MPROS_OK;
}

extern "C" int mpros_plan_batch_submit_dev(mpros_ctx *c, int lane, const double *d_starts3, const double *d_goals3, size_t nq,
                                           int32_t *d_status, double *d_cost, int32_t *d_path_len, double *d_paths, int path_cap,
                                           int64_t *d_stats)
{
    int rc = check_planner_ready(c, "mpros_plan_batch_submit_dev");
    if (rc)
        return rc;
    if (nq == 0)
        return MPROS_OK;
    rc = check_plan_args(c, "mpros_plan_batch_submit_dev", d_starts3, d_goals3, nq, d_status, d_cost, d_path_len, d_paths, path_cap);
    if (rc)
        return rc;
    MPROS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st;
    rc = lane_stream(c, lane, &st);
    if (rc)
        return rc;
    return plan_batch_dev_impl(c, lane, st, d_starts3, d_goals3, (int)nq, d_status, d_cost, d_path_len, d_paths, path_cap, d_stats);
}

extern "C" int mpros_plan_batch_dev(mpros_ctx *c, const double *d_starts3, const double *d_goals3, size_t nq, int32_t *d_status,
                                    double *d_cost, int32_t *d_path_len, double *d_paths, int path_cap, int64_t *d_stats)
{
    return mpros_plan_batch_submit_dev(c, 0, d_starts3, d_goals3, nq, d_status, d_cost, d_path_len, d_paths, path_cap, d_stats);
}

extern "C" int mpros_plan_batch_submit(mpros_ctx *c, int lane, const double *starts3, const double *goals3, size_t nq, int32_t *status,
                                       double *cost, int32_t *path_len, double *paths, int path_cap, int64_t *stats)
{
    int rc = check_planner_ready(c, "mpros_plan_batch_submit");
    if (rc)
        return rc;
    if (nq == 0)
        return MPROS_OK;
    rc = check_plan_args(c, "mpros_plan_batch_submit", starts3, goals3, nq, status, cost, path_len, paths, path_cap);
    if (rc)
        return rc;
    MPROS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st;
    rc = lane_stream(c, lane, &st);
    if (rc)
        return rc;
    size_t pose_b = nq * 24, path_b = nq * (size_t)path_cap * 5 * 8;
    size_t total = 2 * align_up(pose_b, 256) + align_up(nq * 4, 256) * 2 + align_up(nq * 8, 256) + align_up(path_b, 256) +
                   align_up(nq * 48, 256) + 1024;
    if (lane == 0)
        rc = ensure_stage(c, total);
    else if (total > c->lane_stage_bytes[lane])
    {
        // the lane's previous batch may still read its staging area
        MPROS_CUDA(cudaStreamSynchronize(st));
        if (c->lane_stage[lane])
            MPROS_CUDA(cudaFree(c->lane_stage[lane]));
        c->lane_stage[lane] = nullptr;
        c->lane_stage_bytes[lane] = 0;
        MPROS_CUDA(cudaMalloc(&c->lane_stage[lane], total));
        c->lane_stage_bytes[lane] = total;
    }
    if (rc)
        return rc;
    char *base = static_cast<char *>(lane == 0 ? c->stage : c->lane_stage[lane]);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        char *p = base + o;
        o += align_up(bytes, 256);
        return p;
    };
    double *d_s = (double *)take(pose_b), *d_g = (double *)take(pose_b);
    int32_t *d_status = (int32_t *)take(nq * 4), *d_len = (int32_t *)take(nq * 4);
    double *d_cost = (double *)take(nq * 8);
    double *d_paths = path_cap ? (double *)take(path_b) : nullptr;
    int64_t *d_stats = (int64_t *)take(nq * 48);
    MPROS_CUDA(cudaMemcpyAsync(d_s, starts3, pose_b, cudaMemcpyHostToDevice, st));
    MPROS_CUDA(cudaMemcpyAsync(d_g, goals3, pose_b, cudaMemcpyHostToDevice, st));
    rc = plan_batch_dev_impl(c, lane, st, d_s, d_g, (int)nq, d_status, d_cost, d_len, d_paths, path_cap, d_stats);
    if (rc)
        return rc;
    MPROS_CUDA(cudaMemcpyAsync(status, d_status, nq * 4, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaMemcpyAsync(cost, d_cost, nq * 8, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaMemcpyAsync(path_len, d_len, nq * 4, cudaMemcpyDeviceToHost, st));
    if (path_cap)
        MPROS_CUDA(cudaMemcpyAsync(paths, d_paths, path_b, cudaMemcpyDeviceToHost, st));
    if (stats)
        MPROS_CUDA(cudaMemcpyAsync(stats, d_stats, nq * 48, cudaMemcpyDeviceToHost, st));
    return MPROS_OK;
}

extern "C" int mpros_plan_batch(mpros_ctx *c, const double *starts3, const double *goals3, size_t nq, int32_t *status, double *cost,
                                int32_t *path_len, double *paths, int path_cap, int64_t *stats)
{
    int rc = mpros_plan_batch_submit(c, 0, starts3, goals3, nq, status, cost, path_len, paths, path_cap, stats);
    if (rc || nq == 0)
        return rc;
    return mpros_plan_batch_wait(c, 0);
}
