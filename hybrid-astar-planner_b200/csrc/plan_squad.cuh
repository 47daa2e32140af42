// Whole-query planner, a SQUAD of W queries per CTA: one warp per query for the search itself, the footprint walks and the
// heuristics of all W queries evaluated together by the whole CTA. First launch of a batch (plan.cu: plan_solo_pass); the
// team kernel (plan_team.cuh) behind it finishes what the squads hand over.
//
// Why (ncu of the plain warp-per-query kernel, plan_solo.cuh, profiles/r02_plan_solo_kernel_full.csv): 5 400 warp-instructions
// per iteration, 60 % of them in the OMPL-equivalent Reeds-Shepp distance of the children that need a heuristic - evaluated
// with 4-12 of 32 lanes busy (lane = child x symmetry image), eleven formula bodies one after the other - and the 24
// resident warps of an SM spread over ~50 KB of hot code (instruction-cache hit rate 64 %, stall_no_instruction 13 per
// issue, issue slots 22 % busy). Here every round of the CTA has the same phases:
//   1a. each warp advances its own query by one step: pop, Reeds-Shepp shot or the end poses of the expansion, pre-classified
//       by their own cell (map_device.cuh: footprint_preclass);
//   1w. the end poses left undecided by ALL warps are walked one per warp (squad_walks);
//   1b. closed-set probes, g, goal-map values; the children that need a heuristic are POSTED into shared memory;
//   2.  round A, whole CTA: the two CSC words of every posted child, work item = (child, word) on a quad of lanes (lane =
//       symmetry image), items ordered by word so that the 32 lanes of a warp run the SAME closed form on eight children;
//   3.  each warp applies the exact skip rule to its own children (rho * min(CSC) <= h1 for all of them: done) and posts the
//       others for round B;
//   4.  round B, whole CTA: the other nine words, same layout;
//   5.  each warp combines the family minima of its children and inserts them (closed set, node pool, open list).
// All warps of an SM execute the same phase at the same time, so the instruction cache holds one phase instead of all
// (hit rate 91 %, stall_no_instruction 0.46), and the formula code runs with full warps (3.6 warp-passes per expansion
// instead of 11). Work per warp and round is bounded so that one query cannot hold up the other 23: kSquadPops open-list
// removals, kSquadBfsSteps wavefront levels / kSquadBfsWords scanned words (a query still waiting for a goal-map value keeps
// its children in registers and posts nothing).
//
// A query leaves the squad pass for the team pass WITH ITS STATE (squad_hand_over, plan_team.cuh: team_resume) when the
// tickets have run out and its CTA has thinned out, or when no workspace has room for it; a query that outgrows the warp's
// small workspace moves into a grown one first (squad_grow). With other batches in flight (PlanBatchArgs::migrate) the
// queries of a thinned-out squad are adopted by idle warps of squads that are still well filled instead. DESIGN.md 4.
//
// Semantics and arithmetic are those of plan_solo.cuh / plan_team.cuh: same device functions on the same operands in the
// same order, results bit-identical (tests/test_gpu_synthetic_fullsize.py::test_solo_and_team_passes_give_identical_plans).
#pragma once
#include "plan_solo.cuh"

namespace mpros
{

constexpr int kSquadMaxReq = 8;    // children of one expansion that can ask for a heuristic (2 * n_steer <= 8; else plan_solo_kernel)
constexpr int kSquadBfsSteps = 4;  // goal-wavefront levels one warp advances per round at most ...
constexpr int kSquadBfsWords = 256; // ... and no further level once that many words have been scanned in the round
constexpr int kSquadPops = 2;      // open-list entries one warp removes per round at most (CLOSE entries cost a removal each)
constexpr int kHandOff = 0x7ffffffe; // `finish` value: the query goes to the team pass together with its state

struct SquadReq
{
    double inx, iny, dth, nxb, nyb, sphi, cphi; // normalised goal seen from the child (plan_team.cuh, closed-set warp phase C)
    double h1v;                                  // goal-map value * resolution
    unsigned long long fam[3];                   // minima over images and words per family (plain, CCSC, CCSCC), as ordered bit patterns
    unsigned long long pad;
};

template <int W>
struct SquadShared
{
    SoloShared q[W];
    SquadReq req[W * kSquadMaxReq]; // slots w * kSquadMaxReq .. belong to warp w
    short list_a[W * kSquadMaxReq]; // posted slots of this round
    short list_b[W * kSquadMaxReq]; // slots that need round B
    // end poses whose footprint test needs the probe walk (not decided by their own cell), and the verdicts
    double walk[W * kSquadMaxReq][4]; // {x, y, sin yaw, cos yaw}
    unsigned char walk_hit[W * kSquadMaxReq];
    // counters of round r live at index r % 3: thread 0 clears the set of round r + 1 before barrier (1) of round r, when the
    // last readers (round r - 2) are long gone and the next writers (round r + 1) have not started
    int n_a[3], n_b[3], n_done[3], n_w[3];
};

// Phase profile of the squad kernel (developer build, -DMPROS_PROFILE): lane 0 of every warp sums cycles per phase in
// registers and adds them to g_prof when its CTA ends. Slots: 0 own steps (1a + 1b), 1 wait (1), 2 round A, 3 wait (2),
// 4 skip rule + wait (3), 5 round B, 6 wait (4), 7 inserts / results, 24 wait (w1), 25 walk pass, 26 wait (w2); counters:
// 8 warp-rounds, 9 warp-rounds with a round B, 10 requests, 11 round-B requests, 12 expansions, 13 warp-rounds waiting for a
// goal-map value, 14 idle / done warp-rounds, 15 shots, 16 warp-rounds in INIT; 17-22 cycles inside the own step (pop, node
// and primitives, pre-classification, -, closed set and g, goal-map values). A warp blocked at a barrier shows up in the
// slot AFTER the barrier for the most part (the clock read issues late), so read "wait" and the phase behind it together.
#ifdef MPROS_PROFILE
__device__ unsigned long long g_prof_hist[64]; // own-step duration of a warp-round: 8 causes x 5 duration classes (see SquadProf::hist)
#endif
struct SquadProf
{
#ifdef MPROS_PROFILE
    // what made this warp-round's own step long? cause: 0 plain expansion, 1 stale pops, 2 >= 2 probe walks, 3 wavefront steps,
    // 4 shot, 5 query start (ticket / init), 6 no work; duration classes < 16 k, < 32 k, < 64 k, < 128 k, >= 128 k cycles
    __device__ __forceinline__ void hist(int cause)
    {
        const long long d = Prof::now() - t0;
        const int b = d < 16384 ? 0 : d < 32768 ? 1 : d < 65536 ? 2 : d < 131072 ? 3 : 4;
        if ((threadIdx.x & 31) == 0)
            atomicAdd(&g_prof_hist[cause * 5 + b], 1ull);
    }
    unsigned long long acc[32];
    long long t0;
    __device__ __forceinline__ void init()
    {
        for (int k = 0; k < 32; ++k)
            acc[k] = 0;
        t0 = Prof::now();
    }
    __device__ __forceinline__ void mark(int k)
    {
        const long long t1 = Prof::now();
        acc[k] += (unsigned long long)(t1 - t0);
        t0 = t1;
    }
    __device__ __forceinline__ void add(int k, unsigned long long v) { acc[k] += v; }
    // sub-phase of phase 1: the time since the last mark / sub goes to slot k AND stays inside slot 0 (t0 is not moved)
    long long ts;
    __device__ __forceinline__ void sub0() { ts = Prof::now(); }
    __device__ __forceinline__ void sub(int k)
    {
        const long long t1 = Prof::now();
        acc[k] += (unsigned long long)(t1 - ts);
        ts = t1;
    }
    __device__ __forceinline__ void flush()
    {
        if ((threadIdx.x & 31) == 0)
            for (int k = 0; k < 32; ++k)
                atomicAdd(&g_prof[k], acc[k]);
    }
#else
    __device__ __forceinline__ void hist(int) {}
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ void mark(int) {}
    __device__ __forceinline__ void add(int, unsigned long long) {}
    __device__ __forceinline__ void sub0() {}
    __device__ __forceinline__ void sub(int) {}
    __device__ __forceinline__ void flush() {}
#endif
};

// one formula on eight children per warp: item i of `total` = (slot list[i % n], word k0 + i / n), lane = 4 * item + image
template <int W>
__device__ __forceinline__ void squad_formula_pass(SquadShared<W> &S, const short *list, int n, int k0, int nk)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = n * nk;
    for (int base = warp * 8; base < total; base += W * 8)
    {
        const int i = base + (lane >> 2), q = lane & 3;
        const bool live = i < total;
        double Lf = DBL_MAX;
        int slot = 0, k = k0;
        if (live)
        {
            k = k0 + i / n;
            slot = list[i - (k - k0) * n];
            const SquadReq &r = S.req[slot];
            const bool fx = (q == 1 || q == 3), fy = (q == 2 || q == 3), fp = (q == 1 || q == 2);
            const double inx = r.inx, iny = r.iny, dth = r.dth, nxb = r.nxb, nyb = r.nyb, sphi = r.sphi;
            Lf = ompl_formula(k, fx ? -inx : inx, fy ? -iny : iny, fp ? -dth : dth, fx ? -nxb : nxb, fy ? -nyb : nyb, fp ? -sphi : sphi, r.cphi);
        }
        Lf = fmin(Lf, __shfl_xor_sync(kFull, Lf, 1));
        Lf = fmin(Lf, __shfl_xor_sync(kFull, Lf, 2));
        if (live && q == 0 && Lf < DBL_MAX)
            atomicMin(&S.req[slot].fam[ompl_formula_family(k)], (unsigned long long)__double_as_longlong(Lf));
    }
}

// The collective part of a round. Lane c of a warp passes child c (needmask: the children that want a heuristic) with its
// pose, sin / cos of its yaw, sin / cos of (goal yaw - yaw) and goal-map value; returns h of child c on lane c. Every warp of
// the CTA calls it exactly once per round (needmask 0: nothing to ask). all_done: every warp reported `done`.
template <int W>
__device__ __forceinline__ double squad_heuristics(const MapView &m, const PlannerView &p, SquadShared<W> &S, unsigned needmask, double gx, double gy,
                                                   double gyaw, double x, double y, double yaw, double sn, double cs, double sphi, double cphi, int h1,
                                                   bool done, int rnd, int &n_done, SquadProf &pf)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cur3 = rnd % 3, nxt3 = (rnd + 1) % 3;
    if (threadIdx.x == 0)
        S.n_a[nxt3] = 0, S.n_b[nxt3] = 0, S.n_done[nxt3] = 0, S.n_w[nxt3] = 0;
    const bool mine = (needmask >> lane) & 1u;
    const int rank = __popc(needmask & ((1u << lane) - 1u)), cnt = __popc(needmask);
    const int slot = warp * kSquadMaxReq + rank;
    const double h1v = (double)h1 * m.res; // getGoalCost: cost_ * map_resolution_
    if (mine)
    {
        SquadReq &r = S.req[slot];
        const double dx = gx - x, dy = gy - y;
        r.inx = dm::div_(cs * dx + sn * dy, p.min_r);
        r.iny = dm::div_(-sn * dx + cs * dy, p.min_r);
        r.dth = gyaw - yaw;
        r.nxb = r.inx * cphi + r.iny * sphi;
        r.nyb = r.inx * sphi - r.iny * cphi;
        r.sphi = sphi, r.cphi = cphi, r.h1v = h1v;
        r.fam[0] = r.fam[1] = r.fam[2] = (unsigned long long)__double_as_longlong(DBL_MAX);
    }
    if (lane == 0)
    {
        if (cnt)
        {
            const int at = atomicAdd(&S.n_a[cur3], cnt);
            for (int j = 0; j < cnt; ++j)
                S.list_a[at + j] = (short)(warp * kSquadMaxReq + j);
        }
        if (done)
            atomicAdd(&S.n_done[cur3], 1);
    }
    pf.mark(0);
    __syncthreads(); // (1) every request of this round is posted
    pf.mark(1);
    const int n_a = S.n_a[cur3];
    n_done = S.n_done[cur3];
    if (n_a == 0)
        return 0.0;
    squad_formula_pass<W>(S, S.list_a, n_a, 0, 2); // round A: CSC
    pf.mark(2);
    __syncthreads(); // (2)
    pf.mark(3);
    pf.add(10, cnt);
    // exact skip (plan_team.cuh: heur_rounds): when rho * min(CSC) <= h1 for every asking child of this query the other nine
    // words cannot change max(h1, h2)
    bool ok = true;
    if (mine)
        ok = p.min_r * __longlong_as_double((long long)S.req[slot].fam[0]) <= h1v;
    const bool skip = __all_sync(kFull, ok);
    if (!skip && lane == 0)
    {
        const int at = atomicAdd(&S.n_b[cur3], cnt);
        for (int j = 0; j < cnt; ++j)
            S.list_b[at + j] = (short)(warp * kSquadMaxReq + j);
    }
    __syncthreads(); // (3)
    pf.mark(4);
    const int n_b = S.n_b[cur3];
    if (n_b)
    {
        squad_formula_pass<W>(S, S.list_b, n_b, 2, 9); // round B: CCC, CCCC, CCSC, CCSCC
        pf.mark(5);
        __syncthreads(); // (4)
        pf.mark(6);
        pf.add(9, 1);
        pf.add(11, skip ? 0 : cnt);
    }
    double h = 0.0;
    if (mine)
    {
        const SquadReq &r = S.req[slot];
        h = fmax(h1v, ompl_combine(p.min_r, __longlong_as_double((long long)r.fam[0]), __longlong_as_double((long long)r.fam[1]),
                                   __longlong_as_double((long long)r.fam[2]))) *
            (1 + 1e-3);
    }
    return h;
}

// The probe walks of a round, whole CTA: every warp posts the end poses its own pre-classification left undecided
// (walk_mask, lane = child), the poses are walked one per warp (warp_footprint_collision: one lane per probe) whoever
// posted them, and every warp gets the verdicts of its children back as a mask of FREE poses. A warp with six undecided
// children no longer holds up the other 23: the round's walks take the time of one or two. Every warp of the CTA calls this
// exactly once per round.
template <int W>
__device__ __forceinline__ unsigned squad_walks(const MapView &m, const PlannerView &p, SquadShared<W> &S, unsigned walk_mask, double x, double y,
                                                double sn, double cs, int rnd, SquadProf &pf)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cur3 = rnd % 3;
    const bool mine = (walk_mask >> lane) & 1u;
    const int rank = __popc(walk_mask & ((1u << lane) - 1u)), cnt = __popc(walk_mask);
    int base = 0;
    if (lane == 0 && cnt)
        base = atomicAdd(&S.n_w[cur3], cnt);
    base = __shfl_sync(kFull, base, 0);
    if (mine)
    {
        double *o = S.walk[base + rank];
        o[0] = x, o[1] = y, o[2] = sn, o[3] = cs;
    }
    pf.mark(0);
    __syncthreads(); // (w1) every pose of this round is posted
    pf.mark(24);
    const int n_w = S.n_w[cur3];
    if (n_w == 0)
        return 0u;
    for (int i = warp; i < n_w; i += W)
    {
        const double *r = S.walk[i];
        const bool hit = warp_footprint_collision(m, p, r[0], r[1], r[2], r[3]);
        if (lane == 0)
            S.walk_hit[i] = hit;
    }
    pf.mark(25);
    __syncthreads(); // (w2)
    pf.mark(26);
    return __ballot_sync(kFull, mine && !S.walk_hit[base + rank]);
}

// goal-map values with a bounded number of wavefront steps; final = every requested value is known
__device__ __forceinline__ int squad_h1_lookup(const MapView &m, TeamBfs &b, bool want, int i, int j, int &v, int max_steps, int &steps)
{
    if (want && v < 0)
        v = team_bfs_peek(m, b, i, j);
    steps = 0;
    int words = 0;
    while (__any_sync(kFull, want && v < 0) && steps < max_steps && words < kSquadBfsWords)
    {
        words += solo_bfs_step(m, b);
        ++steps;
        if (want && v < 0)
            v = team_bfs_peek(m, b, i, j);
    }
    return !__any_sync(kFull, want && v < 0);
}

constexpr int kNotReady = (int)0xfefefefe; // migration-list entry not written yet (the host fills the list with this byte pattern)
constexpr int kMidFlag = 0x40000000; // hand-off entry: the workspace index refers to the pool of grown workspaces (mid_pool)

// one free workspace of `pl`, or -1 (no waiting: one look at every entry, 32 at a time)
__device__ __forceinline__ int pool_try_acquire_warp(const WorkspacePool &pl, int first)
{
    const int lane = threadIdx.x & 31;
    if (pl.n <= 0)
        return -1;
    int base = first % pl.n;
    for (int seen = 0; seen < pl.n; seen += 32)
    {
        int i = base + lane;
        i = i >= pl.n ? i - pl.n : i;
        const bool free_seen = seen + lane < pl.n && *reinterpret_cast<volatile int *>(&pl.owner[i]) == 0;
        unsigned cand = __ballot_sync(kFull, free_seen);
        while (cand)
        {
            const int l = __ffs(cand) - 1;
            cand &= cand - 1;
            int mine = -1;
            if (lane == l && atomicCAS(&pl.owner[i], 0, 1) == 0)
                mine = i;
            const int got = __shfl_sync(kFull, mine, l);
            if (got >= 0)
            {
                __threadfence();
                return got;
            }
        }
        base = base + 32 >= pl.n ? base + 32 - pl.n : base + 32;
    }
    return -1;
}

// Hand a query over to the team pass with its state: the workspace the query lives in (node pool, open list, closed set,
// wavefront window) travels with it, the rest goes into the ResumeHeader. Called between two iterations (open list
// consistent, nothing half inserted). A query in a GROWN workspace (mid_slot >= 0) simply leaves it to the team pass; a query
// in the warp's own small workspace needs another one for the warp to go on with (custody limit, plan_team.cuh): false when
// there is none - nothing has changed then.
__device__ __noinline__ bool squad_hand_over(const PlanBatchArgs &a, SoloShared &S, const OpenHeap &heap, int q, int &ws_slot, uint32_t &gen,
                                             int mid_slot, uint32_t mid_gen, uint32_t n_nodes, int iteration, int n_shots, long long n_exp,
                                             long long n_coll, bool goal_found, uint32_t goal_parent, double goal_g, int si, int sj, bool adoptable)
{
    const int lane = threadIdx.x & 31;
    const bool mid = mid_slot >= 0;
    // adoptable: onto the migration list, where idle warps of well-filled squads look for work (throughput policy); the team
    // pass takes what nobody has adopted. When that list is nearly full the query goes to the team pass directly.
    bool to_mig = adoptable && a.migrate;
    if (to_mig)
    {
        unsigned n = 0;
        if (lane == 0)
            n = *reinterpret_cast<volatile unsigned *>(a.mig_n);
        if ((int)__shfl_sync(kFull, n, 0) >= a.mig_cap - 8192)
            to_mig = false;
    }
    int fresh = -1;
    if (!mid)
    {
        fresh = pool_swap_acquire_warp(a.solo_pool, a.solo_spare0 + (int)((unsigned)q % (unsigned)max(1, a.solo_spare_n)));
        if (fresh < 0)
            return false;
    }
    const WarpWorkspaceLayout &L = mid ? a.Lm : a.Ls;
    char *ws = mid ? a.mid_pool.base + (size_t)mid_slot * L.stride : a.solo_pool.base + (size_t)ws_slot * L.stride;
    // the top of the open list lives in shared memory: complete the global arrays
    for (int k = lane; k < heap.size && k < kSoloHeapTop; k += 32)
        heap.key[k] = S.hkey[k], heap.id[k] = S.hid[k];
    if (lane == 0)
    {
        ResumeHeader *H = reinterpret_cast<ResumeHeader *>(ws + L.hdr_off);
        H->tag = mid ? mid_gen : gen, H->gen = mid ? mid_gen : gen, H->heap_size = heap.size, H->n_nodes = n_nodes;
        H->iteration = iteration, H->n_shots = n_shots, H->goal_found = goal_found, H->goal_parent = goal_parent;
        H->n_exp = n_exp, H->n_coll = n_coll, H->goal_g = goal_g, H->si = si, H->sj = sj;
        H->bfs = S.bfs;
    }
    __syncwarp();
    uint32_t fresh_gen = 0;
    if (lane == 0)
    {
        __threadfence();
        const int wsv = mid ? (mid_slot | kMidFlag) : ws_slot;
        a.status[q] = ST_OVERFLOW; // overwritten by whoever finishes the query
        if (to_mig)
        {
            const unsigned at = atomicAdd(a.mig_n, 1u);
            a.mig[at] = q;
            __threadfence(); // the workspace entry is what an adopting warp waits for (kNotReady until now)
            *reinterpret_cast<volatile int32_t *>(&a.mig_ws[at]) = wsv;
        }
        else
        {
            const unsigned at = atomicAdd(a.handoff_n, 1u);
            a.handoff[at] = q;
            a.handoff_ws[at] = wsv;
        }
        if (!mid)
            fresh_gen = a.solo_pool.gen[fresh];
    }
    if (!mid)
    {
        ws_slot = fresh;
        gen = __shfl_sync(kFull, fresh_gen, 0);
    }
    return true;
}

// A query has outgrown the warp's small workspace: move it into a free workspace of the mid pool (four times the nodes) and go
// on there - the same warp, the same squad, no other kernel involved. The warp copies the node pool and the open list,
// re-hashes the live closed-set slots under the new workspace's tag and copies the initialised rows of the wavefront window
// (its CTA waits at the next barrier meanwhile: about a millisecond, a few times per CTA and batch). Returns the mid
// workspace, -1 when none is free.
__device__ __noinline__ int squad_grow(const PlanBatchArgs &a, SoloShared &S, const char *src, uint32_t old_tag, int heap_size, uint32_t n_nodes,
                                       int hint, uint32_t &mid_gen)
{
    const int lane = threadIdx.x & 31;
    const int ms = pool_try_acquire_warp(a.mid_pool, hint);
    if (ms < 0)
        return -1;
    const WarpWorkspaceLayout &Ls = a.Ls, &Lm = a.Lm;
    uint32_t g0 = 0;
    if (lane == 0)
        g0 = a.mid_pool.gen[ms];
    mid_gen = __shfl_sync(kFull, g0, 0) + 1;
    char *dst = a.mid_pool.base + (size_t)ms * Lm.stride;
    {
        const uint4 *sn = reinterpret_cast<const uint4 *>(src + Ls.node_off);
        uint4 *dn = reinterpret_cast<uint4 *>(dst + Lm.node_off);
        const size_t n16 = (size_t)n_nodes * (sizeof(PlanNode) / 16);
        for (size_t k = lane; k < n16; k += 128)
        {
            uint4 v[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (k + 32 * t < n16)
                    v[t] = sn[k + 32 * t];
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (k + 32 * t < n16)
                    dn[k + 32 * t] = v[t];
        }
    }
    {
        const unsigned long long *sk = reinterpret_cast<const unsigned long long *>(src + Ls.hkey_off);
        const uint32_t *sid = reinterpret_cast<const uint32_t *>(src + Ls.hid_off);
        unsigned long long *dk = reinterpret_cast<unsigned long long *>(dst + Lm.hkey_off);
        uint32_t *did = reinterpret_cast<uint32_t *>(dst + Lm.hid_off);
        for (int k = kSoloHeapTop + lane; k < heap_size; k += 32) // the first kSoloHeapTop entries stay in shared memory
            dk[k] = sk[k], did[k] = sid[k];
    }
    {
        const ClosedSlot *st = reinterpret_cast<const ClosedSlot *>(src + Ls.table_off);
        ClosedSlot *dt = reinterpret_cast<ClosedSlot *>(dst + Lm.table_off);
        const uint32_t dmask = (uint32_t)Lm.table_slots - 1;
        const unsigned long long new_tag = mid_gen;
        for (int k0 = lane; k0 < Ls.table_slots; k0 += 128)
        {
            ClosedSlot e[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (k0 + 32 * t < Ls.table_slots)
                    e[t] = st[k0 + 32 * t];
#pragma unroll
            for (int t = 0; t < 4; ++t)
            {
                if (k0 + 32 * t >= Ls.table_slots || (e[t].key >> 34) != (unsigned long long)old_tag)
                    continue;
                const unsigned long long idx = e[t].key & ((1ull << 34) - 1ull);
                const unsigned long long want = (new_tag << 34) | idx;
                uint32_t sl = hash_index(idx) & dmask;
                while (true)
                {
                    const unsigned long long cur = dt[sl].key;
                    if ((cur >> 34) != new_tag && atomicCAS(&dt[sl].key, cur, want) == cur)
                        break;
                    sl = (sl + 1) & dmask;
                }
                dt[sl].g = e[t].g;
                dt[sl].node = e[t].node;
            }
        }
    }
    {
        const TeamBfs &b = S.bfs;
        uint32_t *db = reinterpret_cast<uint32_t *>(dst + Lm.blocked_off), *d0 = reinterpret_cast<uint32_t *>(dst + Lm.f0_off),
                 *d1 = reinterpret_cast<uint32_t *>(dst + Lm.f1_off);
        uint4 *dd = reinterpret_cast<uint4 *>(dst + Lm.dist_off);
        if (b.ilo <= b.ihi)
        {
            const size_t w0 = (size_t)(b.ilo - b.r0) * b.wwords, nw = (size_t)(b.ihi - b.ilo + 1) * b.wwords;
            const uint4 *sd = reinterpret_cast<const uint4 *>(b.dist);
            for (size_t k = lane; k < nw; k += 32)
            {
                db[w0 + k] = b.blocked[w0 + k];
                d0[w0 + k] = b.f0[w0 + k];
                d1[w0 + k] = b.f1[w0 + k];
            }
            for (size_t k = lane; k < nw * 4; k += 32) // 32 x uint16 per word = 4 x 16 bytes
                dd[w0 * 4 + k] = sd[w0 * 4 + k];
        }
        __syncwarp();
        if (lane == 0)
        {
            S.bfs.blocked = db, S.bfs.f0 = d0, S.bfs.f1 = d1, S.bfs.dist = reinterpret_cast<uint16_t *>(dd);
            S.bfs.tag = mid_gen;
        }
    }
    __syncwarp();
    return ms;
}

#ifndef MPROS_SQUAD_WARPS
#define MPROS_SQUAD_WARPS 24 // queries per CTA
#endif
#ifndef MPROS_SQUAD_MINB
#define MPROS_SQUAD_MINB 1 // CTAs per SM
#endif
template <int W, int MINB>
__global__ void __launch_bounds__(32 * W, MINB) plan_squad_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p,
                                                                   const __grid_constant__ PlanBatchArgs a)
{
    __shared__ SquadShared<W> SQ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SoloShared &S = SQ.q[warp];
    if (threadIdx.x < 3)
        SQ.n_a[threadIdx.x] = 0, SQ.n_b[threadIdx.x] = 0, SQ.n_done[threadIdx.x] = 0, SQ.n_w[threadIdx.x] = 0;
    __syncthreads();
    const WarpWorkspaceLayout &L = a.Ls;
    const int P = 2 * p.n_steer;
    const unsigned total = (unsigned)a.nq;
    // workspace of this warp (taken with the first ticket)
    int ws_slot = -1;
    uint32_t gen = 0;
    char *ws = nullptr;
    PlanNode *nodes = nullptr;
    double *scratch = nullptr;
    OpenHeap heap;
    heap.skey = S.hkey, heap.sid = S.hid, heap.scap = kSoloHeapTop, heap.size = 0, heap.cap = L.node_cap;
    heap.key = nullptr, heap.id = nullptr;
    ClosedSet closed;
    closed.slot = nullptr, closed.mask = (uint32_t)L.table_slots - 1, closed.tag = 0;
    // query state
    enum
    {
        IDLE,    // no query: draw a ticket
        INIT,    // wavefront on its way to the start cell
        RUN,     // pop and expand
        WAIT_H1, // children computed, a goal-map value is not final yet
        DONE     // no tickets left
    };
    int state = IDLE, q = -1;
    int si = -1, sj = -1;
    double s_start = 0, c_start = 1;
    uint32_t n_nodes = 0, goal_parent = kNoNode, cur = 0;
    int iteration = 0, n_shots = 0;
    long long n_exp = 0, n_coll = 0;
    bool goal_found = false;
    double goal_g = -1.0;
    // children of the current expansion (lane = primitive), kept across rounds while a goal-map value is pending
    double nx = 0, ny = 0, nyaw = 0, steer = 0, direc = 1, ns = 0, nc = 1, sphi = 0, cphi = 1, g = 0, pr_old_cost = 0;
    bool need = false, pr_found = false;
    uint32_t pr_old = kNoNode, pr_slot = 0;
    unsigned long long idx = 0;
    int ci = -1, cj = -1, h1 = -1;
    double cn_x = 0, cn_y = 0; // position of the expanded node (stop radius)
    unsigned long long cur_key = 0; // open-list key the expanded node was popped with
    int exp_shots = 0, exp_checked = 0; // what the failed shot before the current expansion added to the statistics
    int n_done = 0;                 // warps of this CTA without work in the previous round
    int mid_slot = -1;              // grown workspace the current query has moved into (squad_grow), -1: the warp's own small one
    uint32_t mid_gen = 0;
    int node_cap = L.node_cap;
    auto bind_ws = [&]() {
        const bool mid = mid_slot >= 0;
        const WarpWorkspaceLayout &B = mid ? a.Lm : a.Ls;
        ws = mid ? a.mid_pool.base + (size_t)mid_slot * B.stride : a.solo_pool.base + (size_t)ws_slot * B.stride;
        nodes = reinterpret_cast<PlanNode *>(ws + B.node_off);
        scratch = reinterpret_cast<double *>(ws + B.traj_off);
        heap.key = reinterpret_cast<unsigned long long *>(ws + B.hkey_off);
        heap.id = reinterpret_cast<uint32_t *>(ws + B.hid_off);
        heap.cap = B.node_cap;
        closed.slot = reinterpret_cast<ClosedSlot *>(ws + B.table_off);
        closed.mask = (uint32_t)B.table_slots - 1;
        node_cap = B.node_cap;
    };

    SquadProf pf;
    pf.init();
    for (int rnd = 0;; ++rnd)
    {
        // ================= phase 1: every warp advances its own query =================
        unsigned needmask = 0; // children (or the start node, lane 0) posted for a heuristic in this round
        bool post_start = false;
        int r_bfs = 0, r_stale = 0, r_walks = 0, r_cause = 6; // profile: what this warp's step consisted of
        bool fresh = false;                      // a node was popped for expansion in this round
        unsigned free_mask = 0, walk_mask = 0;   // its children: end pose surely free / undecided (lane = primitive)
        PlanNode cn;                             // the popped node
        bool exp_done = false;   // an expansion completes in this round: its children are inserted in phase 5
        int finish = 0x7fffffff; // status to report when the query ends in this round (0x7fffffff: it goes on)
        if (state == DONE && a.migrate && n_done < W - W / 3)
        {
            // The tickets have run out, this CTA is still well filled and this warp has nothing to do: adopt a query from the
            // migration list, in place - the workspace that holds its state becomes this warp's.
            int got = -1;
            if (lane == 0)
            {
                const unsigned n = *reinterpret_cast<volatile unsigned *>(a.mig_n);
                const unsigned t = *reinterpret_cast<volatile unsigned *>(a.mig_taken);
                if (t < n && atomicCAS(a.mig_taken, t, t + 1) == t)
                    got = (int)t;
            }
            got = __shfl_sync(kFull, got, 0);
            if (got >= 0)
            {
                int wsv = kNotReady;
                if (lane == 0)
                {
                    while ((wsv = *reinterpret_cast<volatile int32_t *>(&a.mig_ws[got])) == kNotReady)
                        __nanosleep(100);
                    __threadfence();
                }
                wsv = __shfl_sync(kFull, wsv, 0);
                q = a.mig[got];
                if (wsv & kMidFlag)
                    mid_slot = wsv & ~kMidFlag;
                else
                {
                    // the custody workspace becomes this warp's own; its old one goes back to the pool
                    if (lane == 0)
                    {
                        if (ws_slot >= 0)
                            pool_release(a.solo_pool, ws_slot, gen);
                        atomicSub(a.solo_pool.custody, 1);
                    }
                    ws_slot = wsv;
                }
                bind_ws();
                const ResumeHeader *H = reinterpret_cast<const ResumeHeader *>(ws + (mid_slot >= 0 ? a.Lm.hdr_off : a.Ls.hdr_off));
                if (mid_slot >= 0)
                    mid_gen = H->gen;
                else
                    gen = H->gen;
                closed.tag = H->tag;
                heap.size = H->heap_size;
                n_nodes = H->n_nodes, iteration = H->iteration, n_shots = H->n_shots, goal_found = H->goal_found != 0, goal_parent = H->goal_parent;
                n_exp = H->n_exp, n_coll = H->n_coll, goal_g = H->goal_g, si = H->si, sj = H->sj;
                for (int k = lane; k < heap.size && k < kSoloHeapTop; k += 32)
                    S.hkey[k] = heap.key[k], S.hid[k] = heap.id[k];
                if (lane == 0)
                {
                    S.sx = a.starts[3 * q], S.sy = a.starts[3 * q + 1], S.syaw = a.starts[3 * q + 2];
                    S.gx = a.goals[3 * q], S.gy = a.goals[3 * q + 1], S.gyaw = a.goals[3 * q + 2];
                    S.bfs = H->bfs; // its pointers refer to the workspace that came with the query
                }
                __syncwarp();
                exp_shots = 0, exp_checked = 0;
                state = RUN;
            }
        }
        if (state == IDLE)
        {
            unsigned t = 0;
            if (lane == 0)
                t = atomicAdd(a.solo_counter, 1u);
            t = __shfl_sync(kFull, t, 0);
            if (t >= total)
                state = DONE;
            else
            {
                q = (int)t;
                if (ws_slot < 0)
                {
                    if (lane == 0)
                    {
                        ws_slot = pool_acquire(a.solo_pool, (int)(blockIdx.x * W + warp));
                        gen = a.solo_pool.gen[ws_slot];
                    }
                    ws_slot = __shfl_sync(kFull, ws_slot, 0);
                    gen = __shfl_sync(kFull, gen, 0);
                    bind_ws();
                }
                ++gen;
                closed.tag = gen;
                heap.size = 0;
                n_nodes = 0, goal_parent = kNoNode, iteration = 0, n_shots = 0, n_exp = 0, n_coll = 0, goal_found = false, goal_g = -1.0;
                __syncwarp();
                if (lane == 0)
                {
                    S.sx = a.starts[3 * q], S.sy = a.starts[3 * q + 1], S.syaw = a.starts[3 * q + 2];
                    S.gx = a.goals[3 * q], S.gy = a.goals[3 * q + 1], S.gyaw = a.goals[3 * q + 2];
                    S.bfs.blocked = reinterpret_cast<uint32_t *>(ws + L.blocked_off);
                    S.bfs.f0 = reinterpret_cast<uint32_t *>(ws + L.f0_off);
                    S.bfs.f1 = reinterpret_cast<uint32_t *>(ws + L.f1_off);
                    S.bfs.dist = reinterpret_cast<uint16_t *>(ws + L.dist_off);
                    S.bfs.cells = 0;
                }
                __syncwarp();
                // ---- initialize (hybrid_a_star.cpp:52-118)
                dm::sincos_(S.syaw, &s_start, &c_start);
                bool init_ok = !warp_footprint_collision(m, p, S.sx, S.sy, s_start, c_start); // "Start pose in collision"
                if (init_ok)
                {
                    double s, c;
                    dm::sincos_(S.gyaw, &s, &c);
                    init_ok = !warp_footprint_collision(m, p, S.gx, S.gy, s, c); // "Goal pose in collision"
                }
                int gi = -1, gj = -1;
                if (init_ok)
                {
                    const bool s_in = pos_to_index_ds(m, S.sx, S.sy, si, sj);
                    const bool g_in = pos_to_index_ds(m, S.gx, S.gy, gi, gj);
                    init_ok = s_in && g_in; // goal outside the map: no goal map (nav_map.cpp:297-302)
                }
                if (init_ok)
                {
                    solo_bfs_start(m, S.bfs, gi, gj, gen);
                    h1 = -1;
                    state = INIT;
                }
                else
                    finish = ST_NOT_INIT;
            }
        }
        else if (state == INIT)
        {
            // start node: steer 0, direction 1, g = 0 (:83, :211-215); its heuristic goes through the CTA's rounds like a child's
            if (squad_h1_lookup(m, S.bfs, lane == 0, si, sj, h1, kSquadBfsSteps, r_bfs))
            {
                h1 = __shfl_sync(kFull, h1, 0);
                dm::sincos_(S.gyaw - S.syaw, &sphi, &cphi);
                nx = S.sx, ny = S.sy, nyaw = S.syaw, ns = s_start, nc = c_start;
                needmask = 1u;
                post_start = true;
            }
        }
        else if (state == RUN || state == WAIT_H1)
        {
            // Tail of the batch: the tickets have run out and most warps of this CTA are done. The queries still running are
            // the long ones; a whole team finishes them faster than a nearly empty squad, and the team pass cannot start before
            // this kernel has ended: hand them over, state and all.
            // Throughput policy (a.migrate: other batches are in flight, cost per iteration counts): the hand-over is ADOPTABLE -
            // idle warps of squads that are still well filled take such queries over in place (below), so the long queries stay
            // in squad form while the thinned-out CTAs leave their SMs - and only a query older than a.age_limit iterations goes
            // to the team pass for good (those are the ones that run into the iteration limit; a squad would step them for seconds).
            const bool thin = n_done >= W - W / 3, aged = a.migrate && n_done > 0 && iteration >= a.age_limit;
            if (state == RUN && a.handoff_ws && (thin || aged) && iteration >= 32 && heap.size > 0 &&
                squad_hand_over(a, S, heap, q, ws_slot, gen, mid_slot, mid_gen, n_nodes, iteration, n_shots, n_exp, n_coll, goal_found, goal_parent,
                                goal_g, si, sj, !aged))
            {
                finish = kHandOff;
            }
            else if (state == RUN)
            {
                // ---- Plan() (hybrid_a_star.cpp:682-777): next open node that is not CLOSE (at most kSquadPops removals per round:
                // a run of CLOSE entries is finished in the next one)
                bool have = false;
                pf.sub0();
                for (int pops = 0; heap.size > 0 && pops < kSquadPops; ++pops)
                {
                    if (iteration > p.max_iteration)
                    {
                        finish = ST_LIMIT;
                        break;
                    }
                    cur = heap_pop_warp(heap, cur_key);
                    cn = nodes[cur];
                    if (!cn.closed)
                    {
                        have = true;
                        break;
                    }
                    ++r_stale;
                }
                if (!have && finish == 0x7fffffff && heap.size == 0)
                    finish = goal_found ? ST_FOUND : ST_OPEN_EMPTY;
                pf.sub(17);
                if (have)
                {
                    cn_x = cn.x, cn_y = cn.y;
                    // -- genRSPath (:853-900)
                    bool shot_ok = false;
                    const double ddx = cn.x - S.gx, ddy = cn.y - S.gy;
                    exp_shots = 0, exp_checked = 0;
                    if (ddx * ddx + ddy * ddy < p.rs_range * p.rs_range)
                    {
                        double rs_cost;
                        int checked;
                        const int np = solo_shot(m, p, a, S, cn, scratch, rs_cost, checked);
                        ++n_shots;
                        pf.add(15, 1);
                        n_coll += checked;
                        exp_shots = 1, exp_checked = checked;
                        if (np >= 0)
                        {
                            if (goal_parent == kNoNode || goal_g > cn.g + rs_cost)
                            {
                                goal_parent = cur;
                                goal_g = cn.g + rs_cost;
                            }
                            shot_ok = true;
                        }
                    }
                    if (shot_ok)
                    {
                        goal_found = true;
                        if (!a.optimal || dm::hypot_(S.gx - cn.x, S.gy - cn.y) <= 1.0)
                            finish = ST_FOUND; // first successful shot ends the search (:725-728); stop radius (hybrid_a_star.h:345)
                        else
                            ++iteration;
                    }
                    else
                    {
                        // -- expandNode (:304-402), lane = primitive; lanes P..2P-1 shadow the children for the second sin/cos pair
                        fresh = true;
                        ++n_exp;
                        n_coll += P;
                        const int c = lane < P ? lane : (lane < 2 * P ? lane - P : 0);
                        primitive_end_pose(p, c, cn.x, cn.y, cn.yaw, cn.s, cn.c, nx, ny, nyaw, steer, direc);
                        ns = 0, nc = 1;
                        if (lane < 2 * P)
                            dm::sincos_(lane < P ? nyaw : S.gyaw - nyaw, &ns, &nc);
                        sphi = __shfl_sync(kFull, ns, (lane + P) & 31);
                        cphi = __shfl_sync(kFull, nc, (lane + P) & 31);
                        pf.sub(18);
                        // footprint tests: most end poses are decided by their own cell, the others go to the CTA's walk pass
                        const int cls = lane < P ? footprint_preclass(m, p, a.unsafe_tw, a.center_probe, nx, ny, nyaw) : 1;
                        free_mask = __ballot_sync(kFull, cls == 0);
                        walk_mask = __ballot_sync(kFull, cls == 2);
                        r_walks = __popc(walk_mask);
                        pf.sub(19);
                    }
                }
            }
        }

        r_cause = state == DONE ? 6 : (state == INIT || state == IDLE || finish == ST_NOT_INIT) ? 5 : (exp_shots && !fresh) ? 4 : r_bfs ? 3 : r_stale ? 1 : (fresh ? 0 : 6);
        pf.hist(r_cause);
        // ================= phase 1w: the probe walks of every warp's undecided end poses, whole CTA =================
        free_mask |= squad_walks<W>(m, p, SQ, walk_mask, nx, ny, ns, nc, rnd, pf);

        // ================= phase 1b: closed set, g, goal-map values =================
        if (state == RUN || state == WAIT_H1)
        {
            bool expanding = state == WAIT_H1;
            if (fresh)
            {
                expanding = true;
                pf.sub0();
                // node cell, closed-set entry, g (lane = child)
                need = false, pr_found = false, pr_old = kNoNode, pr_slot = 0xffffffffu - lane, idx = 0, ci = -1, cj = -1, g = 0, pr_old_cost = 0;
                if (lane < P && ((free_mask >> lane) & 1u) && pos_to_index_ds(m, nx, ny, ci, cj))
                {
                    const float voro_f = __ldg(m.voronoi + (size_t)ci * m.W + cj);
                    idx = cal_index(m, p, ci, cj, yaw_to_idx(p, nyaw), direc);
                    pr_slot = closed_find(closed, idx, pr_found, pr_old_cost, pr_old);
                    g = node_cost(p, cn.g, (double)cn.dir, cn.steer, (double)voro_f, steer, direc);
                    need = pr_old_cost > g; // an earlier sibling on the same slot may still beat it; resolved at insert
                }
                h1 = -1;
                pf.sub(21);
            }
            if (expanding)
            {
                pf.sub0();
                const bool h1_final = squad_h1_lookup(m, S.bfs, need, ci, cj, h1, kSquadBfsSteps, r_bfs);
                pf.sub(22);
                if (h1_final)
                {
                    needmask = __ballot_sync(kFull, need);
                    exp_done = true;
                    state = RUN;
                }
                else
                    state = WAIT_H1; // the children stay in registers; nothing is posted in this round
            }
        }

        // ================= phases 2-4: heuristics of every posted child, whole CTA =================
        const double h = squad_heuristics<W>(m, p, SQ, needmask, S.gx, S.gy, S.gyaw, nx, ny, nyaw, ns, nc, sphi, cphi, h1, state == DONE, rnd, n_done, pf);
        const bool all_done = n_done == W;
        pf.add(8, 1);
        pf.add(12, exp_done ? 1 : 0);
        pf.add(13, state == WAIT_H1 ? 1 : 0);
        pf.add(14, (state == DONE || state == IDLE) ? 1 : 0);
        pf.add(16, state == INIT ? 1 : 0);

        // ================= phase 5: inserts =================
        if (post_start)
        {
            const double h0 = __shfl_sync(kFull, h, 0);
            if (lane == 0)
            {
                PlanNode n;
                n.x = S.sx, n.y = S.sy, n.yaw = S.syaw, n.steer = 0.0, n.g = 0.0, n.s = s_start, n.c = c_start, n.parent = kNoNode, n.dir = 1, n.closed = 0,
                n.pad = 0;
                nodes[0] = n;
                // cost_set_/node_set_ seeds (:107-112); the goal entry is written second and wins a shared cell
                bool found;
                double og;
                uint32_t on;
                const unsigned long long sidx = cal_index(m, p, si, sj, yaw_to_idx(p, S.syaw), 1.0);
                const uint32_t s0 = closed_find(closed, sidx, found, og, on);
                closed_store(closed, s0, sidx, 0.0, 0);
                const unsigned long long gidx = cal_index(m, p, S.bfs.gi, S.bfs.gj, yaw_to_idx(p, S.gyaw), 1.0);
                const uint32_t s1 = closed_find(closed, gidx, found, og, on);
                closed_store(closed, s1, gidx, 10000.0, kGoalNode);
                heap_push_lane0(heap, (unsigned long long)__double_as_longlong(0.0 + h0), 0);
            }
            heap.size = 1;
            n_nodes = 1;
            state = RUN;
            __syncwarp();
        }
        // the node pool must take all children of the expansion, or none (a handed-over query resumes between two iterations)
        bool overflow = !post_start && needmask && n_nodes + __popc(needmask) > (uint32_t)node_cap;
        if (!post_start && needmask && !overflow)
        {
            // closed-set compare / insert in primitive order (see plan_solo.cuh)
            const unsigned same = __match_any_sync(kFull, need ? pr_slot : 0xffffffffu - lane);
            const bool conflict = __any_sync(kFull, need && (same & (same - 1)) != 0);
            if (!conflict)
            {
                const int cnt = __popc(needmask);
                if (n_nodes + cnt > (uint32_t)node_cap)
                    overflow = true;
                else
                {
                    const int rank = __popc(needmask & ((1u << lane) - 1u));
                    const uint32_t id = n_nodes + rank;
                    const unsigned long long key = (unsigned long long)__double_as_longlong(g + h);
                    bool up = false;
                    const int pos = heap.size + rank;
                    if (need)
                    {
                        if (pr_found && pr_old < kGoalNode)
                            nodes[pr_old].closed = 1;
                        closed_store(closed, pr_slot, idx, g, id);
                        PlanNode nn;
                        nn.x = nx, nn.y = ny, nn.yaw = nyaw, nn.steer = steer, nn.g = g, nn.s = ns, nn.c = nc, nn.parent = cur;
                        nn.dir = (int8_t)(direc > 0 ? 1 : -1), nn.closed = 0, nn.pad = 0;
                        nodes[id] = nn;
                        if (pos > 0)
                        {
                            const int parent = (pos - 1) >> 5;
                            up = parent >= heap.size ? true : heap_less(key, id, hk(heap, parent), hi(heap, parent));
                        }
                    }
                    if (!__any_sync(kFull, up) && heap.size > 0)
                    {
                        if (need)
                            hset(heap, pos, key, id);
                        heap.size += cnt;
                    }
                    else
                    {
                        unsigned todo = needmask;
                        while (todo)
                        {
                            const int b = __ffs(todo) - 1;
                            todo &= todo - 1;
                            const unsigned long long k = __shfl_sync(kFull, key, b);
                            const uint32_t kid = __shfl_sync(kFull, id, b);
                            if (lane == 0)
                                heap_push_lane0(heap, k, kid);
                            heap.size++;
                            __syncwarp();
                        }
                    }
                    n_nodes += cnt;
                }
            }
            else if (!solo_insert_sequential(nodes, closed, heap, needmask, cur, n_nodes, node_cap, idx, g, h, nx, ny, nyaw, ns, nc, steer, direc))
                overflow = true;
            __syncwarp();
        }
        // end of an expansion (inserts done, or nothing to insert): overflow, stop radius, ++iteration (:725-745)
        if (exp_done)
        {
            if (overflow)
            {
                // The query outgrew the small workspace. Undo the pop (same key, same id: the open list is as it was but for
                // CLOSE entries dropped on the way, which no pop would ever have returned) and hand the query over with its
                // state; without a free workspace to go on with, the team pass starts it afresh.
                if (lane == 0)
                    heap_push_lane0(heap, cur_key, cur);
                heap.size++;
                __syncwarp();
                --n_exp;
                n_coll -= P + exp_checked; // the team pass repeats this step, shot included
                n_shots -= exp_shots;
                int grown = -1;
                if (mid_slot < 0 && a.mid_pool.n > 0)
                    grown = squad_grow(a, S, ws, gen, heap.size, n_nodes, q, mid_gen);
                if (grown >= 0)
                {
                    // the query goes on in a workspace four times the size; the step that did not fit is repeated there
                    mid_slot = grown;
                    closed.tag = mid_gen;
                    bind_ws();
                }
                else if (a.handoff_ws && squad_hand_over(a, S, heap, q, ws_slot, gen, mid_slot, mid_gen, n_nodes, iteration, n_shots, n_exp, n_coll,
                                                         goal_found, goal_parent, goal_g, si, sj, false)) // no squad has room for it: team pass
                    finish = kHandOff;
                else
                    finish = ST_OVERFLOW;
            }
            else if (a.optimal && dm::hypot_(S.gx - cn_x, S.gy - cn_y) <= 1.0) // stop_radius (hybrid_a_star.h:345)
                finish = goal_found ? ST_FOUND : ST_OPEN_EMPTY;
            else
                ++iteration;
        }

        if (finish != 0x7fffffff)
        {
            // ---- the query ends: hand-off, or status + raw path_ of setPath (hybrid_a_star.cpp:779-811)
            if (finish == kHandOff)
            {
                // squad_hand_over has listed the query
            }
            else if (finish == ST_OVERFLOW && a.handoff)
            {
                if (lane == 0)
                {
                    const unsigned at = atomicAdd(a.handoff_n, 1u);
                    a.handoff[at] = q;
                    if (a.handoff_ws)
                        a.handoff_ws[at] = -1;
                    a.status[q] = ST_OVERFLOW;
                }
            }
            else
            {
                int out_len = 0;
                if (finish == ST_FOUND)
                    out_len = solo_write_path(a, S, nodes, goal_parent, q, scratch);
                if (lane == 0)
                {
                    a.status[q] = finish;
                    a.cost[q] = (finish == ST_FOUND) ? goal_g : -1.0;
                    a.path_len[q] = out_len;
                    if (a.stats)
                    {
                        const bool init = finish != ST_NOT_INIT;
                        a.stats[6 * q + 0] = iteration;
                        a.stats[6 * q + 1] = n_exp * P;
                        a.stats[6 * q + 2] = n_shots;
                        a.stats[6 * q + 3] = n_coll;
                        a.stats[6 * q + 4] = init ? S.bfs.cells : 0;
                        a.stats[6 * q + 5] = n_nodes;
                    }
                }
            }
            __syncwarp();
            if (mid_slot >= 0)
            {
                // a finished query gives its grown workspace back; a handed-over one has left it to the team pass
                if (finish != kHandOff && lane == 0)
                    pool_release(a.mid_pool, mid_slot, mid_gen);
                mid_slot = -1;
            }
            if (ws_slot >= 0)
                bind_ws();
            state = IDLE;
            q = -1;
        }
        pf.mark(7);
        if (all_done)
            break;
    }
    pf.flush();
    __syncwarp();
    if (ws_slot >= 0 && lane == 0)
        pool_release(a.solo_pool, ws_slot, gen);
}

} // namespace mpros
