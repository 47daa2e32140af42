// Whole-query planner, one CTA ("team" of NW warps) per query.
//
// Why a team and not a warp per query: A* is a dependent chain (pop -> expand -> insert), so the per-iteration LATENCY
// sets both the tail of a batch (queries that run into the 100 000-iteration limit) and, with a fixed number of
// resident queries, the throughput. The first version (one warp per query) was instruction-fetch bound: eight
// independent warps per SM each walking ~170 KB of code (ncu: stall_no_instruction 4.2 per issue, issue active 16 %).
// Here all warps of a CTA walk the SAME iteration, with roles:
//   * compute warps 0..NW-3 : one warp per motion primitive (footprint test, one lane per probe), OMPL formulas, RS shot
//   * closed-set warp NW-2  : one lane per child: node cell, g, closed-set probe (DRAM latency hidden behind the
//                             footprint tests), symmetry images for the OMPL distance, later the inserts
//   * heap warp NW-1        : owns the open list. It publishes the next node as soon as it is known and restores the
//                             heap (sift-down of the removed root, sift-up of last iteration's inserts) IN THE BACKGROUND
//                             while the other warps expand, so the heap's dependent DRAM accesses leave the critical path.
// New children go to a small shared-memory insertion buffer; pop takes the minimum of (heap root, buffer), so a child
// that is the next best node never touches the heap.
//
// Reference semantics are unchanged: one popped node per step in std::multimap order ((f, insertion id) total order),
// children in steer x direction order, strict '>' prune, first successful shot ends the search
// (hybrid_a_star.cpp:682-777).
#pragma once
#include "planner_device.cuh"

namespace mpros
{

struct TeamBfs
{
    uint32_t *blocked, *f0, *f1;
    uint16_t *dist;
    uint32_t tag;
    int r0, w0, rows, wwords;
    int gi, gj;
    int level, done;
    int br0, br1, bw0, bw1;
    int ilo, ihi;                  // rows [ilo, ihi] (map coordinates) hold this query's bitmaps; others are stale
    int nr0, nr1, nw0, nw1, found; // per-level reduction targets
    long long cells;
};

template <int NW>
struct TeamShared
{
    // query
    double sx, sy, syaw, gx, gy, gyaw;
    int si, sj, gi, gj;
    int q, init_ok, init_hit[2];
    int ws_slot;      // workspace of the pool this team holds
    uint32_t ws_gen;  // its generation counter when the team took it
    // search state
    int iteration, flag, need_shot, shot_ok, overflow, goal_found, stop;
    uint32_t n_nodes, cur_id, goal_parent;
    double goal_g;
    int goal_traj_n, chain;
    long long n_prim, n_shots, n_coll;
    PlanNode cur;
    // children of the current expansion (slot = primitive index)
    int nslots;
    double c_x[kMaxPrim], c_y[kMaxPrim], c_yaw[kMaxPrim], c_s[kMaxPrim], c_c[kMaxPrim], c_steer[kMaxPrim], c_dir[kMaxPrim];
    double c_g[kMaxPrim], c_h[kMaxPrim];
    int c_i[kMaxPrim], c_j[kMaxPrim], c_h1[kMaxPrim];
    unsigned long long c_idx[kMaxPrim];
    unsigned char c_alive[kMaxPrim], c_cand[kMaxPrim], c_need[kMaxPrim];
    int h_list[kMaxPrim], h_n, bfs_more, h_skip;
    unsigned long long hmin[kMaxPrim][3];
    double k_x[kMaxPrim], k_y[kMaxPrim], k_yaw[kMaxPrim]; // closed-set warp's own copy of the child poses
    double sc[2][kMaxPrim][2];                            // {sin, cos} of the child yaw / of goal yaw - child yaw
    double hf[11][kMaxPrim];                              // OMPL base formula k on child c (min over the 4 images)
    int h_skipg[(kMaxPrim + 7) / 8];                      // group of 8 entries settled by the CSC round alone
    double img[kMaxPrim][4][7]; // per child x symmetry image: {x, y, phi, xb, yb, sin phi, cos phi} of the normalised goal
    // open list: top three levels of the 32-ary heap + the insertion buffers (double buffered by iteration parity)
    unsigned long long heap_key[kHeapTop];
    uint32_t heap_id[kHeapTop];
    unsigned long long buf_key[2][kMaxPrim];
    uint32_t buf_id[2][kMaxPrim];
    int buf_n[2];
    PlanNode buf_node[2][kMaxPrim];  // node records of the buffered children: the pop never waits for global memory
    PlanNode root_node;              // node record of the heap root, fetched in the background once the heap is restored
    uint32_t root_node_id;
    uint32_t closed_marks[kMaxPrim]; // nodes marked CLOSE by this iteration's inserts (newer than root_node)
    int n_marks;
    // Reeds-Shepp shot
    double cand_cost[NW];
    int cand_idx[NW];
    RsPath cand_path[NW];
    RsPath best;
    double best_cost;
    int traj_n, first_hit;
    unsigned char traj_hit[kTrajCap];
    double traj[kTrajCap * 5]; // kTrajCap x 3 {x, y, yaw}, then kTrajCap x 2 {steer, dir}
    TeamBfs bfs;
#ifdef MPROS_PROFILE
    unsigned long long prof[24];
#endif
};

// ---- workspace pool -------------------------------------------------------------------------------------------------------
// The per-query workspaces (node pool, open list, closed set, goal-wavefront window: ~110 MB at the 100 000-iteration limit)
// belong to a POOL of the context, not to a launch: a team takes a free one when its CTA starts and gives it back when the
// CTA exits. The register file admits a fixed number of resident teams per device (two per SM), so one pool of that many
// workspaces serves any number of batches in flight ("lanes"): the CTAs of the next batch's kernel take over the
// workspaces - and the SM slots - the previous batch frees. Stale contents are harmless: closed-set slots carry a
// generation tag that is unique per (workspace, query) - the generation counter lives with the workspace - and every other
// array is written by a query before that query reads it.
__device__ __forceinline__ int pool_acquire(const WorkspacePool &pl, int first)
{
    int i = first % pl.n, tries = 0;
    while (atomicCAS(&pl.owner[i], 0, 1) != 0)
    {
        i = i + 1 == pl.n ? 0 : i + 1;
        if (++tries >= pl.n) // every workspace is taken by a resident team of another batch: wait for one to finish
        {
            tries = 0;
            __nanosleep(2000);
        }
    }
    __threadfence(); // acquire: the previous owner's writes (and its generation counter) are visible
    return i;
}
__device__ __forceinline__ void pool_release(const WorkspacePool &pl, int i, uint32_t gen)
{
    pl.gen[i] = gen;
    __threadfence(); // release
    atomicExch(&pl.owner[i], 0);
}

// Non-blocking form for a whole warp that wants to put its present workspace into custody (see WorkspacePool::custody) and
// go on with another one: -1 when the custody limit is reached. Below the limit a free workspace exists; the 32 lanes look
// for it together.
__device__ __forceinline__ int pool_swap_acquire_warp(const WorkspacePool &pl, int first)
{
    const int lane = threadIdx.x & 31;
    int ok = 0;
    if (lane == 0)
    {
        ok = atomicAdd(pl.custody, 1) < pl.custody_max;
        if (!ok)
            atomicSub(pl.custody, 1);
    }
    if (!__shfl_sync(0xffffffffu, ok, 0))
        return -1;
    int base = first % pl.n;
    while (true)
    {
        int i = base + lane;
        i = i >= pl.n ? i - pl.n : i;
        const bool free_seen = lane < pl.n && *reinterpret_cast<volatile int *>(&pl.owner[i]) == 0;
        unsigned cand = __ballot_sync(0xffffffffu, free_seen);
        int got = -1;
        while (cand && got < 0)
        {
            const int l = __ffs(cand) - 1;
            cand &= cand - 1;
            int mine = -1;
            if (lane == l && atomicCAS(&pl.owner[i], 0, 1) == 0)
                mine = i;
            got = __shfl_sync(0xffffffffu, mine, l);
        }
        if (got >= 0)
        {
            __threadfence();
            return got;
        }
        base = base + 32 >= pl.n ? base + 32 - pl.n : base + 32;
    }
}

// State of a query that the squad pass (plan_squad.cuh) hands to the team pass TOGETHER with its workspace: the search goes
// on where it stopped, between two iterations. The node pool, the open list (complete in the global arrays), the closed set
// and the goal-wavefront window are in the small workspace; this header (at WarpWorkspaceLayout::hdr_off) holds the rest.
struct ResumeHeader
{
    uint32_t tag, gen; // closed-set tag of the query in the small workspace; generation counter to release the workspace with
    int heap_size;
    uint32_t n_nodes;
    int iteration, n_shots, goal_found;
    uint32_t goal_parent;
    long long n_exp, n_coll;
    double goal_g;
    int si, sj;
    TeamBfs bfs; // pointers refer to the small workspace
};
static_assert(sizeof(ResumeHeader) <= 512, "ResumeHeader does not fit the layout's header region");

// A CTA holds one team (plan_team_kernel, plan_cluster_kernel) or two (plan_duo_kernel); everything below addresses threads
// and barriers relative to the team: thread index 0 .. NW*32-1, named barriers 1 + 3 t (whole team), 2 + 3 t (compute group:
// every warp except the heap warp), 3 + 3 t (footprint warps -> closed-set warp, one way), t = team index inside the CTA.
template <int NW>
__device__ __forceinline__ int team_tid() { return (int)(threadIdx.x % (NW * 32)); }
template <int NW>
__device__ __forceinline__ int team_index() { return (int)(threadIdx.x / (NW * 32)); }
template <int NW>
__device__ __forceinline__ void team_sync_all()
{
    asm volatile("bar.sync %0, %1;" ::"r"(1 + 3 * team_index<NW>()), "r"(NW * 32) : "memory");
}
template <int NW>
__device__ __forceinline__ void team_sync()
{
    asm volatile("bar.sync %0, %1;" ::"r"(2 + 3 * team_index<NW>()), "r"((NW - 1) * 32) : "memory");
}

// Two teams of one CTA (plan_duo_kernel) keep their iterations in step, softly: the heap warp of a team that reaches the top
// of an iteration waits - for a bounded number of cycles - until the other team has reached its next one too. Both teams
// then walk the same ~50 KB of iteration code at the same time and share the instruction-cache fills (co-resident teams in
// separate CTAs run out of phase and evict each other's lines: 6.5 us per iteration alone, ~10 us with a neighbour).
struct DuoSync
{
    volatile unsigned count[2]; // iterations started by each team
    volatile int active[2];     // inside the search loop
};

// The 11 OMPL base formulas (rs_device.cuh: ompl_formula) are evaluated in two rounds. Round A: the two CSC formulas
// (k = 0, 1) on warps 0 and 1; any valid candidate is an upper bound of the RS distance, so when rho * min(CSC) <= h1 for
// every child the heuristic is max(h1, h2) = h1 exactly and round B is skipped. Round B: the other nine formulas spread
// over the NW-1 compute-group warps, heaviest (CCCC, with tauOmega) alone.
template <int NW>
__device__ __forceinline__ int team_formula_b(int warp, int slot);
template <>
__device__ __forceinline__ int team_formula_b<8>(int warp, int slot)
{
    const signed char tab[7][2] = {{4, -1}, {5, -1}, {6, -1}, {7, -1}, {2, 8}, {3, 9}, {10, -1}};
    return tab[warp][slot];
}
template <>
__device__ __forceinline__ int team_formula_b<4>(int warp, int slot)
{
    const signed char tab[3][3] = {{4, 6, 2}, {5, 7, 3}, {10, 8, 9}};
    return tab[warp][slot];
}
template <>
__device__ __forceinline__ int team_formula_b<6>(int warp, int slot)
{
    // five compute-group warps (four footprint warps + the closed-set warp); the CCSCC formula (10) is the longest
    const signed char tab[5][2] = {{4, 6}, {5, 7}, {2, 8}, {3, 9}, {10, -1}};
    return tab[warp][slot];
}
template <int NW>
__device__ __forceinline__ int team_formula_b_slots() { return NW == 4 ? 3 : 2; }

// ---- on-demand goal wavefront, compute-group version (same semantics as planner_device.cuh: GoalBfs) -----------------
template <int NW>
__device__ __forceinline__ void team_bfs_init_rows(const MapView &m, TeamBfs &b, int ra, int rb)
{
    // The initialised rows form one contiguous range [ilo, ihi] that only grows; rows outside it still hold a previous
    // query's bits and are never read.
    const int tid = team_tid<NW>(), CT = (NW - 1) * 32;
    ra = max(ra, b.r0);
    rb = min(rb, b.r0 + b.rows - 1);
    const int ilo = b.ilo, ihi = b.ihi;
    if (ra >= ilo && rb <= ihi)
        return;
    const int wwords = b.wwords;
    // new rows: [ra, ilo - 1] and [ihi + 1, rb] (the first call has ilo > ihi: everything is new)
    const int lo_n = (ilo > ihi) ? (rb - ra + 1) : max(0, ilo - ra), hi_n = (ilo > ihi) ? 0 : max(0, rb - ihi);
    for (int k = tid; k < (lo_n + hi_n) * wwords; k += CT)
    {
        int rr = k / wwords, w = k % wwords;
        int r = rr < lo_n ? ra + rr : ihi + 1 + (rr - lo_n);
        size_t o = (size_t)(r - b.r0) * wwords + w;
        b.blocked[o] = __ldg(m.ds_bits + (size_t)r * m.pitch + b.w0 + w);
        b.f0[o] = 0;
        b.f1[o] = 0;
    }
    team_sync<NW>();
    if (tid == 0)
    {
        b.ilo = (ilo > ihi) ? ra : min(ilo, ra);
        b.ihi = (ilo > ihi) ? rb : max(ihi, rb);
    }
    team_sync<NW>();
}

template <int NW>
__device__ __noinline__ void team_bfs_start(const MapView &m, TeamBfs &b, int gi, int gj, uint32_t tag)
{
    const int tid = team_tid<NW>();
    if (tid == 0)
    {
        const int wpr = (m.W + 31) >> 5;
        b.tag = tag;
        b.gi = gi;
        b.gj = gj;
        // only cells within 999 steps can ever hold a value < 1000
        b.r0 = max(0, gi - (kHugeInt - 1));
        int r1 = min(m.H - 1, gi + (kHugeInt - 1));
        b.w0 = max(0, gj - (kHugeInt - 1)) >> 5;
        int w1 = min(wpr - 1, (gj + (kHugeInt - 1)) >> 5);
        b.rows = r1 - b.r0 + 1;
        b.wwords = w1 - b.w0 + 1;
        b.level = 0;
        b.done = 0;
        b.cells = 1;
        b.br0 = b.br1 = gi;
        b.bw0 = b.bw1 = gj >> 5;
        b.ilo = 1, b.ihi = 0; // nothing initialised yet
    }
    team_sync<NW>();
    team_bfs_init_rows<NW>(m, b, gi - 2, gi + 2);
    if (tid == 0)
    {
        size_t o = (size_t)(gi - b.r0) * b.wwords + ((gj >> 5) - b.w0);
        b.f0[o] = 1u << (gj & 31);
        b.blocked[o] |= 1u << (gj & 31);
    }
    team_sync<NW>();
}

template <int NW>
__device__ __forceinline__ void team_bfs_step(const MapView &m, TeamBfs &b)
{
    const int tid = team_tid<NW>(), CT = (NW - 1) * 32;
    const int level = b.level + 1;
    if (level >= kHugeInt)
    {
        team_sync<NW>();
        if (tid == 0)
            b.done = 1;
        team_sync<NW>();
        return;
    }
    const int wr1 = b.r0 + b.rows - 1, ww1 = b.w0 + b.wwords - 1;
    const int sr0 = max(b.r0, b.br0 - 1), sr1 = min(wr1, b.br1 + 1);
    const int sw0 = max(b.w0, b.bw0 - 1), sw1 = min(ww1, b.bw1 + 1);
    team_bfs_init_rows<NW>(m, b, sr0 - 2, sr1 + 2);
    if (tid == 0)
    {
        b.nr0 = 1 << 30, b.nr1 = -1, b.nw0 = 1 << 30, b.nw1 = -1, b.found = 0;
    }
    team_sync<NW>();
    const int nw = sw1 - sw0 + 1;
    const int total = (sr1 - sr0 + 1) * nw;
    uint32_t *fprev = (level & 1) ? b.f0 : b.f1;
    uint32_t *fnext = (level & 1) ? b.f1 : b.f0;
    const int wpr = (m.W + 31) >> 5;
    int nr0 = 1 << 30, nr1 = -1, nw0 = 1 << 30, nw1 = -1, found = 0;
    const int rows = b.rows, wwords = b.wwords, r0 = b.r0, w0 = b.w0;
    for (int k = tid; k < total; k += CT)
    {
        int r = sr0 + k / nw, w = sw0 + k % nw;
        int lr = r - r0, lw = w - w0;
        size_t o = (size_t)lr * wwords + lw;
        uint32_t c = fprev[o];
        uint32_t nb = (c << 1) | (c >> 1);
        if (lw > 0)
            nb |= fprev[o - 1] >> 31;
        if (lw + 1 < wwords)
            nb |= fprev[o + 1] << 31;
        if (lr > 0)
            nb |= fprev[o - wwords];
        if (lr + 1 < rows)
            nb |= fprev[o + wwords];
        uint32_t valid = (w == wpr - 1 && (m.W & 31)) ? ((1u << (m.W & 31)) - 1u) : 0xffffffffu;
        uint32_t bl = b.blocked[o];
        uint32_t nbits = nb & ~bl & valid;
        fnext[o] = nbits;
        if (nbits)
        {
            b.blocked[o] = bl | nbits;
            found += __popc(nbits);
            nr0 = min(nr0, r), nr1 = max(nr1, r), nw0 = min(nw0, w), nw1 = max(nw1, w);
            uint16_t *drow = b.dist + o * 32;
            while (nbits)
            {
                int t = __ffs(nbits) - 1;
                nbits &= nbits - 1;
                drow[t] = (uint16_t)level;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        nr0 = min(nr0, __shfl_xor_sync(kFull, nr0, o));
        nr1 = max(nr1, __shfl_xor_sync(kFull, nr1, o));
        nw0 = min(nw0, __shfl_xor_sync(kFull, nw0, o));
        nw1 = max(nw1, __shfl_xor_sync(kFull, nw1, o));
        found += __shfl_xor_sync(kFull, found, o);
    }
    if ((tid & 31) == 0 && found)
    {
        atomicMin(&b.nr0, nr0);
        atomicMax(&b.nr1, nr1);
        atomicMin(&b.nw0, nw0);
        atomicMax(&b.nw1, nw1);
        atomicAdd(&b.found, found);
    }
    team_sync<NW>();
    if (tid == 0)
    {
        b.level = level;
        b.cells += b.found;
        if (b.found == 0)
            b.done = 1;
        else
        {
            b.br0 = min(b.br0, b.nr0), b.br1 = max(b.br1, b.nr1), b.bw0 = min(b.bw0, b.nw0), b.bw1 = max(b.bw1, b.nw1);
        }
    }
    team_sync<NW>();
}

// goal_cost_map_ value of cell (i, j); -1 while not final yet
__device__ __forceinline__ int team_bfs_peek(const MapView &m, const TeamBfs &b, int i, int j)
{
    if (i == b.gi && j == b.gj)
        return 0;
    if (bit_at(m.ds_bits, m.pitch, i, j))
        return kHugeInt; // occupied cells are never entered
    int lr = i - b.r0, lw = (j >> 5) - b.w0;
    if (lr < 0 || lr >= b.rows || lw < 0 || lw >= b.wwords)
        return kHugeInt; // further than 999 steps
    if (i >= b.ilo && i <= b.ihi)
    {
        size_t o = (size_t)lr * b.wwords + lw;
        if ((b.blocked[o] >> (j & 31)) & 1u) // visited (not occupied, checked above)
            return b.dist[o * 32 + (j & 31)];
    }
    return b.done ? kHugeInt : -1;
}

// goal-map values for the slots flagged c_need whose c_h1 is still unknown (< 0): advance the wavefront until final
template <int NW>
__device__ __noinline__ void team_h1_lookup(const MapView &m, TeamShared<NW> &S, Prof &pf)
{
    const int tid = team_tid<NW>();
    while (true)
    {
        if (tid == 0)
            S.bfs_more = 0;
        team_sync<NW>();
        if (tid < S.nslots && S.c_need[tid] && S.c_h1[tid] < 0)
        {
            int v = team_bfs_peek(m, S.bfs, S.c_i[tid], S.c_j[tid]);
            S.c_h1[tid] = v;
            if (v < 0)
                S.bfs_more = 1;
        }
        team_sync<NW>();
        if (!S.bfs_more)
            break;
        if (team_tid<NW>() == 0) pf.mark(3);
        team_bfs_step<NW>(m, S.bfs);
        if (team_tid<NW>() == 0) pf.mark(4);
        if (team_tid<NW>() == 0) pf.count(13);
    }
    if (team_tid<NW>() == 0) pf.mark(3);
}

// symmetry image q of the normalised goal seen from pose (x, y, yaw) into slot c (consumed by team_heuristics)
template <int NW>
__device__ __forceinline__ void team_image(const PlannerView &p, TeamShared<NW> &S, int c, int q, double px, double py, double pyaw)
{
    double x, y, phi, xb, yb, sphi, cphi;
    ompl_image(q, p.min_r, px, py, pyaw, S.gx, S.gy, S.gyaw, x, y, phi, xb, yb);
    dm::sincos_(phi, &sphi, &cphi);
    double *o = S.img[c][q];
    o[0] = x, o[1] = y, o[2] = phi, o[3] = xb, o[4] = yb, o[5] = sphi, o[6] = cphi;
}

// heuristics (hybrid_a_star.cpp:248-264) of the slots flagged c_need: h = max(h1 * res, RS distance) * 1.001
template <int NW>
__device__ __noinline__ void team_heuristics(const MapView &m, const PlannerView &p, TeamShared<NW> &S, Prof &pf)
{
    const int tid = team_tid<NW>(), warp = tid >> 5, lane = tid & 31;
    if (tid == 0)
    {
        int n = 0;
        for (int c = 0; c < S.nslots; ++c)
            if (S.c_need[c])
                S.h_list[n++] = c;
        S.h_n = n;
    }
    if (tid < kMaxPrim * 3)
        S.hmin[tid / 3][tid % 3] = (unsigned long long)__double_as_longlong(DBL_MAX);
    team_sync<NW>();
    const int hn = S.h_n;
    for (int base = 0; base < hn; base += 8)
    {
        const int e = base + (lane >> 2), q = lane & 3;
        const bool live = e < hn;
        const int c = live ? S.h_list[e] : 0;
        const double *im = S.img[c][q];
        // round A: CSC
        if (warp < 2)
        {
            double L = live ? ompl_formula(warp, im[0], im[1], im[2], im[3], im[4], im[5], im[6]) : DBL_MAX;
            L = fmin(L, __shfl_xor_sync(kFull, L, 1));
            L = fmin(L, __shfl_xor_sync(kFull, L, 2));
            if (live && q == 0 && L < DBL_MAX)
                atomicMin(&S.hmin[c][0], (unsigned long long)__double_as_longlong(L));
        }
        team_sync<NW>();
        if (tid == 0)
        {
            bool all = true;
            for (int t = base; t < min(hn, base + 8); ++t)
            {
                int cc = S.h_list[t];
                double ub = p.min_r * __longlong_as_double((long long)S.hmin[cc][0]);
                if (!(ub <= (double)S.c_h1[cc] * m.res))
                    all = false;
            }
            S.h_skip = all;
        }
        team_sync<NW>();
        if (team_tid<NW>() == 0) pf.mark(5);
        // round B: CCC, CCCC, CCSC, CCSCC
        if (!S.h_skip)
        {
            if (team_tid<NW>() == 0) pf.count(14);
#pragma unroll 1
            for (int s = 0; s < team_formula_b_slots<NW>(); ++s)
            {
                int k = team_formula_b<NW>(warp, s);
                if (k < 0)
                    continue;
                double L = live ? ompl_formula(k, im[0], im[1], im[2], im[3], im[4], im[5], im[6]) : DBL_MAX;
                L = fmin(L, __shfl_xor_sync(kFull, L, 1));
                L = fmin(L, __shfl_xor_sync(kFull, L, 2));
                if (live && q == 0 && L < DBL_MAX)
                    atomicMin(&S.hmin[c][ompl_formula_family(k)], (unsigned long long)__double_as_longlong(L));
            }
        }
        team_sync<NW>();
        if (team_tid<NW>() == 0) pf.mark(6);
    }
    if (tid < hn)
    {
        int c = S.h_list[tid];
        double h1 = (double)S.c_h1[c] * m.res; // getGoalCost: cost_ * map_resolution_
        double h2 = ompl_combine(p.min_r, __longlong_as_double((long long)S.hmin[c][0]), __longlong_as_double((long long)S.hmin[c][1]),
                                 __longlong_as_double((long long)S.hmin[c][2]));
        S.c_h[c] = fmax(h1, h2) * (1 + 1e-3);
    }
    team_sync<NW>();
    if (team_tid<NW>() == 0) pf.mark(7);
}

// Reeds-Shepp shot from (x, y, yaw, dir, steer) to the goal: FindOptimalPath + GenTraj into S.traj (compute group).
template <int NW>
__device__ __noinline__ void team_rs_solve(const RsConfig &rs, TeamShared<NW> &S, double x, double y, double yaw, int dir, double steer)
{
    const int tid = team_tid<NW>(), warp = tid >> 5, lane = tid & 31;
    constexpr int CW = NW - 1;
    // NewGoalPoint (rs_curve.cpp:51-58)
    double dx = S.gx - x, dy = S.gy - y;
    double cs, sn;
    dm::sincos_(yaw, &sn, &cs);
    double gx = (dx * cs + dy * sn) / rs.radius;
    double gy = (-dx * sn + dy * cs) / rs.radius;
    double gphi = S.gyaw - yaw;
    // candidate c = 4 f + j; warp w owns formulas w, w + CW, ...; lanes 4 g + j
    const int groups = (12 + CW - 1) / CW;
    double my_cost = 0;
    int my_c = 1 << 20;
    RsPath my;
    my.n = 0;
    {
        int g = lane >> 2, j = lane & 3;
        int f = warp + CW * g;
        if (g < groups && f < 12)
        {
            double cx = (j == 1 || j == 3) ? -gx : gx;
            double cy = (j == 2 || j == 3) ? -gy : gy;
            double ph = (j == 1 || j == 2) ? -gphi : gphi;
            rs_formula(f, cx, cy, ph, my);
            if (j == 1 || j == 3)
                rs_timeflip(my);
            if (j == 2 || j == 3)
                rs_reflect(my);
            double len = rs_path_length(rs, my, dir, steer);
            bool nan_seg = false;
#pragma unroll
            for (int k = 0; k < 5; ++k)
                if (k < my.n && isnan(my.s[k].len))
                    nan_seg = true;
            if ((len != 0.0) && !nan_seg)
            {
                my_cost = len;
                my_c = 4 * f + j;
            }
        }
    }
    double bc = my_cost;
    int bi = my_c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        double oc = __shfl_xor_sync(kFull, bc, o);
        int oi = __shfl_xor_sync(kFull, bi, o);
        bool take = (oi < (1 << 20)) && (bi >= (1 << 20) || oc < bc || (oc == bc && oi < bi));
        if (take)
        {
            bc = oc;
            bi = oi;
        }
    }
    if (my_c == bi && bi < (1 << 20))
    {
        S.cand_cost[warp] = bc;
        S.cand_idx[warp] = bi;
        S.cand_path[warp] = my;
    }
    else if (lane == 0 && bi >= (1 << 20))
        S.cand_idx[warp] = 1 << 20;
    team_sync<NW>();
    if (tid == 0)
    {
        int bw = -1;
        for (int w = 0; w < CW; ++w)
        {
            if (S.cand_idx[w] >= (1 << 20))
                continue;
            if (bw < 0 || S.cand_cost[w] < S.cand_cost[bw] || (S.cand_cost[w] == S.cand_cost[bw] && S.cand_idx[w] < S.cand_idx[bw]))
                bw = w;
        }
        S.best = S.cand_path[bw < 0 ? 0 : bw];
        if (bw < 0)
            S.best.n = 0;
        S.best_cost = bw < 0 ? 0.0 : S.cand_cost[bw];
        // GenTraj (rs_curve.cpp:165-222): sequential pose integration
        S.traj_n = rs_gen_traj(rs, x, y, yaw, S.best, S.traj, S.traj + 3 * kTrajCap, kTrajCap);
    }
    team_sync<NW>();
}

// genRSPath (hybrid_a_star.cpp:853-900) from the popped node S.cur: Reeds-Shepp solve, sampled trajectory, footprint tests.
// Out of line: shots are rare (a few per query) and their ~15 KB of code must not sit inside the hot loop.
template <int NW>
__device__ __noinline__ void team_shot(const MapView &m, const PlannerView &p, const PlanBatchArgs &a, TeamShared<NW> &S, Prof &pf)
{
    const int tid = team_tid<NW>(), warp = tid >> 5, lane = tid & 31;
    constexpr int CW = NW - 1;
    const PlanNode cn = S.cur;
    team_rs_solve<NW>(a.rs, S, cn.x, cn.y, cn.yaw, (int)cn.dir, cn.steer);
    const int np = S.traj_n;
    if (tid == 0)
    {
        S.n_shots++;
        S.first_hit = np > kTrajCap ? 0 : -1; // longer than the buffer cannot happen inside the shot range
    }
    team_sync<NW>();
    for (int base = 0; base < np && base < kTrajCap && S.first_hit < 0; base += CW)
    {
        int k = base + warp;
        bool hit = false;
        if (k < np)
        {
            double s, c;
            dm::sincos_(S.traj[3 * k + 2], &s, &c);
            hit = warp_footprint_collision(m, p, S.traj[3 * k], S.traj[3 * k + 1], s, c);
        }
        if (lane == 0 && k < kTrajCap)
            S.traj_hit[k] = hit;
        team_sync<NW>();
        if (tid == 0)
        {
            for (int t = base; t < min(np, base + CW); ++t)
                if (S.traj_hit[t])
                {
                    S.first_hit = t;
                    break;
                }
        }
        team_sync<NW>();
    }
    if (tid == 0)
    {
        S.n_coll += (S.first_hit >= 0) ? (S.first_hit + 1) : np;
        if (S.first_hit < 0)
        {
            if (S.goal_parent == kNoNode || S.goal_g > cn.g + S.best_cost)
            {
                S.goal_parent = S.cur_id;
                S.goal_g = cn.g + S.best_cost;
                S.goal_traj_n = np;
            }
            S.shot_ok = 1;
            S.goal_found = 1;
        }
    }
    team_sync<NW>();
    if (tid == 0) pf.count(15);
}

// Insert of the children one by one in primitive order (one thread): the fallback when two children of one expansion
// touch the same closed-set slot (same index, or two new keys probing to the same empty slot). Rare, out of line.
template <int NW>
__device__ __noinline__ void team_insert_sequential(PlanNode *nodes, const ClosedSet &closed, TeamShared<NW> &S, int P, int par_cur, int node_cap)
{
    uint32_t n_nodes = S.n_nodes;
    int nb = 0, nm = 0;
    for (int c = 0; c < P; ++c)
    {
        if (!S.c_need[c])
            continue;
        bool f2;
        double oc;
        uint32_t on;
        uint32_t sl = closed_find(closed, S.c_idx[c], f2, oc, on);
        double tg = S.c_g[c];
        if (!(oc > tg))
            continue;
        if (n_nodes >= (uint32_t)node_cap)
        {
            S.overflow = 1;
            break;
        }
        if (f2 && on < kGoalNode)
        {
            nodes[on].closed = 1;
            S.closed_marks[nm++] = on;
            for (int b2 = 0; b2 < nb; ++b2) // a sibling inserted a moment ago
                if (S.buf_id[par_cur][b2] == on)
                    S.buf_node[par_cur][b2].closed = 1;
        }
        closed_store(closed, sl, S.c_idx[c], tg, n_nodes);
        PlanNode nn;
        nn.x = S.c_x[c], nn.y = S.c_y[c], nn.yaw = S.c_yaw[c], nn.steer = S.c_steer[c], nn.g = tg, nn.parent = S.cur_id;
        nn.s = S.c_s[c], nn.c = S.c_c[c];
        nn.dir = (int8_t)(S.c_dir[c] > 0 ? 1 : -1), nn.closed = 0, nn.pad = 0;
        nodes[n_nodes] = nn;
        S.buf_node[par_cur][nb] = nn;
        S.buf_key[par_cur][nb] = (unsigned long long)__double_as_longlong(tg + S.c_h[c]);
        S.buf_id[par_cur][nb] = n_nodes;
        ++nb;
        n_nodes++;
    }
    S.buf_n[par_cur] = nb;
    S.n_nodes = n_nodes;
    S.n_marks = nm;
}

// ---- heuristics rounds (hybrid_a_star.cpp:248-264), shared by the single-CTA team and the heuristic CTA of a cluster -----
// h = max(h1 * res, RS distance) * 1.001 for the children flagged in needmask. Lanes = (child, symmetry image): child
// c = base + (lane >> 2), image q = lane & 3. Round A: the two CSC formulas on warps 0 and 1; any valid candidate is an
// upper bound of the RS distance, so when rho * min(CSC) <= h1 for every needed child of the group of 8 the other nine
// formulas cannot change max(h1, h2) and round B is skipped. Round B: formulas 2..10 spread over warps 0..NW-2.
struct HeurView
{
    double (*img)[4][7];      // [child][image]{x, y, phi, xb, yb, sin phi, cos phi}
    int *h1;                  // goal-map value per child
    double (*hf)[kMaxPrim];   // [formula][child] minimum over the four images
    int *skipg;               // per group of 8 children: settled by round A alone
};
template <int NW, bool CTA_WIDE>
__device__ __forceinline__ void heur_sync()
{
    if (CTA_WIDE)
        __syncthreads();
    else
        team_sync<NW>();
}
template <int NW, bool CTA_WIDE>
__device__ __forceinline__ void heur_rounds(const MapView &m, const PlannerView &p, const HeurView &V, unsigned needmask, int P, int writer_warp,
                                            Prof &pf)
{
    const int warp = team_tid<NW>() >> 5, lane = team_tid<NW>() & 31;
    for (int base = 0; base < P; base += 8)
    {
        if (((needmask >> base) & 0xffu) == 0)
            continue;
        const int c = base + (lane >> 2), q = lane & 3;
        const bool live = c < P && ((needmask >> c) & 1u);
        const double *im = V.img[live ? c : 0][q];
        if (warp < 2)
        {
            double Lf = live ? ompl_formula(warp, im[0], im[1], im[2], im[3], im[4], im[5], im[6]) : DBL_MAX;
            Lf = fmin(Lf, __shfl_xor_sync(kFull, Lf, 1));
            Lf = fmin(Lf, __shfl_xor_sync(kFull, Lf, 2));
            if (live && q == 0)
                V.hf[warp][c] = Lf;
        }
        heur_sync<NW, CTA_WIDE>(); // (2)
        if (team_tid<NW>() == 0) pf.mark(5);
        bool ok = true;
        {
            const int ce = base + lane;
            if (lane < 8 && ce < P && ((needmask >> ce) & 1u))
            {
                const double ub = p.min_r * fmin(V.hf[0][ce], V.hf[1][ce]);
                ok = ub <= (double)V.h1[ce] * m.res;
            }
        }
#ifdef MPROS_EXPERIMENT_SKIP_B // instruction-cache experiment (wrong heuristics): never run the nine non-CSC formulas
        const bool skip = true || __all_sync(kFull, ok);
#else
        const bool skip = __all_sync(kFull, ok);
#endif
        if (warp == writer_warp && lane == 0)
            V.skipg[base >> 3] = skip;
        if (!skip)
        {
            if (team_tid<NW>() == 0) pf.count(14);
            if (warp < NW - 1)
            {
#pragma unroll 1
                for (int sl = 0; sl < team_formula_b_slots<NW>(); ++sl)
                {
                    const int k = team_formula_b<NW>(warp, sl);
                    if (k < 0)
                        continue;
                    double Lf = live ? ompl_formula(k, im[0], im[1], im[2], im[3], im[4], im[5], im[6]) : DBL_MAX;
                    Lf = fmin(Lf, __shfl_xor_sync(kFull, Lf, 1));
                    Lf = fmin(Lf, __shfl_xor_sync(kFull, Lf, 2));
                    if (live && q == 0)
                        V.hf[k][c] = Lf;
                }
            }
            heur_sync<NW, CTA_WIDE>(); // (3)
        }
        if (team_tid<NW>() == 0) pf.mark(6);
    }
}
// one child's heuristic from the per-formula minima (same combination as the sequential CSC/CCC/CCCC, CCSC, CCSCC passes)
__device__ __forceinline__ double heur_combine(const MapView &m, const PlannerView &p, const HeurView &V, int c)
{
    double plain = fmin(V.hf[0][c], V.hf[1][c]), ccsc = DBL_MAX, ccscc = DBL_MAX;
    if (!V.skipg[c >> 3])
    {
        plain = fmin(fmin(plain, fmin(V.hf[2][c], V.hf[3][c])), fmin(V.hf[4][c], V.hf[5][c]));
        ccsc = fmin(fmin(V.hf[6][c], V.hf[7][c]), fmin(V.hf[8][c], V.hf[9][c]));
        ccscc = V.hf[10][c];
    }
    const double h1v = (double)V.h1[c] * m.res; // getGoalCost: cost_ * map_resolution_
    return fmax(h1v, ompl_combine(p.min_r, plain, ccsc, ccscc)) * (1 + 1e-3);
}

// ---- cluster mode: a team is a CLUSTER of two CTAs on the two SMs of a TPC ------------------------------------------------
// Why: the planner is bound by instruction-cache misses (ncu: sm__icc hit rate 66-72 %, SM time = misses x ~38 cycles): one
// A* iteration walks ~45-54 KB of code (search + eleven FP64 Reeds-Shepp formulas with their libdevice routines) against a
// ~32 KB per-SM instruction cache, so every line is fetched again every iteration. CTA rank 0 ("search": open list,
// expansion, footprint tests, closed set, inserts) and CTA rank 1 ("heuristic": the OMPL-equivalent RS distance) each run
// HALF of that code, and the hardware places rank 0 on the even and rank 1 on the odd SM of a TPC (measured: no SM ever gets
// both ranks), so each SM's instruction working set fits its cache. The children's symmetry images travel to the heuristic
// CTA through distributed shared memory (the search CTA's closed-set warp writes them straight into the peer's shared
// memory), the heuristics come back the same way; two cluster barriers (arrive.release / wait.acquire) per iteration order
// the exchange.
struct HeurShared
{
    double img[kMaxPrim][4][7];
    double hf[11][kMaxPrim];
    int h1[kMaxPrim];
    int skipg[(kMaxPrim + 7) / 8];
    unsigned needmask;
    int P;
    int cmd; // 0: work, 1: exit
#ifdef MPROS_PROFILE
    unsigned long long prof[24];
#endif
};
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
template <typename T>
__device__ __forceinline__ T *cluster_peer(T *local, unsigned rank)
{
    // generic address of the same shared-memory object in CTA `rank` of the cluster
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)local), "r"(rank));
    return reinterpret_cast<T *>(out);
}

// the heuristic CTA (cluster rank 1): serve one request per iteration of the search CTA until told to exit
template <int NW>
__device__ void heur_cta_loop(const MapView &m, const PlannerView &p, HeurShared &H, double *peer_c_h)
{
    const int warp = team_tid<NW>() >> 5, lane = team_tid<NW>() & 31;
    Prof pf;
#ifdef MPROS_PROFILE
    pf.start(H.prof); // marks of the shared heuristics code land here and are not reported
#endif
    HeurView V{H.img, H.h1, H.hf, H.skipg};
    while (true)
    {
        cluster_arrive(); // X: the search CTA has written images, h1, needmask
        cluster_wait();
        if (H.cmd != 0)
            break;
        const unsigned needmask = H.needmask;
        const int P = H.P;
        if (needmask)
        {
            heur_rounds<NW, true>(m, p, V, needmask, P, 0, pf);
            if (warp == 0)
            {
                __syncwarp();
                if (lane < P && ((needmask >> lane) & 1u))
                    peer_c_h[lane] = heur_combine(m, p, V, lane);
            }
        }
        cluster_arrive(); // Y: heuristics are in the search CTA's shared memory
        cluster_wait();
    }
}

// ---- heap warp --------------------------------------------------------------------------------------------------------
// Arg-min over the heap root (lane 31) and the entries of an insertion buffer (lanes 0..n-1), lexicographic on
// (key, id). Returns the winning lane, -1 when nothing is valid.
__device__ __forceinline__ int open_argmin(unsigned long long key, uint32_t id, bool valid)
{
    if (!__any_sync(kFull, valid))
        return -1;
    const uint32_t khi = valid ? (uint32_t)(key >> 32) : 0xffffffffu, klo = (uint32_t)key;
    const uint32_t m1 = __reduce_min_sync(kFull, khi);
    const bool a1 = valid && khi == m1;
    const uint32_t m2 = __reduce_min_sync(kFull, a1 ? klo : 0xffffffffu);
    const bool a2 = a1 && klo == m2;
    const uint32_t m3 = __reduce_min_sync(kFull, a2 ? id : 0xffffffffu);
    return __ffs(__ballot_sync(kFull, a2 && id == m3)) - 1;
}

// Take over a query from the squad pass: copy its state from the small workspace `src` into this team's workspace (node
// pool and open list as they are - both heaps are the same implicit 32-ary layout, only the number of entries kept in shared
// memory differs -, the closed set re-hashed into the bigger table under this team's tag, the initialised rows of the
// wavefront window) and give the small workspace back to its pool. All threads of the team.
template <int NW>
__device__ __noinline__ void team_resume(const PlanBatchArgs &a, int ws_small, char *ws, uint32_t tag, TeamShared<NW> &S, int &heap_size, int P)
{
    const int tid = team_tid<NW>(), NT = NW * 32;
    const bool from_mid = (ws_small & 0x40000000) != 0; // plan_squad.cuh: kMidFlag
    ws_small &= 0x3fffffff;
    const WorkspacePool &spool = from_mid ? a.mid_pool : a.solo_pool;
    const WarpWorkspaceLayout &L = a.L, &Ls = from_mid ? a.Lm : a.Ls;
    const char *src = spool.base + (size_t)ws_small * Ls.stride;
    const ResumeHeader *H = reinterpret_cast<const ResumeHeader *>(src + Ls.hdr_off);
    const uint32_t n_nodes = H->n_nodes;
    const int hsize = H->heap_size;
    {
        const uint4 *sn = reinterpret_cast<const uint4 *>(src + Ls.node_off);
        uint4 *dn = reinterpret_cast<uint4 *>(ws + L.node_off);
        for (size_t k = tid; k < (size_t)n_nodes * (sizeof(PlanNode) / 16); k += NT)
            dn[k] = sn[k];
    }
    {
        const unsigned long long *sk = reinterpret_cast<const unsigned long long *>(src + Ls.hkey_off);
        const uint32_t *sid = reinterpret_cast<const uint32_t *>(src + Ls.hid_off);
        unsigned long long *dk = reinterpret_cast<unsigned long long *>(ws + L.hkey_off);
        uint32_t *did = reinterpret_cast<uint32_t *>(ws + L.hid_off);
        for (int k = tid; k < hsize; k += NT)
        {
            if (k < kHeapTop)
                S.heap_key[k] = sk[k], S.heap_id[k] = sid[k];
            else
                dk[k] = sk[k], did[k] = sid[k];
        }
    }
    {
        // closed set: every live slot of the small table goes into the big one (distinct keys; a slot is claimed with a
        // compare-and-swap on its key word, so an entry lands on the first slot of its probe sequence that was not taken)
        const ClosedSlot *st = reinterpret_cast<const ClosedSlot *>(src + Ls.table_off);
        ClosedSlot *dt = reinterpret_cast<ClosedSlot *>(ws + L.table_off);
        const uint32_t dmask = (uint32_t)L.table_slots - 1;
        const unsigned long long old_tag = H->tag, new_tag = tag;
        for (int k = tid; k < Ls.table_slots; k += NT)
        {
            const ClosedSlot e = st[k];
            if ((e.key >> 34) != old_tag)
                continue;
            const unsigned long long idx = e.key & ((1ull << 34) - 1ull);
            const unsigned long long want = (new_tag << 34) | idx;
            uint32_t sl = hash_index(idx) & dmask;
            while (true)
            {
                const unsigned long long cur = dt[sl].key;
                if ((cur >> 34) != new_tag && atomicCAS(&dt[sl].key, cur, want) == cur)
                    break;
                sl = (sl + 1) & dmask;
            }
            dt[sl].g = e.g;
            dt[sl].node = e.node;
        }
    }
    {
        // goal wavefront: the rows initialised so far (same window geometry: it depends on the map and the goal only)
        const TeamBfs &b = H->bfs;
        if (b.ilo <= b.ihi)
        {
            const size_t w0 = (size_t)(b.ilo - b.r0) * b.wwords, nw = (size_t)(b.ihi - b.ilo + 1) * b.wwords;
            const uint32_t *sb = reinterpret_cast<const uint32_t *>(src + Ls.blocked_off), *s0 = reinterpret_cast<const uint32_t *>(src + Ls.f0_off),
                           *s1 = reinterpret_cast<const uint32_t *>(src + Ls.f1_off);
            uint32_t *db = reinterpret_cast<uint32_t *>(ws + L.blocked_off), *d0 = reinterpret_cast<uint32_t *>(ws + L.f0_off),
                     *d1 = reinterpret_cast<uint32_t *>(ws + L.f1_off);
            const uint4 *sd = reinterpret_cast<const uint4 *>(src + Ls.dist_off);
            uint4 *dd = reinterpret_cast<uint4 *>(ws + L.dist_off);
            for (size_t k = tid; k < nw; k += NT)
            {
                db[w0 + k] = sb[w0 + k];
                d0[w0 + k] = s0[w0 + k];
                d1[w0 + k] = s1[w0 + k];
            }
            for (size_t k = tid; k < nw * 4; k += NT) // 32 x uint16 per word = 4 x 16 bytes
                dd[w0 * 4 + k] = sd[w0 * 4 + k];
        }
    }
    if (tid == 0)
    {
        const TeamBfs keep = S.bfs; // this team's pointers
        S.bfs = H->bfs;
        S.bfs.blocked = keep.blocked, S.bfs.f0 = keep.f0, S.bfs.f1 = keep.f1, S.bfs.dist = keep.dist;
        S.bfs.tag = tag;
        S.si = H->si, S.sj = H->sj, S.gi = H->bfs.gi, S.gj = H->bfs.gj;
        S.iteration = H->iteration;
        S.n_nodes = n_nodes;
        S.n_prim = H->n_exp * P;
        S.n_shots = H->n_shots;
        S.n_coll = H->n_coll;
        S.goal_found = H->goal_found;
        S.goal_parent = H->goal_parent;
        S.goal_g = H->goal_g;
        S.init_ok = 1;
    }
    heap_size = hsize;
    const uint32_t gen_small = H->gen;
    __threadfence();
    team_sync_all<NW>(); // every thread has read what it needs from the small workspace
    if (tid == 0)
    {
        pool_release(spool, ws_small, gen_small);
        if (!from_mid)
            atomicSub(a.solo_pool.custody, 1);
    }
}

#ifndef MPROS_DUO_WAIT
#define MPROS_DUO_WAIT 6000 // cycles a team waits at most for its partner at the top of an iteration (plan_duo_kernel)
#endif
template <int NW, bool CL>
__device__ void plan_team_query(const MapView &m, const PlannerView &p, const PlanBatchArgs &a, int q, char *ws, uint32_t tag,
                                TeamShared<NW> &S, HeurShared *Hpeer, DuoSync *D = nullptr, int ws_small = -1)
{
    const int tid = team_tid<NW>(), warp = tid >> 5, lane = tid & 31;

    constexpr int HEAP_W = NW - 1;   // heap warp
    constexpr int CLOSED_W = NW - 2; // closed-set warp
    constexpr int COL_W = NW - 2;    // warps 0..COL_W-1 run the footprint tests
    const bool is_heap = warp == HEAP_W;
    const WarpWorkspaceLayout &L = a.L;
    PlanNode *nodes = reinterpret_cast<PlanNode *>(ws + L.node_off);
    OpenHeap heap;
    heap.key = reinterpret_cast<unsigned long long *>(ws + L.hkey_off);
    heap.id = reinterpret_cast<uint32_t *>(ws + L.hid_off);
    heap.skey = S.heap_key;
    heap.sid = S.heap_id;
    heap.scap = kHeapTop;
    heap.size = 0;
    heap.cap = L.node_cap;
    ClosedSet closed;
    closed.slot = reinterpret_cast<ClosedSlot *>(ws + L.table_off);
    closed.mask = (uint32_t)L.table_slots - 1;
    closed.tag = tag;
    Prof pf; // phase profile (developer build): thread 0, lane 0 of the heap warp and lane 0 of the closed-set warp write
#ifdef MPROS_PROFILE
    if (tid < 24)
        S.prof[tid] = 0;
    team_sync_all<NW>();
    const bool pf_on = tid == 0 || (lane == 0 && (warp == HEAP_W || warp == CLOSED_W));
    pf.start(S.prof);
    (void)pf_on;
#endif

    if (tid == 0)
    {
        S.sx = a.starts[3 * q], S.sy = a.starts[3 * q + 1], S.syaw = a.starts[3 * q + 2];
        S.gx = a.goals[3 * q], S.gy = a.goals[3 * q + 1], S.gyaw = a.goals[3 * q + 2];
        S.init_ok = 0;
        S.init_hit[0] = S.init_hit[1] = 0;
        S.iteration = 0, S.overflow = 0, S.goal_found = 0, S.stop = 0, S.flag = 0, S.shot_ok = 0;
        S.n_nodes = 0, S.goal_parent = kNoNode, S.goal_g = -1.0, S.goal_traj_n = 0, S.chain = 0;
        S.n_prim = S.n_shots = S.n_coll = 0;
        S.buf_n[0] = S.buf_n[1] = 0;
        S.root_node_id = kNoNode;
        S.n_marks = 0;
        S.bfs.blocked = reinterpret_cast<uint32_t *>(ws + L.blocked_off);
        S.bfs.f0 = reinterpret_cast<uint32_t *>(ws + L.f0_off);
        S.bfs.f1 = reinterpret_cast<uint32_t *>(ws + L.f1_off);
        S.bfs.dist = reinterpret_cast<uint16_t *>(ws + L.dist_off);
        S.bfs.cells = 0;
    }
    team_sync_all<NW>();

    // ---- a query handed over by the squad pass goes on from its state; otherwise:
    // ---- initialize (hybrid_a_star.cpp:52-118), compute group only; the heap warp waits at the barrier below ----
    if (ws_small >= 0)
        team_resume<NW>(a, ws_small, ws, tag, S, heap.size, 2 * p.n_steer);
    else if (!is_heap)
    {
        if (warp < 2)
        {
            double px = warp ? S.gx : S.sx, py = warp ? S.gy : S.sy, pyaw = warp ? S.gyaw : S.syaw;
            double s, c;
            dm::sincos_(pyaw, &s, &c);
            bool hit = warp_footprint_collision(m, p, px, py, s, c);
            if (lane == 0)
                S.init_hit[warp] = hit;
        }
        team_sync<NW>();
        if (tid == 0)
        {
            bool ok = !S.init_hit[0] && !S.init_hit[1]; // "Start pose in collision" / "Goal pose in collision"
            if (ok)
            {
                bool s_in = pos_to_index_ds(m, S.sx, S.sy, S.si, S.sj);
                bool g_in = pos_to_index_ds(m, S.gx, S.gy, S.gi, S.gj);
                ok = s_in && g_in; // goal outside the map: no goal map (nav_map.cpp:297-302)
            }
            S.init_ok = ok;
        }
        team_sync<NW>();
        if (S.init_ok)
        {
            team_bfs_start<NW>(m, S.bfs, S.gi, S.gj, tag);
            // start node: steer 0, direction 1, g = 0 (:83, :211-215); its heuristic through the generic slot machinery
            if (tid == 0)
            {
                S.nslots = 1;
                S.c_i[0] = S.si, S.c_j[0] = S.sj;
                S.c_need[0] = 1;
                S.c_h1[0] = -1;
            }
            if (tid >= 32 && tid < 36)
                team_image<NW>(p, S, 0, tid - 32, S.sx, S.sy, S.syaw);
            team_sync<NW>();
            team_h1_lookup<NW>(m, S, pf);
            team_heuristics<NW>(m, p, S, pf);
            if (tid == 0)
            {
                PlanNode n;
                n.x = S.sx, n.y = S.sy, n.yaw = S.syaw, n.steer = 0.0, n.g = 0.0, n.parent = kNoNode, n.dir = 1, n.closed = 0, n.pad = 0;
                dm::sincos_(S.syaw, &n.s, &n.c);
                nodes[0] = n;
                // cost_set_/node_set_ seeds (:107-112); the goal entry is written second and wins a shared cell
                bool found;
                double og;
                uint32_t on;
                unsigned long long sidx = cal_index(m, p, S.si, S.sj, yaw_to_idx(p, S.syaw), 1.0);
                uint32_t s0 = closed_find(closed, sidx, found, og, on);
                closed_store(closed, s0, sidx, 0.0, 0);
                unsigned long long gidx = cal_index(m, p, S.gi, S.gj, yaw_to_idx(p, S.gyaw), 1.0);
                uint32_t s1 = closed_find(closed, gidx, found, og, on);
                closed_store(closed, s1, gidx, 10000.0, kGoalNode);
                // the start node enters the open list through the insertion buffer of "iteration -1"
                S.buf_key[1][0] = (unsigned long long)__double_as_longlong(0.0 + S.c_h[0]);
                S.buf_id[1][0] = 0;
                S.buf_node[1][0] = n;
                S.buf_n[1] = 1;
                S.n_nodes = 1;
            }
        }
    }
    team_sync_all<NW>();
    const bool init_ok = S.init_ok != 0;
    pf.restart();

    // ---- Plan() (hybrid_a_star.cpp:682-777) ----
    const int P = 2 * p.n_steer;
    int bg_dropped = 0; // heap warp: stale roots removed ahead of time in the previous iteration's background phase
    unsigned duo_seen = 0; // heap warp, lane 0: the partner team's iteration count at this team's last iteration top
    if (D && init_ok && tid == 0)
        D->active[team_index<NW>()] = 1;
    if (D && is_heap && lane == 0)
        duo_seen = D->count[team_index<NW>() ^ 1];
    while (init_ok)
    {
        const int it = S.iteration;
        const int par_prev = (it + 1) & 1, par_cur = it & 1; // insertion buffers: last iteration's / this iteration's
        if (is_heap)
        {
            // -- pop: minimum of (heap root, last iteration's inserts). flag 0 continue, 1 iteration limit, 2 open set empty
            int flag = 0;
            uint32_t cur = 0;
            PlanNode cn;
            bool root_taken = false;
            int nbuf = S.buf_n[par_prev];
            unsigned long long bkey = lane < nbuf ? S.buf_key[par_prev][lane] : 0ull;
            uint32_t bid = lane < nbuf ? S.buf_id[par_prev][lane] : 0u;
            bool bvalid = lane < nbuf;
            if (it > p.max_iteration)
            {
                // the reference tests `while (!open_map_.empty())` BEFORE the limit (hybrid_a_star.cpp:702-709), and its
                // multimap still holds the CLOSE entries this warp dropped from the root in the last background phase
                flag = (heap.size > 0 || nbuf > 0 || bg_dropped > 0) ? 1 : 2;
            }
            else
            {
                if (nbuf == 32)
                {
                    // lane 31 is needed for the heap root: flush the whole buffer first (only with 16 steering angles)
                    for (int k = 0; k < 32; ++k)
                    {
                        if (lane == 0)
                            heap_push_lane0(heap, S.buf_key[par_prev][k], S.buf_id[par_prev][k]);
                        heap.size++;
                        __syncwarp();
                    }
                    bvalid = false;
                }
                while (true)
                {
                    unsigned long long key = bkey;
                    uint32_t id = bid;
                    bool valid = bvalid;
                    if (lane == 31 && heap.size > 0)
                    {
                        key = S.heap_key[0]; // the heap top lives in shared memory
                        id = S.heap_id[0];
                        valid = true;
                    }
                    const int win = open_argmin(key, id, valid);
                    if (win < 0)
                    {
                        flag = 2;
                        break;
                    }
                    cur = __shfl_sync(kFull, id, win);
                    const bool from_heap = (win == 31) && heap.size > 0;
                    if (!from_heap)
                        cn = S.buf_node[par_prev][win]; // a child of the last expansion: still in shared memory
                    else if (S.root_node_id == cur)
                    {
                        cn = S.root_node; // fetched in the background; CLOSE marks newer than the fetch are in closed_marks
                        const bool marked = lane < S.n_marks && S.closed_marks[lane] == cur;
                        if (__any_sync(kFull, marked))
                            cn.closed = 1;
                    }
                    else
                        cn = nodes[cur];
                    if (!from_heap && lane == win)
                        bvalid = false; // consumed straight from the buffer, never touches the heap
                    if (!cn.closed)
                    {
                        root_taken = from_heap;
                        break;
                    }
                    if (from_heap)
                    {
                        unsigned long long fk;
                        heap_pop_warp(heap, fk); // closed node: drop it now and look again
                    }
                }
            }
            if (is_heap && lane == 0) pf.mark(10);
            if (D)
            {
                // soft step with the partner team (DuoSync): wait, bounded, until it has started its next iteration too
                if (lane == 0)
                {
                    const int me = team_index<NW>(), other = me ^ 1;
                    D->count[me] = D->count[me] + 1;
                    const long long t0 = clock64();
                    while (D->active[other] && D->count[other] == duo_seen && clock64() - t0 < (long long)MPROS_DUO_WAIT)
                    {
                    }
                    duo_seen = D->count[other];
                }
                __syncwarp();
                if (is_heap && lane == 0) pf.mark(23);
            }
            if (lane == 0)
            {
                S.n_marks = 0;
                S.flag = flag;
                if (flag == 0)
                {
                    S.cur = cn;
                    S.cur_id = cur;
                    double ddx = cn.x - S.gx, ddy = cn.y - S.gy;
                    S.need_shot = (ddx * ddx + ddy * ddy) < p.rs_range * p.rs_range;
                    S.shot_ok = 0;
                }
            }
            team_sync_all<NW>(); // (a) the popped node is public
            if (CL && flag == 0)
                cluster_arrive(); // X of this iteration: the heap warp contributes nothing to the request, so it arrives at once
            if (flag == 0)
            {
                // -- background: restore the heap while the other warps expand
                if (root_taken)
                {
                    unsigned long long fk;
                    heap_pop_warp(heap, fk);
                }
                if (is_heap && lane == 0) pf.mark(21);
                unsigned todo = __ballot_sync(kFull, bvalid);
                if (todo)
                {
                    // common case: none of the new entries beats its parent (children mostly cost more than what is already
                    // open), so all of them are appended at once; one parent read per lane instead of a chain per entry
                    const int rank = __popc(todo & ((1u << lane) - 1u));
                    const int pos = heap.size + rank;
                    bool up = false;
                    if (bvalid && pos > 0)
                    {
                        const int parent = (pos - 1) >> 5;
                        up = parent >= heap.size ? true : heap_less(bkey, bid, hk(heap, parent), hi(heap, parent));
                    }
                    if (!__any_sync(kFull, up) && heap.size > 0)
                    {
                        if (bvalid)
                            hset(heap, pos, bkey, bid);
                        heap.size += __popc(todo);
                        todo = 0;
                        __syncwarp();
                    }
                }
                while (todo)
                {
                    const int b = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const unsigned long long k = __shfl_sync(kFull, bkey, b);
                    const uint32_t kid = __shfl_sync(kFull, bid, b);
                    if (lane == 0)
                        heap_push_lane0(heap, k, kid);
                    heap.size++;
                    __syncwarp();
                    if (is_heap && lane == 0) pf.count(14);
                }
                if (is_heap && lane == 0) pf.mark(22);
                // still in the background: drop roots that are already CLOSE (stale entries of nodes that were re-inserted
                // with a smaller g; each costs a full sift-down) and fetch the surviving root's node record, so that the next
                // pop is an arg-min over shared memory
                bg_dropped = 0;
                while (heap.size > 0)
                {
                    const uint32_t rid = S.heap_id[0];
                    if (lane < 4)
                        reinterpret_cast<uint4 *>(&S.root_node)[lane] = reinterpret_cast<const uint4 *>(&nodes[rid])[lane];
                    __syncwarp();
                    if (!S.root_node.closed)
                    {
                        if (lane == 0)
                            S.root_node_id = rid;
                        break;
                    }
                    unsigned long long fk;
                    heap_pop_warp(heap, fk);
                    ++bg_dropped;
                    if (is_heap && lane == 0) pf.count(13);
                }
                if (heap.size == 0 && lane == 0)
                    S.root_node_id = kNoNode;
                __syncwarp();
            }
            if (is_heap && lane == 0) pf.mark(11);
            if (CL && flag == 0)
            {
                // the cluster barriers count every thread of both CTAs: X (request, arrived above) and Y (heuristics)
                cluster_wait();
                cluster_arrive();
                cluster_wait();
            }
        }
        else
        {
            team_sync_all<NW>(); // (a)
            if (tid == 0) pf.mark(0);
            if (tid == 0) pf.count(12);
        }
        if (S.flag != 0)
            break;
        if (!is_heap)
        {
            const PlanNode cn = S.cur;
            if (tid == 0)
                S.buf_n[par_cur] = 0;
            // -- genRSPath (:853-900)
            if (S.need_shot)
                team_shot<NW>(m, p, a, S, pf);
            if (tid == 0) pf.mark(1);
            pf.restart();
            if (S.shot_ok)
            {
                if (!a.optimal && tid == 0)
                    S.stop = 1; // first successful shot ends the search (:725-728)
                if (CL)
                {
                    // no expansion in this iteration: an empty request keeps the two CTAs in step
                    if (tid == 0)
                    {
                        Hpeer->needmask = 0;
                        Hpeer->cmd = 0;
                    }
                    cluster_arrive();
                    cluster_wait();
                    cluster_arrive();
                    cluster_wait();
                }
            }
            else
            {
                // -- expandNode (:304-402). Three barriers of the compute group: children, CSC round, (other formulas).
                //   warps 0..COL_W-1 : one primitive each: end pose, sin/cos, footprint test (one lane per probe)
                //   closed-set warp  : one lane per child: node cell, closed-set slot, Voronoi cost and goal-map value with ALL
                //                      global loads issued together, the symmetry images computed while they are in flight
                bool pr_found = false;
                double pr_old_cost = 0;
                uint32_t pr_old = kNoNode, pr_slot = 0xffffffffu - lane;
                if (warp < COL_W)
                {
                    // (i) end pose and both sin/cos pairs of every child of this warp: lane 0 the child yaw (footprint test,
                    // node record), lane 1 goal yaw - child yaw (symmetry images) - one call, two lanes
                    for (int c = warp; c < P; c += COL_W)
                    {
                        if (lane < 2)
                        {
                            double nx, ny, nyaw, steer, direc, sn, cs;
                            primitive_end_pose(p, c, cn.x, cn.y, cn.yaw, cn.s, cn.c, nx, ny, nyaw, steer, direc);
                            dm::sincos_(lane ? S.gyaw - nyaw : nyaw, &sn, &cs);
                            S.sc[lane][c][0] = sn, S.sc[lane][c][1] = cs;
                            if (lane == 0)
                            {
                                S.c_x[c] = nx, S.c_y[c] = ny, S.c_yaw[c] = nyaw, S.c_s[c] = sn, S.c_c[c] = cs;
                                S.c_steer[c] = steer, S.c_dir[c] = direc;
                            }
                        }
                    }
                    __syncwarp();
                    // the closed-set warp builds the images from these as soon as all footprint warps have arrived
                    asm volatile("bar.arrive %0, %1;" ::"r"(3 + 3 * team_index<NW>()), "r"((COL_W + 1) * 32) : "memory");
                    if (tid == 0) pf.mark(16);
                    // (ii) footprint tests, one lane per probe
                    for (int c = warp; c < P; c += COL_W)
                    {
                        const bool hit = warp_footprint_collision(m, p, S.c_x[c], S.c_y[c], S.c_s[c], S.c_c[c]);
                        if (lane == 0)
                            S.c_alive[c] = hit ? 0 : 1;
                    }
                    if (tid == 0) pf.mark(17);
                }
                else if (warp == CLOSED_W)
                {
                    pf.restart();
                    // phase A (lane = child): cell, key, and every global load the child needs, issued back to back
                    double k_steer = 0, k_direc = 1;
                    int ci = -1, cj = -1;
                    bool inmap = false, isgoal = false, inwin = false;
                    unsigned long long idx = 0;
                    uint32_t hs = 0, first_node = kNoNode, dsw = 0, blk = 0;
                    uint4 first = make_uint4(0, 0, 0, 0);
                    float voro_f = 0.f;
                    uint16_t dv = 0;
                    if (lane < P)
                    {
                        double nx, ny, nyaw;
                        primitive_end_pose(p, lane, cn.x, cn.y, cn.yaw, cn.s, cn.c, nx, ny, nyaw, k_steer, k_direc);
                        S.k_x[lane] = nx, S.k_y[lane] = ny, S.k_yaw[lane] = nyaw;
                        inmap = pos_to_index_ds(m, nx, ny, ci, cj);
                        if (inmap)
                        {
                            voro_f = __ldg(m.voronoi + (size_t)ci * m.W + cj);
                            idx = cal_index(m, p, ci, cj, yaw_to_idx(p, nyaw), k_direc);
                            hs = hash_index(idx) & closed.mask;
                            first = *reinterpret_cast<const uint4 *>(&closed.slot[hs]);
                            first_node = closed.slot[hs].node;
                            // goal-map value (team_bfs_peek) with its three loads hoisted
                            const TeamBfs &b = S.bfs;
                            isgoal = (ci == b.gi && cj == b.gj);
                            dsw = __ldg(m.ds_bits + (size_t)ci * m.pitch + (cj >> 5));
                            const int lr = ci - b.r0, lw = (cj >> 5) - b.w0;
                            inwin = !(lr < 0 || lr >= b.rows || lw < 0 || lw >= b.wwords);
                            if (inwin && ci >= b.ilo && ci <= b.ihi)
                            {
                                const size_t o = (size_t)lr * b.wwords + lw;
                                blk = b.blocked[o];
                                dv = b.dist[o * 32 + (cj & 31)];
                            }
                        }
                    }
                    __syncwarp();
                    if (warp == CLOSED_W && lane == 0) pf.mark(18);
                    // phase B: the sin/cos pairs come from the footprint warps (child yaw and goal yaw - child yaw)
                    asm volatile("bar.sync %0, %1;" ::"r"(3 + 3 * team_index<NW>()), "r"((COL_W + 1) * 32) : "memory");
                    // phase C: the four symmetry images of the normalised goal per child (ompl_image; sin(-x) = -sin(x) and
                    // cos(-x) = cos(x) hold bit for bit, so the image's own sin/cos phi follow from those of dth)
                    for (int t = lane; t < 4 * P; t += 32)
                    {
                        const int c = t >> 2, q = t & 3;
                        const double dx = S.gx - S.k_x[c], dy = S.gy - S.k_y[c], dth = S.gyaw - S.k_yaw[c];
                        const double sn = S.sc[0][c][0], cs = S.sc[0][c][1];
                        const double inx = dm::div_(cs * dx + sn * dy, p.min_r), iny = dm::div_(-sn * dx + cs * dy, p.min_r);
                        const double sphi = S.sc[1][c][0], cphi = S.sc[1][c][1];
                        const double nxb = inx * cphi + iny * sphi, nyb = inx * sphi - iny * cphi;
                        const bool fx = (q == 1 || q == 3), fy = (q == 2 || q == 3), fp = (q == 1 || q == 2);
                        double *o = CL ? Hpeer->img[c][q] : S.img[c][q]; // cluster mode: straight into the heuristic CTA's shared memory
                        o[0] = fx ? -inx : inx, o[1] = fy ? -iny : iny, o[2] = fp ? -dth : dth, o[3] = fx ? -nxb : nxb, o[4] = fy ? -nyb : nyb;
                        o[5] = fp ? -sphi : sphi, o[6] = cphi;
                    }
                    // phase D (lane = child): consume the loads: g, closed-set compare, goal-map value
                    bool cand = false;
                    int h1 = -1;
                    if (inmap)
                    {
                        const double g = node_cost(p, cn.g, (double)cn.dir, cn.steer, (double)voro_f, k_steer, k_direc);
                        const unsigned long long want = (closed.tag << 34) | idx;
                        const unsigned long long k0 = ((unsigned long long)first.y << 32) | first.x;
                        if (k0 == want)
                        {
                            pr_found = true;
                            pr_old_cost = __longlong_as_double((long long)(((unsigned long long)first.w << 32) | first.z));
                            pr_old = first_node;
                            pr_slot = hs;
                        }
                        else if ((k0 >> 34) != closed.tag)
                        {
                            pr_found = false;
                            pr_old_cost = __longlong_as_double(0x7ff0000000000000ll);
                            pr_old = kNoNode;
                            pr_slot = hs;
                        }
                        else // first slot taken by another key: keep probing (rare at this load factor)
                            pr_slot = closed_find(closed, idx, pr_found, pr_old_cost, pr_old);
                        cand = pr_old_cost > g; // an earlier sibling on the same slot may still beat it; resolved at insert
                        S.c_g[lane] = g, S.c_idx[lane] = idx;
                        if (cand)
                        {
                            if (isgoal)
                                h1 = 0;
                            else if ((dsw >> (cj & 31)) & 1u)
                                h1 = kHugeInt; // occupied cells are never entered
                            else if (!inwin)
                                h1 = kHugeInt; // further than 999 steps
                            else if ((blk >> (cj & 31)) & 1u)
                                h1 = dv; // visited (blk is 0 for rows that are not initialised yet)
                            else
                                h1 = S.bfs.done ? kHugeInt : -1;
                        }
                    }
                    if (lane < P)
                    {
                        // a cell outside the map cannot pass the footprint test; the reference would index outside its arrays
                        S.c_i[lane] = ci, S.c_j[lane] = cj;
                        S.c_cand[lane] = cand;
                        S.c_h1[lane] = h1;
                    }
                    if (lane == 0)
                    {
                        if (a.trace)
                        {
                            // tree_ of the reference (hybrid_a_star.cpp:352) holds P entries per expansion, all with this parent
                            const long long k = S.n_prim / P;
                            if (k < a.trace_cap)
                                a.trace[3 * k] = cn.x, a.trace[3 * k + 1] = cn.y, a.trace[3 * k + 2] = cn.yaw;
                            *a.trace_n = k + 1;
                        }
                        S.nslots = P;
                        S.n_prim += P;
                        S.n_coll += P;
                    }
                    if (warp == CLOSED_W && lane == 0) pf.mark(19);
                }
                team_sync<NW>(); // (1) children public
                if (tid == 0) pf.mark(2);
                // every warp derives the set of children that need a heuristic (alive, and cheaper than the closed set's entry)
                bool need = lane < P && S.c_alive[lane] && S.c_cand[lane];
                if (__any_sync(kFull, need && S.c_h1[lane] < 0))
                {
                    // some goal-map value is not final yet: advance the wavefront (rare; barriers inside)
                    if (tid < P)
                        S.c_need[tid] = need;
                    team_h1_lookup<NW>(m, S, pf);
                }
                if (tid == 0) pf.mark(3);
                const unsigned needmask = __ballot_sync(kFull, need);
                if (CL)
                {
                    // hand the request to the heuristic CTA (rank 1): goal-map values and the set of children; the images are
                    // already there. X: request complete; Y: S.c_h holds the heuristics.
                    if (warp == 0)
                    {
                        if (lane < P)
                            Hpeer->h1[lane] = S.c_h1[lane];
                        if (lane == 0)
                        {
                            Hpeer->needmask = needmask;
                            Hpeer->P = P;
                            Hpeer->cmd = 0;
                        }
                    }
                    cluster_arrive();
                    cluster_wait();
                    if (tid == 0) pf.mark(5);
                    cluster_arrive();
                    cluster_wait();
                    if (tid == 0) pf.mark(6);
                }
                else
                {
                    HeurView V{S.img, S.c_h1, S.hf, S.h_skipg};
                    heur_rounds<NW, false>(m, p, V, needmask, P, CLOSED_W, pf);
                }
                // closed-set compare / insert in primitive order (closed-set warp). Common case: the children that need an
                // insert touch distinct hash slots -> slot / node stores and close-marks run one lane per child, the new open
                // entries go to the insertion buffer. Children that collide on a slot (same index, or two new keys probing
                // to the same empty slot) fall back to the strictly sequential walk.
                if (warp == CLOSED_W)
                {
                    __syncwarp();
                    pf.restart();
                    if (need)
                    {
                        if (!CL)
                        {
                            HeurView V{S.img, S.c_h1, S.hf, S.h_skipg};
                            S.c_h[lane] = heur_combine(m, p, V, lane);
                        }
                        S.c_need[lane] = 1;
                    }
                    else if (lane < P)
                        S.c_need[lane] = 0;
                    __syncwarp();
                    uint32_t n_nodes = S.n_nodes;
                    const uint32_t slot = need ? pr_slot : 0xffffffffu - lane;
                    const unsigned same = __match_any_sync(kFull, slot);
                    const bool conflict = __any_sync(kFull, need && (same & (same - 1)) != 0);
                    if (!conflict)
                    {
                        const bool ins = need && pr_old_cost > S.c_g[lane];
                        const unsigned insmask = __ballot_sync(kFull, ins);
                        const int rank = __popc(insmask & ((1u << lane) - 1u));
                        const uint32_t id = n_nodes + rank;
                        if (n_nodes + __popc(insmask) > (uint32_t)L.node_cap)
                        {
                            if (lane == 0)
                                S.overflow = 1;
                        }
                        else
                        {
                            const bool mk = ins && pr_found && pr_old < kGoalNode;
                            const unsigned mkmask = __ballot_sync(kFull, mk);
                            if (mk)
                                S.closed_marks[__popc(mkmask & ((1u << lane) - 1u))] = pr_old;
                            if (lane == 0)
                                S.n_marks = __popc(mkmask);
                            if (ins)
                            {
                                if (mk)
                                    nodes[pr_old].closed = 1;
                                closed_store(closed, slot, S.c_idx[lane], S.c_g[lane], id);
                                PlanNode nn;
                                nn.x = S.c_x[lane], nn.y = S.c_y[lane], nn.yaw = S.c_yaw[lane], nn.steer = S.c_steer[lane], nn.g = S.c_g[lane];
                                nn.s = S.c_s[lane], nn.c = S.c_c[lane];
                                nn.parent = S.cur_id, nn.dir = (int8_t)(S.c_dir[lane] > 0 ? 1 : -1), nn.closed = 0, nn.pad = 0;
                                nodes[id] = nn;
                                S.buf_node[par_cur][rank] = nn;
                                S.buf_key[par_cur][rank] = (unsigned long long)__double_as_longlong(S.c_g[lane] + S.c_h[lane]);
                                S.buf_id[par_cur][rank] = id;
                            }
                            if (lane == 0)
                            {
                                S.buf_n[par_cur] = __popc(insmask);
                                S.n_nodes = n_nodes + __popc(insmask);
                            }
                        }
                    }
                    else if (lane == 0)
                        team_insert_sequential<NW>(nodes, closed, S, P, par_cur, L.node_cap);
                    __syncwarp();
                    if (lane == 0)
                    {
                        if (S.overflow)
                            S.stop = 1;
                        if (a.optimal && dm::hypot_(S.gx - cn.x, S.gy - cn.y) <= 1.0) // stop_radius (hybrid_a_star.h:345)
                            S.stop = 1;
                        if (!S.stop)
                            S.iteration = it + 1; // every break of the reference's loop precedes ++iteration (:725-745)
                    }
                    if (warp == CLOSED_W && lane == 0) pf.mark(20);
                }
                if (tid == 0) pf.mark(8);
            }
            if (S.shot_ok && tid == 0)
            {
                if (a.optimal && dm::hypot_(S.gx - cn.x, S.gy - cn.y) <= 1.0) // stop_radius (hybrid_a_star.h:345)
                    S.stop = 1;
                if (!S.stop)
                    S.iteration = it + 1;
            }
        }
        team_sync_all<NW>(); // (b) inserts are in the buffer, the heap is consistent again
        if (!is_heap)
        {
            if (tid == 0) pf.mark(9);
        }
        else
            pf.restart();
        if (S.stop)
            break;
    }
    if (D && tid == 0)
        D->active[team_index<NW>()] = 0; // the partner stops waiting for this team
#ifdef MPROS_PROFILE
    team_sync_all<NW>();
    if (tid < 24 && S.prof[tid])
        atomicAdd(&g_prof[tid], S.prof[tid]);
#endif

    // ---- status + raw path_ of setPath (hybrid_a_star.cpp:779-811) ----
    int status = ST_NOT_INIT;
    if (init_ok)
    {
        if (S.overflow)
            status = ST_OVERFLOW;
        else if (S.goal_found)
            status = ST_FOUND;
        else if (S.flag == 1)
            status = ST_LIMIT;
        else
            status = ST_OPEN_EMPTY;
    }
    int out_len = 0;
    if (status == ST_FOUND)
    {
        if (a.optimal) // the best shot may be older than the last one: re-generate it (deterministic)
        {
            if (!is_heap)
            {
                const PlanNode gp = nodes[S.goal_parent];
                team_rs_solve<NW>(a.rs, S, gp.x, gp.y, gp.yaw, (int)gp.dir, gp.steer);
                if (tid == 0)
                    S.goal_traj_n = S.traj_n;
            }
            team_sync_all<NW>();
        }
        if (tid == 0)
        {
            int chain = 0;
            uint32_t c = S.goal_parent;
            while (nodes[c].parent != kNoNode)
            {
                ++chain;
                c = nodes[c].parent;
            }
            S.chain = chain;
            if (a.paths && a.path_cap > 0)
            {
                double *dst = a.paths + (size_t)q * a.path_cap * 5;
                c = S.goal_parent;
                int pos = chain;
                double last_dir = 1.0;
                while (nodes[c].parent != kNoNode)
                {
                    const PlanNode n = nodes[c];
                    if (pos < a.path_cap)
                    {
                        double *o = dst + 5 * pos;
                        o[0] = n.x, o[1] = n.y, o[2] = n.yaw, o[3] = n.steer, o[4] = (double)n.dir;
                    }
                    last_dir = (double)n.dir; // ends as the direction of the node next to the start
                    --pos;
                    c = n.parent;
                }
                // start point: steer 0, direction of the first path point (:792-807)
                double start_dir = chain > 1 ? last_dir : (S.goal_traj_n > 0 ? S.traj[3 * kTrajCap + 1] : 1.0);
                dst[0] = S.sx, dst[1] = S.sy, dst[2] = S.syaw, dst[3] = 0.0, dst[4] = start_dir;
            }
        }
        team_sync_all<NW>();
        const int chain = S.chain;
        out_len = 1 + chain + S.goal_traj_n;
        if (a.paths && a.path_cap > 0)
        {
            double *dst = a.paths + (size_t)q * a.path_cap * 5;
            for (int k = tid; k < S.goal_traj_n && k < kTrajCap; k += NW * 32)
            {
                int pos = 1 + chain + k;
                if (pos < a.path_cap)
                {
                    double *o = dst + 5 * pos;
                    o[0] = S.traj[3 * k], o[1] = S.traj[3 * k + 1], o[2] = S.traj[3 * k + 2];
                    o[3] = S.traj[3 * kTrajCap + 2 * k], o[4] = S.traj[3 * kTrajCap + 2 * k + 1];
                }
            }
        }
    }
    if (tid == 0)
    {
        a.status[q] = status;
        a.cost[q] = (status == ST_FOUND) ? S.goal_g : -1.0;
        a.path_len[q] = out_len;
        if (a.stats)
        {
            a.stats[6 * q + 0] = S.iteration;
            a.stats[6 * q + 1] = S.n_prim;
            a.stats[6 * q + 2] = S.n_shots;
            a.stats[6 * q + 3] = S.n_coll;
            a.stats[6 * q + 4] = S.bfs.cells;
            a.stats[6 * q + 5] = S.n_nodes;
        }
    }
    team_sync_all<NW>();
}

template <int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) plan_team_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p,
                                                                   const __grid_constant__ PlanBatchArgs a)
{
    __shared__ TeamShared<NW> S;
    // a workspace is taken with the first ticket: the CTAs of a pass that finds its hand-off list (nearly) empty leave at once
    // instead of waiting for workspaces other batches hold
    int slot = -1;
    char *ws = nullptr;
    uint32_t gen = 0;
    // work list: what is left of the squad pass's migration list (entries nobody adopted), then the team list
    const int mig_first = a.mig_n ? (int)*a.mig_taken : 0, mig_rem = a.mig_n ? (int)*a.mig_n - mig_first : 0;
    const int total = mig_rem + (a.n_todo_dev ? (int)*a.n_todo_dev : (a.todo ? a.n_todo : a.nq));
    while (true)
    {
        team_sync_all<NW>();
        if (team_tid<NW>() == 0)
            S.q = (int)(a.shared_counter ? atomicAdd_system(a.counter, 1u) : atomicAdd(a.counter, 1u));
        team_sync_all<NW>();
        const int t = S.q;
        if (t >= total)
            break;
        if (slot < 0)
        {
            if (team_tid<NW>() == 0)
            {
                S.ws_slot = pool_acquire(a.pool, (int)blockIdx.x);
                S.ws_gen = a.pool.gen[S.ws_slot];
            }
            team_sync_all<NW>();
            slot = S.ws_slot;
            ws = a.pool.base + (size_t)slot * a.L.stride;
            gen = S.ws_gen;
        }
        int q, ws_small;
        if (t < mig_rem)
            q = a.mig[mig_first + t], ws_small = a.mig_ws[mig_first + t];
        else
        {
            const int u = t - mig_rem;
            q = a.todo ? a.todo[u] : u;
            ws_small = a.todo_ws ? a.todo_ws[u] : -1;
        }
        ++gen;
        plan_team_query<NW, false>(m, p, a, q, ws, gen, S, nullptr, nullptr, ws_small);
    }
    team_sync_all<NW>();
    if (slot >= 0 && team_tid<NW>() == 0)
        pool_release(a.pool, slot, gen);
}

// Two teams per CTA, one CTA per SM (see DuoSync): blockIdx.x * 2 + team index = workspace; each team pulls its own queries.
template <int NW>
__global__ void __launch_bounds__(NW * 64, 1) plan_duo_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p,
                                                               const __grid_constant__ PlanBatchArgs a)
{
    extern __shared__ __align__(16) unsigned char duo_smem[];
    constexpr size_t kTeamBytes = (sizeof(TeamShared<NW>) + 15) / 16 * 16;
    const int team = team_index<NW>();
    TeamShared<NW> &S = *reinterpret_cast<TeamShared<NW> *>(duo_smem + team * kTeamBytes);
    DuoSync *D = reinterpret_cast<DuoSync *>(duo_smem + 2 * kTeamBytes);
    if (threadIdx.x == 0)
    {
        D->count[0] = D->count[1] = 0;
        D->active[0] = D->active[1] = 0;
    }
    __syncthreads();
    if (team_tid<NW>() == 0)
    {
        S.ws_slot = pool_acquire(a.pool, (int)blockIdx.x * 2 + team);
        S.ws_gen = a.pool.gen[S.ws_slot];
    }
    team_sync_all<NW>();
    const int slot = S.ws_slot;
    char *ws = a.pool.base + (size_t)slot * a.L.stride;
    uint32_t gen = S.ws_gen;
    const int total = a.todo ? a.n_todo : a.nq;
    while (true)
    {
        team_sync_all<NW>();
        if (team_tid<NW>() == 0)
            S.q = (int)(a.shared_counter ? atomicAdd_system(a.counter, 1u) : atomicAdd(a.counter, 1u));
        team_sync_all<NW>();
        const int t = S.q;
        if (t >= total)
            break;
        int q = a.todo ? a.todo[t] : t;
        ++gen;
        plan_team_query<NW, false>(m, p, a, q, ws, gen, S, nullptr, D);
    }
    team_sync_all<NW>();
    if (team_tid<NW>() == 0)
        pool_release(a.pool, slot, gen);
}
template <int NW>
constexpr size_t duo_smem_bytes() { return 2 * ((sizeof(TeamShared<NW>) + 15) / 16 * 16) + sizeof(DuoSync) + 16; }

// Cluster mode (see "cluster mode" above): CTA pair per team, blockIdx.x >> 1 = team, cluster rank = role.
template <int NW, int MINB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NW * 32, MINB)
    plan_cluster_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p, const __grid_constant__ PlanBatchArgs a)
{
    // the search CTA uses S, the heuristic CTA uses H: same storage (each CTA of the pair has its own shared memory)
    __shared__ union ClusterShared
    {
        TeamShared<NW> S;
        HeurShared H;
    } U;
    TeamShared<NW> &S = U.S;
    HeurShared &H = U.H;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    // both CTAs must be running before either touches the other's shared memory
    cluster_arrive();
    cluster_wait();
    if (rank == 1)
    {
        heur_cta_loop<NW>(m, p, H, cluster_peer(&S.c_h[0], 0));
        return;
    }
    HeurShared *Hpeer = cluster_peer(&H, 1);
    if (team_tid<NW>() == 0)
    {
        S.ws_slot = pool_acquire(a.pool, (int)(blockIdx.x >> 1));
        S.ws_gen = a.pool.gen[S.ws_slot];
    }
    team_sync_all<NW>();
    const int slot = S.ws_slot;
    char *ws = a.pool.base + (size_t)slot * a.L.stride;
    uint32_t gen = S.ws_gen;
    const int total = a.todo ? a.n_todo : a.nq;
    while (true)
    {
        team_sync_all<NW>();
        if (team_tid<NW>() == 0)
            S.q = (int)(a.shared_counter ? atomicAdd_system(a.counter, 1u) : atomicAdd(a.counter, 1u));
        team_sync_all<NW>();
        const int t = S.q;
        if (t >= total)
            break;
        int q = a.todo ? a.todo[t] : t;
        ++gen;
        plan_team_query<NW, true>(m, p, a, q, ws, gen, S, Hpeer);
    }
    team_sync_all<NW>();
    if (team_tid<NW>() == 0)
        pool_release(a.pool, slot, gen);
    // release the heuristic CTA: the last X of this cluster carries the exit command
    if (team_tid<NW>() == 0)
        Hpeer->cmd = 1;
    cluster_arrive();
    cluster_wait();
}

} // namespace mpros
