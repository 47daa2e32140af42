// Whole-query planner, one WARP per query ("solo"): the throughput form of the search.
//
// plan_team.cuh spends a whole CTA on one query to shorten the dependent chain pop -> expand -> insert (6.5-10 us per
// iteration): that is what the TAIL of a batch needs (a query that runs into the 100 000-iteration limit), but the eight
// warps of a team wait for each other most of the time (issue slots 22 % busy), so the device finishes only ~28 M
// iterations/s. Most queries are short (median 326 iterations on the synthetic set; 99 % below 23 000): for them latency
// does not matter and a warp per query - 32 independent queries per SM hiding each other's FP64 chains and memory
// latency, like expand_batch_kernel does for single expansions - finishes several times more iterations per second.
//
// Division of labour (plan.cu: plan_batch_dev_impl): every query starts here with a SMALL workspace (node pool of
// a.Ls.node_cap records). A query that outgrows its pool is appended to a hand-off list in device memory and run again
// by the team kernel, which is launched behind this one on the same stream and reads the list's length from the device
// (no host round trip between the two).
//
// Reference semantics are those of the team kernel: one popped node per step in std::multimap order ((f, insertion id)),
// children in steer x direction order, strict '>' prune, first successful shot ends the search
// (hybrid_a_star.cpp:682-777). The arithmetic is the same device code (primitive_end_pose, node_cost, ompl_formula,
// rs_find_optimal_quad, rs_gen_traj_regs, the fixed-point footprint walk behind footprint_preclass), so status, iteration
// count, cost and path are bit-identical to the team kernel's (tests/test_gpu_synthetic_fullsize.py).
#pragma once
#include "plan_team.cuh"

namespace mpros
{

constexpr int kSoloHeapTop = 1 + 32; // heap levels 0-1 live in shared memory (396 bytes per warp)

struct SoloShared
{
    double sx, sy, syaw, gx, gy, gyaw;
    TeamBfs bfs;
    unsigned long long hkey[kSoloHeapTop];
    uint32_t hid[kSoloHeapTop];
};

// ---- on-demand goal wavefront, one warp (same window / contiguous-row scheme as team_bfs_*) ------------------------------
__device__ __noinline__ void solo_bfs_init_rows(const MapView &m, TeamBfs &b, int ra, int rb)
{
    const int lane = threadIdx.x & 31;
    ra = max(ra, b.r0);
    rb = min(rb, b.r0 + b.rows - 1);
    const int ilo = b.ilo, ihi = b.ihi;
    if (ra >= ilo && rb <= ihi)
        return;
    const int wwords = b.wwords;
    const int lo_n = (ilo > ihi) ? (rb - ra + 1) : max(0, ilo - ra), hi_n = (ilo > ihi) ? 0 : max(0, rb - ihi);
    for (int k = lane; k < (lo_n + hi_n) * wwords; k += 32)
    {
        int rr = k / wwords, w = k % wwords;
        int r = rr < lo_n ? ra + rr : ihi + 1 + (rr - lo_n);
        size_t o = (size_t)(r - b.r0) * wwords + w;
        b.blocked[o] = __ldg(m.ds_bits + (size_t)r * m.pitch + b.w0 + w);
        b.f0[o] = 0;
        b.f1[o] = 0;
    }
    __syncwarp();
    if (lane == 0)
    {
        b.ilo = (ilo > ihi) ? ra : min(ilo, ra);
        b.ihi = (ilo > ihi) ? rb : max(ihi, rb);
    }
    __syncwarp();
}

__device__ __noinline__ void solo_bfs_start(const MapView &m, TeamBfs &b, int gi, int gj, uint32_t tag)
{
    const int lane = threadIdx.x & 31;
    if (lane == 0)
    {
        const int wpr = (m.W + 31) >> 5;
        b.tag = tag;
        b.gi = gi;
        b.gj = gj;
        b.r0 = max(0, gi - (kHugeInt - 1)); // only cells within 999 steps can ever hold a value < 1000
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
    __syncwarp();
    solo_bfs_init_rows(m, b, gi - 2, gi + 2);
    if (lane == 0)
    {
        size_t o = (size_t)(gi - b.r0) * b.wwords + ((gj >> 5) - b.w0);
        b.f0[o] = 1u << (gj & 31);
        b.blocked[o] |= 1u << (gj & 31);
    }
    __syncwarp();
}

// One level of the wavefront. Returns the number of words scanned (the caller's measure of the work done). The scan takes
// four words per lane at a time with all their loads issued before the first result is used: the loop is a chain of global
// memory round trips otherwise (a level 50 cells out is ~400 words: 13 dependent trips, ~20 k cycles for ONE level).
__device__ __noinline__ int solo_bfs_step(const MapView &m, TeamBfs &b)
{
    const int lane = threadIdx.x & 31;
    const int level = b.level + 1;
    if (level >= kHugeInt)
    {
        __syncwarp();
        if (lane == 0)
            b.done = 1;
        __syncwarp();
        return 0;
    }
    const int wr1 = b.r0 + b.rows - 1, ww1 = b.w0 + b.wwords - 1;
    const int sr0 = max(b.r0, b.br0 - 1), sr1 = min(wr1, b.br1 + 1);
    const int sw0 = max(b.w0, b.bw0 - 1), sw1 = min(ww1, b.bw1 + 1);
    solo_bfs_init_rows(m, b, sr0 - 2, sr1 + 2);
    const int nw = sw1 - sw0 + 1;
    const int total = (sr1 - sr0 + 1) * nw;
    const uint32_t *__restrict__ fprev = (level & 1) ? b.f0 : b.f1;
    uint32_t *__restrict__ fnext = (level & 1) ? b.f1 : b.f0;
    uint32_t *__restrict__ blocked = b.blocked;
    const int wpr = (m.W + 31) >> 5;
    int nr0 = 1 << 30, nr1 = -1, nw0 = 1 << 30, nw1 = -1, found = 0;
    const int rows = b.rows, wwords = b.wwords, r0 = b.r0, w0 = b.w0;
    constexpr int U = 4;
    for (int k0 = lane; k0 < total; k0 += 32 * U)
    {
        uint32_t c[U], lf[U], rt[U], up[U], dn[U], bl[U];
        int rr[U], ww[U];
        size_t oo[U];
#pragma unroll
        for (int t = 0; t < U; ++t)
        {
            const int k = k0 + 32 * t;
            const bool act = k < total;
            const int kk = act ? k : 0;
            const int r = sr0 + kk / nw, w = sw0 + kk % nw;
            const int lr = r - r0, lw = w - w0;
            const size_t o = (size_t)lr * wwords + lw;
            rr[t] = act ? r : -1, ww[t] = w, oo[t] = o;
            c[t] = act ? fprev[o] : 0u;
            lf[t] = (act && lw > 0) ? fprev[o - 1] : 0u;
            rt[t] = (act && lw + 1 < wwords) ? fprev[o + 1] : 0u;
            up[t] = (act && lr > 0) ? fprev[o - wwords] : 0u;
            dn[t] = (act && lr + 1 < rows) ? fprev[o + wwords] : 0u;
            bl[t] = act ? blocked[o] : 0xffffffffu;
        }
#pragma unroll
        for (int t = 0; t < U; ++t)
        {
            if (rr[t] < 0)
                continue;
            const uint32_t nb = (c[t] << 1) | (c[t] >> 1) | (lf[t] >> 31) | (rt[t] << 31) | up[t] | dn[t];
            const uint32_t valid = (ww[t] == wpr - 1 && (m.W & 31)) ? ((1u << (m.W & 31)) - 1u) : 0xffffffffu;
            uint32_t nbits = nb & ~bl[t] & valid;
            fnext[oo[t]] = nbits;
            if (nbits)
            {
                blocked[oo[t]] = bl[t] | nbits;
                found += __popc(nbits);
                nr0 = min(nr0, rr[t]), nr1 = max(nr1, rr[t]), nw0 = min(nw0, ww[t]), nw1 = max(nw1, ww[t]);
                uint16_t *drow = b.dist + oo[t] * 32;
                while (nbits)
                {
                    const int bit = __ffs(nbits) - 1;
                    nbits &= nbits - 1;
                    drow[bit] = (uint16_t)level;
                }
            }
        }
    }
    nr0 = __reduce_min_sync(kFull, nr0);
    nr1 = __reduce_max_sync(kFull, nr1);
    nw0 = __reduce_min_sync(kFull, nw0);
    nw1 = __reduce_max_sync(kFull, nw1);
    found = __reduce_add_sync(kFull, found);
    __syncwarp();
    if (lane == 0)
    {
        b.level = level;
        b.cells += found;
        if (found == 0)
            b.done = 1;
        else
        {
            b.br0 = min(b.br0, nr0), b.br1 = max(b.br1, nr1), b.bw0 = min(b.bw0, nw0), b.bw1 = max(b.bw1, nw1);
        }
    }
    __syncwarp();
    return total;
}

// goal-map value per lane (want = false: no request); advances the wavefront until every requested cell is final
__device__ __forceinline__ int solo_h1_lookup(const MapView &m, TeamBfs &b, bool want, int i, int j)
{
    int v = want ? team_bfs_peek(m, b, i, j) : 0;
    while (__any_sync(kFull, v < 0))
    {
        solo_bfs_step(m, b);
        if (v < 0)
            v = team_bfs_peek(m, b, i, j);
    }
    return v;
}

// ---- heuristics (hybrid_a_star.cpp:248-264) of the children flagged in needmask -------------------------------------------
// Lane c < P passes child c: pose, sin / cos of its yaw, sin / cos of (goal yaw - yaw), goal-map value. Work layout
// lane = 4 * (c - base) + q, q = symmetry image, eight children per round; the same two rounds and the same exact skip as
// plan_team.cuh: heur_rounds + heur_combine (round A: the two CSC words; round B, only when rho * min(CSC) > h1 for some
// needed child of the group: the other nine). Returns h of child c on lane c.
__device__ __noinline__ double solo_heuristics(const MapView &m, const PlannerView &p, unsigned needmask, int P, double gx, double gy, double gyaw,
                                               double x, double y, double yaw, double sn, double cs, double sphi, double cphi, int h1)
{
    const int lane = threadIdx.x & 31;
    double hval = 0;
    for (int base = 0; base < P; base += 8)
    {
        if (((needmask >> base) & 0xffu) == 0)
            continue;
        const int c = base + (lane >> 2), q = lane & 3;
        const bool live = c < P && ((needmask >> c) & 1u);
        const int src = c < 32 ? c : 0;
        const double cx = __shfl_sync(kFull, x, src), cy = __shfl_sync(kFull, y, src), cyaw = __shfl_sync(kFull, yaw, src);
        const double csn = __shfl_sync(kFull, sn, src), ccs = __shfl_sync(kFull, cs, src);
        const double csphi = __shfl_sync(kFull, sphi, src), ccphi = __shfl_sync(kFull, cphi, src);
        const int ch1 = __shfl_sync(kFull, h1, src);
        // symmetry image q of the normalised goal (plan_team.cuh, closed-set warp phase C)
        const double dx = gx - cx, dy = gy - cy, dth = gyaw - cyaw;
        const double inx = dm::div_(ccs * dx + csn * dy, p.min_r), iny = dm::div_(-csn * dx + ccs * dy, p.min_r);
        const double nxb = inx * ccphi + iny * csphi, nyb = inx * csphi - iny * ccphi;
        const bool fx = (q == 1 || q == 3), fy = (q == 2 || q == 3), fp = (q == 1 || q == 2);
        const double X = fx ? -inx : inx, Y = fy ? -iny : iny, PHI = fp ? -dth : dth, XB = fx ? -nxb : nxb, YB = fy ? -nyb : nyb;
        const double SP = fp ? -csphi : csphi, CP = ccphi;
        double plain = DBL_MAX, ccsc = DBL_MAX, ccscc = DBL_MAX;
        if (live)
            plain = fmin(ompl_formula(0, X, Y, PHI, XB, YB, SP, CP), ompl_formula(1, X, Y, PHI, XB, YB, SP, CP));
        plain = fmin(plain, __shfl_xor_sync(kFull, plain, 1));
        plain = fmin(plain, __shfl_xor_sync(kFull, plain, 2));
        const double h1v = (double)ch1 * m.res; // getGoalCost: cost_ * map_resolution_
        const bool ok = !live || (p.min_r * plain <= h1v);
        if (!__all_sync(kFull, ok))
        {
#pragma unroll 1
            for (int k = 2; k <= 10; ++k)
            {
                double Lf = live ? ompl_formula(k, X, Y, PHI, XB, YB, SP, CP) : DBL_MAX;
                Lf = fmin(Lf, __shfl_xor_sync(kFull, Lf, 1));
                Lf = fmin(Lf, __shfl_xor_sync(kFull, Lf, 2));
                if (k < 6)
                    plain = fmin(plain, Lf);
                else if (k < 10)
                    ccsc = fmin(ccsc, Lf);
                else
                    ccscc = Lf;
            }
        }
        const double h = fmax(h1v, ompl_combine(p.min_r, plain, ccsc, ccscc)) * (1 + 1e-3);
        const int owner = lane - base;
        const double back = __shfl_sync(kFull, h, (owner >= 0 && owner < 8) ? owner * 4 : 0);
        if (owner >= 0 && owner < 8)
            hval = back;
    }
    return hval;
}

// ---- pathInCollision of up to 32 sampled poses held one per lane: index of the first colliding pose, -1 when none ------
// (footprint_preclass decides most poses by their own cell; the undecided ones before the first sure hit get the warp-wide
// probe walk in order, so the result is the index the reference's sequential loop stops at)
__device__ __forceinline__ int solo_first_hit32(const MapView &m, const PlannerView &p, const PlanBatchArgs &a, bool pose, double px, double py,
                                                double pyaw)
{
    const int cls = pose ? footprint_preclass(m, p, a.unsafe_tw, a.center_probe, px, py, pyaw) : 0;
    unsigned hits = __ballot_sync(kFull, cls == 1);
    unsigned walk = __ballot_sync(kFull, cls == 2);
    if (walk)
    {
        double sn = 0, cs = 1;
        if (cls == 2)
            dm::sincos_(pyaw, &sn, &cs);
        while (walk)
        {
            const int t = __ffs(walk) - 1;
            if (hits && t > __ffs(hits) - 1)
                break; // an earlier pose already collides
            walk &= walk - 1;
            if (warp_footprint_collision(m, p, __shfl_sync(kFull, px, t), __shfl_sync(kFull, py, t), __shfl_sync(kFull, sn, t),
                                         __shfl_sync(kFull, cs, t)))
            {
                hits |= 1u << t;
                break;
            }
        }
    }
    return hits ? __ffs(hits) - 1 : -1;
}

// one-lane GenTraj of a candidate into the warp's scratch (trajectories of more than 32 poses; rare inside the shot range)
__device__ __noinline__ int solo_traj_long(const RsConfig &rs, const RsWord &w, double x0, double y0, double yaw0, double *scratch)
{
    const int lane = threadIdx.x & 31;
    int np = 0;
    if (lane == 0)
    {
        RsPath path;
        path.n = w.n;
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (k < w.n)
                rs_word_seg(w, k, path.s[k].len, path.s[k].steer, path.s[k].dir);
        np = rs_gen_traj(rs, x0, y0, yaw0, path, scratch, scratch + 3 * kTrajCap, kTrajCap);
    }
    np = __shfl_sync(kFull, np, 0);
    __syncwarp();
    return np;
}

// genRSPath (hybrid_a_star.cpp:853-900) from node cn: Reeds-Shepp solve, sampled trajectory, footprint tests. Returns the
// number of poses when the trajectory is free (>= 0), -1 when it collides; cost = RS cost; checked = footprint tests the
// reference's loop would have run (statistics). Out of line: a few shots per query.
__device__ __noinline__ int solo_shot(const MapView &m, const PlannerView &p, const PlanBatchArgs &a, const SoloShared &S, const PlanNode &cn,
                                      double *scratch, double &cost, int &checked)
{
    const int lane = threadIdx.x & 31;
    const RsWord w = rs_find_optimal_quad(a.rs, cn.x, cn.y, cn.yaw, (int)cn.dir, cn.steer, S.gx, S.gy, S.gyaw, cost);
    double px, py, pyaw;
    int pst, pdr;
    int np = rs_gen_traj_regs(a.rs, cn.x, cn.y, cn.yaw, w, px, py, pyaw, pst, pdr);
    int first = -1;
    if (np <= 32)
        first = solo_first_hit32(m, p, a, lane < np, px, py, pyaw);
    else
    {
        np = solo_traj_long(a.rs, w, cn.x, cn.y, cn.yaw, scratch);
        if (np > kTrajCap)
            first = 0; // longer than the buffer cannot happen inside the shot range; blocked rather than truncated
        for (int k0 = 0; k0 < np && k0 < kTrajCap && first < 0; k0 += 32)
        {
            const int k = k0 + lane;
            const bool pose = k < np;
            const int f = solo_first_hit32(m, p, a, pose, pose ? scratch[3 * k] : 0.0, pose ? scratch[3 * k + 1] : 0.0, pose ? scratch[3 * k + 2] : 0.0);
            if (f >= 0)
                first = k0 + f;
        }
    }
    checked = first >= 0 ? first + 1 : np;
    return first >= 0 ? -1 : np;
}

// raw path_ of setPath (hybrid_a_star.cpp:779-811): start, expanded nodes, then the RS goal path, re-generated from the
// stored parent (deterministic). Returns the path length.
__device__ __noinline__ int solo_write_path(const PlanBatchArgs &a, const SoloShared &S, const PlanNode *nodes, uint32_t goal_parent, int q,
                                            double *scratch)
{
    const int lane = threadIdx.x & 31;
    const PlanNode gp = nodes[goal_parent];
    double cost;
    const RsWord w = rs_find_optimal_quad(a.rs, gp.x, gp.y, gp.yaw, (int)gp.dir, gp.steer, S.gx, S.gy, S.gyaw, cost);
    double px, py, pyaw;
    int pst, pdr;
    int np = rs_gen_traj_regs(a.rs, gp.x, gp.y, gp.yaw, w, px, py, pyaw, pst, pdr);
    const bool in_regs = np <= 32;
    if (!in_regs)
        np = solo_traj_long(a.rs, w, gp.x, gp.y, gp.yaw, scratch);
    int chain = 0;
    if (lane == 0)
    {
        uint32_t c = goal_parent;
        while (nodes[c].parent != kNoNode)
        {
            ++chain;
            c = nodes[c].parent;
        }
    }
    chain = __shfl_sync(kFull, chain, 0);
    if (a.paths && a.path_cap > 0)
    {
        double *dst = a.paths + (size_t)q * a.path_cap * 5;
        // direction of the first trajectory pose (goal_path_.front().back())
        const double first_dir = in_regs ? (double)__shfl_sync(kFull, pdr, 0) : scratch[3 * kTrajCap + 1];
        if (lane == 0)
        {
            uint32_t c = goal_parent;
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
            const double start_dir = chain > 1 ? last_dir : (np > 0 ? first_dir : 1.0);
            dst[0] = S.sx, dst[1] = S.sy, dst[2] = S.syaw, dst[3] = 0.0, dst[4] = start_dir;
        }
        for (int k = lane; k < np && k < kTrajCap; k += 32)
        {
            const int pos = 1 + chain + k;
            if (pos < a.path_cap)
            {
                double *o = dst + 5 * pos;
                if (in_regs)
                    o[0] = px, o[1] = py, o[2] = pyaw, o[3] = (double)pst * a.rs.steer_angle, o[4] = (double)pdr;
                else
                    o[0] = scratch[3 * k], o[1] = scratch[3 * k + 1], o[2] = scratch[3 * k + 2], o[3] = scratch[3 * kTrajCap + 2 * k],
                    o[4] = scratch[3 * kTrajCap + 2 * k + 1];
            }
        }
    }
    return 1 + chain + np;
}

// children in primitive order, one at a time (lane 0): the fallback when two children of one expansion touch the same
// closed-set slot. Returns false when the node pool is full.
__device__ __noinline__ bool solo_insert_sequential(PlanNode *nodes, const ClosedSet &closed, OpenHeap &heap, unsigned needmask, uint32_t cur,
                                                    uint32_t &n_nodes, int node_cap, unsigned long long idx, double g, double h, double nx, double ny,
                                                    double nyaw, double ns, double nc, double steer, double direc)
{
    const int lane = threadIdx.x & 31;
    bool full = false;
    while (needmask && !full)
    {
        const int c = __ffs(needmask) - 1;
        needmask &= needmask - 1;
        const unsigned long long tidx = __shfl_sync(kFull, idx, c);
        const double tg = __shfl_sync(kFull, g, c), th = __shfl_sync(kFull, h, c);
        const double tx = __shfl_sync(kFull, nx, c), ty = __shfl_sync(kFull, ny, c), tyaw = __shfl_sync(kFull, nyaw, c);
        const double ts = __shfl_sync(kFull, ns, c), tc = __shfl_sync(kFull, nc, c);
        const double tsteer = __shfl_sync(kFull, steer, c), tdir = __shfl_sync(kFull, direc, c);
        int inserted = 0;
        if (lane == 0)
        {
            bool found;
            double oc;
            uint32_t on;
            const uint32_t sl = closed_find(closed, tidx, found, oc, on);
            if (oc > tg)
            {
                if (n_nodes >= (uint32_t)node_cap)
                    inserted = -1;
                else
                {
                    if (found && on < kGoalNode)
                        nodes[on].closed = 1;
                    closed_store(closed, sl, tidx, tg, n_nodes);
                    PlanNode nn;
                    nn.x = tx, nn.y = ty, nn.yaw = tyaw, nn.steer = tsteer, nn.g = tg, nn.s = ts, nn.c = tc, nn.parent = cur;
                    nn.dir = (int8_t)(tdir > 0 ? 1 : -1), nn.closed = 0, nn.pad = 0;
                    nodes[n_nodes] = nn;
                    heap_push_lane0(heap, (unsigned long long)__double_as_longlong(tg + th), n_nodes);
                    inserted = 1;
                }
            }
        }
        inserted = __shfl_sync(kFull, inserted, 0);
        if (inserted < 0)
            full = true;
        else if (inserted)
        {
            heap.size++;
            n_nodes++;
        }
        __syncwarp();
    }
    return !full;
}

__device__ void plan_solo_query(const MapView &m, const PlannerView &p, const PlanBatchArgs &a, int q, char *ws, uint32_t tag, SoloShared &S)
{
    const int lane = threadIdx.x & 31;
    const WarpWorkspaceLayout &L = a.Ls;
    PlanNode *nodes = reinterpret_cast<PlanNode *>(ws + L.node_off);
    double *scratch = reinterpret_cast<double *>(ws + L.traj_off);
    OpenHeap heap;
    heap.key = reinterpret_cast<unsigned long long *>(ws + L.hkey_off);
    heap.id = reinterpret_cast<uint32_t *>(ws + L.hid_off);
    heap.skey = S.hkey;
    heap.sid = S.hid;
    heap.scap = kSoloHeapTop;
    heap.size = 0;
    heap.cap = L.node_cap;
    ClosedSet closed;
    closed.slot = reinterpret_cast<ClosedSlot *>(ws + L.table_off);
    closed.mask = (uint32_t)L.table_slots - 1;
    closed.tag = tag;
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

    // ---- initialize (hybrid_a_star.cpp:52-118) ----
    bool init_ok;
    double s_start, c_start;
    {
        dm::sincos_(S.syaw, &s_start, &c_start);
        init_ok = !warp_footprint_collision(m, p, S.sx, S.sy, s_start, c_start); // "Start pose in collision"
        if (init_ok)
        {
            double s, c;
            dm::sincos_(S.gyaw, &s, &c);
            init_ok = !warp_footprint_collision(m, p, S.gx, S.gy, s, c); // "Goal pose in collision"
        }
    }
    int si = -1, sj = -1, gi = -1, gj = -1;
    if (init_ok)
    {
        const bool s_in = pos_to_index_ds(m, S.sx, S.sy, si, sj);
        const bool g_in = pos_to_index_ds(m, S.gx, S.gy, gi, gj);
        init_ok = s_in && g_in; // goal outside the map: no goal map (nav_map.cpp:297-302)
    }
    uint32_t n_nodes = 0;
    int iteration = 0, n_shots = 0;
    long long n_exp = 0, n_coll = 0;
    bool goal_found = false, limit = false, overflow = false;
    uint32_t goal_parent = kNoNode;
    double goal_g = -1.0;
    const int P = 2 * p.n_steer;
    if (init_ok)
    {
        solo_bfs_start(m, S.bfs, gi, gj, tag);
        // start node: steer 0, direction 1, g = 0 (:83, :211-215)
        const int h1s = solo_h1_lookup(m, S.bfs, lane == 0, si, sj);
        double sphi, cphi;
        dm::sincos_(S.gyaw - S.syaw, &sphi, &cphi);
        const double h0 = solo_heuristics(m, p, 1u, 1, S.gx, S.gy, S.gyaw, S.sx, S.sy, S.syaw, s_start, c_start, sphi, cphi, h1s);
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
            const unsigned long long gidx = cal_index(m, p, gi, gj, yaw_to_idx(p, S.gyaw), 1.0);
            const uint32_t s1 = closed_find(closed, gidx, found, og, on);
            closed_store(closed, s1, gidx, 10000.0, kGoalNode);
            heap_push_lane0(heap, (unsigned long long)__double_as_longlong(0.0 + h0), 0);
        }
        heap.size = 1;
        n_nodes = 1;
        __syncwarp();

        // ---- Plan() (hybrid_a_star.cpp:682-777) ----
        while (heap.size > 0)
        {
            if (iteration > p.max_iteration)
            {
                limit = true;
                break;
            }
            unsigned long long fk;
            const uint32_t cur = heap_pop_warp(heap, fk);
            const PlanNode cn = nodes[cur];
            if (cn.closed)
                continue;
            // -- genRSPath (:853-900)
            bool shot_ok = false;
            {
                const double ddx = cn.x - S.gx, ddy = cn.y - S.gy;
                if (ddx * ddx + ddy * ddy < p.rs_range * p.rs_range)
                {
                    double rs_cost;
                    int checked;
                    const int np = solo_shot(m, p, a, S, cn, scratch, rs_cost, checked);
                    ++n_shots;
                    n_coll += checked;
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
            }
            if (shot_ok)
            {
                goal_found = true;
                if (!a.optimal)
                    break; // first successful shot ends the search (:725-728)
            }
            else
            {
                // -- expandNode (:304-402), lane = primitive. Lanes P..2P-1 shadow the children once more for the second
                // sin/cos pair (goal yaw - child yaw, for the symmetry images), so both come out of one call.
                ++n_exp;
                n_coll += P;
                const bool two_in_one = 2 * P <= 32;
                const int c = lane < P ? lane : ((two_in_one && lane < 2 * P) ? lane - P : 0);
                double nx, ny, nyaw, steer, direc;
                primitive_end_pose(p, c, cn.x, cn.y, cn.yaw, cn.s, cn.c, nx, ny, nyaw, steer, direc);
                double ns = 0, nc = 1, sphi = 0, cphi = 1;
                if (two_in_one)
                {
                    if (lane < 2 * P)
                        dm::sincos_(lane < P ? nyaw : S.gyaw - nyaw, &ns, &nc);
                    sphi = __shfl_sync(kFull, ns, (lane + P) & 31);
                    cphi = __shfl_sync(kFull, nc, (lane + P) & 31);
                }
                else if (lane < P)
                {
                    dm::sincos_(nyaw, &ns, &nc);
                    dm::sincos_(S.gyaw - nyaw, &sphi, &cphi);
                }
                // footprint tests: most end poses are decided by their own cell; the others get the warp-wide probe walk
                const int cls = lane < P ? footprint_preclass(m, p, a.unsafe_tw, a.center_probe, nx, ny, nyaw) : 1;
                unsigned free_mask = __ballot_sync(kFull, cls == 0);
                unsigned walk_mask = __ballot_sync(kFull, cls == 2);
                while (walk_mask)
                {
                    const int t = __ffs(walk_mask) - 1;
                    walk_mask &= walk_mask - 1;
                    if (!warp_footprint_collision(m, p, __shfl_sync(kFull, nx, t), __shfl_sync(kFull, ny, t), __shfl_sync(kFull, ns, t),
                                                  __shfl_sync(kFull, nc, t)))
                        free_mask |= 1u << t;
                }
                // node cell, closed-set entry, g (lane = child)
                bool need = false, pr_found = false;
                double g = 0, pr_old_cost = 0;
                uint32_t pr_old = kNoNode, pr_slot = 0xffffffffu - lane;
                unsigned long long idx = 0;
                int ci = -1, cj = -1;
                if (lane < P && ((free_mask >> lane) & 1u) && pos_to_index_ds(m, nx, ny, ci, cj))
                {
                    // (a cell outside the map cannot pass the footprint test; the reference would index outside its arrays)
                    const float voro_f = __ldg(m.voronoi + (size_t)ci * m.W + cj);
                    idx = cal_index(m, p, ci, cj, yaw_to_idx(p, nyaw), direc);
                    pr_slot = closed_find(closed, idx, pr_found, pr_old_cost, pr_old);
                    g = node_cost(p, cn.g, (double)cn.dir, cn.steer, (double)voro_f, steer, direc);
                    need = pr_old_cost > g; // an earlier sibling on the same slot may still beat it; resolved at insert
                }
                const unsigned needmask = __ballot_sync(kFull, need);
                if (needmask)
                {
                    const int h1 = solo_h1_lookup(m, S.bfs, need, ci, cj);
                    const double h = solo_heuristics(m, p, needmask, P, S.gx, S.gy, S.gyaw, nx, ny, nyaw, ns, nc, sphi, cphi, h1);
                    // closed-set compare / insert in primitive order. Common case: the children touch distinct slots -> one lane
                    // per child; children that collide on a slot (same index, or two new keys probing to the same empty slot)
                    // take the strictly sequential walk.
                    const unsigned same = __match_any_sync(kFull, need ? pr_slot : 0xffffffffu - lane);
                    const bool conflict = __any_sync(kFull, need && (same & (same - 1)) != 0);
                    if (!conflict)
                    {
                        const int cnt = __popc(needmask);
                        if (n_nodes + cnt > (uint32_t)L.node_cap)
                            overflow = true;
                        else
                        {
                            const int rank = __popc(needmask & ((1u << lane) - 1u));
                            const uint32_t id = n_nodes + rank;
                            const unsigned long long key = (unsigned long long)__double_as_longlong(g + h);
                            // all new entries appended at once when none of them beats its parent in the heap (children mostly
                            // cost more than what is already open); otherwise one by one
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
                    else if (!solo_insert_sequential(nodes, closed, heap, needmask, cur, n_nodes, L.node_cap, idx, g, h, nx, ny, nyaw, ns, nc, steer,
                                                     direc))
                        overflow = true;
                    __syncwarp();
                }
                if (overflow)
                    break;
            }
            if (a.optimal && dm::hypot_(S.gx - cn.x, S.gy - cn.y) <= 1.0) // stop_radius (hybrid_a_star.h:345)
                break;
            ++iteration; // every break of the reference's loop precedes ++iteration (:725-745)
        }
    }

    if (overflow && a.handoff)
    {
        // the query outgrew the solo workspace: the team kernel runs it again with a full-size one
        if (lane == 0)
        {
            const unsigned at = atomicAdd(a.handoff_n, 1u);
            a.handoff[at] = q;
            if (a.handoff_ws)
                a.handoff_ws[at] = -1;
            a.status[q] = ST_OVERFLOW;
        }
        __syncwarp();
        return;
    }
    int status = ST_NOT_INIT;
    if (init_ok)
        status = overflow ? ST_OVERFLOW : goal_found ? ST_FOUND : limit ? ST_LIMIT : ST_OPEN_EMPTY;
    int out_len = 0;
    if (status == ST_FOUND)
        out_len = solo_write_path(a, S, nodes, goal_parent, q, scratch);
    if (lane == 0)
    {
        a.status[q] = status;
        a.cost[q] = (status == ST_FOUND) ? goal_g : -1.0;
        a.path_len[q] = out_len;
        if (a.stats)
        {
            a.stats[6 * q + 0] = iteration;
            a.stats[6 * q + 1] = n_exp * P;
            a.stats[6 * q + 2] = n_shots;
            a.stats[6 * q + 3] = n_coll;
            a.stats[6 * q + 4] = S.bfs.cells;
            a.stats[6 * q + 5] = n_nodes;
        }
    }
    __syncwarp();
}

#ifndef MPROS_SOLO_WARPS
#define MPROS_SOLO_WARPS 1 // warps per CTA: with one, every warp leaves the SM as soon as the tickets run out
#endif
#ifndef MPROS_SOLO_MINB
#define MPROS_SOLO_MINB 24 // resident CTAs per SM asked of the compiler (register budget 65536 / (32 * warps * MINB))
#endif
template <int W, int MINB>
__global__ void __launch_bounds__(32 * W, MINB) plan_solo_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p,
                                                                  const __grid_constant__ PlanBatchArgs a)
{
    __shared__ SoloShared SS[W];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    SoloShared &S = SS[w];
    int slot = -1;
    uint32_t gen = 0;
    const unsigned total = (unsigned)a.nq;
    while (true)
    {
        unsigned t = 0;
        if (lane == 0)
            t = atomicAdd(a.solo_counter, 1u);
        t = __shfl_sync(kFull, t, 0);
        if (t >= total)
            break;
        if (slot < 0) // a workspace is taken only by a warp that has work
        {
            if (lane == 0)
            {
                slot = pool_acquire(a.solo_pool, (int)(blockIdx.x * W + w));
                gen = a.solo_pool.gen[slot];
            }
            slot = __shfl_sync(kFull, slot, 0);
            gen = __shfl_sync(kFull, gen, 0);
        }
        ++gen;
        plan_solo_query(m, p, a, (int)t, a.solo_pool.base + (size_t)slot * a.Ls.stride, gen, S);
    }
    __syncwarp();
    if (slot >= 0 && lane == 0)
        pool_release(a.solo_pool, slot, gen);
}

} // namespace mpros
