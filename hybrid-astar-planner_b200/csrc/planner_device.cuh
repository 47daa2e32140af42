// Device-resident Hybrid A* building blocks (one WARP per query):
//   * warp-cooperative footprint test (probes of one pose spread over the lanes, one ballot = verdict),
//   * open list: 32-ary min-heap keyed on (f, insertion id)  == std::multimap<double,NodePtr> pop order (FIFO ties),
//   * closed set: tagged open-addressing hash replacing the reference's dense cost_set_/node_set_ arrays,
//   * obstacle-aware holonomic heuristic computed ON DEMAND: the goal wavefront (N4) is advanced only until the
//     cells the current expansion asks for are final (values < current level are final in a unit-weight BFS),
//   * node expansion E1-E4 and the Reeds-Shepp shot R4.
#pragma once
#include "map_device.cuh"
#include "rs_device.cuh"

namespace mpros
{

constexpr uint32_t kNoNode = 0xffffffffu;
constexpr uint32_t kGoalNode = 0xfffffffeu;
constexpr int kTrajCap = 256;
constexpr int kMaxPrim = 2 * kMaxSteer;

struct __align__(16) PlanNode
{
    double x, y, yaw, steer, g;
    double s, c; // sin / cos of yaw (already computed for the footprint test when the node was a child)
    uint32_t parent;
    int8_t dir;
    uint8_t closed;
    uint16_t pad;
};

// ---- warp-cooperative footPrintInCollision4 (hybrid_a_star.cpp:617-680) ---------------------------------------------
// All lanes pass the same pose; probe k of the pose is evaluated by lane (k mod 32). Verdict is warp-uniform.
// Fast path: the fixed-point walk of map_device.cuh (Q(k) = Q(0) + k * dQ over the padded tile-word layers) with the same
// preconditions and the same error bound; if the pose is out of its range or ANY probe lies within kFxEps of a cell edge
// the reference arithmetic below decides.
__device__ __noinline__ bool warp_footprint_collision(const MapView &m, const PlannerView &p, double px, double py, double s,
                                                         double c)
{
    const int lane = threadIdx.x & 31;
    double x0 = c * p.longi_x0 + px, y0 = s * p.longi_x0 + py;
    double x1 = c * p.longi_x1 + px, y1 = s * p.longi_x1 + py;
    const int total = p.n_check + 1 + 4;
    if (fabs(x0) + fabs(y0) < m.fx_span - p.fx_extent)
    {
        const double kTwo32 = 4294967296.0;
        const double sc = m.inv_res0 * kTwo32;
        const double jc = m.ocos * sc, js = m.osin * sc;
        const double off = (double)kTwMargin * kTwo32;
        const double ax0 = x0 - m.ox, ay0 = y0 - m.oy, ax1 = x1 - m.ox, ay1 = y1 - m.oy;
        const double fj0 = ax0 * jc + ay0 * js + off, fi0 = ay0 * jc - ax0 * js + off;
        const double fj1 = ax1 * jc + ay1 * js + off, fi1 = ay1 * jc - ax1 * js + off;
        const long long qj0 = __double2ll_rd(fj0), qi0 = __double2ll_rd(fi0);
        const long long ej = __double2ll_rd(fj1), ei = __double2ll_rd(fi1);
        const unsigned lim_j = (unsigned)(m.W0 + 2 * kTwMargin - 2 * p.fx_slack), lim_i = (unsigned)(m.H0 + 2 * kTwMargin - 2 * p.fx_slack);
        const bool in = ((unsigned)((int)(qj0 >> 32) - p.fx_slack) < lim_j) & ((unsigned)((int)(ej >> 32) - p.fx_slack) < lim_j) &
                        ((unsigned)((int)(qi0 >> 32) - p.fx_slack) < lim_i) & ((unsigned)((int)(ei >> 32) - p.fx_slack) < lim_i);
        if (in)
        {
            const long long dj = __double2ll_rn((fj1 - fj0) * p.inv_n_check), di = __double2ll_rn((fi1 - fi0) * p.inv_n_check);
            uint32_t any = 0, edge = 0xffffffffu;
            for (int base = 0; base < total; base += 32)
            {
                const int k = base + lane;
                if (k <= p.n_check)
                {
                    // the sequential walk adds dQ k times; k * dQ is the same integer
                    const long long qj = qj0 + (long long)k * dj, qi = qi0 + (long long)k * di;
                    any |= tw_probe(m.infl_tw, m.tw_pitch, qj, qi);
                    edge = tw_edge(edge, qj, qi);
                }
                else if (k < total)
                {
                    const int q = k - p.n_check - 1;
                    const double cx = p.corner_x[q], cy = p.corner_y[q];
                    const double X = cx * c - cy * s + px, Y = cx * s + cy * c + py;
                    const double ax = X - m.ox, ay = Y - m.oy;
                    const long long cj = __double2ll_rd(ax * jc + ay * js + off), ci = __double2ll_rd(ay * jc - ax * js + off);
                    any |= tw_probe(m.raw_tw, m.tw_pitch, cj, ci);
                    edge = tw_edge(edge, cj, ci);
                }
            }
            if (!__any_sync(kFull, edge < 2 * kFxEps))
                return __any_sync(kFull, (any & 1u) != 0);
        }
    }
    double n = (double)p.n_check;
    double ox = (x1 - x0) / n, oy = (y1 - y0) / n;
    for (int base = 0; base < total; base += 32)
    {
        int k = base + lane;
        bool hit = false;
        if (k <= p.n_check)
        {
            double kk = (double)k;
            hit = point_hits_orig(m, m.infl_bits, ox * kk + x0, oy * kk + y0);
        }
        else if (k < total)
        {
            int q = k - p.n_check - 1;
            double cx = p.corner_x[q], cy = p.corner_y[q];
            hit = point_hits_orig(m, m.raw_bits, cx * c - cy * s + px, cx * s + cy * c + py);
        }
        if (__any_sync(kFull, hit))
            return true;
    }
    return false;
}

// ---- open list --------------------------------------------------------------------------------------------------------
// f >= 0 always, so the IEEE bit pattern orders like the value. Ties pop in insertion order (node ids grow with
// insertion), which is what std::multimap::begin() + insert-at-upper-bound give the reference.
constexpr int kHeapTop = 1 + 32 + 1024; // heap levels 0-2 live in shared memory
struct OpenHeap
{
    unsigned long long *key; // global arrays (entries >= scap)
    uint32_t *id;
    unsigned long long *skey; // shared-memory copy of the first scap entries (the hot top of the heap)
    uint32_t *sid;
    int scap;
    int size, cap;
};
__device__ __forceinline__ unsigned long long hk(const OpenHeap &h, int i) { return i < h.scap ? h.skey[i] : h.key[i]; }
__device__ __forceinline__ uint32_t hi(const OpenHeap &h, int i) { return i < h.scap ? h.sid[i] : h.id[i]; }
__device__ __forceinline__ void hset(const OpenHeap &h, int i, unsigned long long k, uint32_t id)
{
    if (i < h.scap)
    {
        h.skey[i] = k;
        h.sid[i] = id;
    }
    else
    {
        h.key[i] = k;
        h.id[i] = id;
    }
}
__device__ __forceinline__ bool heap_less(unsigned long long ka, uint32_t ia, unsigned long long kb, uint32_t ib)
{
    return ka < kb || (ka == kb && ia < ib);
}
// executed by ONE thread; caller makes the writes visible (__syncwarp / __syncthreads) before the next pop
__device__ __noinline__ void heap_push_lane0(OpenHeap &h, unsigned long long k, uint32_t id)
{
    int pos = h.size;
    while (pos > 0)
    {
        int parent = (pos - 1) >> 5;
        unsigned long long pk = hk(h, parent);
        uint32_t pi = hi(h, parent);
        if (!heap_less(k, id, pk, pi))
            break;
        hset(h, pos, pk, pi);
        pos = parent;
    }
    hset(h, pos, k, id);
}
// warp-cooperative pop; returns the id with the smallest (key, id); h.size must be > 0.
// Bottom-up variant: the hole left by the root walks down along the smallest children to a leaf WITHOUT looking at the
// element that will fill it (the heap's last entry, almost always larger than everything on the way), so that entry's load
// overlaps the whole descent instead of heading the dependent chain; it is then placed by a (rarely needed) sift-up.
// Any valid heap gives the same pop order: the order is fixed by the (key, id) total order alone.
__device__ __noinline__ uint32_t heap_pop_warp(OpenHeap &h, unsigned long long &out_key)
{
    const int lane = threadIdx.x & 31;
    unsigned long long rk = hk(h, 0);
    uint32_t ri = hi(h, 0);
    out_key = rk;
    int n = --h.size;
    if (n == 0)
        return ri;
    unsigned long long lk = hk(h, n); // in flight during the descent
    uint32_t li = hi(h, n);
    int pos = 0;
    while ((pos << 5) + 1 < n)
    {
        int child = (pos << 5) + 1 + lane;
        unsigned long long ck = ~0ull;
        uint32_t ci = 0xffffffffu;
        if (child < n)
        {
            ck = hk(h, child);
            ci = hi(h, child);
        }
        // lexicographic arg-min over (key hi, key lo, id) with three warp reductions (redux.sync)
        const uint32_t khi = (uint32_t)(ck >> 32), klo = (uint32_t)ck;
        const uint32_t m1 = __reduce_min_sync(kFull, khi);
        const bool a1 = khi == m1;
        const uint32_t m2 = __reduce_min_sync(kFull, a1 ? klo : 0xffffffffu);
        const bool a2 = a1 && klo == m2;
        const uint32_t m3 = __reduce_min_sync(kFull, a2 ? ci : 0xffffffffu);
        const unsigned win = __ballot_sync(kFull, a2 && ci == m3);
        if (lane == 0)
            hset(h, pos, ((unsigned long long)m1 << 32) | m2, m3);
        pos = (pos << 5) + 1 + (__ffs(win) - 1);
    }
    // place the last entry: sift it up from the hole (lane 0; the parents on the way were written above, so make them visible)
    __syncwarp();
    if (lane == 0)
    {
        while (pos > 0)
        {
            const int parent = (pos - 1) >> 5;
            const unsigned long long pk = hk(h, parent);
            const uint32_t pi = hi(h, parent);
            if (!heap_less(lk, li, pk, pi))
                break;
            hset(h, pos, pk, pi);
            pos = parent;
        }
        hset(h, pos, lk, li);
    }
    __syncwarp();
    return ri;
}

// ---- closed set -------------------------------------------------------------------------------------------------------
// cost_set_[index] / node_set_[index] (hybrid_a_star.cpp:103-112, 380-394) as a hash: slot = {tag<<34 | index, g, node}.
// The tag (query generation) makes stale slots of earlier queries read as empty, so the table is never cleared.
struct __align__(32) ClosedSlot
{
    unsigned long long key; // tag << 34 | index
    double g;               // cost_set_[index]
    uint32_t node;          // node_set_[index]
    uint32_t pad[3];
};
struct ClosedSet
{
    ClosedSlot *slot;
    uint32_t mask; // slots - 1 (power of two)
    unsigned long long tag;
};
__device__ __forceinline__ uint32_t hash_index(unsigned long long idx)
{
    idx ^= idx >> 29;
    idx *= 0xbf58476d1ce4e5b9ull;
    idx ^= idx >> 32;
    return (uint32_t)idx;
}
// single-thread find-or-insert-slot: returns the slot index; found says whether the key was present, and then g / node
// are the stored values (one 32-byte sector per probe)
__device__ __noinline__ uint32_t closed_find(const ClosedSet &t, unsigned long long idx, bool &found, double &g, uint32_t &node)
{
    unsigned long long want = (t.tag << 34) | idx;
    uint32_t s = hash_index(idx) & t.mask;
    while (true)
    {
        const uint4 raw = *reinterpret_cast<const uint4 *>(&t.slot[s]); // key + g
        unsigned long long k = ((unsigned long long)raw.y << 32) | raw.x;
        if (k == want)
        {
            found = true;
            g = __longlong_as_double((long long)(((unsigned long long)raw.w << 32) | raw.z));
            node = t.slot[s].node;
            return s;
        }
        if ((k >> 34) != t.tag)
        {
            found = false;
            g = __longlong_as_double(0x7ff0000000000000ll);
            node = kNoNode;
            return s;
        }
        s = (s + 1) & t.mask;
    }
}
__device__ __forceinline__ void closed_store(const ClosedSet &t, uint32_t s, unsigned long long idx, double g, uint32_t node)
{
    ClosedSlot v;
    v.key = (t.tag << 34) | idx;
    v.g = g;
    v.node = node;
    v.pad[0] = v.pad[1] = v.pad[2] = 0;
    t.slot[s] = v;
}

// ---- on-demand goal wavefront (N4 semantics, nav_map.cpp:279-349) -------------------------------------------------------
struct GoalBfs
{
    uint32_t *blocked, *f0, *f1; // [rows x wwords] window bitmaps (blocked = occupied | visited)
    uint16_t *dist;              // [rows x wwords*32]
    uint32_t *row_tag;           // [rows] == tag when the row's bitmaps are initialised for this query
    uint32_t tag;
    int r0, w0, rows, wwords; // window origin (row, word) and extent
    int gi, gj;               // goal cell
    int level;                // levels <= level are final
    int done;                 // frontier exhausted or level 999 reached
    int br0, br1, bw0, bw1;   // bbox (rows, words; map coordinates) of every visited cell so far
    long long cells;          // cells discovered so far (statistics)
};

__device__ __noinline__ void bfs_init_rows(const MapView &m, GoalBfs &b, int ra, int rb)
{
    const int lane = threadIdx.x & 31;
    ra = max(ra, b.r0);
    rb = min(rb, b.r0 + b.rows - 1);
    for (int r = ra; r <= rb; ++r)
    {
        int lr = r - b.r0;
        if (b.row_tag[lr] == b.tag)
            continue;
        __syncwarp();
        for (int w = lane; w < b.wwords; w += 32)
        {
            size_t o = (size_t)lr * b.wwords + w;
            b.blocked[o] = __ldg(m.ds_bits + (size_t)r * m.pitch + b.w0 + w);
            b.f0[o] = 0;
            b.f1[o] = 0;
        }
        if (lane == 0)
            b.row_tag[lr] = b.tag;
        __syncwarp();
    }
}

__device__ __forceinline__ void bfs_start(const MapView &m, GoalBfs &b, int gi, int gj, uint32_t tag)
{
    const int lane = threadIdx.x & 31;
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
    bfs_init_rows(m, b, gi - 2, gi + 2);
    if (lane == 0)
    {
        size_t o = (size_t)(gi - b.r0) * b.wwords + ((gj >> 5) - b.w0);
        b.f0[o] = 1u << (gj & 31);
        b.blocked[o] |= 1u << (gj & 31);
    }
    __syncwarp();
}

// advance the wavefront by one level (warp-cooperative)
__device__ __noinline__ void bfs_step(const MapView &m, GoalBfs &b)
{
    const int lane = threadIdx.x & 31;
    const int level = b.level + 1;
    if (level >= kHugeInt)
    {
        b.done = 1;
        return;
    }
    const int wr1 = b.r0 + b.rows - 1, ww1 = b.w0 + b.wwords - 1;
    const int sr0 = max(b.r0, b.br0 - 1), sr1 = min(wr1, b.br1 + 1);
    const int sw0 = max(b.w0, b.bw0 - 1), sw1 = min(ww1, b.bw1 + 1);
    bfs_init_rows(m, b, sr0 - 2, sr1 + 2);
    const int nw = sw1 - sw0 + 1;
    const int total = (sr1 - sr0 + 1) * nw;
    uint32_t *fprev = (level & 1) ? b.f0 : b.f1;
    uint32_t *fnext = (level & 1) ? b.f1 : b.f0;
    const int wpr = (m.W + 31) >> 5;
    int nr0 = 1 << 30, nr1 = -1, nw0 = 1 << 30, nw1 = -1;
    int found = 0;
    for (int k = lane; k < total; k += 32)
    {
        int r = sr0 + k / nw, w = sw0 + k % nw;
        int lr = r - b.r0, lw = w - b.w0;
        size_t o = (size_t)lr * b.wwords + lw;
        uint32_t c = fprev[o];
        uint32_t nb = (c << 1) | (c >> 1);
        if (lw > 0)
            nb |= fprev[o - 1] >> 31;
        if (lw + 1 < b.wwords)
            nb |= fprev[o + 1] << 31;
        if (lr > 0)
            nb |= fprev[o - b.wwords];
        if (lr + 1 < b.rows)
            nb |= fprev[o + b.wwords];
        uint32_t valid = (w == wpr - 1 && (m.W & 31)) ? ((1u << (m.W & 31)) - 1u) : 0xffffffffu;
        uint32_t bl = b.blocked[o];
        uint32_t nbits = nb & ~bl & valid;
        fnext[o] = nbits;
        if (nbits)
        {
            b.blocked[o] = bl | nbits;
            found += __popc(nbits);
            nr0 = min(nr0, r), nr1 = max(nr1, r), nw0 = min(nw0, w), nw1 = max(nw1, w);
            uint16_t *drow = b.dist + ((size_t)lr * b.wwords + lw) * 32;
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
    __syncwarp();
    b.level = level;
    b.cells += found;
    if (nr1 < 0)
        b.done = 1; // empty frontier
    else
    {
        b.br0 = min(b.br0, nr0), b.br1 = max(b.br1, nr1), b.bw0 = min(b.bw0, nw0), b.bw1 = max(b.bw1, nw1);
    }
}

// goal_cost_map_ value of cell (i, j); -1 while not final yet. Per-lane.
__device__ __forceinline__ int bfs_peek(const MapView &m, const GoalBfs &b, int i, int j)
{
    if (i == b.gi && j == b.gj)
        return 0;
    if (bit_at(m.ds_bits, m.pitch, i, j))
        return kHugeInt; // occupied cells are never entered
    int lr = i - b.r0, lw = (j >> 5) - b.w0;
    if (lr < 0 || lr >= b.rows || lw < 0 || lw >= b.wwords)
        return kHugeInt; // further than 999 steps
    if (b.row_tag[lr] == b.tag)
    {
        size_t o = (size_t)lr * b.wwords + lw;
        if ((b.blocked[o] >> (j & 31)) & 1u) // visited (not occupied, checked above)
            return b.dist[o * 32 + (j & 31)];
    }
    return b.done ? kHugeInt : -1;
}

// Warp-collective lookup: every lane may ask for one cell (want = false -> no request). Advances the wavefront until
// all requested cells are final.
__device__ __forceinline__ int bfs_lookup_warp(const MapView &m, GoalBfs &b, bool want, int i, int j)
{
    int v = want ? bfs_peek(m, b, i, j) : 0;
    while (__any_sync(kFull, v < 0))
    {
        bfs_step(m, b);
        if (v < 0)
            v = bfs_peek(m, b, i, j);
    }
    return v;
}

// hybrid_a_star.cpp:266-279 / :292-302 / :281-290
__device__ __forceinline__ double regular_yaw(double yaw)
{
    if (yaw <= -kPi)
        return yaw + 2 * kPi;
    if (yaw > kPi)
        return yaw - 2 * kPi;
    return yaw;
}
__device__ __forceinline__ int yaw_to_idx(const PlannerView &p, double yaw)
{
    double yaw_pos = yaw;
    if (yaw < 0)
        yaw_pos = yaw + 2 * kPi;
    return (int)(yaw_pos / p.yaw_resol);
}
__device__ __forceinline__ unsigned long long cal_index(const MapView &m, const PlannerView &p, int i, int j, int iyaw, double direc)
{
    int idx_direc = direc > 0 ? 0 : 1;
    return ((unsigned long long)((long long)i * m.W + j) * p.num_yaw + (unsigned long long)iyaw) * 2ull + idx_direc;
}

// One motion primitive (hybrid_a_star.cpp:320-350): prim = 2 * steer_index + direction_index. The vehicle-frame
// displacement and yaw change are per-primitive constants (PlannerView::prim_*); only the rotation into the map frame
// depends on the node.
__device__ __forceinline__ void primitive_end_pose(const PlannerView &p, int prim, double cx, double cy, double cyaw, double s_yaw,
                                                   double c_yaw, double &nx, double &ny, double &nyaw, double &steer, double &direc)
{
    steer = p.steer_set[prim >> 1];
    direc = (prim & 1) ? -1.0 : 1.0;
    const double dx = p.prim_dx[prim], dy = p.prim_dy[prim];
    nx = dx * c_yaw - dy * s_yaw + cx;
    ny = dx * s_yaw + dy * c_yaw + cy;
    nyaw = (steer == 0) ? cyaw : regular_yaw(cyaw + p.prim_dyaw[prim]);
}

// setNodeCost (hybrid_a_star.cpp:208-246)
__device__ __forceinline__ double node_cost(const PlannerView &p, double parent_g, double parent_dir, double parent_steer, double voro,
                                            double steer, double direc)
{
    double path_cost = p.step_size;
    double direc_cost = 0;
    if (direc < 0)
        direc_cost += p.pen_back * p.step_size;
    if (direc != parent_dir)
        direc_cost += p.pen_gear;
    double steer_cost = 0;
    if (steer != 0)
        steer_cost += p.pen_steer * fabs(steer);
    if (steer != parent_steer)
        steer_cost += p.pen_steer_change * fabs(steer - parent_steer);
    return parent_g + voro + path_cost + direc_cost + steer_cost;
}

} // namespace mpros
