// N4: NavMap::generateGoalMap (nav_map.cpp:279-349) — the obstacle-aware holonomic heuristic.
//
// Reference semantics: unit-weight 4-connected relaxation from the goal cell over the DOWN-SAMPLED, un-inflated
// map; every cell starts at HUGE_INT = 1000; a neighbour is relaxed iff neighbour > current + 1, so stored values
// never exceed 999; the goal cell is 0 even when occupied; occupied / outside cells are never entered. The loop
// is label-correcting (broken heap), but the fixed point is plain BFS distance capped at 999.
//
// GPU form: level-synchronous, BIT-PARALLEL wavefront. One 32-bit word = 32 cells; one level is
//   new = (F<<1 | F>>1 | F_up | F_down) & ~blocked;  blocked |= new;  dist[new cells] = level
// over the words inside the L1 ball of the current level around the goal (the only words that can change).
// All <= 999 levels run inside ONE cooperative kernel (grid.sync between levels), frontier double-buffered.
#include <cooperative_groups.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "mpros_internal.cuh"

namespace cg = cooperative_groups;

namespace mpros
{

__global__ void __launch_bounds__(256) goal_bfs_kernel(const uint32_t *__restrict__ occ, int H, int W, int pitch, int gi, int gj,
                                                       uint16_t *__restrict__ dist, uint32_t *__restrict__ blocked,
                                                       uint32_t *__restrict__ f0, uint32_t *__restrict__ f1,
                                                       unsigned int *__restrict__ level_count)
{
    cg::grid_group grid = cg::this_grid();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    const int wpr = (W + 31) >> 5;

    // prologue: dist = 1000 everywhere, blocked = occupancy, frontiers empty
    const size_t cells = (size_t)H * W;
    for (size_t k = tid; k < cells; k += nthreads)
        dist[k] = (uint16_t)kHugeInt;
    const size_t words = (size_t)H * pitch;
    for (size_t k = tid; k < words; k += nthreads)
    {
        blocked[k] = occ[k];
        f0[k] = 0;
        f1[k] = 0;
    }
    for (size_t k = tid; k < (size_t)kHugeInt; k += nthreads)
        level_count[k] = 0;
    grid.sync();
    if (tid == 0)
    {
        size_t w = (size_t)gi * pitch + (gj >> 5);
        f0[w] = 1u << (gj & 31);
        blocked[w] |= 1u << (gj & 31);
        dist[(size_t)gi * W + gj] = 0; // goal cell is 0 even if occupied (nav_map.cpp:304)
    }
    grid.sync();

    uint32_t *fprev = f0, *fnext = f1;
    for (int level = 1; level < kHugeInt; ++level)
    {
        // words that can change at this level: the L1 ball of radius `level` around the goal
        int r0 = max(0, gi - level), r1 = min(H - 1, gi + level);
        int w0 = max(0, (gj - level) >> 5), w1 = min(wpr - 1, (gj + level) >> 5);
        if (gj - level < 0)
            w0 = 0;
        int nw = w1 - w0 + 1;
        size_t total = (size_t)(r1 - r0 + 1) * nw;
        unsigned int found = 0;
        for (size_t k = tid; k < total; k += nthreads)
        {
            int i = r0 + (int)(k / nw), w = w0 + (int)(k % nw);
            size_t idx = (size_t)i * pitch + w;
            uint32_t c = fprev[idx];
            uint32_t nb = (c << 1) | (c >> 1);
            if (w > 0)
                nb |= fprev[idx - 1] >> 31;
            if (w + 1 < wpr)
                nb |= fprev[idx + 1] << 31;
            if (i > 0)
                nb |= fprev[idx - pitch];
            if (i + 1 < H)
                nb |= fprev[idx + pitch];
            uint32_t valid = (w == wpr - 1 && (W & 31)) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;
            uint32_t nw_bits = nb & ~blocked[idx] & valid;
            fnext[idx] = nw_bits;
            if (nw_bits)
            {
                blocked[idx] |= nw_bits;
                found += __popc(nw_bits);
                uint16_t *drow = dist + (size_t)i * W + ((size_t)w << 5);
                uint32_t b = nw_bits;
                while (b)
                {
                    int t = __ffs(b) - 1;
                    b &= b - 1;
                    drow[t] = (uint16_t)level;
                }
            }
        }
        if (found)
            atomicAdd(&level_count[level], found);
        grid.sync();
        if (level_count[level] == 0)
            break; // uniform across the grid: frontier exhausted
        uint32_t *t = fprev;
        fprev = fnext;
        fnext = t;
    }
}

// Cluster form of the same wavefront for windows that fit the shared memory of one thread-block cluster (every BASELINE
// config: 1024^2 needs 55 KB per CTA, the 2000^2 Arena stand-in at ds = 1 205 KB). Only cells within 999 steps of the goal
// can ever hold a value < 1000, so the wavefront lives in the WINDOW rows [gi-999, gi+999] x words [(gj-999)>>5, (gj+999)>>5].
//  * The window is cut into 8 row bands, one per CTA of the cluster; `blocked` and both frontier buffers of a band stay in
//    that CTA's shared memory for all levels. Distances go straight to global memory when a cell is discovered.
//  * Every band carries K halo rows above and below (copies of the neighbour CTAs' rows, read through distributed shared
//    memory). Row e of level L+1 depends on rows e-1, e, e+1 of level L, so after an exchange the CTA can run K levels on
//    its own: at sub-level j the halo is exact except its outer j rows, the band itself stays exact. The cluster meets
//    only every K levels (two hardware cluster barriers + the halo copy); in between a __syncthreads() separates levels.
//  * A lane moves 128 cells (uint4) of a row per step, over the bounding box of the L1 ball of the current level: the level
//    time is the instruction stream of the 8 SMs, and the vector form runs a quarter of the instructions of a word per lane. (A variant with per-row
//    masks of non-zero frontier words touched fewer words but ran more instructions per warp and level and was slower:
//    3.75 ms against 2.45 ms for the plain form with one cluster barrier per level; the level time is the dependent
//    instruction chain of one warp, not the number of words.)
// Termination: at every exchange each CTA posts "found something since the last one" into a slot of every CTA.
constexpr int kGoalCluster = 8;     // portable cluster size
constexpr int kGoalClusterMax = 16; // non-portable size, used when the device can co-schedule it (one GPC)
constexpr int kGoalThreads = 1024;
constexpr int kGoalMaxHalo = 8;
constexpr size_t kGoalMaxSmem = 220 * 1024;

__global__ void fill_u16_kernel(uint16_t *__restrict__ p, size_t n, uint16_t v)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    const uint32_t vv = (uint32_t)v | ((uint32_t)v << 16);
    uint32_t *p32 = reinterpret_cast<uint32_t *>(p); // cudaMalloc alignment
    for (size_t k = tid; k < n / 2; k += nt)
        p32[k] = vv;
    if (tid == 0 && (n & 1))
        p[n - 1] = v;
}

// words per row in shared memory: a multiple of 4 so that a lane moves 128 cells (uint4) at a time
__host__ __device__ inline int goal_row_stride(int nw) { return (nw + 3) & ~3; }

__global__ void __launch_bounds__(kGoalThreads, 1)
    goal_bfs_cluster_kernel(const uint32_t *__restrict__ occ, int H, int W, int pitch, int gi, int gj, uint16_t *__restrict__ dist, int r0,
                            int rows, int w0, int nw, int band, int K)
{
    extern __shared__ __align__(16) uint32_t gsm[];
    __shared__ int slots[2][kGoalClusterMax];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int csize = (int)cluster.num_blocks(); // 8 or 16 CTAs (launch attribute)
    const int tid = threadIdx.x;
    const int wpr = (W + 31) >> 5;
    const int ext = band + 2 * K;       // rows held by this CTA: K halo + band + K halo
    const int nws = goal_row_stride(nw); // row stride in words
    const int G = nws >> 2;             // uint4 groups per row
    const int ext_words = ext * nws;
    uint32_t *blocked = gsm, *fa = gsm + ext_words, *fb = fa + ext_words;
    const int row_first = rank * band;  // window row of the band's first row; ext row e <-> window row row_first - K + e
    const int my_rows = max(0, min(band, rows - row_first));

    // prologue: rows outside the window, padding words and the bits beyond the last column are blocked for good, so the
    // level loop needs no validity masks
    const uint32_t last_valid = (W & 31) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;
    for (int idx = tid; idx < ext_words; idx += kGoalThreads)
    {
        const int e = idx / nws, w = idx - e * nws;
        const int gr = row_first - K + e;
        uint32_t b = 0xffffffffu;
        if (gr >= 0 && gr < rows && w < nw)
        {
            b = occ[(size_t)(r0 + gr) * pitch + (w0 + w)];
            if (w0 + w == wpr - 1)
                b |= ~last_valid;
        }
        blocked[idx] = b;
        fa[idx] = 0;
        fb[idx] = 0;
    }
    if (tid < 2 * kGoalClusterMax)
        (&slots[0][0])[tid] = 0;
    __syncthreads();
    {
        const int ge = gi - r0 - row_first + K; // goal row in ext coordinates (it may lie in a halo)
        if (tid == 0 && ge >= 0 && ge < ext)
        {
            const int gw = (gj >> 5) - w0;
            fa[ge * nws + gw] = 1u << (gj & 31);
            blocked[ge * nws + gw] |= 1u << (gj & 31);
            if (ge >= K && ge < K + my_rows)
                dist[(size_t)gi * W + gj] = 0; // goal cell is 0 even if occupied (nav_map.cpp:304)
        }
    }
    cluster.sync();

    uint32_t *const up_base = rank > 0 ? cluster.map_shared_rank(gsm, rank - 1) : nullptr;
    uint32_t *const dn_base = rank + 1 < csize ? cluster.map_shared_rank(gsm, rank + 1) : nullptr;
    int found = 0; // in the band, since the last exchange
    for (int level = 1; level < kHugeInt; ++level)
    {
        const bool odd = level & 1; // odd levels read fa, write fb
        const uint32_t *fprev = odd ? fa : fb;
        uint32_t *fnext = odd ? fb : fa;
        // the L1 ball of radius `level`: rows in ext coordinates, uint4 groups in window coordinates. Every group inside it
        // is rewritten at every level (the ball only grows), so the buffer written two levels ago holds nothing stale.
        const int lo = max(0, gi - level - r0 - row_first + K), hi = min(ext - 1, gi + level - r0 - row_first + K);
        const int gl = max(0, (max(0, gj - level) >> 5) - w0) >> 2, gh = min(nw - 1, ((gj + level) >> 5) - w0) >> 2;
        // lanes per row: the power of two that covers the groups in range; the other lanes take further rows
        const int ga = gh - gl + 1;
        const int lpr_log = ga <= 1 ? 0 : 32 - __clz(ga - 1);
        const int rows_per_pass = kGoalThreads >> lpr_log;
        const int g = gl + (tid & ((1 << lpr_log) - 1));
        if (g <= gh)
            for (int e = lo + (tid >> lpr_log); e <= hi; e += rows_per_pass)
            {
                const int idx = e * nws + 4 * g;
                const uint4 c = *reinterpret_cast<const uint4 *>(fprev + idx);
                uint4 u = make_uint4(0, 0, 0, 0), d = make_uint4(0, 0, 0, 0);
                if (e > 0)
                    u = *reinterpret_cast<const uint4 *>(fprev + idx - nws);
                if (e + 1 < ext)
                    d = *reinterpret_cast<const uint4 *>(fprev + idx + nws);
                const uint32_t lw = g > 0 ? fprev[idx - 1] : 0u, rw = g + 1 < G ? fprev[idx + 4] : 0u;
                const uint4 bl = *reinterpret_cast<const uint4 *>(blocked + idx);
                uint4 fr;
                fr.x = ((c.x << 1) | (c.x >> 1) | (lw >> 31) | (c.y << 31) | u.x | d.x) & ~bl.x;
                fr.y = ((c.y << 1) | (c.y >> 1) | (c.x >> 31) | (c.z << 31) | u.y | d.y) & ~bl.y;
                fr.z = ((c.z << 1) | (c.z >> 1) | (c.y >> 31) | (c.w << 31) | u.z | d.z) & ~bl.z;
                fr.w = ((c.w << 1) | (c.w >> 1) | (c.z >> 31) | (rw << 31) | u.w | d.w) & ~bl.w;
                *reinterpret_cast<uint4 *>(fnext + idx) = fr;
                if (fr.x | fr.y | fr.z | fr.w)
                {
                    *reinterpret_cast<uint4 *>(blocked + idx) = make_uint4(bl.x | fr.x, bl.y | fr.y, bl.z | fr.z, bl.w | fr.w);
                    if (e >= K && e < K + my_rows)
                    {
                        found = 1;
                        uint16_t *drow = dist + (size_t)(r0 + row_first - K + e) * W + ((size_t)(w0 + 4 * g) << 5);
                        const uint32_t fw[4] = {fr.x, fr.y, fr.z, fr.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                        {
                            uint32_t b = fw[q];
                            while (b)
                            {
                                const int t = __ffs(b) - 1;
                                b &= b - 1;
                                drow[32 * q + t] = (uint16_t)level;
                            }
                        }
                    }
                }
            }
        if (level % K != 0)
        {
            __syncthreads();
            continue;
        }
        // ---- exchange: termination vote, then the K halo rows on each side from the neighbours' bands
        const int par = (level / K) & 1;
        const int any = __syncthreads_or(found);
        found = 0;
        if (tid < csize)
            *cluster.map_shared_rank(&slots[par][rank], tid) = any;
        cluster.sync(); // every band is final for this level, the votes are visible
        int tot = 0;
        for (int t = 0; t < csize; ++t)
            tot |= slots[par][t];
        if (!tot)
            break; // uniform across the cluster: no band found anything in the last K levels, the frontier is exhausted
        {
            const size_t f_off = (size_t)(fnext - gsm); // this level's output = next level's input
            if (up_base) // my ext rows [0, K) <- rows [band, band + K) of rank - 1 (the last K rows of its band)
                for (int idx = tid; idx < K * nws; idx += kGoalThreads)
                {
                    blocked[idx] = up_base[band * nws + idx];
                    fnext[idx] = up_base[f_off + band * nws + idx];
                }
            if (dn_base) // my ext rows [K + band, K + band + K) <- rows [K, 2K) of rank + 1 (the first K rows of its band)
            {
                const int dst = (K + band) * nws;
                for (int idx = tid; idx < K * nws; idx += kGoalThreads)
                {
                    blocked[dst + idx] = dn_base[K * nws + idx];
                    fnext[dst + idx] = dn_base[f_off + K * nws + idx];
                }
            }
        }
        cluster.sync(); // halos copied before any band changes again
    }
    cluster.sync(); // nobody leaves while a neighbour may still read its shared memory
}

int map_build_goal(Ctx *c)
{
    // posToIndex on the down-sampled grid + pointInMap (nav_map.cpp:295-302)
    double dx = c->goal_x - c->ox, dy = c->goal_y - c->oy;
    double qj = (dx * c->ocos + dy * c->osin) / c->res;
    double qi = (dx * -c->osin + dy * c->ocos) / c->res;
    bool inside = qi > -1.0 && qi < (double)c->H && qj > -1.0 && qj < (double)c->W;
    if (!inside)
    {
        c->get_goal = false; // "Goal point out of map, goal cost map not generated"
        return MPROS_OK;
    }
    int gi = (int)qi, gj = (int)qj;
    // cluster form when the <= 1999 x 1999-cell window around the goal fits the cluster's shared memory
    {
        const int wpr = (c->W + 31) >> 5;
        const int r0 = std::max(0, gi - (kHugeInt - 1)), r1 = std::min(c->H - 1, gi + (kHugeInt - 1));
        const int w0 = std::max(0, gj - (kHugeInt - 1)) >> 5, w1 = std::min(wpr - 1, (gj + (kHugeInt - 1)) >> 5);
        const int rows = r1 - r0 + 1, nw = w1 - w0 + 1;
        const bool force_grid = c->opt.goal_mode == 2; // "grid": the cooperative kernel; 1 = "cluster8": the portable cluster size
        static int attr_set[16] = {}; // 0 = not yet, 1 = 8 CTAs only, 2 = 16 CTAs available
        if (!force_grid && !attr_set[c->device & 15])
        {
            MPROS_CUDA(cudaFuncSetAttribute(goal_bfs_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGoalMaxSmem));
            attr_set[c->device & 15] = 1;
            if (cudaFuncSetAttribute(goal_bfs_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess)
            {
                cudaLaunchConfig_t probe = {};
                probe.gridDim = dim3(kGoalClusterMax);
                probe.blockDim = dim3(kGoalThreads);
                probe.dynamicSmemBytes = 64 * 1024;
                cudaLaunchAttribute pa[1];
                pa[0].id = cudaLaunchAttributeClusterDimension;
                pa[0].val.clusterDim.x = kGoalClusterMax;
                pa[0].val.clusterDim.y = 1;
                pa[0].val.clusterDim.z = 1;
                probe.attrs = pa;
                probe.numAttrs = 1;
                int n_clusters = 0;
                if (cudaOccupancyMaxActiveClusters(&n_clusters, goal_bfs_cluster_kernel, &probe) == cudaSuccess && n_clusters >= 1)
                    attr_set[c->device & 15] = 2;
            }
            (void)cudaGetLastError();
        }
        int csize = (attr_set[c->device & 15] == 2 && c->opt.goal_mode != 1) ? kGoalClusterMax : kGoalCluster;
        auto plan_for = [&](int cs, int &band_o, int &halo_o) {
            band_o = (rows + cs - 1) / cs;
            // halo depth: as deep as the shared memory allows, never deeper than a band (a halo is a copy of the neighbour's band rows)
            halo_o = std::min(kGoalMaxHalo, band_o);
            auto smem_for = [&](int k) { return (size_t)3 * (band_o + 2 * k) * goal_row_stride(nw) * sizeof(uint32_t); };
            while (halo_o > 1 && smem_for(halo_o) > kGoalMaxSmem)
                halo_o >>= 1;
            return smem_for(halo_o);
        };
        int band = 0, halo = 0;
        size_t smem = plan_for(csize, band, halo);
        if (!force_grid && smem <= kGoalMaxSmem)
        {
            const size_t cells = (size_t)c->H * c->W;
            fill_u16_kernel<<<c->n_sm * 4, 256, 0, c->stream>>>(c->goal, cells, (uint16_t)kHugeInt);
            MPROS_LAUNCH_CHECK(c);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(csize);
            cfg.blockDim = dim3(kGoalThreads);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = c->stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = csize;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            const uint32_t *occ = c->ds_bits;
            uint16_t *dist = c->goal;
            int H = c->H, W = c->W, pitch = c->pitch;
            MPROS_CUDA(cudaLaunchKernelEx(&cfg, goal_bfs_cluster_kernel, occ, H, W, pitch, gi, gj, dist, r0, rows, w0, nw, band, halo));
            c->launches++;
            return MPROS_OK;
        }
    }
    size_t words = (size_t)c->H * c->pitch;
    size_t need = words * 4 * 3 + kHugeInt * sizeof(unsigned int) + 64;
    int rc = ensure_scratch(c, need);
    if (rc)
        return rc;
    uint32_t *blocked = static_cast<uint32_t *>(c->scratch);
    uint32_t *f0 = blocked + words, *f1 = f0 + words;
    unsigned int *level_count = reinterpret_cast<unsigned int *>(f1 + words);

    int per_sm = 0;
    MPROS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, goal_bfs_kernel, 256, 0));
    if (per_sm < 1)
        per_sm = 1;
    if (per_sm > 4)
        per_sm = 4; // the per-level work is small; fewer CTAs make grid.sync cheaper
    int grid = c->n_sm * per_sm;
    const uint32_t *occ = c->ds_bits;
    int H = c->H, W = c->W, pitch = c->pitch;
    uint16_t *dist = c->goal;
    void *args[] = {&occ, &H, &W, &pitch, &gi, &gj, &dist, &blocked, &f0, &f1, &level_count};
    MPROS_CUDA(cudaLaunchCooperativeKernel((void *)goal_bfs_kernel, dim3(grid), dim3(256), args, 0, c->stream));
    c->launches++;
    return MPROS_OK;
}

} // namespace mpros

using namespace mpros;

extern "C" int mpros_map_update_goal(mpros_ctx *c, double goal_x, double goal_y)
{
    if (!c)
    {
        set_error("mpros_map_update_goal: NULL context");
        return MPROS_ERR_INVALID;
    }
    if (!c->get_map)
    {
        // the reference would run generateGoalMap on an empty map; there is nothing to compute
        set_error("mpros_map_update_goal: no map set");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    c->goal_x = goal_x;
    c->goal_y = goal_y;
    c->get_goal = true; // nav_map.h:208
    int rc = map_build_goal(c);
    if (rc)
        return rc;
    MPROS_CUDA(cudaStreamSynchronize(c->stream));
    return MPROS_OK;
}
