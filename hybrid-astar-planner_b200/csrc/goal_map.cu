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
// config: 1024^2 needs 48 KB per CTA, the 2000^2 Arena stand-in at ds = 1 189 KB). Only cells within 999 steps of the goal
// can ever hold a value < 1000, so the wavefront lives in the WINDOW rows [gi-999, gi+999] x words [(gj-999)>>5, (gj+999)>>5].
// The window is cut into 8 row bands, one per CTA of the cluster; `blocked` and both frontier buffers of a band stay in
// that CTA's shared memory for all levels, the row above / below a band is read from the neighbour CTA through
// distributed shared memory, and ONE hardware cluster barrier per level replaces the grid-wide software barrier of the
// cooperative kernel (~1 us instead of ~10 us per level; <= 999 levels). Distances go straight to global memory when a
// cell is discovered. Termination: every CTA posts "found something" into a slot of every CTA before the barrier.
constexpr int kGoalCluster = 8;
constexpr int kGoalThreads = 1024;

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

__global__ void __cluster_dims__(kGoalCluster, 1, 1) __launch_bounds__(kGoalThreads, 1)
    goal_bfs_cluster_kernel(const uint32_t *__restrict__ occ, int H, int W, int pitch, int gi, int gj, uint16_t *__restrict__ dist, int r0,
                            int rows, int w0, int nw, int band)
{
    extern __shared__ uint32_t gsm[];
    __shared__ int slots[2][kGoalCluster];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    const int wpr = (W + 31) >> 5;
    const int band_words = band * nw;
    uint32_t *blocked = gsm, *fa = gsm + band_words, *fb = fa + band_words;
    const int row_first = rank * band;                               // window row of local row 0
    const int my_rows = max(0, min(band, rows - row_first));         // the last bands may be short or empty
    const bool have_up = rank > 0 && my_rows > 0;                    // CTA rank-1 always holds `band` rows when this one has any
    const bool have_down = rank + 1 < kGoalCluster && row_first + band < rows;
    const uint32_t *up_a = have_up ? cluster.map_shared_rank(fa, rank - 1) + (size_t)(band - 1) * nw : nullptr;
    const uint32_t *up_b = have_up ? cluster.map_shared_rank(fb, rank - 1) + (size_t)(band - 1) * nw : nullptr;
    const uint32_t *dn_a = have_down ? cluster.map_shared_rank(fa, rank + 1) : nullptr;
    const uint32_t *dn_b = have_down ? cluster.map_shared_rank(fb, rank + 1) : nullptr;

    for (int k = tid; k < band_words; k += kGoalThreads)
    {
        const int li = k / nw, w = k - li * nw;
        blocked[k] = li < my_rows ? occ[(size_t)(r0 + row_first + li) * pitch + (w0 + w)] : 0xffffffffu;
        fa[k] = 0;
        fb[k] = 0;
    }
    if (tid < 2 * kGoalCluster)
        (&slots[0][0])[tid] = 0;
    __syncthreads();
    {
        const int gl = gi - r0 - row_first; // goal row, local
        if (tid == 0 && gl >= 0 && gl < my_rows)
        {
            const int k = gl * nw + ((gj >> 5) - w0);
            fa[k] = 1u << (gj & 31);
            blocked[k] |= 1u << (gj & 31);
            dist[(size_t)gi * W + gj] = 0; // goal cell is 0 even if occupied (nav_map.cpp:304)
        }
    }
    cluster.sync();

    const uint32_t last_valid = (W & 31) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;
    int found = 0; // since the last termination test
    for (int level = 1; level < kHugeInt; ++level)
    {
        const bool odd = level & 1; // odd levels read fa, write fb
        const uint32_t *fprev = odd ? fa : fb;
        uint32_t *fnext = odd ? fb : fa;
        const uint32_t *up = odd ? up_a : up_b, *dn = odd ? dn_a : dn_b;
        // the L1 ball of radius `level`, in window / local coordinates
        const int lo = max(max(0, gi - level - r0), row_first) - row_first;
        const int hi = min(min(rows - 1, gi + level - r0), row_first + my_rows - 1) - row_first;
        const int wl = max(0, (max(0, gj - level) >> 5) - w0), wh = min(nw - 1, ((gj + level) >> 5) - w0);
        const int nwa = wh - wl + 1;
        const int total = (hi >= lo && nwa > 0) ? (hi - lo + 1) * nwa : 0;
        for (int k = tid; k < total; k += kGoalThreads)
        {
            const int li = lo + k / nwa, w = wl + k % nwa;
            const int idx = li * nw + w;
            const uint32_t c = fprev[idx];
            uint32_t nb = (c << 1) | (c >> 1);
            if (w > 0)
                nb |= fprev[idx - 1] >> 31;
            if (w + 1 < nw)
                nb |= fprev[idx + 1] << 31;
            if (li > 0)
                nb |= fprev[idx - nw];
            else if (up)
                nb |= up[w];
            if (li + 1 < my_rows)
                nb |= fprev[idx + nw];
            else if (dn)
                nb |= dn[w];
            const uint32_t valid = (w0 + w == wpr - 1) ? last_valid : 0xffffffffu;
            const uint32_t fresh = nb & ~blocked[idx] & valid;
            fnext[idx] = fresh;
            if (fresh)
            {
                blocked[idx] |= fresh;
                found = 1;
                uint16_t *drow = dist + (size_t)(r0 + row_first + li) * W + ((size_t)(w0 + w) << 5);
                uint32_t b = fresh;
                while (b)
                {
                    const int t = __ffs(b) - 1;
                    b &= b - 1;
                    drow[t] = (uint16_t)level;
                }
            }
        }
        // termination is tested every 4th level only (an exhausted frontier produces nothing, the extra levels are empty):
        // three of four levels cost one cluster barrier and nothing else
        const bool check = (level & 3) == 0;
        const int par = (level >> 2) & 1;
        if (check)
        {
            const int any = __syncthreads_or(found);
            found = 0;
            if (tid < kGoalCluster)
                *cluster.map_shared_rank(&slots[par][rank], tid) = any;
        }
        cluster.sync();
        if (check)
        {
            int tot = 0;
#pragma unroll
            for (int t = 0; t < kGoalCluster; ++t)
                tot |= slots[par][t];
            if (!tot)
                break; // uniform across the cluster: frontier exhausted
        }
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
        const int band = (rows + kGoalCluster - 1) / kGoalCluster;
        const size_t smem = (size_t)3 * band * nw * sizeof(uint32_t);
        const char *mode = std::getenv("MPROS_GOAL_MODE"); // "grid": force the cooperative kernel (tests compare the two)
        if (smem <= 200 * 1024 && !(mode && std::strcmp(mode, "grid") == 0))
        {
            static bool attr_set[16] = {};
            if (!attr_set[c->device & 15])
            {
                MPROS_CUDA(cudaFuncSetAttribute(goal_bfs_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr_set[c->device & 15] = true;
            }
            const size_t cells = (size_t)c->H * c->W;
            fill_u16_kernel<<<148 * 4, 256, 0, c->stream>>>(c->goal, cells, (uint16_t)kHugeInt);
            MPROS_LAUNCH_CHECK(c);
            goal_bfs_cluster_kernel<<<kGoalCluster, kGoalThreads, smem, c->stream>>>(c->ds_bits, c->H, c->W, c->pitch, gi, gj, c->goal, r0, rows,
                                                                                 w0, nw, band);
            MPROS_LAUNCH_CHECK(c);
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
    cudaDeviceProp prop;
    MPROS_CUDA(cudaGetDeviceProperties(&prop, c->device));
    if (per_sm < 1)
        per_sm = 1;
    if (per_sm > 4)
        per_sm = 4; // the per-level work is small; fewer CTAs make grid.sync cheaper
    int grid = prop.multiProcessorCount * per_sm;
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
