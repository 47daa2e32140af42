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
