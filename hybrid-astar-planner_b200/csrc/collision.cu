// C1/C2: batched HybridAstar::footPrintInCollision4 / pathInCollision (hybrid_a_star.cpp:404-430, 617-680).
//
// One thread per pose. The probes of one pose fall on a <= (n+1)-cell line plus 4 corners, i.e. a handful of
// 32-byte sectors of the bit-packed layers, which are L2-resident (8 MiB per 8192^2 layer). The kernel is bound
// by FP64 issue (two IEEE divisions per probe) and L2 sector throughput, not by HBM; the algorithmic HBM traffic
// is the 24 B pose read + 1 B verdict write.
#include <cstdlib>
#include <cstring>

#include "map_device.cuh"

namespace mpros
{

#ifndef MPROS_C1_MINB
#define MPROS_C1_MINB 4
#endif
__global__ void __launch_bounds__(256, MPROS_C1_MINB) collision_poses_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p, const double *__restrict__ xyyaw, size_t n,
                                                              uint8_t *__restrict__ verdicts)
{
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    {
        double x = __ldg(xyyaw + 3 * k), y = __ldg(xyyaw + 3 * k + 1), yaw = __ldg(xyyaw + 3 * k + 2);
        verdicts[k] = footprint_in_collision4(m, p, x, y, yaw) ? 1 : 0;
    }
}

// Two-phase form of the same batch (default). Most uniformly distributed poses are decided by the cell of the pose alone:
//   * SURELY FREE: no cell any probe can fall into is occupied or outside the map (one bit of Ctx::unsafe_tw, built by
//     map_build_unsafe from both collision layers and the footprint) -> false;
//   * SURE HIT: the axis line's middle probe (k = n_check / 2, when n_check is even and the axis symmetric) is the pose
//     itself up to ~1e-12 cells of rounding, so if the pose lies at least 2^-22 cells away from every cell edge and its
//     cell is occupied in the inflated layer the reference's probe hits it too -> true.
// Neither needs sin/cos or the probe walk. Everything is warp-local (no block barrier, no atomics): a warp classifies
// 32 x kC1PosesPerLane poses (all pose loads, then all probe loads in flight together), appends the undecided ones to its
// own ring in shared memory, and runs the exact footprint test (fixed-point walk with its own exact fallback) on FULL
// rounds of 32 entries; what is left (< 32) waits for the next chunk. Only ~30 % of uniform poses reach the walk and the
// warps that run it are full.
constexpr int kC1PosesPerLane = 4;
constexpr int kC1Ring = 256; // entries per warp ring (power of two, > 31 + 32 * kC1PosesPerLane)
__global__ void __launch_bounds__(256, MPROS_C1_MINB) collision_poses_2phase_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p,
                                                                                      const uint32_t *__restrict__ unsafe_tw, int center_probe,
                                                                                      const double *__restrict__ xyyaw, size_t n,
                                                                                      uint8_t *__restrict__ verdicts)
{
    constexpr int kChunk = 32 * kC1PosesPerLane;
    __shared__ uint32_t ring_all[8][kC1Ring];
    const int lane = threadIdx.x & 31;
    uint32_t *ring = ring_all[threadIdx.x >> 5];
    const double kTwo32 = 4294967296.0;
    const double sc = m.inv_res0 * kTwo32;
    const double jc = m.ocos * sc, js = m.osin * sc;
    const double off = (double)kTwMargin * kTwo32;
    const double span = m.fx_span - p.fx_extent;
    const unsigned lim_j = (unsigned)(m.W0 + 2 * kTwMargin - 2 * p.fx_slack), lim_i = (unsigned)(m.H0 + 2 * kTwMargin - 2 * p.fx_slack);
    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    unsigned head = 0, tail = 0; // warp-uniform ring positions
    for (size_t chunk = warp_global; chunk < n_chunks; chunk += n_warps)
    {
        const size_t base = chunk * kChunk;
        double x[kC1PosesPerLane], y[kC1PosesPerLane], yaw[kC1PosesPerLane];
#pragma unroll
        for (int r = 0; r < kC1PosesPerLane; ++r)
        {
            const size_t k = base + r * 32 + lane;
            const bool have = k < n;
            x[r] = have ? __ldg(xyyaw + 3 * k) : 0.0;
            y[r] = have ? __ldg(xyyaw + 3 * k + 1) : 0.0;
            yaw[r] = have ? __ldg(xyyaw + 3 * k + 2) : 0.0;
        }
        uint32_t us[kC1PosesPerLane], hit[kC1PosesPerLane];
        bool pre[kC1PosesPerLane], centred[kC1PosesPerLane];
#pragma unroll
        for (int r = 0; r < kC1PosesPerLane; ++r)
        {
            // same preconditions as the fixed-point walk (pose within the error bound's range, finite) + a finite yaw
            pre[r] = fabs(x[r]) + fabs(y[r]) < span && fabs(yaw[r]) <= 1.7976931348623157e308;
            const double ax = x[r] - m.ox, ay = y[r] - m.oy;
            long long qj = __double2ll_rd(ax * jc + ay * js + off), qi = __double2ll_rd(ay * jc - ax * js + off);
            pre[r] = pre[r] && (((unsigned)((int)(qj >> 32) - p.fx_slack) < lim_j) & ((unsigned)((int)(qi >> 32) - p.fx_slack) < lim_i));
            if (!pre[r])
                qj = qi = (long long)kTwMargin << 32; // any valid address; the result is not used
            us[r] = tw_probe(unsafe_tw, m.tw_pitch, qj, qi);
            hit[r] = tw_probe(m.infl_tw, m.tw_pitch, qj, qi);
            centred[r] = tw_edge(0xffffffffu, qj, qi) >= 2 * kFxEps;
        }
#pragma unroll
        for (int r = 0; r < kC1PosesPerLane; ++r)
        {
            const size_t k = base + r * 32 + lane;
            bool undecided = k < n;
            if (undecided && pre[r])
            {
                if (!(us[r] & 1u))
                {
                    verdicts[k] = 0;
                    undecided = false;
                }
                else if (center_probe && (hit[r] & 1u) && centred[r])
                {
                    verdicts[k] = 1;
                    undecided = false;
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, undecided);
            if (undecided)
                ring[(tail + __popc(mask & ((1u << lane) - 1u))) & (kC1Ring - 1)] = (uint32_t)k;
            tail += __popc(mask);
        }
        __syncwarp();
        while (tail - head >= 32)
        {
            const size_t k = ring[(head + lane) & (kC1Ring - 1)];
            verdicts[k] = footprint_in_collision4(m, p, __ldg(xyyaw + 3 * k), __ldg(xyyaw + 3 * k + 1), __ldg(xyyaw + 3 * k + 2)) ? 1 : 0;
            head += 32;
        }
        __syncwarp();
    }
    if (lane < tail - head)
    {
        const size_t k = ring[(head + lane) & (kC1Ring - 1)];
        verdicts[k] = footprint_in_collision4(m, p, __ldg(xyyaw + 3 * k), __ldg(xyyaw + 3 * k + 1), __ldg(xyyaw + 3 * k + 2)) ? 1 : 0;
    }
}

// ---- the two-phase kernel with the pose stream staged through shared memory by the bulk-copy engine (default) -------------
// Same two rules, same verdicts. What changed against collision_poses_2phase_kernel (kept above for unaligned buffers):
//   * a warp's tile of 128 poses (3072 contiguous bytes) arrives by ONE cp.async.bulk (1-D TMA) into a per-warp double
//     buffer, completion on an mbarrier; the copy of tile i+2 is issued as soon as tile i has been read into registers, so
//     two tiles per warp are always in flight without holding a register (the old form issued twelve strided 8-byte loads
//     per lane and waited for them: stall_long_scoreboard 13 per issue);
//   * classification reads ONE 8-byte word pair {not-surely-free, inflated} (Ctx::cls_tw) instead of two scattered words;
//   * verdicts of a tile leave as 32 four-byte stores (ballot -> byte expansion) instead of 128 one-byte stores; the
//     undecided poses are overwritten later by the exact walk (same warp, ordered by __syncwarp).
constexpr int kC1Tile = 128;                 // poses per warp tile
constexpr int kC1TileBytes = kC1Tile * 24;   // 3072
constexpr int kC1Ring2 = 192;                // ring entries per warp (> 31 left over + 128 new), indexed modulo
constexpr int kC1Warps = 8;
// STAGES tiles of shared memory per warp: 2 = double buffer (two tiles in flight), 1 = single buffer (the next tile is copied
// while the current one is processed from registers), 0 = no staging (strided loads straight from global memory).
// Shared memory is carved out of the L1 cache the map probes live in: 2 stages x 8 warps x 4 CTAs leave 28 KB of L1 per SM
// (measured: L1 hit rate 49 % -> 23 %, 0.27 -> 0.58 ms), so the default is the smallest carve-out that still prefetches.
#ifndef MPROS_C1_STAGES
#define MPROS_C1_STAGES 0
#endif
constexpr int kC1DefaultStages = MPROS_C1_STAGES;
template <int STAGES>
__host__ __device__ constexpr int c1_warp_smem() { return STAGES * kC1TileBytes + kC1Ring2 * 4 + 16 * (STAGES > 0 ? STAGES : 1); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#ifndef MPROS_C1_STREAM
#define MPROS_C1_STREAM 1
#endif
__device__ __forceinline__ double c1_load(const double *p) { return MPROS_C1_STREAM ? __ldcs(p) : __ldg(p); }
__device__ __forceinline__ void c1_store(uint32_t *p, uint32_t v)
{
    if (MPROS_C1_STREAM)
        __stcs(p, v);
    else
        *p = v;
}

template <int STAGES>
__global__ void __launch_bounds__(kC1Warps * 32, MPROS_C1_MINB) collision_poses_tma_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p,
                                                                                          const uint2 *__restrict__ cls_tw, int center_probe,
                                                                                          const double *__restrict__ xyyaw, size_t n,
                                                                                          uint8_t *__restrict__ verdicts, int bulk_ok)
{
    extern __shared__ __align__(128) unsigned char c1_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NS = STAGES > 0 ? STAGES : 1;
    unsigned char *mine = c1_smem + (size_t)warp * c1_warp_smem<STAGES>();
    double *tile[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k)
        tile[k] = reinterpret_cast<double *>(mine + k * kC1TileBytes);
    uint32_t *ring = reinterpret_cast<uint32_t *>(mine + STAGES * kC1TileBytes);
    uint64_t *bar = reinterpret_cast<uint64_t *>(mine + STAGES * kC1TileBytes + kC1Ring2 * 4);
    if (STAGES > 0)
    {
        if (lane == 0)
        {
#pragma unroll
            for (int k = 0; k < NS; ++k)
                mbar_init(&bar[k], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    const double kTwo32 = 4294967296.0;
    const double sc = m.inv_res0 * kTwo32;
    const double jc = m.ocos * sc, js = m.osin * sc;
    const double off = (double)kTwMargin * kTwo32;
    const double span = m.fx_span - p.fx_extent;
    const unsigned lim_j = (unsigned)(m.W0 + 2 * kTwMargin - 2 * p.fx_slack), lim_i = (unsigned)(m.H0 + 2 * kTwMargin - 2 * p.fx_slack);
    const size_t n_tiles = (n + kC1Tile - 1) / kC1Tile;
    const size_t warp_global = (size_t)blockIdx.x * kC1Warps + warp, n_warps = (size_t)gridDim.x * kC1Warps;

    // fill one stage with tile t: one bulk copy when the tile is whole (and the buffer 16-byte aligned), else by the lanes
    auto issue = [&](size_t t, int s) {
        if (STAGES == 0 || t >= n_tiles)
            return;
        const size_t base = t * kC1Tile;
        const size_t have = n - base < (size_t)kC1Tile ? n - base : (size_t)kC1Tile;
        if (bulk_ok && have == (size_t)kC1Tile)
        {
            if (lane == 0)
            {
                mbar_expect_tx(&bar[s], kC1TileBytes);
                bulk_g2s(tile[s], xyyaw + 3 * base, kC1TileBytes, &bar[s]);
            }
        }
        else
        {
            for (int k = lane; k < (int)have * 3; k += 32)
                tile[s][k] = __ldg(xyyaw + 3 * base + k);
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&bar[s]);
        }
    };
#pragma unroll
    for (int k = 0; k < NS; ++k)
        issue(warp_global + k * n_warps, k);
    unsigned head = 0, tail = 0; // warp-uniform ring positions (monotonic; slots are taken modulo kC1Ring2)
    unsigned it = 0;
    for (size_t t = warp_global; t < n_tiles; t += n_warps, ++it)
    {
        const int s = STAGES == 2 ? (it & 1) : 0;
        if (STAGES > 0)
            mbar_wait(&bar[s], (STAGES == 2 ? (it >> 1) : it) & 1u);
        const size_t base = t * kC1Tile;
        // the tile's poses go straight from shared memory into their classification address (24-byte stride: conflict-free
        // 8-byte accesses); nothing of the pose itself is kept, the stage can be refilled right after
        uint32_t widx[4];
        unsigned sh[4]; // bit position inside the tile word | pre << 8 | centred << 9
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
            const bool have = base + r * 32 + lane < n;
            double x, y, yaw;
            if (STAGES > 0)
            {
                const double *q = tile[s] + 3 * (r * 32 + lane);
                x = q[0], y = q[1], yaw = q[2];
            }
            else
            {
                // streamed once: evict-first, so that the pose stream does not push the map's tile words out of L1
                const double *q = xyyaw + 3 * (base + r * 32 + lane);
                x = have ? c1_load(q) : 0.0, y = have ? c1_load(q + 1) : 0.0, yaw = have ? c1_load(q + 2) : 0.0;
            }
            // same preconditions as the fixed-point walk (pose within the error bound's range, finite) + a finite yaw
            bool pre = have && fabs(x) + fabs(y) < span && fabs(yaw) <= 1.7976931348623157e308;
            const double ax = x - m.ox, ay = y - m.oy;
            long long qj = __double2ll_rd(ax * jc + ay * js + off), qi = __double2ll_rd(ay * jc - ax * js + off);
            pre = pre && (((unsigned)((int)(qj >> 32) - p.fx_slack) < lim_j) & ((unsigned)((int)(qi >> 32) - p.fx_slack) < lim_i));
            if (!pre)
                qj = qi = (long long)kTwMargin << 32; // any valid address; the result is not used
            const int hj = (int)(qj >> 32), hi = (int)(qi >> 32);
#if MPROS_TW_BLOCK16
            widx[r] = (uint32_t)(((((hi >> 4) * m.tw_pitch + (hj >> 4)) << 3) | ((hi >> 1) & 6)) | ((hj >> 3) & 1));
#else
            widx[r] = (uint32_t)((hi >> 2) * m.tw_pitch + (hj >> 3));
#endif
            const bool centred = tw_edge(0xffffffffu, qj, qi) >= 2 * kFxEps;
            sh[r] = ((((unsigned)hi << 3) & 24u) | ((unsigned)hj & 7u)) | (pre ? 256u : 0u) | (centred ? 512u : 0u);
        }
        if (STAGES > 0)
        {
            __syncwarp();
            issue(t + NS * n_warps, s); // the stage is free again: prefetch NS tiles ahead
        }
        uint2 cw[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
            cw[r] = __ldg(cls_tw + widx[r]);
        unsigned hitmask[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
        {
            const bool have = base + r * 32 + lane < n;
            const bool pre = (sh[r] & 256u) != 0, centred = (sh[r] & 512u) != 0;
            const bool unsafe = (cw[r].x >> (sh[r] & 31u)) & 1u, infl = (cw[r].y >> (sh[r] & 31u)) & 1u;
            const bool free_ = pre && !unsafe;
            const bool hit = pre && unsafe && center_probe && infl && centred;
            const bool undecided = have && !free_ && !hit;
            hitmask[r] = __ballot_sync(0xffffffffu, hit);
            const unsigned um = __ballot_sync(0xffffffffu, undecided);
            if (undecided)
                ring[(tail + __popc(um & ((1u << lane) - 1u))) % kC1Ring2] = (uint32_t)(base + r * 32 + lane);
            tail += __popc(um);
        }
        // lane l stores the verdict bytes of poses base + 4 l .. 4 l + 3 = bits 4 (l & 7) .. + 3 of hitmask[l >> 3]
        {
            const unsigned hm = lane < 8 ? hitmask[0] : lane < 16 ? hitmask[1] : lane < 24 ? hitmask[2] : hitmask[3];
            const unsigned bits = (hm >> ((lane & 7) << 2)) & 0xfu;
            const uint32_t word = (bits & 1u) | ((bits & 2u) << 7) | ((bits & 4u) << 14) | ((bits & 8u) << 21);
            const size_t k0 = base + 4 * (size_t)lane;
            if (k0 + 4 <= n)
                c1_store(reinterpret_cast<uint32_t *>(verdicts + k0), word);
            else
                for (int b = 0; b < 4; ++b)
                    if (k0 + b < n)
                        verdicts[k0 + b] = (uint8_t)((word >> (8 * b)) & 1u);
        }
        __syncwarp(); // ring entries visible; the packed store precedes the walk's byte stores below
        while (tail - head >= 32)
        {
            const size_t k = ring[(head + lane) % kC1Ring2];
            verdicts[k] = footprint_in_collision4(m, p, __ldg(xyyaw + 3 * k), __ldg(xyyaw + 3 * k + 1), __ldg(xyyaw + 3 * k + 2)) ? 1 : 0;
            head += 32;
        }
        __syncwarp();
    }
    if (lane < tail - head)
    {
        const size_t k = ring[(head + lane) % kC1Ring2];
        verdicts[k] = footprint_in_collision4(m, p, __ldg(xyyaw + 3 * k), __ldg(xyyaw + 3 * k + 1), __ldg(xyyaw + 3 * k + 2)) ? 1 : 0;
    }
}

// One warp per path: lanes stride over the poses, ballot gives the OR with an early exit per 32-pose round.
__global__ void __launch_bounds__(256) collision_paths_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p, const double *__restrict__ xyyaw,
                                                              const int64_t *__restrict__ offsets, size_t n_paths,
                                                              uint8_t *__restrict__ verdicts, const uint32_t *__restrict__ unsafe_tw, int center_probe)
{
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t q = warp; q < n_paths; q += nwarps)
    {
        int64_t b = offsets[q], e = offsets[q + 1];
        bool hit = false;
        for (int64_t k0 = b; k0 < e && !hit; k0 += 32)
        {
            int64_t k = k0 + lane;
            bool h = false;
            if (k < e)
            {
                // decided by the pose's own cell where possible (footprint_preclass), the probe walk otherwise
                const double x = xyyaw[3 * k], y = xyyaw[3 * k + 1], yaw = xyyaw[3 * k + 2];
                const int cls = footprint_preclass(m, p, unsafe_tw, center_probe, x, y, yaw);
                h = cls == 1 || (cls == 2 && footprint_in_collision4(m, p, x, y, yaw));
            }
            hit = __any_sync(0xffffffffu, h);
        }
        if (lane == 0)
            verdicts[q] = hit ? 1 : 0;
    }
}

static int check_ready(Ctx *c, const char *who)
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

static int launch_poses(Ctx *c, cudaStream_t st, const double *d_xyyaw, size_t n, uint8_t *d_verdicts)
{
    // grid = multiple of the SM count, 8 resident CTAs of 256 threads each; grid-stride covers the rest
    size_t blocks = (n + 255) / 256;
    size_t cap = (size_t)c->n_sm * 32;
    if (blocks > cap)
        blocks = cap;
    // MPROS_C1_MODE=walk: every pose through the probe walk (the single-phase kernel), for A/B measurements and tests
    if (c->opt.c1_walk)
    {
        collision_poses_kernel<<<(unsigned)blocks, 256, 0, st>>>(c->view(), c->pv, d_xyyaw, n, d_verdicts);
        MPROS_LAUNCH_CHECK(c);
        return MPROS_OK;
    }
    int rc = map_build_unsafe(c);
    if (rc)
        return rc;
    size_t chunks = (n + 256 * kC1PosesPerLane - 1) / (256 * kC1PosesPerLane); // CTAs of 8 warps, one warp chunk each at least
    // persistent: one wave of resident CTAs, so that a warp sees many chunks and its last, partly filled round is rare
    if (chunks > (size_t)c->n_sm * MPROS_C1_MINB)
        chunks = (size_t)c->n_sm * MPROS_C1_MINB;
    if (n >= (1ull << 32)) // ring entries are 32-bit pose indices
    {
        collision_poses_kernel<<<(unsigned)blocks, 256, 0, st>>>(c->view(), c->pv, d_xyyaw, n, d_verdicts);
        MPROS_LAUNCH_CHECK(c);
        return MPROS_OK;
    }
    if (c->opt.c1_ldg || ((uintptr_t)d_verdicts & 3u))
    {
        // the form without shared-memory staging: option "c1_mode" = "2phase_ldg", or a verdict buffer that is not 4-byte aligned
        collision_poses_2phase_kernel<<<(unsigned)chunks, 256, 0, st>>>(c->view(), c->pv, c->unsafe_tw, c->unsafe_center_probe, d_xyyaw, n,
                                                                             d_verdicts);
        MPROS_LAUNCH_CHECK(c);
        return MPROS_OK;
    }
    size_t tiles = (n + (size_t)kC1Tile * kC1Warps - 1) / ((size_t)kC1Tile * kC1Warps);
    if (tiles > (size_t)c->n_sm * MPROS_C1_MINB)
        tiles = (size_t)c->n_sm * MPROS_C1_MINB; // persistent: one wave of resident CTAs
    const int bulk_ok = ((uintptr_t)d_xyyaw & 15u) == 0; // cp.async.bulk needs 16-byte aligned source and size
    const int stages = c->opt.c1_stages >= 0 ? c->opt.c1_stages : kC1DefaultStages;
    if (!c->c1_attr_set)
    {
        MPROS_CUDA(cudaFuncSetAttribute(collision_poses_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1Warps * c1_warp_smem<1>()));
        MPROS_CUDA(cudaFuncSetAttribute(collision_poses_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1Warps * c1_warp_smem<2>()));
        c->c1_attr_set = true;
    }
    const MapView mv = c->view();
    if (stages == 2)
        collision_poses_tma_kernel<2><<<(unsigned)tiles, kC1Warps * 32, kC1Warps * c1_warp_smem<2>(), st>>>(mv, c->pv, c->cls_tw, c->unsafe_center_probe,
                                                                                                          d_xyyaw, n, d_verdicts, bulk_ok);
    else if (stages == 1)
        collision_poses_tma_kernel<1><<<(unsigned)tiles, kC1Warps * 32, kC1Warps * c1_warp_smem<1>(), st>>>(mv, c->pv, c->cls_tw, c->unsafe_center_probe,
                                                                                                          d_xyyaw, n, d_verdicts, bulk_ok);
    else
        collision_poses_tma_kernel<0><<<(unsigned)tiles, kC1Warps * 32, kC1Warps * c1_warp_smem<0>(), st>>>(mv, c->pv, c->cls_tw, c->unsafe_center_probe,
                                                                                                          d_xyyaw, n, d_verdicts, bulk_ok);
    MPROS_LAUNCH_CHECK(c);
    return MPROS_OK;
}

} // namespace mpros

using namespace mpros;

extern "C" int mpros_collision_check_poses_dev(mpros_ctx *c, const double *d_xyyaw, size_t n, uint8_t *d_verdicts)
{
    int rc = check_ready(c, "mpros_collision_check_poses_dev");
    if (rc)
        return rc;
    if (n == 0)
        return MPROS_OK;
    if (!d_xyyaw || !d_verdicts)
    {
        set_error("mpros_collision_check_poses_dev: NULL buffer");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    return launch_poses(c, c->stream, d_xyyaw, n, d_verdicts);
}

extern "C" int mpros_collision_check_poses(mpros_ctx *c, const double *xyyaw, size_t n, uint8_t *verdicts)
{
    int rc = check_ready(c, "mpros_collision_check_poses");
    if (rc)
        return rc;
    if (n == 0)
        return MPROS_OK;
    if (!xyyaw || !verdicts)
    {
        set_error("mpros_collision_check_poses: NULL buffer");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    size_t in_bytes = n * 24;
    rc = ensure_stage(c, in_bytes + n);
    if (rc)
        return rc;
    double *d_in = static_cast<double *>(c->stage);
    uint8_t *d_out = reinterpret_cast<uint8_t *>(c->stage) + in_bytes;
    // Pieces of 2^20 poses alternate between the context's stream and a second one: the upload of piece i+1 (24 B per pose
    // over PCIe, the whole cost of this entry point) runs while piece i is checked and its verdicts travel back on the
    // other copy engine. Pieces are multiples of the kernel's tile, so every piece keeps the aligned bulk-copy path.
    constexpr size_t kPiece = (size_t)1 << 20;
    if (n <= kPiece)
    {
        MPROS_CUDA(cudaMemcpyAsync(d_in, xyyaw, in_bytes, cudaMemcpyHostToDevice, c->stream));
        rc = launch_poses(c, c->stream, d_in, n, d_out);
        if (rc)
            return rc;
        MPROS_CUDA(cudaMemcpyAsync(verdicts, d_out, n, cudaMemcpyDeviceToHost, c->stream));
        MPROS_CUDA(cudaStreamSynchronize(c->stream));
        return MPROS_OK;
    }
    if (!c->aux_stream)
    {
        MPROS_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        MPROS_CUDA(cudaEventCreateWithFlags(&c->aux_event, cudaEventDisableTiming));
    }
    if (!c->opt.c1_walk)
    {
        rc = map_build_unsafe(c); // on the context's stream, before either stream needs it
        if (rc)
            return rc;
    }
    MPROS_CUDA(cudaEventRecord(c->aux_event, c->stream));
    MPROS_CUDA(cudaStreamWaitEvent(c->aux_stream, c->aux_event, 0));
    int piece = 0;
    for (size_t k0 = 0; k0 < n; k0 += kPiece, ++piece)
    {
        const size_t cnt = n - k0 < kPiece ? n - k0 : kPiece;
        cudaStream_t st = (piece & 1) ? c->aux_stream : c->stream;
        MPROS_CUDA(cudaMemcpyAsync(d_in + 3 * k0, xyyaw + 3 * k0, cnt * 24, cudaMemcpyHostToDevice, st));
        rc = launch_poses(c, st, d_in + 3 * k0, cnt, d_out + k0);
        if (rc)
            return rc;
        MPROS_CUDA(cudaMemcpyAsync(verdicts + k0, d_out + k0, cnt, cudaMemcpyDeviceToHost, st));
    }
    MPROS_CUDA(cudaStreamSynchronize(c->aux_stream));
    MPROS_CUDA(cudaStreamSynchronize(c->stream));
    return MPROS_OK;
}

extern "C" int mpros_collision_check_paths(mpros_ctx *c, const double *xyyaw, const int64_t *offsets, size_t n_paths,
                                           uint8_t *verdicts)
{
    int rc = check_ready(c, "mpros_collision_check_paths");
    if (rc)
        return rc;
    if (n_paths == 0)
        return MPROS_OK;
    if (!offsets || !verdicts)
    {
        set_error("mpros_collision_check_paths: NULL buffer");
        return MPROS_ERR_INVALID;
    }
    int64_t n_poses = offsets[n_paths];
    if (n_poses < 0 || (n_poses > 0 && !xyyaw))
    {
        set_error("mpros_collision_check_paths: bad offsets / NULL poses");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    size_t pose_bytes = (size_t)n_poses * 24, off_bytes = (n_paths + 1) * 8;
    rc = ensure_stage(c, pose_bytes + off_bytes + n_paths + 16);
    if (rc)
        return rc;
    char *base = static_cast<char *>(c->stage);
    double *d_in = reinterpret_cast<double *>(base);
    int64_t *d_off = reinterpret_cast<int64_t *>(base + pose_bytes);
    uint8_t *d_out = reinterpret_cast<uint8_t *>(base + pose_bytes + off_bytes);
    if (pose_bytes)
        MPROS_CUDA(cudaMemcpyAsync(d_in, xyyaw, pose_bytes, cudaMemcpyHostToDevice, c->stream));
    MPROS_CUDA(cudaMemcpyAsync(d_off, offsets, off_bytes, cudaMemcpyHostToDevice, c->stream));
    size_t blocks = (n_paths * 32 + 255) / 256;
    if (blocks > (size_t)c->n_sm * 32)
        blocks = (size_t)c->n_sm * 32;
    rc = map_build_unsafe(c);
    if (rc)
        return rc;
    collision_paths_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(c->view(), c->pv, d_in, d_off, n_paths, d_out, c->unsafe_tw, c->unsafe_center_probe);
    MPROS_LAUNCH_CHECK(c);
    MPROS_CUDA(cudaMemcpyAsync(verdicts, d_out, n_paths, cudaMemcpyDeviceToHost, c->stream));
    MPROS_CUDA(cudaStreamSynchronize(c->stream));
    return MPROS_OK;
}
