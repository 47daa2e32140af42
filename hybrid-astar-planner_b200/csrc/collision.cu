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

static int launch_poses(Ctx *c, const double *d_xyyaw, size_t n, uint8_t *d_verdicts)
{
    // grid = multiple of the SM count, 8 resident CTAs of 256 threads each; grid-stride covers the rest
    size_t blocks = (n + 255) / 256;
    size_t cap = (size_t)c->n_sm * 32;
    if (blocks > cap)
        blocks = cap;
    // MPROS_C1_MODE=walk: every pose through the probe walk (the single-phase kernel), for A/B measurements and tests
    if (c->opt.c1_walk)
    {
        collision_poses_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(c->view(), c->pv, d_xyyaw, n, d_verdicts);
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
        collision_poses_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(c->view(), c->pv, d_xyyaw, n, d_verdicts);
        MPROS_LAUNCH_CHECK(c);
        return MPROS_OK;
    }
    collision_poses_2phase_kernel<<<(unsigned)chunks, 256, 0, c->stream>>>(c->view(), c->pv, c->unsafe_tw, c->unsafe_center_probe, d_xyyaw, n,
                                                                         d_verdicts);
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
    return launch_poses(c, d_xyyaw, n, d_verdicts);
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
    MPROS_CUDA(cudaMemcpyAsync(d_in, xyyaw, in_bytes, cudaMemcpyHostToDevice, c->stream));
    rc = launch_poses(c, d_in, n, d_out);
    if (rc)
        return rc;
    MPROS_CUDA(cudaMemcpyAsync(verdicts, d_out, n, cudaMemcpyDeviceToHost, c->stream));
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
