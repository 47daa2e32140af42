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
// Neither needs sin/cos or the probe walk. Phase 1 classifies kC1PosesPerThread poses per thread and compacts the
// undecided ones into shared memory; phase 2 runs the exact footprint test (fixed-point walk with its own exact
// fallback) on the dense list, so the warps of phase 2 are full although only ~30 % of the poses reach it.
constexpr int kC1PosesPerThread = 4;
__global__ void __launch_bounds__(256, MPROS_C1_MINB) collision_poses_2phase_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p,
                                                                                      const uint32_t *__restrict__ unsafe_tw, int center_probe,
                                                                                      const double *__restrict__ xyyaw, size_t n,
                                                                                      uint8_t *__restrict__ verdicts)
{
    constexpr int kChunk = 256 * kC1PosesPerThread;
    __shared__ uint16_t list[kChunk];
    __shared__ int count;
    const int tid = threadIdx.x, lane = tid & 31;
    const double kTwo32 = 4294967296.0;
    const double sc = m.inv_res0 * kTwo32;
    const double jc = m.ocos * sc, js = m.osin * sc;
    const double off = (double)kTwMargin * kTwo32;
    const double span = m.fx_span - p.fx_extent;
    const unsigned lim_j = (unsigned)(m.W0 + 2 * kTwMargin - 2 * p.fx_slack), lim_i = (unsigned)(m.H0 + 2 * kTwMargin - 2 * p.fx_slack);
    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    for (size_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x)
    {
        const size_t base = chunk * kChunk;
        if (tid == 0)
            count = 0;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kC1PosesPerThread; ++r)
        {
            const int kl = r * 256 + tid;
            const size_t k = base + kl;
            bool undecided = false;
            if (k < n)
            {
                const double x = __ldg(xyyaw + 3 * k), y = __ldg(xyyaw + 3 * k + 1), yaw = __ldg(xyyaw + 3 * k + 2);
                undecided = true;
                // same preconditions as the fixed-point walk (pose within the error bound's range, finite) + a finite yaw
                if (fabs(x) + fabs(y) < span && fabs(yaw) <= 1.7976931348623157e308)
                {
                    const double ax = x - m.ox, ay = y - m.oy;
                    const long long qj = __double2ll_rd(ax * jc + ay * js + off), qi = __double2ll_rd(ay * jc - ax * js + off);
                    const bool in = ((unsigned)((int)(qj >> 32) - p.fx_slack) < lim_j) & ((unsigned)((int)(qi >> 32) - p.fx_slack) < lim_i);
                    if (in)
                    {
                        const uint32_t us = tw_probe(unsafe_tw, m.tw_pitch, qj, qi), hit = tw_probe(m.infl_tw, m.tw_pitch, qj, qi);
                        if (!(us & 1u))
                        {
                            verdicts[k] = 0;
                            undecided = false;
                        }
                        else if (center_probe && (hit & 1u) && tw_edge(0xffffffffu, qj, qi) >= 2 * kFxEps)
                        {
                            verdicts[k] = 1;
                            undecided = false;
                        }
                    }
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, undecided);
            if (mask)
            {
                int slot = 0;
                if (lane == __ffs(mask) - 1)
                    slot = atomicAdd(&count, __popc(mask));
                slot = __shfl_sync(0xffffffffu, slot, __ffs(mask) - 1);
                if (undecided)
                    list[slot + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)kl;
            }
        }
        __syncthreads();
        const int cnt = count;
        for (int t = tid; t < cnt; t += 256)
        {
            const size_t k = base + list[t];
            verdicts[k] = footprint_in_collision4(m, p, __ldg(xyyaw + 3 * k), __ldg(xyyaw + 3 * k + 1), __ldg(xyyaw + 3 * k + 2)) ? 1 : 0;
        }
        __syncthreads();
    }
}

// One warp per path: lanes stride over the poses, ballot gives the OR with an early exit per 32-pose round.
__global__ void __launch_bounds__(256) collision_paths_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p, const double *__restrict__ xyyaw,
                                                              const int64_t *__restrict__ offsets, size_t n_paths,
                                                              uint8_t *__restrict__ verdicts)
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
                h = footprint_in_collision4(m, p, xyyaw[3 * k], xyyaw[3 * k + 1], xyyaw[3 * k + 2]);
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
    // grid = multiple of the 148 SMs, 8 resident CTAs of 256 threads each; grid-stride covers the rest
    size_t blocks = (n + 255) / 256;
    size_t cap = 148 * 32;
    if (blocks > cap)
        blocks = cap;
    // MPROS_C1_MODE=walk: every pose through the probe walk (the single-phase kernel), for A/B measurements and tests
    const char *mode = std::getenv("MPROS_C1_MODE");
    if (mode && std::strcmp(mode, "walk") == 0)
    {
        collision_poses_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(c->view(), c->pv, d_xyyaw, n, d_verdicts);
        MPROS_LAUNCH_CHECK(c);
        return MPROS_OK;
    }
    int rc = map_build_unsafe(c);
    if (rc)
        return rc;
    size_t chunks = (n + 256 * kC1PosesPerThread - 1) / (256 * kC1PosesPerThread);
    if (chunks > cap)
        chunks = cap;
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
    if (blocks > 148 * 32)
        blocks = 148 * 32;
    collision_paths_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(c->view(), c->pv, d_in, d_off, n_paths, d_out);
    MPROS_LAUNCH_CHECK(c);
    MPROS_CUDA(cudaMemcpyAsync(verdicts, d_out, n_paths, cudaMemcpyDeviceToHost, c->stream));
    MPROS_CUDA(cudaStreamSynchronize(c->stream));
    return MPROS_OK;
}
