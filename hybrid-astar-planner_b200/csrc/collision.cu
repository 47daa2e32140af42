// C1/C2: batched HybridAstar::footPrintInCollision4 / pathInCollision (hybrid_a_star.cpp:404-430, 617-680).
//
// One thread per pose. The probes of one pose fall on a <= (n+1)-cell line plus 4 corners, i.e. a handful of
// 32-byte sectors of the bit-packed layers, which are L2-resident (8 MiB per 8192^2 layer). The kernel is bound
// by FP64 issue (two IEEE divisions per probe) and L2 sector throughput, not by HBM; the algorithmic HBM traffic
// is the 24 B pose read + 1 B verdict write.
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
    collision_poses_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(c->view(), c->pv, d_xyyaw, n, d_verdicts);
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
