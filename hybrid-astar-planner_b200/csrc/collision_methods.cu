// the three footprint tests of the public HybridAstar interface that no caller of the reference
// uses (method 4, footPrintInCollision4, is the hot one: collision.cu).
//   method 1  footPrintInCollision   hybrid_a_star.cpp:432-483  raw-map cells on the Bresenham lines (LineIterator,
//                                    mpros_utils/line_iterator.h) between the grid cells of the four rectangle corners
//   method 2  footPrintInCollision2  :485-575  winding-angle sum of the obstacle cells inside the circumscribed square
//   method 3  footPrintInCollision3  :577-615  AND of the raw map with a rasterised footprint kernel per yaw bin
//                                    (Footprint, mpros_utils/footprint.h:33-245), window outside the map = collision
// One warp per pose. Methods 1 and 3 are integer / bit work after the cell index of the pose (reference arithmetic, IEEE
// division); method 2 is FP64 transcendental work whose verdict is a threshold on a sum of angles (2 pi per obstacle cell
// inside the polygon, ~0 per cell outside), so last-ulp differences of the device libm cannot move it except for a cell
// that lies exactly on an edge.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "map_device.cuh"

namespace mpros
{
namespace
{
// ---- host: Footprint::configure / setKernel (footprint.h:33-126, 186-228) ------------------------------------------
// LineIterator's point set in closed form: after the steep / order swaps point k has x = x1 + k and
// y = y1 + ystep * max(0, ceil((k * deltay - deltax / 2) / deltax)) (line_iterator.h:31-65).
struct HostLine
{
    std::vector<int> xs, ys;
    HostLine(int x1, int y1, int x2, int y2)
    {
        const bool steep = std::abs(y2 - y1) > std::abs(x2 - x1);
        if (steep)
        {
            std::swap(x1, y1);
            std::swap(x2, y2);
        }
        if (x1 > x2)
        {
            std::swap(x1, x2);
            std::swap(y1, y2);
        }
        const int deltax = x2 - x1, deltay = std::abs(y2 - y1);
        int error = deltax / 2;
        const int ystep = y1 < y2 ? 1 : -1;
        int y = y1;
        for (int x = x1; x <= x2; ++x)
        {
            xs.push_back(steep ? y : x);
            ys.push_back(steep ? x : y);
            error -= deltay;
            if (error < 0)
            {
                y += ystep;
                error += deltax;
            }
        }
    }
};

// all_kernel_ as bit rows: kernel[yaw][i][w] holds columns 32 w .. 32 w + 31 of row i
void build_footprint_kernels(double len, double wid, double xoff, double reso, double yaw_reso, int &num_yaw, int &ksize, int &origin,
                             std::vector<uint32_t> &bits)
{
    const double yoff = 0.0;
    num_yaw = (int)std::ceil(2 * M_PI / yaw_reso);
    const double x_front = len / 2 + xoff, x_back = len / 2 - xoff, y_left = wid / 2 + yoff, y_right = wid / 2 - yoff;
    ksize = (int)(2 * std::ceil(std::hypot(std::fmax(x_front, x_back), std::fmax(y_left, y_right)) / reso) + 1);
    const double origin_d = ksize / 2; // integer division, stored in a double member (footprint.h:49)
    origin = (int)origin_d;
    const int kw = (ksize + 31) / 32;
    bits.assign((size_t)num_yaw * ksize * kw, 0u);
    std::vector<std::vector<int8_t>> k;
    for (int iy = 0; iy < num_yaw; ++iy)
    {
        const double yaw = iy * yaw_reso, sn = std::sin(yaw), cs = std::cos(yaw);
        // rotate = {{c, s}, {-s, c}}; corners = {{xf, -xb, -xb, xf}, {yl, yl, -yr, -yr}}; (rotate * corners) / reso + origin
        const double cx[4] = {x_front, -x_back, -x_back, x_front}, cy[4] = {y_left, y_left, -y_right, -y_right};
        int xv[4], yv[4];
        for (int q = 0; q < 4; ++q)
        {
            double r0 = cs * cx[q];
            r0 = r0 + sn * cy[q];
            double r1 = -sn * cx[q];
            r1 = r1 + cs * cy[q];
            xv[q] = (int)(r0 / reso + origin_d);
            yv[q] = (int)(r1 / reso + origin_d);
        }
        // fillRectangular(k, y_vec, x_vec): the roles of x and y are swapped by the call (footprint.h:121)
        k.assign(ksize, std::vector<int8_t>(ksize, 0));
        const int *X = yv, *Y = xv;
        HostLine line1(X[0], Y[0], X[1], Y[1]), line2(X[1], Y[1], X[2], Y[2]);
        const int x1 = X[1], y1 = Y[1];
        for (size_t a = 0; a < line1.xs.size(); ++a)
            for (size_t b = 0; b < line2.xs.size(); ++b)
            {
                const int ii = line1.xs[a] + (line2.xs[b] - x1), jj = line1.ys[a] + (line2.ys[b] - y1);
                if (ii >= 0 && ii < ksize && jj >= 0 && jj < ksize) // outside would be out-of-bounds writes in the reference
                    k[ii][jj] = 1;
            }
        for (int i = 1; i < ksize - 1; ++i) // "fill the missed grids", in place and in this order
            for (int j = 1; j < ksize - 1; ++j)
                if (k[i - 1][j] && k[i + 1][j] && k[i][j - 1] && k[i][j + 1])
                    k[i][j] = 1;
        for (int i = 0; i < ksize; ++i)
            for (int j = 0; j < ksize; ++j)
                if (k[i][j])
                    bits[((size_t)iy * ksize + i) * kw + (j >> 5)] |= 1u << (j & 31);
    }
}
} // namespace

struct MethodArgs
{
    const double *xyyaw;
    size_t n;
    uint8_t *verdicts;
    // method 2
    int r_in_pixel;
    double fx[4], fy[4]; // footprint_x_ / footprint_y_ (hybrid_a_star.h:388-396: built from the DEFAULT vehicle size)
    // method 3
    const uint32_t *kernels;
    int num_yaw, ksize, origin;
    double yaw_resolution;
};

// NavMap::posToIndexOrig (nav_map.h:159-166): truncating conversion; far-away values saturate like x86's cvttsd2si
__device__ __forceinline__ int trunc_index(double q)
{
    if (!(q > -2147483649.0 && q < 2147483648.0))
        return (int)0x80000000;
    return (int)q;
}
__device__ __forceinline__ void pos_to_index_orig(const MapView &m, double x, double y, int &i, int &j)
{
    double qi, qj;
    pos_to_grid(m, x, y, m.res0, qi, qj);
    i = trunc_index(qi);
    j = trunc_index(qj);
}
// NavMap::pointInCollisionOrig (nav_map.cpp:259-267)
__device__ __forceinline__ bool cell_hits_raw(const MapView &m, int i, int j)
{
    if (i < 0 || i >= m.H0 || j < 0 || j >= m.W0)
        return true;
    return bit_at(m.raw_bits, m.pitch0, i, j);
}

// ---- method 1 ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) collision_method1_kernel(const __grid_constant__ MapView m, const __grid_constant__ PlannerView p,
                                                                const __grid_constant__ MethodArgs a)
{
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t q = warp; q < a.n; q += nwarps)
    {
        const double px = a.xyyaw[3 * q], py = a.xyyaw[3 * q + 1], yaw = a.xyyaw[3 * q + 2];
        const double c = cos(yaw), s = sin(yaw);
        int ci[4], cj[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const double x = p.corner_x[k], y = p.corner_y[k];
            pos_to_index_orig(m, x * c - y * s + px, x * s + y * c + py, ci[k], cj[k]);
        }
        bool hit = false;
        for (int e = 0; e < 4 && !hit; ++e)
        {
            // LineIterator(idx_i1, idx_j1, idx_i2, idx_j2): "x" is the row index, "y" the column index
            int x1 = ci[e], y1 = cj[e], x2 = ci[(e + 1) & 3], y2 = cj[(e + 1) & 3];
            const bool steep = abs(y2 - y1) > abs(x2 - x1);
            if (steep)
            {
                int t = x1; x1 = y1; y1 = t;
                t = x2; x2 = y2; y2 = t;
            }
            if (x1 > x2)
            {
                int t = x1; x1 = x2; x2 = t;
                t = y1; y1 = y2; y2 = t;
            }
            const long long deltax = (long long)x2 - x1, deltay = llabs((long long)y2 - y1), e0 = deltax / 2;
            const int ystep = y1 < y2 ? 1 : -1;
            for (long long base = 0; base <= deltax && !hit; base += 32)
            {
                const long long k = base + lane;
                bool h = false;
                if (k <= deltax)
                {
                    long long cstep = 0;
                    if (k * deltay > e0 && deltax > 0)
                        cstep = (k * deltay - e0 + deltax - 1) / deltax;
                    const int x = (int)(x1 + k), y = (int)(y1 + ystep * cstep);
                    h = steep ? cell_hits_raw(m, y, x) : cell_hits_raw(m, x, y);
                }
                hit = __any_sync(0xffffffffu, h);
            }
        }
        if (lane == 0)
            a.verdicts[q] = hit ? 1 : 0;
    }
}

// ---- method 2 ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) collision_method2_kernel(const __grid_constant__ MapView m, const __grid_constant__ MethodArgs a)
{
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t q = warp; q < a.n; q += nwarps)
    {
        const double x = a.xyyaw[3 * q], y = a.xyyaw[3 * q + 1], yaw = a.xyyaw[3 * q + 2];
        int pos_i, pos_j;
        pos_to_index_orig(m, x, y, pos_i, pos_j);
        const int r = a.r_in_pixel;
        // (int)fmin(pos + r, size) / (int)fmax(pos - r, 0) with the additions done in int like the reference
        const int max_i = (int)fmin((double)(pos_i + r), (double)m.H0), max_j = (int)fmin((double)(pos_j + r), (double)m.W0);
        const int min_i = (int)fmax((double)(pos_i - r), 0.0), min_j = (int)fmax((double)(pos_j - r), 0.0);
        const double s = sin(yaw), c = cos(yaw);
        const long long wj = (long long)max_j - min_j + 1, wi = (long long)max_i - min_i + 1;
        double sum = 0.0;
        bool any = false;
        if (wi > 0 && wj > 0)
            for (long long t = lane; t < wi * wj; t += 32)
            {
                const int i = min_i + (int)(t / wj), j = min_j + (int)(t % wj);
                if (!cell_hits_raw(m, i, j))
                    continue;
                any = true;
                // indexToPosOrig (nav_map.h:193-200)
                const double dx = ((double)j + 0.5) * m.res0, dy = ((double)i + 0.5) * m.res0;
                const double ob_x = (dx * m.ocos - dy * m.osin) + m.ox, ob_y = (dx * m.osin + dy * m.ocos) + m.oy;
                const double obx = c * (ob_x - x) + s * (ob_y - y), oby = -s * (ob_x - x) + c * (ob_y - y);
                for (int e = 0; e < 3; ++e) // footprint_x_.size() - 1 edges: the closing edge is not visited
                {
                    const double dx1 = a.fx[e] - obx, dy1 = a.fy[e] - oby, dx2 = a.fx[e + 1] - obx, dy2 = a.fy[e + 1] - oby;
                    const double theta1 = atan2(dy1, dx1);
                    double cos_angle = (dx1 * dx2 + dy1 * dy2) / (hypot(dx1, dy1) * hypot(dx2, dy2));
                    if (cos_angle >= 1.0)
                        cos_angle = 1.0;
                    if (-sin(theta1) * dx2 + cos(theta1) * dy2 > 0)
                        sum += acos(cos_angle);
                    else
                        sum -= acos(cos_angle);
                }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
        any = __any_sync(0xffffffffu, any);
        if (lane == 0)
            a.verdicts[q] = (any && sum >= 3.14159265358979323846) ? 1 : 0;
    }
}

// ---- method 3 ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t map_bits32(const MapView &m, int i, int j0)
{
    // 32 raw-map cells of row i starting at column j0 (0 <= j0, j0 + 31 may run past the row: padding bits are zero)
    const uint32_t *row = m.raw_bits + (size_t)i * m.pitch0;
    const int w = j0 >> 5, sh = j0 & 31;
    const uint32_t lo = row[w];
    const uint32_t hi = (w + 1 < m.pitch0) ? row[w + 1] : 0u;
    return __funnelshift_r(lo, hi, sh);
}
__global__ void __launch_bounds__(256) collision_method3_kernel(const __grid_constant__ MapView m, const __grid_constant__ MethodArgs a)
{
    const int lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int kw = (a.ksize + 31) / 32;
    for (size_t q = warp; q < a.n; q += nwarps)
    {
        const double x = a.xyyaw[3 * q], y = a.xyyaw[3 * q + 1];
        double yaw = a.xyyaw[3 * q + 2];
        // Footprint::regularYaw (footprint.h:166-177) + getKernel's bin
        const double twopi = 2.0 * 3.14159265358979323846;
        while (yaw < 0)
            yaw = fmod(yaw, -twopi) + twopi;
        while (yaw > twopi)
            yaw = fmod(yaw, twopi);
        int iy = (int)floor(yaw / a.yaw_resolution);
        iy = min(max(iy, 0), a.num_yaw - 1); // yaw == 2 pi exactly indexes past the reference's table
        int i0, j0;
        pos_to_index_orig(m, x, y, i0, j0);
        const long long top = (long long)i0 - a.origin, left = (long long)j0 - a.origin;
        bool hit = top < 0 || top + a.ksize > m.H0 || left < 0 || left + a.ksize > m.W0; // any window cell outside the map
        if (!hit)
        {
            const uint32_t *kern = a.kernels + (size_t)iy * a.ksize * kw;
            bool h = false;
            for (int r = lane; r < a.ksize; r += 32)
                for (int w = 0; w < kw; ++w)
                    h |= (map_bits32(m, (int)top + r, (int)left + 32 * w) & kern[(size_t)r * kw + w]) != 0;
            hit = __any_sync(0xffffffffu, h);
        }
        if (lane == 0)
            a.verdicts[q] = hit ? 1 : 0;
    }
}
} // namespace mpros

using namespace mpros;

extern "C" int mpros_collision_check_poses_method(mpros_ctx *c, int method, const double *xyyaw, size_t n, uint8_t *verdicts)
{
    if (!c)
    {
        set_error("mpros_collision_check_poses_method: NULL context");
        return MPROS_ERR_INVALID;
    }
    if (!c->get_map || !c->planner_configured)
    {
        set_error("mpros_collision_check_poses_method: needs a map and mpros_planner_config first");
        return MPROS_ERR_INVALID;
    }
    if (method == 4)
        return mpros_collision_check_poses(c, xyyaw, n, verdicts);
    if (method < 1 || method > 3)
    {
        set_error("mpros_collision_check_poses_method: method must be 1, 2, 3 or 4");
        return MPROS_ERR_INVALID;
    }
    if (n == 0)
        return MPROS_OK;
    if (!xyyaw || !verdicts)
    {
        set_error("mpros_collision_check_poses_method: NULL buffer");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    MethodArgs a;
    std::memset(&a, 0, sizeof(a));
    std::vector<uint32_t> kbits;
    if (method == 3)
    {
        // footprint_kernel_.configure(V_LENGTH_, V_WIDTH_, V_STEER_OFFSET_, 0, map_resolution_orig_, 2 pi / num_yaw_) (:204)
        a.yaw_resolution = 2.0 * M_PI / c->pp.yaw_discretization;
        build_footprint_kernels(c->pp.length, c->pp.width, c->pp.steer_offset, c->res0, a.yaw_resolution, a.num_yaw, a.ksize, a.origin, kbits);
    }
    const size_t in_bytes = n * 24, kern_bytes = kbits.size() * 4;
    const size_t kern_off = (in_bytes + n + 255) / 256 * 256;
    int rc = ensure_stage(c, kern_off + kern_bytes + 256);
    if (rc)
        return rc;
    char *base = static_cast<char *>(c->stage);
    MPROS_CUDA(cudaMemcpyAsync(base, xyyaw, in_bytes, cudaMemcpyHostToDevice, st));
    if (kern_bytes)
        MPROS_CUDA(cudaMemcpyAsync(base + kern_off, kbits.data(), kern_bytes, cudaMemcpyHostToDevice, st));
    a.xyyaw = reinterpret_cast<const double *>(base);
    a.n = n;
    a.verdicts = reinterpret_cast<uint8_t *>(base + in_bytes);
    a.kernels = reinterpret_cast<const uint32_t *>(base + kern_off);
    {
        // footprint_r_ (hybrid_a_star.cpp:202) from the configured corners; footprint_x_/y_ from the header's defaults
        const PlannerView &v = c->pv;
        const double n0 = std::sqrt(v.corner_x[0] * v.corner_x[0] + v.corner_y[0] * v.corner_y[0]);
        const double n2 = std::sqrt(v.corner_x[2] * v.corner_x[2] + v.corner_y[2] * v.corner_y[2]);
        a.r_in_pixel = (int)std::ceil(std::fmax(n0, n2) / c->res0);
        const double L = 4.2, W = 2.0, off = 0.0;
        const double fx[4] = {L / 2 + off, L / 2 + off, -L / 2 + off, -L / 2 + off}, fy[4] = {W / 2, -W / 2, -W / 2, W / 2};
        for (int k = 0; k < 4; ++k)
            a.fx[k] = fx[k], a.fy[k] = fy[k];
    }
    size_t blocks = (n * 32 + 255) / 256;
    if (blocks > (size_t)c->n_sm * 16)
        blocks = (size_t)c->n_sm * 16;
    if (method == 1)
        collision_method1_kernel<<<(unsigned)blocks, 256, 0, st>>>(c->view(), c->pv, a);
    else if (method == 2)
        collision_method2_kernel<<<(unsigned)blocks, 256, 0, st>>>(c->view(), a);
    else
        collision_method3_kernel<<<(unsigned)blocks, 256, 0, st>>>(c->view(), a);
    MPROS_LAUNCH_CHECK(c);
    MPROS_CUDA(cudaMemcpyAsync(verdicts, a.verdicts, n, cudaMemcpyDeviceToHost, st));
    MPROS_CUDA(cudaStreamSynchronize(st));
    return MPROS_OK;
}
