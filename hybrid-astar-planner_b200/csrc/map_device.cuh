// Device-side map lookups and the method-4 footprint test, shared by the collision, Reeds-Shepp,
// expansion and planner kernels. All pose math is FP64 with IEEE division and no FMA contraction
// (the library is compiled with --fmad=false) so that cell indices match the reference's.
#pragma once
#include "mpros_internal.cuh"

namespace mpros
{

// NavMap::posToIndexOrig / posToIndex (nav_map.h:144-167): subtract origin, rotate by -map_yaw, DIVIDE by the
// resolution, truncate toward zero. Returned as doubles-before-truncation so that the in-map test can be done
// without an out-of-range float->int conversion (x86 yields INT_MIN there, i.e. "outside").
__device__ __forceinline__ void pos_to_grid(const MapView &m, double x, double y, double res, double &qi, double &qj)
{
    double dx = x - m.ox;
    double dy = y - m.oy;
    qj = (dx * m.ocos + dy * m.osin) / res;
    qi = (dx * -m.osin + dy * m.ocos) / res;
}

// int(q) lies in [0, n) <=> q > -1 && q < n (truncation toward zero; NaN fails both -> outside)
__device__ __forceinline__ bool in_range(double q, int n) { return q > -1.0 && q < (double)n; }

__device__ __forceinline__ bool bit_at(const uint32_t *bits, int pitch, int i, int j)
{
    return (__ldg(bits + (size_t)i * pitch + (j >> 5)) >> (j & 31)) & 1u;
}

// NavMap::pointInCollisionOrig / pointInCollisionINFLOrig on a world point (nav_map.cpp:259-277):
// outside the map counts as collision.
__device__ __forceinline__ bool point_hits_orig(const MapView &m, const uint32_t *bits, double x, double y)
{
    double qi, qj;
    pos_to_grid(m, x, y, m.res0, qi, qj);
    if (!in_range(qi, m.H0) || !in_range(qj, m.W0))
        return true;
    return bit_at(bits, m.pitch0, (int)qi, (int)qj);
}

// NavMap::posToIndex on the downsampled grid (node cells, hybrid_a_star_node.h:45). Returns false when the
// reference would index outside the arrays (undefined behaviour there).
__device__ __forceinline__ bool pos_to_index_ds(const MapView &m, double x, double y, int &i, int &j)
{
    double qi, qj;
    pos_to_grid(m, x, y, m.res, qi, qj);
    bool ok = in_range(qi, m.H) && in_range(qj, m.W);
    i = ok ? (int)qi : -1;
    j = ok ? (int)qj : -1;
    return ok;
}

// HybridAstar::footPrintInCollision4 (hybrid_a_star.cpp:617-680).
//  (i)  axis line: endpoints R(yaw) * (longi_x{0,1}, 0) + p, n_check steps, points orient * i + col0,
//       i = 0..n, tested in the INFLATED map;
//  (ii) the 4 rectangle corners x*c - y*s + px, x*s + y*c + py tested in the RAW map.
// The verdict is an OR, so evaluation order / early exit do not change it.
__device__ __forceinline__ bool footprint_in_collision4(const MapView &m, const PlannerView &p, double px, double py,
                                                        double yaw)
{
    double s, c;
    sincos(yaw, &s, &c);
    // Eigen 2x2 * 2x2: a(i,0)*b(0,j) + a(i,1)*b(1,j) with b(1,j) == 0 -> the second term is a signed zero
    double x0 = c * p.longi_x0 + px, y0 = s * p.longi_x0 + py;
    double x1 = c * p.longi_x1 + px, y1 = s * p.longi_x1 + py;
    double n = (double)p.n_check;
    double ox = (x1 - x0) / n, oy = (y1 - y0) / n;
    for (int k = 0; k <= p.n_check; ++k)
    {
        double kk = (double)k;
        if (point_hits_orig(m, m.infl_bits, ox * kk + x0, oy * kk + y0))
            return true;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        double cx = p.corner_x[k], cy = p.corner_y[k];
        if (point_hits_orig(m, m.raw_bits, cx * c - cy * s + px, cx * s + cy * c + py))
            return true;
    }
    return false;
}

} // namespace mpros
