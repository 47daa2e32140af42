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
//
// Reference arithmetic, probe by probe (FP64, two IEEE divisions per probe): the exact path, used as the fallback.
static __device__ __noinline__ bool footprint_in_collision4_exact(const MapView &m, const PlannerView &p, double px, double py, double s, double c)
{
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
#pragma unroll 1
    for (int k = 0; k < 4; ++k)
    {
        double cx = p.corner_x[k], cy = p.corner_y[k];
        if (point_hits_orig(m, m.raw_bits, cx * c - cy * s + px, cx * s + cy * c + py))
            return true;
    }
    return false;
}

// Fast path. Only trunc(q) and the comparisons of q with -1 and the map size are ever used of a probe's grid
// coordinates q = (rotated offset) / res. In real arithmetic q is LINEAR along the axis line, so the probes are walked in
// 32.32 fixed point: Q(k) = Q(0) + k * dQ, cell = high word, no FP64 and no division per probe (~18 integer
// instructions instead of ~80 FP64 ones). The reference's rounded q and this Q differ by at most
//   128 * 2^-53 * (|x| + |y| + |origin| + footprint) / res  (FP64 roundings on both sides)  +  (n_check + 1) * 2^-32
// cells, i.e. < 2^-24 + 2^-23 under the two preconditions checked below (pose within 2^22 cells of the origin,
// n_check <= 512). If ANY probe's fractional part lies within kFxEps = 2^-22 of a cell edge the whole pose is decided by
// the exact path (probability ~3e-5 per pose); everywhere else both values have the same trunc() and the same in-map
// verdict. Verdicts stay bit-exact.
//
// The probes read the PADDED TILE-WORD copies of the two collision layers (MapView::raw_tw / infl_tw, built by
// build_tilewords_kernel): a margin of kTwMargin cells around the map reads "occupied" (outside the map = collision,
// nav_map.cpp:259-277) except the ring at index -1, which repeats cell 0 (truncation toward zero: q in (-1, 0) is cell 0).
// So the loop needs no bounds test and no clamp. A 32-bit word holds 4 rows x 8 columns, a 32-byte sector 4 x 64 cells.
constexpr uint32_t kFxEps = 1u << 10; // 2^-22 cell in units of 2^-32

__device__ __forceinline__ uint32_t tw_probe(const uint32_t *__restrict__ tw, int pitch, long long qj, long long qi)
{
    const int hj = (int)(qj >> 32), hi = (int)(qi >> 32);
#if MPROS_TW_BLOCK16
    // 16 x 16-cell blocks: the eight words of a block (4 word rows x 2 word columns) share one 32-byte sector
    const uint32_t w = __ldg(tw + (((((hi >> 4) * pitch + (hj >> 4)) << 3) | ((hi >> 1) & 6)) | ((hj >> 3) & 1)));
#else
    const uint32_t w = __ldg(tw + ((hi >> 2) * pitch + (hj >> 3)));
#endif
    return w >> ((((unsigned)hi << 3) & 24u) | ((unsigned)hj & 7u));
}
// distance of the fractional parts from the nearest cell edge (wraps: 0 at an edge), running minimum
__device__ __forceinline__ uint32_t tw_edge(uint32_t acc, long long qj, long long qi)
{
    return min(acc, min((uint32_t)qj + kFxEps, (uint32_t)qi + kFxEps));
}

__device__ __forceinline__ bool footprint_in_collision4_sc(const MapView &m, const PlannerView &p, double px, double py, double s, double c)
{
    const double x0 = c * p.longi_x0 + px, y0 = s * p.longi_x0 + py;
    const double x1 = c * p.longi_x1 + px, y1 = s * p.longi_x1 + py;
    // far away, a NaN anywhere in the pose (x0 / y0 carry px, py, sin and cos) or a footprint the fast path cannot take
    if (!(fabs(x0) + fabs(y0) < m.fx_span - p.fx_extent))
        return footprint_in_collision4_exact(m, p, px, py, s, c);
    const double kTwo32 = 4294967296.0;
    const double sc = m.inv_res0 * kTwo32;
    const double jc = m.ocos * sc, js = m.osin * sc; // grid-j row of the scaled rotation; grid-i row is (-js, jc)
    const double off = (double)kTwMargin * kTwo32;   // padded coordinates
    const double ax0 = x0 - m.ox, ay0 = y0 - m.oy, ax1 = x1 - m.ox, ay1 = y1 - m.oy;
    const double fj0 = ax0 * jc + ay0 * js + off, fi0 = ay0 * jc - ax0 * js + off;
    const double fj1 = ax1 * jc + ay1 * js + off, fi1 = ay1 * jc - ax1 * js + off;
    long long qj = __double2ll_rd(fj0), qi = __double2ll_rd(fi0);
    // both axis end points at least fx_slack cells inside the padded array: then every probe and corner is inside it
    {
        const long long ej = __double2ll_rd(fj1), ei = __double2ll_rd(fi1);
        const unsigned lim_j = (unsigned)(m.W0 + 2 * kTwMargin - 2 * p.fx_slack), lim_i = (unsigned)(m.H0 + 2 * kTwMargin - 2 * p.fx_slack);
        const bool in = ((unsigned)((int)(qj >> 32) - p.fx_slack) < lim_j) & ((unsigned)((int)(ej >> 32) - p.fx_slack) < lim_j) &
                        ((unsigned)((int)(qi >> 32) - p.fx_slack) < lim_i) & ((unsigned)((int)(ei >> 32) - p.fx_slack) < lim_i);
        if (!in)
            return footprint_in_collision4_exact(m, p, px, py, s, c);
    }
    const long long dj = __double2ll_rn((fj1 - fj0) * p.inv_n_check), di = __double2ll_rn((fi1 - fi0) * p.inv_n_check);
    uint32_t any = 0, edge = 0xffffffffu;
    const int n = p.n_check;
#ifndef MPROS_C1_UNROLL
#define MPROS_C1_UNROLL 4
#endif
    constexpr int kUnroll = MPROS_C1_UNROLL; // independent probe loads in flight per thread
#pragma unroll kUnroll
    for (int k = 0; k <= n; ++k)
    {
        any |= tw_probe(m.infl_tw, m.tw_pitch, qj, qi);
        edge = tw_edge(edge, qj, qi);
        qj += dj;
        qi += di;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        const double cx = p.corner_x[k], cy = p.corner_y[k];
        const double X = cx * c - cy * s + px, Y = cx * s + cy * c + py; // as the reference computes them
        const double ax = X - m.ox, ay = Y - m.oy;
        const long long cj = __double2ll_rd(ax * jc + ay * js + off), ci = __double2ll_rd(ay * jc - ax * js + off);
        any |= tw_probe(m.raw_tw, m.tw_pitch, cj, ci);
        edge = tw_edge(edge, cj, ci);
    }
    if (edge < 2 * kFxEps) // some probe too close to a cell edge to trust the fixed-point walk
        return footprint_in_collision4_exact(m, p, px, py, s, c);
    return (any & 1u) != 0;
}

// Pre-classification of a pose by its own cell (collision.cu explains the two rules; Ctx::unsafe_tw is built by
// map_build_unsafe): 0 = surely free, 1 = sure hit, 2 = undecided (run the footprint test). yaw is only tested for
// finiteness. unsafe_tw == nullptr classifies nothing.
__device__ __forceinline__ int footprint_preclass(const MapView &m, const PlannerView &p, const uint32_t *__restrict__ unsafe_tw, int center_probe,
                                                  double px, double py, double yaw)
{
    if (!unsafe_tw || !(fabs(px) + fabs(py) < m.fx_span - p.fx_extent) || !(fabs(yaw) <= 1.7976931348623157e308))
        return 2;
    const double kTwo32 = 4294967296.0;
    const double sc = m.inv_res0 * kTwo32;
    const double jc = m.ocos * sc, js = m.osin * sc;
    const double off = (double)kTwMargin * kTwo32;
    const double ax = px - m.ox, ay = py - m.oy;
    const long long qj = __double2ll_rd(ax * jc + ay * js + off), qi = __double2ll_rd(ay * jc - ax * js + off);
    const unsigned lim_j = (unsigned)(m.W0 + 2 * kTwMargin - 2 * p.fx_slack), lim_i = (unsigned)(m.H0 + 2 * kTwMargin - 2 * p.fx_slack);
    if (!(((unsigned)((int)(qj >> 32) - p.fx_slack) < lim_j) & ((unsigned)((int)(qi >> 32) - p.fx_slack) < lim_i)))
        return 2;
    const uint32_t us = tw_probe(unsafe_tw, m.tw_pitch, qj, qi), hit = tw_probe(m.infl_tw, m.tw_pitch, qj, qi);
    if (!(us & 1u))
        return 0;
    if (center_probe && (hit & 1u) && tw_edge(0xffffffffu, qj, qi) >= 2 * kFxEps)
        return 1;
    return 2;
}

__device__ __forceinline__ bool footprint_in_collision4(const MapView &m, const PlannerView &p, double px, double py,
                                                        double yaw)
{
    double s, c;
    sincos(yaw, &s, &c);
    return footprint_in_collision4_sc(m, p, px, py, s, c);
}

} // namespace mpros
