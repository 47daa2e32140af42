// Device-side Reeds-Shepp machinery.
//  * RS generator of mpros_rs_path (rs_curve.cpp / rs_curve_basic.h): the reference's own 12 word formulas x 4
//    symmetries WITH their quirks (path4's asin, PathSeg sign handling, positional penalty mix-up), penalty-weighted
//    first-wins minimum, sequential pose integration (GenTraj).
//  * OMPL-equivalent Reeds-Shepp distance used by the heuristic (rs_statespace.h:50-54).
// Warp-cooperative: one warp works on one query; lanes spread over candidates / symmetry images.
#pragma once
#include <cfloat>

#include "mpros_internal.cuh"

namespace mpros
{

constexpr double kPi = 3.14159265358979323846;
constexpr double kPi2 = 1.57079632679489661923; // M_PI_2
constexpr unsigned kFull = 0xffffffffu;

// Out-of-line wrappers of the FP64 math routines. The planner kernel calls each of them from dozens of sites; inlined
// they blow the code up to ~290 KB and the kernel becomes instruction-fetch bound (ncu: stall_no_instruction dominant).
// One shared copy each keeps the hot loop inside the instruction caches. Results are identical to the inlined calls.
namespace dm
{
__device__ __noinline__ double atan2_(double y, double x) { return atan2(y, x); }
__device__ __noinline__ double sqrt_(double x) { return sqrt(x); }
__device__ __noinline__ double asin_(double x) { return asin(x); }
__device__ __noinline__ double acos_(double x) { return acos(x); }
__device__ __noinline__ double hypot_(double x, double y) { return hypot(x, y); }
__device__ __noinline__ double sin_(double x) { return sin(x); }
__device__ __noinline__ double cos_(double x) { return cos(x); }
__device__ __noinline__ void sincos_(double x, double *s, double *c) { sincos(x, s, c); }
__device__ __noinline__ double div_(double a, double b) { return a / b; }
__device__ __noinline__ double fmod_(double a, double b) { return fmod(a, b); }
} // namespace dm

struct RsConfig
{
    double radius, steer_angle, step;
    // as RSCurve::setPenalty receives them positionally (rs_curve.cpp:314-323)
    double pen_gear, pen_back, pen_steer_change, pen_steer_angle;
};

struct RsSeg
{
    double len;
    int steer, dir;
};
struct RsPath
{
    RsSeg s[5];
    int n;
};

__device__ __forceinline__ void rs_push(RsPath &p, double value, int steer, int dir)
{
    // PathSeg ctor (rs_curve_basic.h:84-88): |value|; direction flips unless value > 0 (also for 0 and NaN)
    p.s[p.n].len = fabs(value);
    p.s[p.n].steer = steer;
    p.s[p.n].dir = (value > 0 ? dir : -dir);
    ++p.n;
}
__device__ __forceinline__ double rs_regular_rad(double t) // rs_curve_basic.h:113-124
{
    if (t > kPi)
        return t - 2 * kPi;
    else if (t < -kPi)
        return t + 2 * kPi;
    return t;
}

// PathFunction::path1..12 (rs_curve.cpp:353-599) in two parts: the closed-form values (t, u, v) of word f - the only
// floating-point work - and the word's fixed segment pattern. Every segment of every word is one of t, u, v or pi/2 with a
// fixed steer and direction, pushed through the PathSeg constructor (rs_curve_basic.h:84-88: |value|, direction flipped
// unless value > 0). A candidate therefore lives in three registers plus the word number; no per-thread segment array.
// dist/theta: PolarCoord of the formula's own argument. Returns false when the word has no solution (empty path).
__device__ __forceinline__ bool rs_formula_tuv(int f, double x, double y, double phi, double &t, double &u, double &v)
{
    double sphi, cphi;
    dm::sincos_(phi, &sphi, &cphi);
    const bool typeA = (f == 0 || f == 2 || f == 3 || f == 4 || f == 7 || f == 9);
    double ax = typeA ? x - sphi : x + sphi;
    double ay = typeA ? y - 1 + cphi : y - 1 - cphi;
    double dist = dm::hypot_(ax, ay), theta = dm::atan2_(ay, ax);
    t = u = v = 0;
    switch (f)
    {
    case 0:
        u = dist, t = theta, v = rs_regular_rad(phi - t);
        return true;
    case 1:
        if (dist * dist >= 4)
        {
            u = dm::sqrt_(dist * dist - 4);
            t = rs_regular_rad(theta + dm::atan2_(2.0, u));
            v = rs_regular_rad(t - phi);
            return true;
        }
        return false;
    case 2:
        if (dist <= 4)
        {
            double A = dm::acos_(dist / 4);
            u = rs_regular_rad(kPi - 2 * A);
            t = rs_regular_rad(kPi2 + A + theta);
            v = rs_regular_rad(phi - t - u);
            return true;
        }
        return false;
    case 3:
        if (dist <= 4)
        {
            double A = dm::acos_(dist / 4);
            u = rs_regular_rad(kPi - 2 * A);
            t = dm::asin_(theta + kPi2 + A); // sic, rs_curve.cpp:418
            v = rs_regular_rad(t + u - phi);
            return true;
        }
        return false;
    case 4:
        if (dist <= 4)
        {
            u = dm::acos_(1 - dist * dist / 8);
            double A = dm::asin_(2 * dm::sin_(u) / dist);
            t = rs_regular_rad(theta + kPi2 - A);
            v = rs_regular_rad(t - u - phi);
            return true;
        }
        return false;
    case 5:
        if (dist <= 4)
        {
            double A;
            if (dist <= 2)
            {
                A = dm::acos_((dist + 2) / 4);
                t = rs_regular_rad(theta + kPi2 + A);
                u = rs_regular_rad(A);
                v = rs_regular_rad(phi - t + 2 * u);
            }
            else
            {
                A = dm::acos_((dist - 2) / 4);
                t = rs_regular_rad(theta + kPi2 - A);
                u = rs_regular_rad(kPi - A);
                v = rs_regular_rad(phi - t + 2 * u);
            }
            return true;
        }
        return false;
    case 6:
    {
        double u1 = (20 - dist * dist) / 16;
        if (dist <= 6 && u1 >= 0 && u1 <= 1)
        {
            u = dm::acos_(u1);
            double A = dm::asin_(2 * dm::sin_(u) / dist);
            t = rs_regular_rad(theta + kPi2 + A);
            v = rs_regular_rad(t - phi);
            return true;
        }
        return false;
    }
    case 7:
        if (dist >= 2)
        {
            u = dm::sqrt_(dist * dist - 4) - 2;
            double A = dm::atan2_(2.0, u + 2);
            t = rs_regular_rad(theta + kPi2 + A);
            v = rs_regular_rad(t - phi + kPi2);
            return true;
        }
        return false;
    case 8:
        if (dist >= 2)
        {
            t = rs_regular_rad(theta + kPi2);
            u = dist - 2;
            v = rs_regular_rad(phi - t - kPi2);
            return true;
        }
        return false;
    case 9:
        if (dist >= 2)
        {
            u = dm::sqrt_(dist * dist - 4) - 2;
            double A = dm::atan2_(u + 2, 2.0);
            t = rs_regular_rad(theta + kPi2 - A);
            v = rs_regular_rad(t - phi - kPi2);
            return true;
        }
        return false;
    case 10:
        if (dist >= 2)
        {
            u = dist - 2;
            t = rs_regular_rad(theta);
            v = rs_regular_rad(phi - t - kPi2);
            return true;
        }
        return false;
    case 11:
        if (dist >= 4)
        {
            u = dm::sqrt_(dist * dist - 4) - 4;
            double A = dm::atan2_(2.0, u + 4);
            t = rs_regular_rad(theta + kPi2 + A);
            v = rs_regular_rad(t - phi);
            return true;
        }
        return false;
    default:
        return false;
    }
}

// Segment pattern of word f: bits [5k, 5k+5) = segment k {value selector (0 t, 1 u, 2 v, 3 pi/2), steer + 1, direction > 0},
// bits [28, 31) = number of segments. L = 1, S = 0, R = -1, F = 1, B = -1 (rs_curve.cpp:353-599, the order of the pushes).
__host__ __device__ constexpr uint32_t rs_seg_code(int sel, int steer, int dir) { return (uint32_t)sel | ((uint32_t)(steer + 1) << 2) | ((dir > 0 ? 1u : 0u) << 4); }
__host__ __device__ constexpr uint32_t rs_word_code(int n, uint32_t a, uint32_t b, uint32_t c, uint32_t d = 0, uint32_t e = 0)
{
    return a | (b << 5) | (c << 10) | (d << 15) | (e << 20) | ((uint32_t)n << 28);
}
__device__ __forceinline__ uint32_t rs_word_pattern(int f)
{
    constexpr int T = 0, U = 1, V = 2, H = 3, L_ = 1, S_ = 0, R_ = -1, F_ = 1, B_ = -1;
    switch (f)
    {
    case 0: return rs_word_code(3, rs_seg_code(T, L_, F_), rs_seg_code(U, S_, F_), rs_seg_code(V, L_, F_));
    case 1: return rs_word_code(3, rs_seg_code(T, L_, F_), rs_seg_code(U, S_, F_), rs_seg_code(V, R_, F_));
    case 2: return rs_word_code(3, rs_seg_code(T, L_, F_), rs_seg_code(U, R_, B_), rs_seg_code(V, L_, F_));
    case 3: return rs_word_code(3, rs_seg_code(T, L_, F_), rs_seg_code(U, R_, B_), rs_seg_code(V, L_, B_));
    case 4: return rs_word_code(3, rs_seg_code(T, L_, F_), rs_seg_code(U, R_, F_), rs_seg_code(V, L_, B_));
    case 5: return rs_word_code(4, rs_seg_code(T, L_, F_), rs_seg_code(U, R_, F_), rs_seg_code(U, L_, B_), rs_seg_code(V, R_, B_));
    case 6: return rs_word_code(4, rs_seg_code(T, L_, F_), rs_seg_code(U, R_, B_), rs_seg_code(U, L_, B_), rs_seg_code(V, R_, F_));
    case 7: return rs_word_code(4, rs_seg_code(T, L_, F_), rs_seg_code(H, R_, B_), rs_seg_code(U, S_, B_), rs_seg_code(V, L_, B_));
    case 8: return rs_word_code(4, rs_seg_code(T, L_, F_), rs_seg_code(H, R_, B_), rs_seg_code(U, S_, B_), rs_seg_code(V, R_, B_));
    case 9: return rs_word_code(4, rs_seg_code(T, L_, F_), rs_seg_code(U, S_, F_), rs_seg_code(H, R_, F_), rs_seg_code(V, L_, B_));
    case 10: return rs_word_code(4, rs_seg_code(T, L_, F_), rs_seg_code(U, S_, F_), rs_seg_code(H, L_, F_), rs_seg_code(V, R_, B_));
    case 11: return rs_word_code(5, rs_seg_code(T, L_, F_), rs_seg_code(H, R_, B_), rs_seg_code(U, S_, B_), rs_seg_code(H, L_, B_), rs_seg_code(V, R_, F_));
    default: return 0;
    }
}

// One Reeds-Shepp candidate in registers: word f under symmetry j (0 plain, 1 timeflip, 2 reflect, 3 both), n = 0 when the
// word has no solution for this goal.
struct RsWord
{
    double t, u, v;
    int f, j, n;
};
// Segment k of a candidate as FindOptimalPath holds it (rs_curve.cpp:111-163): PathSeg ctor, then Timeflip, then Reflect
// (each re-constructs the segment from its non-negative length: zero-length and NaN segments flip once more, :31-49).
__device__ __forceinline__ void rs_word_seg(const RsWord &w, int k, double &len, int &steer, int &dir)
{
    const uint32_t code = rs_word_pattern(w.f) >> (5 * k);
    const int sel = code & 3;
    const double value = sel == 0 ? w.t : (sel == 1 ? w.u : (sel == 2 ? w.v : kPi2));
    steer = (int)((code >> 2) & 3) - 1;
    int d = ((code >> 4) & 1) ? 1 : -1;
    len = fabs(value);
    d = value > 0 ? d : -d;
    if (w.j == 1 || w.j == 3)
    {
        const int e = -d;
        d = len > 0 ? e : -e;
    }
    if (w.j == 2 || w.j == 3)
    {
        steer = -steer;
        d = len > 0 ? d : -d;
    }
    dir = d;
}

// path1..12 into a segment array (the team planner's shot and the word-by-word entry point)
__device__ __noinline__ void rs_formula(int f, double x, double y, double phi, RsPath &p)
{
    RsWord w;
    w.f = f, w.j = 0;
    w.n = rs_formula_tuv(f, x, y, phi, w.t, w.u, w.v) ? (int)(rs_word_pattern(f) >> 28) : 0;
    p.n = w.n;
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (k < w.n)
            rs_word_seg(w, k, p.s[k].len, p.s[k].steer, p.s[k].dir);
}

// Timeflip / Reflect (rs_curve.cpp:31-49): each segment is re-constructed through the PathSeg ctor with its
// (non-negative) length, so zero-length and NaN segments flip their direction once more.
__device__ __forceinline__ void rs_timeflip(RsPath &p)
{
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (k < p.n)
        {
            int d = -p.s[k].dir;
            p.s[k].dir = (p.s[k].len > 0 ? d : -d);
        }
}
__device__ __forceinline__ void rs_reflect(RsPath &p)
{
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (k < p.n)
        {
            p.s[k].steer = -p.s[k].steer;
            p.s[k].dir = (p.s[k].len > 0 ? p.s[k].dir : -p.s[k].dir);
        }
}

// RSCurve::PathLength (rs_curve.cpp:60-97)
__device__ __forceinline__ double rs_path_length(const RsConfig &c, const RsPath &path, int start_direc, double start_steer)
{
    double length = 0, penalty = 0;
    int last_steer = (int)start_steer; // int initialised from a double (:64)
    if (path.n == 0)
        return 10000;
#pragma unroll
    for (int i = 0; i < 5; ++i)
        if (i < path.n)
        {
            const RsSeg &seg = path.s[i];
            if (seg.dir != start_direc)
                penalty += c.pen_gear;
            if (seg.dir < 0)
                penalty += c.pen_back * seg.len * c.radius;
            if (seg.steer != last_steer)
                penalty += c.pen_steer_change * (double)abs(seg.steer - last_steer) * c.steer_angle;
            if (seg.steer != 0)
                penalty += c.pen_steer_angle * (double)abs(seg.steer) * c.steer_angle;
            length += seg.len * c.radius;
            last_steer = seg.steer;
        }
    return (length + penalty) * (1.0 + 1e-3);
}

// RSCurve::PathLength (rs_curve.cpp:60-97) of a candidate in registers, the same operations in the same order as
// rs_path_length; nan_seg = some segment length is NaN (FindOptimalPath drops such candidates, :150).
__device__ __forceinline__ double rs_word_length(const RsConfig &c, const RsWord &w, int start_direc, double start_steer, bool &nan_seg)
{
    nan_seg = false;
    if (w.n == 0)
        return 10000;
    double length = 0, penalty = 0;
    int last_steer = (int)start_steer; // int initialised from a double (:64)
#pragma unroll
    for (int i = 0; i < 5; ++i)
        if (i < w.n)
        {
            double len;
            int steer, dir;
            rs_word_seg(w, i, len, steer, dir);
            nan_seg |= isnan(len);
            if (dir != start_direc)
                penalty += c.pen_gear;
            if (dir < 0)
                penalty += c.pen_back * len * c.radius;
            if (steer != last_steer)
                penalty += c.pen_steer_change * (double)abs(steer - last_steer) * c.steer_angle;
            if (steer != 0)
                penalty += c.pen_steer_angle * (double)abs(steer) * c.steer_angle;
            length += len * c.radius;
            last_steer = steer;
        }
    return (length + penalty) * (1.0 + 1e-3);
}

// RSCurve::FindOptimalPath (rs_curve.cpp:111-163) for EIGHT shots per warp: lane = 4 * shot + symmetry j, and the twelve
// words are walked together, so all 32 lanes evaluate the SAME closed form at the same time (one shot per warp puts eight
// different words on four lanes each: the warp then executes every formula body for an eighth of its lanes). Candidate
// c = 4 f + j in the reference's evaluation order; multimap.begin() = smallest cost, first inserted among equals. The four
// lanes of a quad return the same best candidate; a candidate is three doubles and the word number, all in registers.
__device__ __forceinline__ RsWord rs_find_optimal_quad(const RsConfig &c, double sx, double sy, double sphi, int start_direc,
                                                        double start_steer, double ex, double ey, double ephi, double &cost)
{
    const int j = threadIdx.x & 3;
    // NewGoalPoint (:51-58)
    double dx = ex - sx, dy = ey - sy;
    double cs, sn;
    dm::sincos_(sphi, &sn, &cs);
    const double gx = (dx * cs + dy * sn) / c.radius;
    const double gy = (-dx * sn + dy * cs) / c.radius;
    const double gphi = ephi - sphi;
    const double x = (j == 1 || j == 3) ? -gx : gx;
    const double y = (j == 2 || j == 3) ? -gy : gy;
    const double ph = (j == 1 || j == 2) ? -gphi : gphi;
    RsWord best;
    best.t = best.u = best.v = 0, best.f = 0, best.j = j, best.n = 0;
    double bc = 0;
    int bi = 1 << 20; // "no candidate"
#pragma unroll 1
    for (int f = 0; f < 12; ++f)
    {
        RsWord w;
        w.f = f, w.j = j;
        w.n = rs_formula_tuv(f, x, y, ph, w.t, w.u, w.v) ? (int)(rs_word_pattern(f) >> 28) : 0;
        bool nan_seg;
        const double len = rs_word_length(c, w, start_direc, start_steer, nan_seg);
        const bool ok = (len != 0.0) && !nan_seg; // :150 (NaN cost counts as true but implies a NaN segment)
        if (ok && (bi >= (1 << 20) || len < bc))
        {
            bc = len;
            bi = 4 * f + j;
            best = w;
        }
    }
    // arg-min over the quad on (cost, candidate index); costs are never NaN here
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1)
    {
        const double oc = __shfl_xor_sync(kFull, bc, o);
        const int oi = __shfl_xor_sync(kFull, bi, o);
        RsWord ow;
        ow.t = __shfl_xor_sync(kFull, best.t, o), ow.u = __shfl_xor_sync(kFull, best.u, o), ow.v = __shfl_xor_sync(kFull, best.v, o);
        ow.f = __shfl_xor_sync(kFull, best.f, o), ow.j = __shfl_xor_sync(kFull, best.j, o), ow.n = __shfl_xor_sync(kFull, best.n, o);
        const bool take = (oi < (1 << 20)) && (bi >= (1 << 20) || oc < bc || (oc == bc && oi < bi));
        if (take)
        {
            bc = oc;
            bi = oi;
            best = ow;
        }
    }
    if (bi >= (1 << 20)) // unreachable: word 0 always yields a candidate
        best.n = 0, bc = 0.0;
    cost = bc;
    return best;
}

// RSCurve::GenTraj (rs_curve.cpp:165-222): sequential pose integration, executed by ONE lane.
// out_xyz: n x 3 {x, y, yaw}; out_sd (optional): n x 2 {steer*steer_angle, direc}. Returns the number of poses the
// reference would generate (may exceed cap; only the first cap are stored).
__device__ __noinline__ int rs_gen_traj(const RsConfig &c, double sx, double sy, double syaw, const RsPath &path, double *out_xyz,
                                        double *out_sd, int cap)
{
    const double curve_min_num = 3;
    double curr_x = sx, curr_y = sy, curr_yaw = syaw;
    int n = 0;
#pragma unroll 1
    for (int k = 0; k < path.n; ++k)
    {
        const RsSeg seg = path.s[k];
        int num_pt = (int)fmax(ceil(seg.len * c.radius / c.step), curve_min_num);
        double direc = (double)seg.dir;
        if (seg.steer != 0)
        {
            double dyaw = seg.len / num_pt;
            double sd, cd;
            dm::sincos_(dyaw, &sd, &cd);
            double dx = sd * c.radius * direc;
            double dy = (1 - cd) * c.radius * (double)seg.steer;
#pragma unroll 1
            for (int i = 0; i < num_pt; ++i)
            {
                double sy_, cy_;
                dm::sincos_(curr_yaw, &sy_, &cy_);
                double nx = curr_x + (dx * cy_ - dy * sy_);
                double ny = curr_y + (dx * sy_ + dy * cy_);
                curr_x = nx;
                curr_y = ny;
                curr_yaw += dyaw * direc * (double)seg.steer;
                curr_yaw = rs_regular_rad(curr_yaw);
                if (n < cap)
                {
                    out_xyz[3 * n] = curr_x, out_xyz[3 * n + 1] = curr_y, out_xyz[3 * n + 2] = curr_yaw;
                    if (out_sd)
                        out_sd[2 * n] = (double)seg.steer * c.steer_angle, out_sd[2 * n + 1] = direc;
                }
                ++n;
            }
        }
        else
        {
            double step = seg.len * c.radius / num_pt;
            double sy_, cy_;
            dm::sincos_(curr_yaw, &sy_, &cy_);
#pragma unroll 1
            for (int i = 0; i < num_pt; ++i)
            {
                curr_x += step * cy_ * direc;
                curr_y += step * sy_ * direc;
                if (n < cap)
                {
                    out_xyz[3 * n] = curr_x, out_xyz[3 * n + 1] = curr_y, out_xyz[3 * n + 2] = curr_yaw;
                    if (out_sd)
                        out_sd[2 * n] = (double)seg.steer * c.steer_angle, out_sd[2 * n + 1] = direc;
                }
                ++n;
            }
        }
    }
    return n;
}

// RSCurve::GenTraj by a whole warp, the same operations in the same order as rs_gen_traj (all lanes hold the same candidate
// and start pose; all 32 lanes must call it). GenTraj carries three chains from pose to pose: the yaw (one add + wrap per
// pose: cheap, run by every lane redundantly), sin/cos of the yaw a pose starts from (expensive, independent once the yaws
// are known: ONE lane per pose) and the position (x += increment: sequential floating-point adds, replayed in order over
// warp shuffles). The poses stay in registers: lane n < total returns pose n and the steer / direction of its segment.
// Returns the number of poses; more than 32 (rare inside the 10 m analytic range) produces no poses - the caller takes the
// one-lane routine.
__device__ __forceinline__ int rs_gen_traj_regs(const RsConfig &c, double sx, double sy, double syaw, const RsWord &w, double &px, double &py,
                                                double &pyaw, int &psteer, int &pdir)
{
    const int lane = threadIdx.x & 31;
    const double curve_min_num = 3;
    // per segment: number of poses, and the constants of its pose step (lane k < n evaluates segment k)
    int my_num = 0, my_steer = 0, my_dir = 1;
    double my_a = 0, my_b = 0, my_dyaw = 0; // arc: dx, dy, yaw step; straight: step, unused, 0
    if (lane < w.n)
    {
        double len;
        rs_word_seg(w, lane, len, my_steer, my_dir);
        my_num = (int)fmax(ceil(len * c.radius / c.step), curve_min_num);
        const double direc = (double)my_dir;
        if (my_steer != 0)
        {
            const double dyaw = len / my_num;
            double sd, cd;
            dm::sincos_(dyaw, &sd, &cd);
            my_a = sd * c.radius * direc;
            my_b = (1 - cd) * c.radius * (double)my_steer;
            my_dyaw = dyaw * direc * (double)my_steer;
        }
        else
            my_a = len * c.radius / my_num;
    }
    int total = 0;
    for (int k = 0; k < w.n; ++k)
        total += __shfl_sync(kFull, my_num, k);
    px = sx, py = sy, pyaw = syaw, psteer = 0, pdir = 1;
    if (total > 32)
        return total;
    // yaw chain: lane n keeps the yaw pose n starts from, the yaw it ends with and its segment
    double yaw_from = syaw, yaw_to = syaw, curr_yaw = syaw;
    int my_seg = 0;
    {
        int n = 0;
        for (int k = 0; k < w.n; ++k)
        {
            const int num_pt = __shfl_sync(kFull, my_num, k);
            const double dyaw = __shfl_sync(kFull, my_dyaw, k);
            const bool arc = __shfl_sync(kFull, my_steer, k) != 0;
            for (int i = 0; i < num_pt; ++i, ++n)
            {
                const double before = curr_yaw;
                if (arc)
                {
                    curr_yaw += dyaw;
                    curr_yaw = rs_regular_rad(curr_yaw);
                }
                if (lane == n)
                    yaw_from = before, yaw_to = curr_yaw, my_seg = k;
            }
        }
    }
    // position increments, one lane per pose
    const double a = __shfl_sync(kFull, my_a, my_seg), b = __shfl_sync(kFull, my_b, my_seg);
    const int seg_steer = __shfl_sync(kFull, my_steer, my_seg), seg_dir = __shfl_sync(kFull, my_dir, my_seg);
    double inc_x = 0, inc_y = 0;
    {
        double sy_, cy_;
        dm::sincos_(yaw_from, &sy_, &cy_);
        if (seg_steer != 0)
        {
            inc_x = a * cy_ - b * sy_;
            inc_y = a * sy_ + b * cy_;
        }
        else
        {
            const double direc = (double)seg_dir;
            inc_x = a * cy_ * direc;
            inc_y = a * sy_ * direc;
        }
    }
    // positions: the reference's running sums, in order
    double cx = sx, cy = sy;
    for (int n = 0; n < total; ++n)
    {
        cx = cx + __shfl_sync(kFull, inc_x, n);
        cy = cy + __shfl_sync(kFull, inc_y, n);
        if (lane == n)
            px = cx, py = cy;
    }
    pyaw = yaw_to;
    psteer = seg_steer;
    pdir = seg_dir;
    return total;
}

// ---- OMPL-equivalent Reeds-Shepp distance ---------------------------------------------------
// The oracle's stand-in for ompl::base::ReedsSheppStateSpace::distance is the specification (parity unpinned vs real
// OMPL, see oracle/port/ompl_rs_distance.h); this mirrors it operation for operation.
__device__ __noinline__ double ompl_mod2pi(double x) // out of line: ~25 call sites in the hot formulas (3 KB of instruction-cache footprint inlined)
{
    // v = fmod(x, 2*pi), exact: identity below 2*pi; one exact subtraction (Sterbenz: 2pi <= |x| < 4pi) below 4*pi
    const double twopi = 2.0 * kPi;
    double ax = fabs(x), v;
    if (ax < twopi)
        v = x;
    else if (ax < 2.0 * twopi)
        v = x < 0 ? x + twopi : x - twopi;
    else
        v = dm::fmod_(x, twopi); // rare (|x| >= 4 pi); out of line: inlined it was 30 KB of cold code inside the hot paths
    if (v < -kPi)
        v += twopi;
    else if (v > kPi)
        v -= twopi;
    return v;
}
constexpr double kOmplZero = 10.0 * DBL_EPSILON;

__device__ __noinline__ void ompl_tau_omega(double u, double v, double xi, double eta, double phi, double &tau, double &omega)
{
    double delta = ompl_mod2pi(u - v);
    double su, cu, sd, cd;
    dm::sincos_(u, &su, &cu);
    dm::sincos_(delta, &sd, &cd);
    double A = su - sd, B = cu - cd - 1.;
    double t1 = dm::atan2_(eta * A - xi * B, xi * A + eta * B), t2 = 2. * (cd - dm::cos_(v) - cu) + 3;
    tau = (t2 < 0) ? ompl_mod2pi(t1 + kPi) : ompl_mod2pi(t1);
    omega = ompl_mod2pi(tau - u + v - phi);
}

// All base formulas on one image (x, y, phi) and its "backwards" point; returns candidate lengths per family:
//   csc_ccc_cccc = min over CSC, CCC (normal + backwards), CCCC candidates; ccsc = min over CCSC candidates (without the
//   fixed pi/2); ccscc = CCSCC candidate (without the fixed pi). DBL_MAX = no candidate.
__device__ __noinline__ void ompl_families(double x, double y, double phi, double xb, double yb, double &m_plain, double &m_ccsc,
                                           double &m_ccscc)
{
    m_plain = DBL_MAX;
    m_ccsc = DBL_MAX;
    m_ccscc = DBL_MAX;
    double sphi, cphi;
    dm::sincos_(phi, &sphi, &cphi);
    double t, u, v;
    // ---- arguments shared by several formulas
    const double xm = x - sphi, ym = y - 1. + cphi; // (xi, eta) of LpSpLp / LpRmL / LpRmSmLm
    const double xp = x + sphi, yp = y - 1. - cphi; // (xi, eta) of LpSpRp / CCCC / LpRmSmRm / LpRmSLmRp
    // LpSpLp (8.1)
    {
        u = dm::sqrt_(xm * xm + ym * ym);
        t = dm::atan2_(ym, xm);
        if (t >= -kOmplZero)
        {
            v = ompl_mod2pi(phi - t);
            if (v >= -kOmplZero)
                m_plain = fmin(m_plain, fabs(t) + fabs(u) + fabs(v));
        }
    }
    // LpSpRp (8.2)
    {
        double u1 = dm::sqrt_(xp * xp + yp * yp), t1 = dm::atan2_(yp, xp);
        u1 = u1 * u1;
        if (u1 >= 4.)
        {
            u = dm::sqrt_(u1 - 4.);
            double theta = dm::atan2_(2., u);
            t = ompl_mod2pi(t1 + theta);
            v = ompl_mod2pi(t - phi);
            if (t >= -kOmplZero && v >= -kOmplZero)
                m_plain = fmin(m_plain, fabs(t) + fabs(u) + fabs(v));
        }
    }
    // LpRmL (8.3/8.4) on the point and on the backwards point
#pragma unroll
    for (int b = 0; b < 2; ++b)
    {
        double xi = b ? xb - sphi : xm, eta = b ? yb - 1. + cphi : ym;
        double u1 = dm::sqrt_(xi * xi + eta * eta), theta = dm::atan2_(eta, xi);
        if (u1 <= 4.)
        {
            u = -2. * dm::asin_(.25 * u1);
            t = ompl_mod2pi(theta + .5 * u + kPi);
            v = ompl_mod2pi(phi - t + u);
            if (t >= -kOmplZero && u <= kOmplZero)
                m_plain = fmin(m_plain, fabs(t) + fabs(u) + fabs(v));
        }
    }
    // LpRupLumRm (8.7)
    {
        double rho = .25 * (2. + dm::sqrt_(xp * xp + yp * yp));
        if (rho <= 1.)
        {
            u = dm::acos_(rho);
            ompl_tau_omega(u, -u, xp, yp, phi, t, v);
            if (t >= -kOmplZero && v <= kOmplZero)
                m_plain = fmin(m_plain, fabs(t) + 2. * fabs(u) + fabs(v));
        }
    }
    // LpRumLumRp (8.8)
    {
        double rho = (20. - xp * xp - yp * yp) / 16.;
        if (rho >= 0 && rho <= 1)
        {
            u = -dm::acos_(rho);
            if (u >= -.5 * kPi)
            {
                ompl_tau_omega(u, u, xp, yp, phi, t, v);
                if (t >= -kOmplZero && v >= -kOmplZero)
                    m_plain = fmin(m_plain, fabs(t) + 2. * fabs(u) + fabs(v));
            }
        }
    }
    // LpRmSmLm (8.9) and LpRmSmRm (8.10), on the point and on the backwards point
#pragma unroll
    for (int b = 0; b < 2; ++b)
    {
        double xi = b ? xb - sphi : xm, eta = b ? yb - 1. + cphi : ym;
        double rho = dm::sqrt_(xi * xi + eta * eta), theta = dm::atan2_(eta, xi);
        if (rho >= 2.)
        {
            double r = dm::sqrt_(rho * rho - 4.);
            u = 2. - r;
            t = ompl_mod2pi(theta + dm::atan2_(r, -2.));
            v = ompl_mod2pi(phi - .5 * kPi - t);
            if (t >= -kOmplZero && u <= kOmplZero && v <= kOmplZero)
                m_ccsc = fmin(m_ccsc, fabs(t) + fabs(u) + fabs(v));
        }
        double xi2 = b ? xb + sphi : xp, eta2 = b ? yb - 1. - cphi : yp;
        double rho2 = dm::sqrt_(eta2 * eta2 + xi2 * xi2), theta2 = dm::atan2_(xi2, -eta2); // polar(-eta, xi)
        if (rho2 >= 2.)
        {
            t = theta2;
            u = 2. - rho2;
            v = ompl_mod2pi(t + .5 * kPi - phi);
            if (t >= -kOmplZero && u <= kOmplZero && v <= kOmplZero)
                m_ccsc = fmin(m_ccsc, fabs(t) + fabs(u) + fabs(v));
        }
    }
    // LpRmSLmRp (8.11)
    {
        double rho = dm::sqrt_(xp * xp + yp * yp);
        if (rho >= 2.)
        {
            u = 4. - dm::sqrt_(rho * rho - 4.);
            if (u <= kOmplZero)
            {
                t = ompl_mod2pi(dm::atan2_((4. - u) * xp - 2. * yp, -2. * xp + (u - 4.) * yp));
                v = ompl_mod2pi(t - phi);
                if (t >= -kOmplZero && v >= -kOmplZero)
                    m_ccscc = fmin(m_ccscc, fabs(t) + fabs(u) + fabs(v));
            }
        }
    }
}

// One base formula on one image, for the team planner where the 11 evaluations of ompl_families are spread over the
// warps of a CTA. k is warp-uniform. Returns the candidate length (without the fixed quarter turns) or DBL_MAX.
//   k: 0 LpSpLp  1 LpSpRp  2 LpRmL  3 LpRmL(backwards)  4 LpRupLumRm  5 LpRumLumRp      -> family 0 (plain)
//      6 LpRmSmLm  7 LpRmSmLm(backwards)  8 LpRmSmRm  9 LpRmSmRm(backwards)            -> family 1 (CCSC)
//      10 LpRmSLmRp                                                                   -> family 2 (CCSCC)
// Operation for operation the same arithmetic as ompl_families.
__device__ __forceinline__ int ompl_formula_family(int k) { return k < 6 ? 0 : (k < 10 ? 1 : 2); }
__device__ __noinline__ double ompl_formula(int k, double x, double y, double phi, double xb, double yb, double sphi, double cphi)
{
    double t, u, v;
    const double xm = x - sphi, ym = y - 1. + cphi;
    const double xp = x + sphi, yp = y - 1. - cphi;
    switch (k)
    {
    case 0:
        u = dm::sqrt_(xm * xm + ym * ym);
        t = dm::atan2_(ym, xm);
        if (t >= -kOmplZero)
        {
            v = ompl_mod2pi(phi - t);
            if (v >= -kOmplZero)
                return fabs(t) + fabs(u) + fabs(v);
        }
        return DBL_MAX;
    case 1:
    {
        double u1 = dm::sqrt_(xp * xp + yp * yp), t1 = dm::atan2_(yp, xp);
        u1 = u1 * u1;
        if (u1 >= 4.)
        {
            u = dm::sqrt_(u1 - 4.);
            double theta = dm::atan2_(2., u);
            t = ompl_mod2pi(t1 + theta);
            v = ompl_mod2pi(t - phi);
            if (t >= -kOmplZero && v >= -kOmplZero)
                return fabs(t) + fabs(u) + fabs(v);
        }
        return DBL_MAX;
    }
    case 2:
    case 3:
    {
        double xi = (k == 3) ? xb - sphi : xm, eta = (k == 3) ? yb - 1. + cphi : ym;
        double u1 = dm::sqrt_(xi * xi + eta * eta), theta = dm::atan2_(eta, xi);
        if (u1 <= 4.)
        {
            u = -2. * dm::asin_(.25 * u1);
            t = ompl_mod2pi(theta + .5 * u + kPi);
            v = ompl_mod2pi(phi - t + u);
            if (t >= -kOmplZero && u <= kOmplZero)
                return fabs(t) + fabs(u) + fabs(v);
        }
        return DBL_MAX;
    }
    case 4:
    {
        double rho = .25 * (2. + dm::sqrt_(xp * xp + yp * yp));
        if (rho <= 1.)
        {
            u = dm::acos_(rho);
            ompl_tau_omega(u, -u, xp, yp, phi, t, v);
            if (t >= -kOmplZero && v <= kOmplZero)
                return fabs(t) + 2. * fabs(u) + fabs(v);
        }
        return DBL_MAX;
    }
    case 5:
    {
        double rho = (20. - xp * xp - yp * yp) / 16.;
        if (rho >= 0 && rho <= 1)
        {
            u = -dm::acos_(rho);
            if (u >= -.5 * kPi)
            {
                ompl_tau_omega(u, u, xp, yp, phi, t, v);
                if (t >= -kOmplZero && v >= -kOmplZero)
                    return fabs(t) + 2. * fabs(u) + fabs(v);
            }
        }
        return DBL_MAX;
    }
    case 6:
    case 7:
    {
        double xi = (k == 7) ? xb - sphi : xm, eta = (k == 7) ? yb - 1. + cphi : ym;
        double rho = dm::sqrt_(xi * xi + eta * eta), theta = dm::atan2_(eta, xi);
        if (rho >= 2.)
        {
            double r = dm::sqrt_(rho * rho - 4.);
            u = 2. - r;
            t = ompl_mod2pi(theta + dm::atan2_(r, -2.));
            v = ompl_mod2pi(phi - .5 * kPi - t);
            if (t >= -kOmplZero && u <= kOmplZero && v <= kOmplZero)
                return fabs(t) + fabs(u) + fabs(v);
        }
        return DBL_MAX;
    }
    case 8:
    case 9:
    {
        double xi2 = (k == 9) ? xb + sphi : xp, eta2 = (k == 9) ? yb - 1. - cphi : yp;
        double rho2 = dm::sqrt_(eta2 * eta2 + xi2 * xi2), theta2 = dm::atan2_(xi2, -eta2);
        if (rho2 >= 2.)
        {
            t = theta2;
            u = 2. - rho2;
            v = ompl_mod2pi(t + .5 * kPi - phi);
            if (t >= -kOmplZero && u <= kOmplZero && v <= kOmplZero)
                return fabs(t) + fabs(u) + fabs(v);
        }
        return DBL_MAX;
    }
    default:
    {
        double rho = dm::sqrt_(xp * xp + yp * yp);
        if (rho >= 2.)
        {
            u = 4. - dm::sqrt_(rho * rho - 4.);
            if (u <= kOmplZero)
            {
                t = ompl_mod2pi(dm::atan2_((4. - u) * xp - 2. * yp, -2. * xp + (u - 4.) * yp));
                v = ompl_mod2pi(t - phi);
                if (t >= -kOmplZero && v >= -kOmplZero)
                    return fabs(t) + fabs(u) + fabs(v);
            }
        }
        return DBL_MAX;
    }
    }
}

// image q (0 id, 1 timeflip, 2 reflect, 3 both) of the normalised goal and of its backwards point
__device__ __forceinline__ void ompl_image(int q, double rho, double x1, double y1, double th1, double x2, double y2, double th2, double &x,
                                           double &y, double &phi, double &xb, double &yb)
{
    double dx = x2 - x1, dy = y2 - y1, dth = th2 - th1;
    double c, s;
    dm::sincos_(th1, &s, &c);
    double nx = (c * dx + s * dy) / rho, ny = (-s * dx + c * dy) / rho;
    double sphi, cphi;
    dm::sincos_(dth, &sphi, &cphi);
    double nxb = nx * cphi + ny * sphi, nyb = nx * sphi - ny * cphi;
    const bool fx = (q == 1 || q == 3), fy = (q == 2 || q == 3), fp = (q == 1 || q == 2);
    x = fx ? -nx : nx, y = fy ? -ny : ny, phi = fp ? -dth : dth, xb = fx ? -nxb : nxb, yb = fy ? -nyb : nyb;
}
// combine the three family minima exactly like the sequential CSC/CCC/CCCC, CCSC, CCSCC passes
__device__ __forceinline__ double ompl_combine(double rho, double m_plain, double m_ccsc, double m_ccscc)
{
    double best = m_plain;
    {
        double l0 = best - .5 * kPi;
        if (m_ccsc < l0)
            best = m_ccsc + .5 * kPi;
    }
    {
        double l0 = best - kPi;
        if (m_ccscc < l0)
            best = m_ccscc + kPi;
    }
    return rho * best;
}

// Groups of 4 consecutive lanes cooperate on one (start -> goal) pair: lane (q = lane & 3) evaluates symmetry image q
// of every base formula; the family minima are combined across the group with shuffles. All 4 lanes of the group get
// the result. Inactive groups must still call this (full-warp shuffles) with any finite arguments.
__device__ __noinline__ double ompl_rs_distance_quad(double rho, double x1, double y1, double th1, double x2, double y2,
                                                        double th2)
{
    const int q = threadIdx.x & 3;
    double dx = x2 - x1, dy = y2 - y1, dth = th2 - th1;
    double c, s;
    dm::sincos_(th1, &s, &c);
    double x = (c * dx + s * dy) / rho, y = (-s * dx + c * dy) / rho;
    // backwards point of the un-flipped image; its images follow the same sign pattern
    double sphi, cphi;
    dm::sincos_(dth, &sphi, &cphi);
    double xb = x * cphi + y * sphi, yb = x * sphi - y * cphi;
    const bool fx = (q == 1 || q == 3), fy = (q == 2 || q == 3), fp = (q == 1 || q == 2);
    double m_plain, m_ccsc, m_ccscc;
    ompl_families(fx ? -x : x, fy ? -y : y, fp ? -dth : dth, fx ? -xb : xb, fy ? -yb : yb, m_plain, m_ccsc, m_ccscc);
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1)
    {
        m_plain = fmin(m_plain, __shfl_xor_sync(kFull, m_plain, o));
        m_ccsc = fmin(m_ccsc, __shfl_xor_sync(kFull, m_ccsc, o));
        m_ccscc = fmin(m_ccscc, __shfl_xor_sync(kFull, m_ccscc, o));
    }
    double best = m_plain;
    {
        double l0 = best - .5 * kPi;
        if (m_ccsc < l0)
            best = m_ccsc + .5 * kPi;
    }
    {
        double l0 = best - kPi;
        if (m_ccscc < l0)
            best = m_ccscc + kPi;
    }
    return rho * best;
}

// Upper bound of the Reeds-Shepp distance from the two CSC formulas (k = 0, 1) on the quad's four symmetry images; same
// lane layout and the same image arithmetic as ompl_rs_distance_quad, the same candidates as its plain family holds.
// Any valid candidate is an upper bound of the distance: when it does not exceed the goal-map term h1 of
// max(h1, h2) the other nine formulas cannot change the heuristic (plan_team.cuh uses the same exact skip).
// +inf when neither CSC word applies. All 32 lanes must call it.
__device__ __noinline__ double ompl_csc_upper_bound_quad(double rho, double x1, double y1, double th1, double x2, double y2, double th2)
{
    const int q = threadIdx.x & 3;
    double dx = x2 - x1, dy = y2 - y1, dth = th2 - th1;
    double c, s;
    dm::sincos_(th1, &s, &c);
    double x = (c * dx + s * dy) / rho, y = (-s * dx + c * dy) / rho;
    double sphi, cphi;
    dm::sincos_(dth, &sphi, &cphi);
    double xb = x * cphi + y * sphi, yb = x * sphi - y * cphi;
    const bool fx = (q == 1 || q == 3), fy = (q == 2 || q == 3), fp = (q == 1 || q == 2);
    const double X = fx ? -x : x, Y = fy ? -y : y, PHI = fp ? -dth : dth, XB = fx ? -xb : xb, YB = fy ? -yb : yb;
    const double S = fp ? -sphi : sphi; // sin(-a) = -sin(a), cos(-a) = cos(a) bit for bit
    double L = fmin(ompl_formula(0, X, Y, PHI, XB, YB, S, cphi), ompl_formula(1, X, Y, PHI, XB, YB, S, cphi));
    L = fmin(L, __shfl_xor_sync(kFull, L, 1));
    L = fmin(L, __shfl_xor_sync(kFull, L, 2));
    return L < DBL_MAX ? rho * L : __longlong_as_double(0x7ff0000000000000ll);
}

} // namespace mpros
