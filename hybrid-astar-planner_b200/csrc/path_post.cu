// the last consumer of the footprint test. HybridAstar::setPath's PathPoint post-processing
// (hybrid_a_star.cpp:814-838): updatePathLength / setCusp / updateCurvatureNew (planner_utils.h:35-61, 251-309) and, with
// interpolation_enable, upsampleAndPopulatePath (hybrid_a_star.cpp:902-1112: cubic Bezier arcs between path points,
// cubicBezier / tangentDir / findArcCenter planner_utils.h:72-109, 311-340, 376-410) — what getNewPath() returns.
//
// Shape of the work: for one start point the reference tries end points from the far one (next cusp / path end) towards
// the start until a Bezier arc is collision-free and within 1.5 x max curvature; every try samples the arc and runs the
// footprint test on each sample. The control-point length a try starts from depends only on the GEOMETRY of the try
// before it, never on its verdict, so all tries of one start point are independent: the host lays out every candidate
// arc of the current segment of EVERY path of the batch (closed-form geometry, a few hundred points), ONE launch of the
// batched pathInCollision kernel (collision.cu) gives all verdicts, and the host takes the first acceptable candidate.
// A batch therefore costs (max segments per path) launches, not (paths x segments x tries) like a per-arc loop would.
//
// The geometry runs on the host with the same libm and the same operation order as the reference (two-term sums, no
// FMA contraction on x86-64), so the returned points are bit-identical wherever the verdicts agree (they are bit-exact).
#include <algorithm>
#include <cmath>
#include <limits>

#include "mpros_internal.cuh"

namespace mpros
{
namespace
{
struct V2
{
    double x, y;
};
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
inline V2 operator*(V2 a, double s) { return {a.x * s, a.y * s}; }
inline V2 operator/(V2 a, double s) { return {a.x / s, a.y / s}; }
inline V2 neg(V2 a) { return {-a.x, -a.y}; }
inline double dot(V2 a, V2 b)
{
    double s = a.x * b.x;
    s = s + a.y * b.y;
    return s;
}
inline double norm(V2 a) { return std::sqrt(dot(a, a)); }
inline V2 normalized(V2 a)
{
    double n = norm(a);
    return n > 0.0 ? a / n : a;
}

// PathPoint (mpros_utils/pathpoint.h:7-35), the fields the path carries
struct PP
{
    V2 p;
    double yaw = 0, curv = 0, direc = 0, steering = 0, dist = 0;
    bool cusp = false;
};

// planner_utils.h:35-45
void set_cusp(std::vector<PP> &path)
{
    for (size_t i = 0; i + 1 < path.size(); ++i)
        if (path[i].direc != path[i + 1].direc)
            path[i].cusp = true;
}
// planner_utils.h:52-61
void update_path_length(std::vector<PP> &path)
{
    double len = 0;
    path[0].dist = 0;
    for (size_t i = 1; i < path.size(); ++i)
    {
        len += norm(path[i].p - path[i - 1].p);
        path[i].dist = len;
    }
}
// planner_utils.h:251-309. Needs >= 2 points (the reference reads path[i + 1] for i == 0).
double update_curvature_new(std::vector<PP> &path)
{
    double max_curv = 0.0;
    const size_t n = path.size();
    for (size_t i = 0; i < n; ++i)
    {
        PP &t = path[i];
        V2 prev, pt, next;
        if (i == 0)
        {
            pt = t.p;
            next = path[i + 1].p;
            V2 orient{std::cos(t.yaw), std::sin(t.yaw)};
            V2 d12 = next - pt;
            V2 d02 = normalized(orient) * dot(orient, d12) * 2.0;
            prev = next - d02;
        }
        else if (i == n - 1)
        {
            pt = t.p;
            prev = path[i - 1].p;
            V2 orient{std::cos(t.yaw), std::sin(t.yaw)};
            V2 d01 = pt - prev;
            V2 d02 = normalized(orient) * dot(orient, d01) * 2.0;
            next = prev + d02;
        }
        else
        {
            prev = path[i - 1].p;
            pt = t.p;
            next = t.cusp ? (pt + pt) - path[i + 1].p : path[i + 1].p;
        }
        V2 d1 = normalized(next - prev);
        V2 d2 = normalized(next - pt) - normalized(pt - prev);
        t.curv = (d1.x * d2.y - d1.y * d2.x) / std::pow(norm(d1), 3);
        max_curv = std::fmax(std::abs(t.curv), max_curv);
    }
    return max_curv;
}
// planner_utils.h:311-340
V2 cubic_bezier(V2 p0, V2 p1, V2 p2, V2 p3, double mu)
{
    V2 a, b, c;
    c.x = 3 * (p1.x - p0.x);
    c.y = 3 * (p1.y - p0.y);
    b.x = 3 * (p2.x - p1.x) - c.x;
    b.y = 3 * (p2.y - p1.y) - c.y;
    a.x = p3.x - p0.x - c.x - b.x;
    a.y = p3.y - p0.y - c.y - b.y;
    return {a.x * mu * mu * mu + b.x * mu * mu + c.x * mu + p0.x, a.y * mu * mu * mu + b.y * mu * mu + c.y * mu + p0.y};
}
// planner_utils.h:72-109
V2 find_arc_center(V2 prev, V2 pt, V2 next, bool iscusp)
{
    V2 seg1 = pt - prev, seg2 = next - pt;
    if (iscusp)
    {
        seg2 = neg(seg2);
        next = pt + seg2;
    }
    double det = seg1.x * seg2.y - seg1.y * seg2.x;
    if (std::fabs(det) < 1e-7)
        return {std::numeric_limits<double>::infinity(), std::numeric_limits<double>::infinity()};
    V2 mid1 = (pt + prev) / 2.0, mid2 = (pt + next) / 2.0;
    V2 dir1{-seg1.y, seg1.x}, dir2{-seg2.y, seg2.x};
    double det1 = (mid1.x + dir1.x) * mid1.y - (mid1.y + dir1.y) * mid1.x;
    double det2 = (mid2.x + dir2.x) * mid2.y - (mid2.y + dir2.y) * mid2.x;
    return {(det1 * dir2.x - det2 * dir1.x) / det, (det1 * dir2.y - det2 * dir1.y) / det};
}
// planner_utils.h:376-410
V2 tangent_dir(V2 prev, V2 pt, V2 next, bool iscusp)
{
    V2 center = find_arc_center(prev, pt, next, iscusp);
    V2 d2 = next - pt;
    if (iscusp)
    {
        d2 = neg(d2);
        next = pt + d2;
    }
    if (std::isinf(center.x) || std::isinf(center.y))
        return {next.x - prev.x, next.y - prev.y};
    V2 dir{center.y - pt.y, pt.x - center.x};
    if (dot(dir, d2) < 0)
        dir = neg(dir);
    return dir;
}

struct Candidate
{
    int end_idx;
    double cpl_after; // control_pt_len the next try (or the next segment) starts from
    std::vector<PP> temp;
    bool ok_curv;
};

struct Job
{
    std::vector<PP> orig, out;
    std::vector<int> cusp_idx;
    int i = 0;
    double control_pt_len = 0.0;
    bool done = false;
    int start_idx = 0;
    std::vector<Candidate> cands;
};

// One try of upsampleAndPopulatePath's inner loop (hybrid_a_star.cpp:980-1080) without the footprint tests: the sampled
// arc temp_path (all num_sample points; the reference stops sampling at the first colliding point, which only matters
// for a candidate that is rejected anyway) and the control-point length it leaves behind.
void build_candidate(const Job &J, int start_idx, int end_idx, double cpl_in, const mpros_planner_params &pp, int upsample_factor,
                     Candidate &out)
{
    const int idx_diff = end_idx - start_idx;
    const PP start_pt = J.orig[start_idx], end_pt = J.orig[end_idx];
    const double dist = end_pt.dist - start_pt.dist;
    double control_pt_len = cpl_in;
    if (control_pt_len == 0.0)
        control_pt_len = 0.4 * dist;
    V2 prev_orient{std::cos(start_pt.yaw), std::sin(start_pt.yaw)};
    V2 next_orient{std::cos(end_pt.yaw), std::sin(end_pt.yaw)};
    if (start_pt.cusp)
        prev_orient = prev_orient * -1.0;
    const int num_sample = (int)std::ceil(std::fmin((end_pt.dist - start_pt.dist) / pp.path_point_distance, (double)(upsample_factor * idx_diff)));
    V2 pt1 = start_pt.p + prev_orient * control_pt_len * start_pt.direc;
    V2 pt2 = end_pt.p - next_orient * (0.4 * dist) * end_pt.direc;
    control_pt_len = 0.4 * dist;
    if (num_sample * pp.path_point_distance <= 2 * pp.step_size)
    {
        pt1 = start_pt.p + prev_orient * control_pt_len * start_pt.direc;
        pt2 = end_pt.p - next_orient * (0.1 * dist) * end_pt.direc;
        control_pt_len = 0.1 * dist;
    }
    out.end_idx = end_idx;
    out.cpl_after = control_pt_len;
    std::vector<PP> &temp = out.temp;
    temp.clear();
    temp.push_back(start_pt);
    for (int j = 1; j <= num_sample; ++j)
    {
        const double mu = (double)j / num_sample;
        const V2 np = cubic_bezier(start_pt.p, pt1, pt2, end_pt.p, mu);
        PP q;
        q.p = np;
        q.yaw = 0.0;
        q.direc = end_pt.direc;
        q.steering = end_pt.steering;
        temp.push_back(q);
        if (j > 1)
        {
            V2 orient = tangent_dir(temp[j - 2].p, temp[j - 1].p, np, false);
            if (end_pt.direc < 0)
                orient = neg(orient);
            temp[j - 1].yaw = std::atan2(orient.y, orient.x);
        }
        if (j == num_sample)
        {
            temp.back().yaw = end_pt.yaw;
            temp.back().cusp = end_pt.cusp;
        }
    }
    // a degenerate arc (no sample: coincident points) cannot be evaluated by the reference either (it reads past its
    // one-point temp_path); it is rejected here
    out.ok_curv = false;
    if (num_sample >= 1)
    {
        std::vector<PP> copy = temp; // curvature of the samples; the accepted points get theirs recomputed at the end
        const double max_curv_seg = update_curvature_new(copy);
        out.ok_curv = !(max_curv_seg > 1.5 * pp.max_curvature);
    }
}

// candidates of the segment that starts at J.i (hybrid_a_star.cpp:936-975)
void open_segment(Job &J, const mpros_planner_params &pp, int upsample_factor)
{
    const int n = (int)J.orig.size();
    J.start_idx = J.i;
    J.out.push_back(J.orig[J.start_idx]);
    int end_idx = n - 1;
    for (int c : J.cusp_idx)
        if (c > J.i)
        {
            end_idx = c;
            break;
        }
    const int max_iter = end_idx - J.start_idx;
    J.cands.assign((size_t)std::max(0, max_iter), Candidate());
    double cpl = J.control_pt_len;
    for (int k = 0; k < max_iter; ++k)
    {
        build_candidate(J, J.start_idx, end_idx - k, cpl, pp, upsample_factor, J.cands[k]);
        cpl = J.cands[k].cpl_after;
    }
}
} // namespace
} // namespace mpros

using namespace mpros;

extern "C" int mpros_path_postprocess_batch(mpros_ctx *c, const double *paths5, const int64_t *offsets, size_t n_paths, double *out8,
                                            size_t out_cap_points, int64_t *out_offsets)
{
    if (!c || !offsets || !out_offsets || (n_paths && !paths5))
    {
        set_error("mpros_path_postprocess_batch: NULL argument");
        return MPROS_ERR_INVALID;
    }
    if (!c->get_map || !c->planner_configured)
    {
        set_error("mpros_path_postprocess_batch: needs a map and mpros_planner_config first");
        return MPROS_ERR_INVALID;
    }
    const mpros_planner_params &pp = c->pp;
    const int upsample_factor = (int)std::ceil(pp.step_size / pp.path_point_distance); // hybrid_a_star.cpp:164
    std::vector<Job> jobs(n_paths);
    for (size_t q = 0; q < n_paths; ++q)
    {
        Job &J = jobs[q];
        const int64_t b = offsets[q], e = offsets[q + 1];
        if (e < b)
        {
            set_error("mpros_path_postprocess_batch: offsets must be non-decreasing");
            return MPROS_ERR_INVALID;
        }
        J.orig.resize((size_t)(e - b));
        for (int64_t k = b; k < e; ++k)
        {
            PP &t = J.orig[(size_t)(k - b)];
            t.p = {paths5[5 * k], paths5[5 * k + 1]};
            t.yaw = paths5[5 * k + 2];
            t.steering = paths5[5 * k + 3];
            t.direc = paths5[5 * k + 4];
        }
        if (J.orig.size() < 2)
        {
            // nothing to measure or interpolate (the reference's helpers read path[1]); returned as it is
            J.out = J.orig;
            J.done = true;
            continue;
        }
        update_path_length(J.orig);
        set_cusp(J.orig);
        update_curvature_new(J.orig);
        if (!pp.interpolation_enable)
        {
            J.out = J.orig;
            J.done = true;
            continue;
        }
        for (int i = 0; i + 1 < (int)J.orig.size(); ++i)
            if (J.orig[i].cusp)
                J.cusp_idx.push_back(i);
    }
    // rounds: one segment of every unfinished path per round, one batched footprint launch per round
    std::vector<double> poses;
    std::vector<int64_t> poffs;
    std::vector<uint8_t> verdicts;
    while (true)
    {
        poses.clear();
        poffs.assign(1, 0);
        bool any = false;
        for (Job &J : jobs)
        {
            if (J.done)
                continue;
            any = true;
            open_segment(J, pp, upsample_factor);
            for (const Candidate &cd : J.cands)
            {
                // the poses the reference tests: temp_path[0 .. num_sample - 1] (hybrid_a_star.cpp:1081)
                for (size_t k = 0; k + 1 < cd.temp.size(); ++k)
                {
                    poses.push_back(cd.temp[k].p.x);
                    poses.push_back(cd.temp[k].p.y);
                    poses.push_back(cd.temp[k].yaw);
                }
                poffs.push_back((int64_t)(poses.size() / 3));
            }
        }
        if (!any)
            break;
        verdicts.assign(poffs.size() - 1, 0);
        if (!verdicts.empty())
        {
            int rc = mpros_collision_check_paths(c, poses.data(), poffs.data(), verdicts.size(), verdicts.data());
            if (rc)
                return rc;
        }
        size_t v = 0;
        for (Job &J : jobs)
        {
            if (J.done)
                continue;
            int next_i = J.start_idx + 1; // no arc found: continue from the next point (hybrid_a_star.cpp:966-975)
            double cpl = 0.0;
            bool found = false;
            for (const Candidate &cd : J.cands)
            {
                const bool hit = verdicts[v++] != 0;
                if (found || hit || !cd.ok_curv)
                    continue;
                found = true;
                J.out.insert(J.out.end(), cd.temp.begin() + 1, cd.temp.end() - 1);
                next_i = cd.end_idx;
                cpl = cd.cpl_after;
            }
            J.i = next_i;
            J.control_pt_len = cpl;
            J.cands.clear();
            if (J.i >= (int)J.orig.size() - 1)
            {
                J.out.push_back(J.orig.back());
                update_curvature_new(J.out);
                update_path_length(J.out);
                J.done = true;
            }
        }
    }
    size_t total = 0;
    out_offsets[0] = 0;
    for (size_t q = 0; q < n_paths; ++q)
    {
        total += jobs[q].out.size();
        out_offsets[q + 1] = (int64_t)total;
    }
    if (total > out_cap_points || (total && !out8))
    {
        set_error("mpros_path_postprocess_batch: output buffer too small (out_offsets hold the required sizes)");
        return MPROS_ERR_CAPACITY;
    }
    size_t k = 0;
    for (const Job &J : jobs)
        for (const PP &t : J.out)
        {
            double *o = out8 + 8 * k++;
            o[0] = t.p.x, o[1] = t.p.y, o[2] = t.yaw, o[3] = t.curv, o[4] = t.direc, o[5] = t.steering, o[6] = t.cusp ? 1.0 : 0.0, o[7] = t.dist;
        }
    return MPROS_OK;
}
