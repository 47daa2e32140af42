// TEST INFRASTRUCTURE ONLY — CPU oracle ("port"). Never included, linked or executed by the product path;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load libmpros_oracle.so.
//
// A plain C++ restatement of the reference's algorithm for the hot path, each function citing the reference
// file:line it follows. It is pinned against the reference's own code compiled unmodified (oracle/_ref, see
// tests/test_oracle_port.py) and against the golden fixtures under tests/golden/ generated from that build.
// The OMPL Reeds-Shepp distance (ompl_rs_distance.h) is the one PARITY-UNPINNED piece.
//
// Where the reference is order-dependent (Voronoi region ties, obstacle-side edge flags) this
// restatement implements the canonical, order-free contract the CUDA path also implements, and the tests report
// the measured divergence from the reference.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <unordered_map>
#include <vector>

#include "ompl_rs_distance.h"

namespace port
{
static const int HUGE_INT = 1000; // node.h:4

struct MapParams
{
    double occupied_thresh = 0.68, free_thresh = 0.2; // nav_map.cpp:29-30
    float voro_alpha = 10.0f, voro_dmax = 50.0f;      // :31-32
    int ds = 1;                                       // :33
    double inflation_radius = 0.0;                    // :34
};

struct PlannerParams
{
    // hybrid_a_star.cpp:128-161
    double max_curvature = 0.4, width = 2.0, wheelbase = 3.0, length = 4.2, steer_offset = 0.0, max_steer = 0.35;
    double pen_gear = 10.0, pen_back = 0.0, pen_steer = 0.1, pen_steer_change = 10.0, rs_range = 10.0;
    int num_yaw = 36;
    double step = 1.1, granularity = 0.2;
    int num_steer = 3, max_iter = 10000, optimal = 0;
};

struct Map
{
    MapParams p;
    bool get_map = false, get_goal = false;
    int H0 = 0, W0 = 0, H = 0, W = 0;
    double res0 = 0, res = 0, ox = 0, oy = 0, oyaw = 0, osin = 0, ocos = 1;
    double goal_x = 0, goal_y = 0;
    std::vector<int8_t> static_orig, static_ds, inflate;
    std::vector<int> goal, obs_dist, region, edge_dist;
    std::vector<uint8_t> isedge;
    std::vector<float> voronoi;
    int num_regions = 0;

    // nav_map.h:144-167 — truncation toward zero; out-of-int-range doubles behave as "outside"
    bool posToIndexImpl(double x, double y, double r, int Hh, int Ww, int &i, int &j) const
    {
        double dx = x - ox, dy = y - oy;
        double qj = (dx * ocos + dy * osin) / r;
        double qi = (dx * -osin + dy * ocos) / r;
        bool ok = qi > -1.0 && qi < (double)Hh && qj > -1.0 && qj < (double)Ww;
        i = ok ? (int)qi : -1;
        j = ok ? (int)qj : -1;
        return ok;
    }
    // nav_map.cpp:259-277 on a world point
    bool hitsOrig(const std::vector<int8_t> &layer, double x, double y) const
    {
        int i, j;
        if (!posToIndexImpl(x, y, res0, H0, W0, i, j))
            return true;
        return layer[(size_t)i * W0 + j] != 0;
    }
    bool inCollisionDs(int i, int j) const // nav_map.cpp:249-257
    {
        if (i < 0 || i >= H || j < 0 || j >= W)
            return true;
        return static_ds[(size_t)i * W + j] != 0;
    }
};

// ---- N1: nav_map.cpp:114-131 --------------------------------------------------------------------------------
static void threshold(Map &m, const int8_t *data)
{
    size_t n = (size_t)m.H0 * m.W0;
    for (size_t k = 0; k < n; ++k)
    {
        if ((data[k] > m.p.free_thresh || data[k] < 0) && m.static_orig[k] == 0)
            m.static_orig[k] = 1;
        else if (data[k] < m.p.free_thresh && m.static_orig[k] == 1)
            m.static_orig[k] = 0;
    }
}

// ---- N2: nav_map.cpp:592-652 (off-by-one quirks kept; index == size is UB in the reference, 0 here) --------------
static void downsample(Map &m)
{
    const int ds = m.p.ds;
    m.H = (int)std::ceil(static_cast<double>(m.H0) / ds);
    m.W = (int)std::ceil(static_cast<double>(m.W0) / ds);
    m.res = m.res0 * ds;
    if (ds == 1)
    {
        m.static_ds = m.static_orig;
        return;
    }
    m.static_ds.assign((size_t)m.H * m.W, 0);
    const size_t size0 = m.static_orig.size();
    for (int i = 0; i < m.H; ++i)
        for (int j = 0; j < m.W; ++j)
        {
            int8_t value = 0;
            for (int a = 0; a < ds; ++a)
            {
                int oi = a + i * ds;
                if (oi > m.H0)
                    break;
                for (int b = 0; b < ds; ++b)
                {
                    int oj = b + j * ds;
                    if (oj > m.W0)
                        break;
                    size_t k = (size_t)oi * m.W0 + oj;
                    if (k >= size0)
                        continue;
                    value = (int8_t)std::fmax(value, m.static_orig[k]);
                }
            }
            m.static_ds[(size_t)i * m.W + j] = value;
        }
}

// ---- N3: nav_map.cpp:164-214 — dilation by {(di,dj): (|di|-.5)^2 + (|dj|-.5)^2 <= r^2}, clipped to the map.
// Restated separably: per-row distance to the nearest occupied cell d1, then inflated(i,j) <=> exists |di| <= r with
// d1(i+di, j) <= w(|di|), w(a) = max{b : (a-.5)^2 + (b-.5)^2 <= r^2}. Pinned against the reference's direct loop.
static void inflate(Map &m)
{
    const int r = (int)std::ceil(m.p.inflation_radius / m.res0);
    const double r_sq = r * r;
    std::vector<int> w(r + 1, -1);
    for (int a = 0; a <= r; ++a)
        for (int b = 0; b <= r; ++b)
        {
            double i_dist = std::abs(a) - 0.5, j_dist = std::abs(b) - 0.5;
            if (i_dist * i_dist + j_dist * j_dist <= r_sq)
                w[a] = b;
        }
    const int H = m.H0, W = m.W0;
    const int INF = 1 << 29;
    std::vector<int> d1((size_t)H * W, INF);
    for (int i = 0; i < H; ++i)
    {
        int last = -INF;
        for (int j = 0; j < W; ++j)
        {
            if (m.static_orig[(size_t)i * W + j])
                last = j;
            if (last > -INF)
                d1[(size_t)i * W + j] = j - last;
        }
        last = INF;
        for (int j = W - 1; j >= 0; --j)
        {
            if (m.static_orig[(size_t)i * W + j])
                last = j;
            if (last < INF)
                d1[(size_t)i * W + j] = std::min(d1[(size_t)i * W + j], last - j);
        }
    }
    m.inflate.assign((size_t)H * W, 0);
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j)
        {
            int8_t v = 0;
            for (int di = -r; di <= r && !v; ++di)
            {
                int ii = i + di;
                if (ii < 0 || ii >= H)
                    continue;
                int ww = w[std::abs(di)];
                if (ww >= 0 && d1[(size_t)ii * W + j] <= ww)
                    v = 1;
            }
            m.inflate[(size_t)i * W + j] = v;
        }
}

// ---- N4: nav_map.cpp:279-349 — label-correcting relaxation == BFS capped at 999 -------------------------------
static void generateGoalMap(Map &m)
{
    int gi, gj;
    if (!m.posToIndexImpl(m.goal_x, m.goal_y, m.res, m.H, m.W, gi, gj))
    {
        m.get_goal = false; // :297-302
        return;
    }
    std::vector<int> cost((size_t)m.H * m.W, HUGE_INT);
    cost[(size_t)gi * m.W + gj] = 0; // even if occupied (:304)
    std::deque<int> q;
    q.push_back(gi * m.W + gj);
    const int di[4] = {-1, 0, 1, 0}, dj[4] = {0, 1, 0, -1}; // :351-382
    while (!q.empty())
    {
        int k = q.front();
        q.pop_front();
        int i = k / m.W, j = k % m.W, c = cost[k];
        for (int d = 0; d < 4; ++d)
        {
            int ni = i + di[d], nj = j + dj[d];
            if (m.inCollisionDs(ni, nj))
                continue;
            int &nc = cost[(size_t)ni * m.W + nj];
            if (nc > c + 1) // :329 — values never exceed 999
            {
                nc = c + 1;
                q.push_back(ni * m.W + nj);
            }
        }
    }
    m.goal.swap(cost);
}

// ---- N5-N9: nav_map.cpp:384-590 ----------------------------------------------------------------------------
static void generateVoronoi(Map &m)
{
    const int H = m.H, W = m.W;
    const size_t n = (size_t)H * W;
    const int di[4] = {-1, 0, 1, 0}, dj[4] = {0, 1, 0, -1};
    // N5 markObs / floodFillObs (:384-432): components in raster order of their first cell
    m.region.assign(n, -1);
    int region = 0;
    std::vector<int> stack;
    for (size_t k = 0; k < n; ++k)
    {
        if (!m.static_ds[k] || m.region[k] != -1)
            continue;
        m.region[k] = region;
        stack.push_back((int)k);
        while (!stack.empty())
        {
            int c = stack.back();
            stack.pop_back();
            int i = c / W, j = c % W;
            for (int d = 0; d < 4; ++d)
            {
                int ni = i + di[d], nj = j + dj[d];
                if (ni < 0 || ni >= H || nj < 0 || nj >= W)
                    continue;
                size_t nk = (size_t)ni * W + nj;
                if (m.static_ds[nk] && m.region[nk] == -1)
                {
                    m.region[nk] = region;
                    stack.push_back((int)nk);
                }
            }
        }
        ++region;
    }
    m.num_regions = region;
    // N6 calObsDist (:434-473): L1 distance to the nearest occupied cell (two-pass city-block transform), cap 999;
    // region of a free cell = SMALLEST region id among all nearest obstacle cells (canonical tie-break).
    typedef unsigned long long u64;
    const u64 ONE = 1ull << 32, INF = (1ull << 24) << 32;
    std::vector<u64> key(n, INF);
    for (size_t k = 0; k < n; ++k)
        if (m.static_ds[k])
            key[k] = (u64)(unsigned)m.region[k];
    auto relax = [&](size_t a, size_t b) { // a <- b + 1
        u64 v = key[b] + ONE;
        if (v < key[a])
            key[a] = v;
    };
    auto l1_transform = [&]() {
        for (int i = 0; i < H; ++i)
        {
            for (int j = 1; j < W; ++j)
                relax((size_t)i * W + j, (size_t)i * W + j - 1);
            for (int j = W - 2; j >= 0; --j)
                relax((size_t)i * W + j, (size_t)i * W + j + 1);
        }
        for (int i = 1; i < H; ++i)
            for (int j = 0; j < W; ++j)
                relax((size_t)i * W + j, (size_t)(i - 1) * W + j);
        for (int i = H - 2; i >= 0; --i)
            for (int j = 0; j < W; ++j)
                relax((size_t)i * W + j, (size_t)(i + 1) * W + j);
    };
    l1_transform();
    m.obs_dist.assign(n, HUGE_INT);
    for (size_t k = 0; k < n; ++k)
    {
        unsigned d = (unsigned)(key[k] >> 32);
        if (d < (unsigned)HUGE_INT)
        {
            m.obs_dist[k] = (int)d;
            m.region[k] = (int)(unsigned)(key[k] & 0xffffffffull);
        }
        else
            m.region[k] = -1;
    }
    // N7 findEdge (:475-507): reached free cell is an edge iff an in-map neighbour has another region; any other
    // cell iff it is the first differing neighbour (up, right, down, left) of a reached free neighbour.
    m.isedge.assign(n, 0);
    auto first_differing = [&](int i, int j) {
        int r = m.region[(size_t)i * W + j];
        for (int d = 0; d < 4; ++d)
        {
            int ni = i + di[d], nj = j + dj[d];
            if (ni < 0 || ni >= H || nj < 0 || nj >= W)
                continue;
            if (m.region[(size_t)ni * W + nj] != r)
                return d;
        }
        return -1;
    };
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j)
        {
            size_t k = (size_t)i * W + j;
            if (m.static_ds[k] || m.obs_dist[k] >= HUGE_INT)
                continue;
            int d = first_differing(i, j);
            if (d >= 0)
            {
                m.isedge[k] = 1;
                m.isedge[(size_t)(i + di[d]) * W + j + dj[d]] = 1;
            }
        }
    // N8 calEdgeDist (:509-539): L1 distance to the nearest edge cell through all cells, cap 999
    for (size_t k = 0; k < n; ++k)
        key[k] = m.isedge[k] ? 0ull : INF;
    l1_transform();
    m.edge_dist.assign(n, HUGE_INT);
    for (size_t k = 0; k < n; ++k)
    {
        unsigned d = (unsigned)(key[k] >> 32);
        if (d < (unsigned)HUGE_INT)
            m.edge_dist[k] = (int)d;
    }
    // N9 cost loop (:565-582), float, left to right
    m.voronoi.assign(n, 0.f);
    const float voro_alpha_ = m.p.voro_alpha, voro_d_o_max_ = m.p.voro_dmax;
    for (size_t k = 0; k < n; ++k)
    {
        if (m.static_ds[k])
            m.voronoi[k] = 1;
        else if (m.isedge[k])
            m.voronoi[k] = 0;
        else
        {
            int obs_dist_ = m.obs_dist[k], edge_dist_ = m.edge_dist[k];
            m.voronoi[k] = (voro_alpha_ / (voro_alpha_ + (float)obs_dist_)) *
                           ((float)edge_dist_ / (float)(edge_dist_ + obs_dist_)) *
                           ((float)obs_dist_ - voro_d_o_max_) * ((float)obs_dist_ - voro_d_o_max_) /
                           voro_d_o_max_ / voro_d_o_max_;
        }
    }
}

// nav_map.cpp:42-162 (first map / same geometry; geometry change keeps the linear prefix like vector::resize)
static void setOccupancy(Map &m, const int8_t *data, int W, int H, double res, double ox, double oy, double oyaw)
{
    bool change = !m.get_map || m.H0 != H || m.W0 != W || m.res0 != res || m.ox != ox || m.oy != oy || m.oyaw != oyaw;
    if (change)
    {
        m.H0 = H;
        m.W0 = W;
        m.res0 = res;
        m.ox = ox;
        m.oy = oy;
        m.oyaw = oyaw;
        m.osin = std::sin(oyaw);
        m.ocos = std::cos(oyaw);
        m.static_orig.resize((size_t)H * W, 0);
    }
    threshold(m, data);
    downsample(m);
    inflate(m);
    m.get_map = true;
    if (m.get_goal)
        generateGoalMap(m);
    generateVoronoi(m);
}

// ---- planner constants (hybrid_a_star.cpp:120-206) -----------------------------------------------------------
struct Planner
{
    PlannerParams p;
    const Map *map = nullptr;
    double longi_x0, longi_x1, corner_x[4], corner_y[4];
    int n_check;
    std::vector<double> steer_set;
    double min_r, yaw_resol;

    void config(const Map *m, const PlannerParams &pp)
    {
        map = m;
        p = pp;
        const double L = p.length, Wd = p.width, off = p.steer_offset;
        corner_x[0] = L / 2 + off, corner_y[0] = Wd / 2;
        corner_x[1] = L / 2 + off, corner_y[1] = -Wd / 2;
        corner_x[2] = -L / 2 + off, corner_y[2] = -Wd / 2;
        corner_x[3] = -L / 2 + off, corner_y[3] = Wd / 2;
        longi_x0 = L / 2 + off - 1.0 / 2 * Wd;
        longi_x1 = -L / 2 + off + 1.0 / 2 * Wd;
        n_check = (int)std::ceil((L - Wd) / m->res0);
        min_r = 1.0 / p.max_curvature;
        yaw_resol = 2 * M_PI / p.num_yaw;
        steer_set.clear();
        int num_pos = p.num_steer / 2;
        for (int i = 1; i <= num_pos; ++i)
        {
            steer_set.push_back(p.max_steer / num_pos * i);
            steer_set.push_back(-p.max_steer / num_pos * i);
        }
        steer_set.push_back(0.0);
    }

    // C1: hybrid_a_star.cpp:617-680
    bool collision4(double px, double py, double yaw) const
    {
        double c = std::cos(yaw), s = std::sin(yaw);
        double x0 = c * longi_x0 + px, y0 = s * longi_x0 + py;
        double x1 = c * longi_x1 + px, y1 = s * longi_x1 + py;
        double n = (double)n_check;
        double ox = (x1 - x0) / n, oy = (y1 - y0) / n;
        for (int k = 0; k <= n_check; ++k)
            if (map->hitsOrig(map->inflate, ox * (double)k + x0, oy * (double)k + y0))
                return true;
        for (int k = 0; k < 4; ++k)
        {
            double cx = corner_x[k], cy = corner_y[k];
            if (map->hitsOrig(map->static_orig, cx * c - cy * s + px, cx * s + cy * c + py))
                return true;
        }
        return false;
    }
};

// ---- Reeds-Shepp generator (mpros_rs_path): rs_curve.cpp / rs_curve_basic.h ------------------------------------
struct Seg
{
    double len; // |value|
    int steer;  // LEFT 1, STRAIGHT 0, RIGHT -1
    int dir;    // FORWARD 1, BACKWARD -1
};
struct RsPath
{
    Seg s[5];
    int n = 0;
    void push(double value, int steer, int dir)
    {
        // PathSeg ctor, rs_curve_basic.h:84-88: direction flips unless value > 0 (so also for 0 and NaN)
        s[n].len = std::fabs(value);
        s[n].steer = steer;
        s[n].dir = (value > 0 ? dir : -dir);
        ++n;
    }
};
static inline double regularRad(double t) // rs_curve_basic.h:113-124
{
    if (t > M_PI)
        return t - 2 * M_PI;
    else if (t < -M_PI)
        return t + 2 * M_PI;
    return t;
}
struct Polar // rs_curve_basic.h:56-62
{
    double dist, theta;
    Polar(double x, double y) : dist(std::hypot(x, y)), theta(std::atan2(y, x)) {}
};
enum { L_ = 1, S_ = 0, R_ = -1, F_ = 1, B_ = -1 };

// rs_curve.cpp:353-599, formula by formula (including path4's asin, :418)
static RsPath rs_formula(int f, double x, double y, double phi)
{
    RsPath p;
    switch (f)
    {
    case 0:
    {
        Polar po(x - sin(phi), y - 1 + cos(phi));
        double u = po.dist, t = po.theta, v = regularRad(phi - t);
        p.push(t, L_, F_), p.push(u, S_, F_), p.push(v, L_, F_);
        break;
    }
    case 1:
    {
        Polar po(x + sin(phi), y - 1 - cos(phi));
        if (po.dist * po.dist >= 4)
        {
            double u = sqrt(po.dist * po.dist - 4);
            double t = regularRad(po.theta + std::atan2(2, u));
            double v = regularRad(t - phi);
            p.push(t, L_, F_), p.push(u, S_, F_), p.push(v, R_, F_);
        }
        break;
    }
    case 2:
    {
        Polar po(x - sin(phi), y - 1 + cos(phi));
        if (po.dist <= 4)
        {
            double A = acos(po.dist / 4);
            double u = regularRad(M_PI - 2 * A);
            double t = regularRad(M_PI_2 + A + po.theta);
            double v = regularRad(phi - t - u);
            p.push(t, L_, F_), p.push(u, R_, B_), p.push(v, L_, F_);
        }
        break;
    }
    case 3:
    {
        Polar po(x - sin(phi), y - 1 + cos(phi));
        if (po.dist <= 4)
        {
            double A = acos(po.dist / 4);
            double u = regularRad(M_PI - 2 * A);
            double t = asin(po.theta + M_PI_2 + A);
            double v = regularRad(t + u - phi);
            p.push(t, L_, F_), p.push(u, R_, B_), p.push(v, L_, B_);
        }
        break;
    }
    case 4:
    {
        Polar po(x - sin(phi), y - 1 + cos(phi));
        if (po.dist <= 4)
        {
            double u = acos(1 - po.dist * po.dist / 8);
            double A = asin(2 * sin(u) / po.dist);
            double t = regularRad(po.theta + M_PI_2 - A);
            double v = regularRad(t - u - phi);
            p.push(t, L_, F_), p.push(u, R_, F_), p.push(v, L_, B_);
        }
        break;
    }
    case 5:
    {
        Polar po(x + sin(phi), y - 1 - cos(phi));
        if (po.dist <= 4)
        {
            double A, t, u, v;
            if (po.dist <= 2)
            {
                A = acos((po.dist + 2) / 4);
                t = regularRad(po.theta + M_PI_2 + A);
                u = regularRad(A);
                v = regularRad(phi - t + 2 * u);
            }
            else
            {
                A = acos((po.dist - 2) / 4);
                t = regularRad(po.theta + M_PI_2 - A);
                u = regularRad(M_PI - A);
                v = regularRad(phi - t + 2 * u);
            }
            p.push(t, L_, F_), p.push(u, R_, F_), p.push(u, L_, B_), p.push(v, R_, B_);
        }
        break;
    }
    case 6:
    {
        Polar po(x + sin(phi), y - 1 - cos(phi));
        double u1 = (20 - po.dist * po.dist) / 16;
        if (po.dist <= 6 && u1 >= 0 && u1 <= 1)
        {
            double u = acos(u1);
            double A = asin(2 * sin(u) / po.dist);
            double t = regularRad(po.theta + M_PI_2 + A);
            double v = regularRad(t - phi);
            p.push(t, L_, F_), p.push(u, R_, B_), p.push(u, L_, B_), p.push(v, R_, F_);
        }
        break;
    }
    case 7:
    {
        Polar po(x - sin(phi), y - 1 + cos(phi));
        if (po.dist >= 2)
        {
            double u = sqrt(po.dist * po.dist - 4) - 2;
            double A = atan2(2, u + 2);
            double t = regularRad(po.theta + M_PI_2 + A);
            double v = regularRad(t - phi + M_PI_2);
            p.push(t, L_, F_), p.push(M_PI_2, R_, B_), p.push(u, S_, B_), p.push(v, L_, B_);
        }
        break;
    }
    case 8:
    {
        Polar po(x + sin(phi), y - 1 - cos(phi));
        if (po.dist >= 2)
        {
            double t = regularRad(po.theta + M_PI_2);
            double u = po.dist - 2;
            double v = regularRad(phi - t - M_PI_2);
            p.push(t, L_, F_), p.push(M_PI_2, R_, B_), p.push(u, S_, B_), p.push(v, R_, B_);
        }
        break;
    }
    case 9:
    {
        Polar po(x - sin(phi), y - 1 + cos(phi));
        if (po.dist >= 2)
        {
            double u = sqrt(po.dist * po.dist - 4) - 2;
            double A = atan2(u + 2, 2);
            double t = regularRad(po.theta + M_PI_2 - A);
            double v = regularRad(t - phi - M_PI_2);
            p.push(t, L_, F_), p.push(u, S_, F_), p.push(M_PI_2, R_, F_), p.push(v, L_, B_);
        }
        break;
    }
    case 10:
    {
        Polar po(x + sin(phi), y - 1 - cos(phi));
        if (po.dist >= 2)
        {
            double u = po.dist - 2;
            double t = regularRad(po.theta);
            double v = regularRad(phi - t - M_PI_2);
            p.push(t, L_, F_), p.push(u, S_, F_), p.push(M_PI_2, L_, F_), p.push(v, R_, B_);
        }
        break;
    }
    case 11:
    {
        Polar po(x + sin(phi), y - 1 - cos(phi));
        if (po.dist >= 4)
        {
            double u = sqrt(po.dist * po.dist - 4) - 4;
            double A = atan2(2, u + 4);
            double t = regularRad(po.theta + M_PI_2 + A);
            double v = regularRad(t - phi);
            p.push(t, L_, F_), p.push(M_PI_2, R_, B_), p.push(u, S_, B_), p.push(M_PI_2, L_, B_), p.push(v, R_, F_);
        }
        break;
    }
    }
    return p;
}

struct RsConfig
{
    double radius, steer_angle, step;
    // as RSCurve::setPenalty receives them POSITIONALLY (rs_curve.cpp:314-323): gear, backward, steer_change, steer_angle
    double pen_gear, pen_back, pen_steer_change, pen_steer_angle;
};

// Timeflip / Reflect (rs_curve.cpp:31-49) go through the PathSeg ctor again with the (positive) length, so a
// zero-length or NaN segment flips its direction once more (rs_curve_basic.h:84-88,131-144).
static RsPath rs_timeflip(const RsPath &a)
{
    RsPath r;
    for (int k = 0; k < a.n; ++k)
        r.push(a.s[k].len, a.s[k].steer, -a.s[k].dir);
    return r;
}
static RsPath rs_reflect(const RsPath &a)
{
    RsPath r;
    for (int k = 0; k < a.n; ++k)
        r.push(a.s[k].len, -a.s[k].steer, a.s[k].dir);
    return r;
}

// rs_curve.cpp:60-97
static double rs_path_length(const RsConfig &c, const RsPath &path, int start_direc, double start_steer)
{
    double length = 0, penalty = 0;
    int last_steer = start_steer; // int initialised from a double (:64)
    if (path.n == 0)
        return 10000;
    for (int i = 0; i < path.n; ++i)
    {
        const Seg &seg = path.s[i];
        if (seg.dir != start_direc)
            penalty += c.pen_gear;
        if (seg.dir < 0)
            penalty += c.pen_back * seg.len * c.radius;
        if (seg.steer != last_steer)
            penalty += c.pen_steer_change * std::abs(seg.steer - last_steer) * c.steer_angle;
        if (seg.steer != 0)
            penalty += c.pen_steer_angle * std::abs(seg.steer) * c.steer_angle;
        length += seg.len * c.radius;
        last_steer = seg.steer;
    }
    return (length + penalty) * (1.0 + 1e-3);
}

// FindOptimalPath, rs_curve.cpp:111-163: 12 formulas x {id, timeflip, reflect, both}; multimap.begin() == smallest
// cost, first inserted among equals; candidates with a NaN segment (or cost 0) are dropped.
static double rs_find_optimal(const RsConfig &c, const double *start5, const double *goal3, RsPath &best)
{
    const double sx = start5[0], sy = start5[1], sphi = start5[2];
    const int start_direc = (int)start5[3];
    const double start_steer = start5[4];
    // NewGoalPoint (:51-58)
    double dx = goal3[0] - sx, dy = goal3[1] - sy;
    double gx = (dx * cos(sphi) + dy * sin(sphi)) / c.radius;
    double gy = (-dx * sin(sphi) + dy * cos(sphi)) / c.radius;
    double gphi = goal3[2] - sphi;
    bool have = false;
    double best_len = 0;
    best.n = 0;
    for (int f = 0; f < 12; ++f)
        for (int j = 0; j < 4; ++j)
        {
            RsPath path;
            switch (j)
            {
            case 0:
                path = rs_formula(f, gx, gy, gphi);
                break;
            case 1:
                path = rs_timeflip(rs_formula(f, -gx, gy, -gphi));
                break;
            case 2:
                path = rs_reflect(rs_formula(f, gx, -gy, -gphi));
                break;
            case 3:
                // goal_pt.timeflip().reflect() = (-x, -y, phi)  (rs_curve_basic.h:44-50)
                path = rs_reflect(rs_timeflip(rs_formula(f, -gx, -gy, gphi)));
                break;
            }
            double len = rs_path_length(c, path, start_direc, start_steer);
            bool nan_seg = false;
            for (int k = 0; k < path.n; ++k)
                if (std::isnan(path.s[k].len))
                    nan_seg = true;
            if (len && !nan_seg) // :150 (a NaN cost is "true")
            {
                // std::multimap ordering: strict <, NaN keys never compare less -> kept after everything
                if (!have || len < best_len)
                {
                    have = true;
                    best_len = len;
                    best = path;
                }
            }
        }
    return best_len;
}

// GenTraj, rs_curve.cpp:165-222. out: n x 5 {x, y, yaw, steer*steer_angle, direc}
static int rs_gen_traj(const RsConfig &c, const double *start3, const RsPath &path, double *out, int cap)
{
    const double curve_min_num_ = 3;
    double curr_x = start3[0], curr_y = start3[1], curr_yaw = start3[2];
    int n = 0;
    for (int k = 0; k < path.n; ++k)
    {
        const Seg &seg = path.s[k];
        int steer = seg.steer;
        if (steer != 0)
        {
            int num_pt = (int)std::fmax(std::ceil(seg.len * c.radius / c.step), curve_min_num_);
            double dyaw = seg.len / num_pt;
            double direc = seg.dir;
            for (int i = 0; i < num_pt; ++i)
            {
                double dx = sin(dyaw) * c.radius * direc;
                double dy = (1 - cos(dyaw)) * c.radius * steer;
                // compound assignment: the right-hand side is evaluated with the OLD curr_x/curr_yaw
                double nx = curr_x + (dx * cos(curr_yaw) - dy * sin(curr_yaw));
                double ny = curr_y + (dx * sin(curr_yaw) + dy * cos(curr_yaw));
                curr_x = nx;
                curr_y = ny;
                curr_yaw += dyaw * direc * (double)steer;
                curr_yaw = regularRad(curr_yaw);
                if (n < cap)
                {
                    double *o = out + 5 * n;
                    o[0] = curr_x, o[1] = curr_y, o[2] = curr_yaw, o[3] = (double)steer * c.steer_angle, o[4] = direc;
                }
                ++n;
            }
        }
        else
        {
            double direc = seg.dir;
            int num_pt = (int)std::fmax(std::ceil(seg.len * c.radius / c.step), curve_min_num_);
            double step = seg.len * c.radius / num_pt;
            for (int i = 0; i < num_pt; ++i)
            {
                curr_x += step * cos(curr_yaw) * direc;
                curr_y += step * sin(curr_yaw) * direc;
                if (n < cap)
                {
                    double *o = out + 5 * n;
                    o[0] = curr_x, o[1] = curr_y, o[2] = curr_yaw, o[3] = (double)steer * c.steer_angle, o[4] = direc;
                }
                ++n;
            }
        }
    }
    return n;
}

// ---- Hybrid A* (hybrid_a_star.cpp:52-118, 208-402, 682-900) -----------------------------------------------------
struct Node
{
    double x, y, yaw, steer, dir, g, h, f;
    int i, j, parent;
    bool closed;
};

struct Search
{
    const Planner *pl;
    const Map *map;
    RsConfig rs;
    std::vector<Node> nodes;
    std::multimap<double, int> open;
    std::unordered_map<long long, std::pair<double, int>> cost_set; // index -> (g, node), absent = (+inf, null)
    int start = -1, goal = -1;
    double goal_xyz[3];
    std::vector<double> goal_path; // n x 5
    std::vector<double> tree;      // n x 8 like tree_
    bool initialized = false;
    long long n_rs_shots = 0, n_collision = 0;
    int num_nodes = 0;

    // hybrid_a_star.cpp:266-279
    static void regularYaw(double &yaw)
    {
        if (yaw <= -M_PI)
        {
            yaw += 2 * M_PI;
            return;
        }
        if (yaw > M_PI)
        {
            yaw -= 2 * M_PI;
            return;
        }
    }
    int yawToIdx(double yaw) const // :292-302
    {
        double yaw_pos = yaw;
        if (yaw < 0)
            yaw_pos = yaw + 2 * M_PI;
        return int(yaw_pos / pl->yaw_resol);
    }
    long long calIndex(int i, int j, int iyaw, double direc) const // :281-290 (64-bit to be safe on large maps)
    {
        int idx_direc = direc > 0 ? 0 : 1;
        return ((long long)i * map->W + j) * pl->p.num_yaw * 2 + (long long)iyaw * 2 + idx_direc;
    }
    int makeNode(double x, double y, double yaw, double steer, double dir, int parent)
    {
        Node n;
        n.x = x, n.y = y, n.yaw = yaw, n.steer = steer, n.dir = dir, n.g = 0, n.h = 0, n.f = 0, n.parent = parent, n.closed = false;
        map->posToIndexImpl(x, y, map->res, map->H, map->W, n.i, n.j); // hybrid_a_star_node.h:45
        nodes.push_back(n);
        return (int)nodes.size() - 1;
    }
    void setCost(Node &n, double c)
    {
        n.g = c;
        n.f = n.g + n.h;
    }
    void setHeuristicValue(Node &n, double h)
    {
        n.h = h;
        n.f = n.g + n.h;
    }
    void setNodeCost(int id) // :208-246
    {
        Node &n = nodes[id];
        if (n.parent < 0)
        {
            setCost(n, 0.0);
            return;
        }
        const Node &par = nodes[n.parent];
        const PlannerParams &p = pl->p;
        double pre_cost = par.g;
        double voro_cost = map->voronoi[(size_t)n.i * map->W + n.j];
        double path_cost = p.step;
        double direc_cost = 0;
        if (n.dir < 0)
            direc_cost += p.pen_back * p.step;
        if (n.dir != par.dir)
            direc_cost += p.pen_gear;
        double steer_cost = 0;
        if (n.steer != 0)
            steer_cost += p.pen_steer * std::abs(n.steer);
        if (n.steer != par.steer)
            steer_cost += p.pen_steer_change * std::abs(n.steer - par.steer);
        setCost(n, pre_cost + voro_cost + path_cost + direc_cost + steer_cost);
    }
    double heuristic(const Node &n) const // :248-264
    {
        double h1 = map->goal[(size_t)n.i * map->W + n.j] * map->res; // nav_map.h:231-234
        double h2 = ompl_rs::distance(pl->min_r, n.x, n.y, n.yaw, goal_xyz[0], goal_xyz[1], goal_xyz[2]);
        return std::fmax(h1, h2) * (1 + 1e-3);
    }

    bool initialize(const double *s, const double *g) // :52-118
    {
        if (pl->collision4(s[0], s[1], s[2]))
            return false;
        if (pl->collision4(g[0], g[1], g[2]))
            return false;
        nodes.clear();
        open.clear();
        cost_set.clear();
        start = makeNode(s[0], s[1], s[2], 0, 1, -1);
        goal = makeNode(g[0], g[1], g[2], 0, 1, -1);
        goal_xyz[0] = g[0], goal_xyz[1] = g[1], goal_xyz[2] = g[2];
        if (!map->get_goal)
            return false;
        setNodeCost(start);
        setHeuristicValue(nodes[start], heuristic(nodes[start]));
        setCost(nodes[goal], std::numeric_limits<double>::infinity());
        setHeuristicValue(nodes[goal], 0.0);
        long long si = calIndex(nodes[start].i, nodes[start].j, yawToIdx(nodes[start].yaw), 1);
        long long gi = calIndex(nodes[goal].i, nodes[goal].j, yawToIdx(nodes[goal].yaw), 1);
        cost_set[si] = std::make_pair(nodes[start].g, start);
        cost_set[gi] = std::make_pair(10000.0, goal); // :111-112 (overwrites the start entry if both share a cell)
        num_nodes += 2;
        initialized = true;
        return true;
    }

    void expandNode(int id) // :304-402
    {
        const PlannerParams &p = pl->p;
        const double curr_x = nodes[id].x, curr_y = nodes[id].y, curr_yaw = nodes[id].yaw;
        struct Pt
        {
            double x, y, yaw, steer, direc;
        };
        std::vector<Pt> next_points;
        const double direction_set[2] = {1.0, -1.0};
        for (double steer : pl->steer_set)
            for (double direc : direction_set)
            {
                double dx, dy, dyaw, next_x, next_y, next_yaw;
                if (steer == 0)
                {
                    dx = direc * p.step;
                    dy = 0;
                    next_x = dx * cos(curr_yaw) - dy * sin(curr_yaw) + curr_x;
                    next_y = dx * sin(curr_yaw) + dy * cos(curr_yaw) + curr_y;
                    next_yaw = curr_yaw;
                }
                else
                {
                    double radius = (p.wheelbase / 2 + p.steer_offset) / std::tan(std::abs(steer));
                    double steer_sign = steer / std::abs(steer);
                    dyaw = p.step / radius;
                    dx = sin(dyaw) * radius * direc;
                    dy = (1 - cos(dyaw)) * radius * steer_sign;
                    next_x = dx * cos(curr_yaw) - dy * sin(curr_yaw) + curr_x;
                    next_y = dx * sin(curr_yaw) + dy * cos(curr_yaw) + curr_y;
                    next_yaw = curr_yaw + dyaw * steer_sign * direc;
                    regularYaw(next_yaw);
                }
                double t[8] = {curr_x, curr_y, curr_yaw, next_x, next_y, next_yaw, direc, steer};
                tree.insert(tree.end(), t, t + 8);
                ++n_collision;
                if (!pl->collision4(next_x, next_y, next_yaw))
                    next_points.push_back({next_x, next_y, next_yaw, steer, direc});
            }
        for (const Pt &pt : next_points)
        {
            int nn = makeNode(pt.x, pt.y, pt.yaw, pt.steer, pt.direc, id);
            setNodeCost(nn);
            setHeuristicValue(nodes[nn], heuristic(nodes[nn]));
            num_nodes++;
            long long index = calIndex(nodes[nn].i, nodes[nn].j, yawToIdx(nodes[nn].yaw), nodes[nn].dir);
            auto it = cost_set.find(index);
            double old_cost = it == cost_set.end() ? std::numeric_limits<double>::infinity() : it->second.first;
            if (old_cost > nodes[nn].g)
            {
                if (it != cost_set.end() && it->second.second >= 0)
                    nodes[it->second.second].closed = true;
                cost_set[index] = std::make_pair(nodes[nn].g, nn);
                open.insert(std::make_pair(nodes[nn].f, nn));
            }
        }
    }

    bool genRSPath(int id) // :853-900
    {
        const PlannerParams &p = pl->p;
        const Node &n = nodes[id];
        double goal_dist_sq = std::pow(n.x - goal_xyz[0], 2) + std::pow(n.y - goal_xyz[1], 2);
        if (goal_dist_sq < p.rs_range * p.rs_range)
        {
            ++n_rs_shots;
            double start5[5] = {n.x, n.y, n.yaw, (double)(int)n.dir, n.steer};
            RsPath best;
            double rs_cost = rs_find_optimal(rs, start5, goal_xyz, best);
            std::vector<double> traj(5 * 512);
            int np = rs_gen_traj(rs, start5, best, traj.data(), 512);
            traj.resize((size_t)5 * std::min(np, 512));
            bool hit = false;
            for (int k = 0; k < np && !hit; ++k)
            {
                ++n_collision;
                hit = pl->collision4(traj[5 * k], traj[5 * k + 1], traj[5 * k + 2]);
            }
            if (!hit)
            {
                Node &g = nodes[goal];
                if (g.parent < 0 || g.g > n.g + rs_cost)
                {
                    g.parent = id;
                    setCost(g, n.g + rs_cost);
                    goal_path = traj;
                }
                return true;
            }
        }
        return false;
    }

    // Plan(), :682-777. Returns 1 goal found, 0 iteration limit, -1 open set empty, -2 not initialised
    int plan(int &iterations)
    {
        iterations = 0;
        if (!initialized)
            return -2;
        bool goal_found = false, limit_reached = false;
        int iteration = 0;
        open.insert(std::make_pair(nodes[start].f, start));
        while (!open.empty())
        {
            if (iteration > pl->p.max_iter)
            {
                limit_reached = true;
                break;
            }
            int cur = open.begin()->second;
            open.erase(open.begin());
            if (nodes[cur].closed)
                continue;
            if (genRSPath(cur))
            {
                goal_found = true;
                if (!pl->p.optimal)
                    break;
            }
            else
                expandNode(cur);
            if (pl->p.optimal)
            {
                if (std::hypot(goal_xyz[0] - nodes[cur].x, goal_xyz[1] - nodes[cur].y) <= 1.0)
                    break;
            }
            ++iteration;
        }
        iterations = iteration;
        if (goal_found)
            return 1;
        return limit_reached ? 0 : -1;
    }

    // raw path_ of setPath (:779-811): start ... last expanded node, then the RS goal path
    std::vector<double> rawPath()
    {
        std::vector<double> rev;
        int cur = nodes[goal].parent;
        while (cur >= 0 && nodes[cur].parent >= 0)
        {
            const Node &n = nodes[cur];
            double r[5] = {n.x, n.y, n.yaw, n.steer, n.dir};
            rev.insert(rev.end(), r, r + 5);
            cur = n.parent;
        }
        size_t npts = rev.size() / 5;
        double start_dir;
        if (npts > 1)
            start_dir = rev[rev.size() - 1]; // direction of the last pushed point (:795-799)
        else
            start_dir = goal_path.empty() ? 1.0 : goal_path[4]; // goal_path_.front().back()
        const Node &s = nodes[start];
        double r[5] = {s.x, s.y, s.yaw, 0.0, start_dir};
        rev.insert(rev.end(), r, r + 5);
        std::vector<double> out;
        for (size_t k = rev.size() / 5; k-- > 0;)
            out.insert(out.end(), rev.begin() + 5 * k, rev.begin() + 5 * k + 5);
        out.insert(out.end(), goal_path.begin(), goal_path.end());
        return out;
    }
};

static RsConfig rs_config_from_planner(const Planner &pl)
{
    // hybrid_a_star.cpp:142-143,169-170: setPenalty(GEAR, BACKWARD, STEER_ANGLE, STEER_CHANGE) lands positionally in
    // RSCurve's (gear, backward, steer_change, steer_angle) slots
    RsConfig c;
    c.radius = pl.min_r;
    c.steer_angle = pl.p.max_steer;
    c.step = pl.p.step;
    c.pen_gear = pl.p.pen_gear;
    c.pen_back = pl.p.pen_back;
    c.pen_steer_change = pl.p.pen_steer;       // planner's STEER_ANGLE penalty ends up as RS steer-change weight
    c.pen_steer_angle = pl.p.pen_steer_change; // and vice versa
    return c;
}
} // namespace port

// ================================================================================================================
using namespace port;

extern "C"
{
    void *port_map_create(const double *mp6)
    {
        Map *m = new Map();
        m->p.occupied_thresh = mp6[0];
        m->p.free_thresh = mp6[1];
        m->p.voro_alpha = (float)mp6[2];
        m->p.voro_dmax = (float)mp6[3];
        m->p.ds = (int)mp6[4];
        m->p.inflation_radius = mp6[5];
        return m;
    }
    void port_map_destroy(void *h) { delete static_cast<Map *>(h); }
    void port_map_set_occupancy(void *h, const int8_t *data, int W, int H, double res, double ox, double oy, double oyaw)
    {
        setOccupancy(*static_cast<Map *>(h), data, W, H, res, ox, oy, oyaw);
    }
    void port_map_update_goal(void *h, double gx, double gy)
    {
        Map *m = static_cast<Map *>(h);
        m->goal_x = gx;
        m->goal_y = gy;
        m->get_goal = true;
        generateGoalMap(*m);
    }
    void port_map_info(void *h, int *oi, double *od)
    {
        Map *m = static_cast<Map *>(h);
        oi[0] = m->H, oi[1] = m->W, oi[2] = m->H0, oi[3] = m->W0, oi[4] = m->get_map, oi[5] = m->get_goal, oi[6] = m->num_regions;
        od[0] = m->res, od[1] = m->res0, od[2] = m->ox, od[3] = m->oy, od[4] = m->oyaw, od[5] = m->osin, od[6] = m->ocos;
    }
    long port_map_layer(void *h, int layer, void *dst)
    {
        Map *m = static_cast<Map *>(h);
#define COPY(vec)                                                                \
    do                                                                           \
    {                                                                            \
        if (dst)                                                                 \
            std::memcpy(dst, (vec).data(), (vec).size() * sizeof((vec)[0]));     \
        return (long)(vec).size();                                               \
    } while (0)
        switch (layer)
        {
        case 0:
            COPY(m->static_orig);
        case 1:
            COPY(m->static_ds);
        case 2:
            COPY(m->inflate);
        case 3:
            COPY(m->goal);
        case 4:
            COPY(m->obs_dist);
        case 5:
            COPY(m->region);
        case 6:
            COPY(m->isedge);
        case 7:
            COPY(m->edge_dist);
        case 8:
            COPY(m->voronoi);
        }
#undef COPY
        return -1;
    }
    // replace the Voronoi cost layer (lets tests inject the reference's order-dependent field)
    void port_map_set_voronoi(void *h, const float *v)
    {
        Map *m = static_cast<Map *>(h);
        m->voronoi.assign(v, v + (size_t)m->H * m->W);
    }

    // pp: 17 doubles in PlannerParams order
    void *port_planner_create(void *map, const double *pp)
    {
        PlannerParams p;
        p.max_curvature = pp[0], p.width = pp[1], p.wheelbase = pp[2], p.length = pp[3], p.steer_offset = pp[4], p.max_steer = pp[5];
        p.pen_gear = pp[6], p.pen_back = pp[7], p.pen_steer = pp[8], p.pen_steer_change = pp[9], p.rs_range = pp[10];
        p.num_yaw = (int)pp[11], p.step = pp[12], p.granularity = pp[13], p.num_steer = (int)pp[14], p.max_iter = (int)pp[15];
        p.optimal = (int)pp[16];
        Planner *pl = new Planner();
        pl->config(static_cast<Map *>(map), p);
        return pl;
    }
    void port_planner_destroy(void *h) { delete static_cast<Planner *>(h); }
    void port_planner_collision4(void *h, const double *xyyaw, long n, uint8_t *out)
    {
        Planner *pl = static_cast<Planner *>(h);
        for (long k = 0; k < n; ++k)
            out[k] = pl->collision4(xyyaw[3 * k], xyyaw[3 * k + 1], xyyaw[3 * k + 2]);
    }
    // cfg7: {radius, max_steer, step, gear, back, a3, a4} with a3/a4 in RSCurve::setPenalty's positional slots
    int port_rs_solve(const double *cfg7, const double *start5, const double *goal3, double *seg_out, int *nseg, double *cost,
                      double *traj_out, int traj_cap)
    {
        RsConfig c;
        c.radius = cfg7[0], c.steer_angle = cfg7[1], c.step = cfg7[2], c.pen_gear = cfg7[3], c.pen_back = cfg7[4];
        c.pen_steer_change = cfg7[5], c.pen_steer_angle = cfg7[6];
        RsPath best;
        *cost = rs_find_optimal(c, start5, goal3, best);
        *nseg = best.n;
        for (int k = 0; k < best.n; ++k)
            seg_out[3 * k] = best.s[k].len, seg_out[3 * k + 1] = best.s[k].steer, seg_out[3 * k + 2] = best.s[k].dir;
        return rs_gen_traj(c, start5, best, traj_out, traj_cap);
    }
    double port_rs_distance(double rho, const double *a, const double *b)
    {
        return ompl_rs::distance(rho, a[0], a[1], a[2], b[0], b[1], b[2]);
    }
    // one expansion on a fresh search: parent6 {x,y,yaw,steer,dir,g}. prim_out P x 7 {x,y,yaw,steer,dir,collision,0};
    // child_out <= P x 10 {x,y,yaw,steer,dir,g,h,f,i,j} of inserted children. Returns inserted count, -1 if not initialised.
    int port_expand(void *planner, const double *start3, const double *goal3, const double *parent6, double *prim_out,
                    double *child_out)
    {
        Planner *pl = static_cast<Planner *>(planner);
        Search s;
        s.pl = pl;
        s.map = pl->map;
        s.rs = rs_config_from_planner(*pl);
        if (!s.initialize(start3, goal3))
            return -1;
        int grand = s.makeNode(parent6[0], parent6[1], parent6[2], 0.0, 1.0, -1);
        int node = s.makeNode(parent6[0], parent6[1], parent6[2], parent6[3], parent6[4], grand);
        s.setCost(s.nodes[node], parent6[5]);
        size_t before = s.nodes.size();
        s.expandNode(node);
        int np = (int)(s.tree.size() / 8);
        for (int k = 0; k < np; ++k)
        {
            const double *t = &s.tree[8 * k];
            double *o = prim_out + 7 * k;
            o[0] = t[3], o[1] = t[4], o[2] = t[5], o[3] = t[7], o[4] = t[6];
            o[5] = pl->collision4(t[3], t[4], t[5]);
            o[6] = 0;
        }
        int nc = 0;
        std::vector<char> inserted(s.nodes.size(), 0);
        for (auto &kv : s.open)
            inserted[kv.second] = 1;
        for (size_t k = before; k < s.nodes.size(); ++k)
            if (inserted[k])
            {
                const Node &n = s.nodes[k];
                double *o = child_out + 10 * nc++;
                o[0] = n.x, o[1] = n.y, o[2] = n.yaw, o[3] = n.steer, o[4] = n.dir, o[5] = n.g, o[6] = n.h, o[7] = n.f, o[8] = n.i, o[9] = n.j;
            }
        return nc;
    }
    // whole query. stats (8): {status, iterations, goal_g, n_primitives, num_nodes, path_len, n_rs_shots, n_collision}
    int port_plan_query(void *planner, const double *start3, const double *goal3, double *stats, double *path_out, int path_cap)
    {
        Planner *pl = static_cast<Planner *>(planner);
        Search s;
        s.pl = pl;
        s.map = pl->map;
        s.rs = rs_config_from_planner(*pl);
        s.initialize(start3, goal3);
        int iters = 0;
        int status = s.plan(iters);
        std::vector<double> path;
        if (status == 1)
            path = s.rawPath();
        stats[0] = status;
        stats[1] = iters;
        stats[2] = s.goal >= 0 ? s.nodes[s.goal].g : -1.0;
        stats[3] = (double)(s.tree.size() / 8);
        stats[4] = s.num_nodes;
        stats[5] = (double)(path.size() / 5);
        stats[6] = (double)s.n_rs_shots;
        stats[7] = (double)s.n_collision;
        size_t n = std::min(path.size(), (size_t)path_cap * 5);
        std::memcpy(path_out, path.data(), n * sizeof(double));
        return status;
    }
}
