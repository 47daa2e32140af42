// Implementation of the host-side mirror classes: every computing method forwards to the C ABI (CUDA library).
#include "mpros_facade.h"

#include <cstring>

namespace mpros_b200
{

static void check(int rc, const char *what)
{
    if (rc != MPROS_OK)
        throw MprosFailure(std::string(what) + ": " + mpros_last_error_string());
}

// tf2::getYaw (tf2/impl/utils.h) — the value NavMap stores in map_origin_yaw_ (nav_map.cpp:64-65)
static double yaw_from_quaternion(double x, double y, double z, double w)
{
    double sqx = x * x, sqy = y * y, sqz = z * z, sqw = w * w;
    double sarg = -2 * (x * z - w * y) / (sqx + sqy + sqz + sqw);
    if (sarg <= -0.99999)
        return -2 * std::atan2(y, x);
    if (sarg >= 0.99999)
        return 2 * std::atan2(y, x);
    return std::atan2(2 * (x * y + w * z), sqw + sqx - sqy - sqz);
}

// ---- NavMap ---------------------------------------------------------------------------------------------------------
NavMap::NavMap(int device)
{
    check(mpros_ctx_create(device, &ctx_), "mpros_ctx_create");
    mpros_map_params_default(&params_);
}
NavMap::NavMap(ParamServer &nh, int device) : NavMap(device)
{
#ifdef MPROS_FACADE_ROS
    nh_ = nh;
    map_sub = nh.subscribe("map", 1, &NavMap::getOccupancyMap, this); // nav_map.cpp:18
#else
    (void)nh;
#endif
}
NavMap::~NavMap() { mpros_ctx_destroy(ctx_); }

void NavMap::getOccupancyMap(const OccupancyGrid::ConstPtr &map_msg) { setOccupancyMap(*map_msg); }

void NavMap::config(const ParamServer &nh)
{
    const std::string prefix = "/mpros/map/";
    nh.param<double>(prefix + "occupied_thresh", params_.occupied_thresh, 0.68);
    nh.param<double>(prefix + "free_thresh", params_.free_thresh, 0.2);
    nh.param<float>(prefix + "voro_fallout_rate", params_.voro_fallout_rate, 10.0f);
    nh.param<float>(prefix + "voro_max_region", params_.voro_max_region, 50.0f);
    nh.param<int>(prefix + "downsampling_factor", params_.downsampling_factor, 1);
    nh.param<double>(prefix + "obstacle_inflation_radius", params_.obstacle_inflation_radius, 0.0);
    check(mpros_map_config(ctx_, &params_), "mpros_map_config");
}

void NavMap::refresh()
{
    mpros_map_info i;
    check(mpros_map_get_info(ctx_, &i), "mpros_map_get_info");
    get_map_ = i.get_map;
    get_goal_ = i.get_goal;
    map_height_ = i.map_height, map_width_ = i.map_width;
    map_height_orig_ = i.map_height_orig, map_width_orig_ = i.map_width_orig;
    map_resolution_ = i.map_resolution, map_resolution_orig_ = i.map_resolution_orig;
    map_origin_x_ = i.map_origin_x, map_origin_y_ = i.map_origin_y, map_origin_yaw_ = i.map_origin_yaw;
    map_origin_sin_ = i.map_origin_sin, map_origin_cos_ = i.map_origin_cos;
    x_max_ = i.x_max, x_min_ = i.x_min, y_max_ = i.y_max, y_min_ = i.y_min;
}

void NavMap::setOccupancyMap(const OccupancyGrid &og)
{
    if (grid_stamp_nsec(og) == 0) // "is initial map" -> ignored (nav_map.cpp:48-52)
        return;
#ifdef MPROS_FACADE_ROS
    const double qx = og.info.origin.orientation.x, qy = og.info.origin.orientation.y, qz = og.info.origin.orientation.z,
                 qw = og.info.origin.orientation.w;
    const int width = (int)og.info.width, height = (int)og.info.height;
    const double res = (double)og.info.resolution, ox = og.info.origin.position.x, oy = og.info.origin.position.y;
#else
    const double qx = og.origin_qx, qy = og.origin_qy, qz = og.origin_qz, qw = og.origin_qw;
    const int width = (int)og.width, height = (int)og.height;
    const double res = (double)og.resolution, ox = og.origin_x, oy = og.origin_y;
#endif
    const double yaw = yaw_from_quaternion(qx, qy, qz, qw);
    check(mpros_map_set_occupancy(ctx_, og.data.data(), width, height, res, ox, oy, yaw), "mpros_map_set_occupancy");
    origin_q_[0] = qx, origin_q_[1] = qy, origin_q_[2] = qz, origin_q_[3] = qw;
    goal_cache_ok_ = voro_cache_ok_ = false;
    refresh();
}

// vecToMsg (nav_map.cpp:654-691): the scaled int8 data come from the device, the geometry from the public members
OccupancyGrid NavMap::layerMsg(int layer, bool is_original)
{
    OccupancyGrid msg;
    size_t n = 0;
    if (get_map_)
        check(mpros_map_layer_msg(ctx_, layer, nullptr, 0, &n), "mpros_map_layer_msg");
#ifdef MPROS_FACADE_ROS
    msg.header.frame_id = "map";
    if (!get_map_ || n == 0)
        return msg;
    msg.info.height = is_original ? map_height_orig_ : map_height_;
    msg.info.width = is_original ? map_width_orig_ : map_width_;
    msg.info.resolution = is_original ? map_resolution_orig_ : map_resolution_;
    msg.info.origin.position.x = map_origin_x_;
    msg.info.origin.position.y = map_origin_y_;
    msg.info.origin.orientation.x = origin_q_[0], msg.info.origin.orientation.y = origin_q_[1];
    msg.info.origin.orientation.z = origin_q_[2], msg.info.origin.orientation.w = origin_q_[3];
#else
    msg.frame_id = "map";
    if (!get_map_ || n == 0)
        return msg;
    msg.height = is_original ? map_height_orig_ : map_height_;
    msg.width = is_original ? map_width_orig_ : map_width_;
    msg.resolution = (float)(is_original ? map_resolution_orig_ : map_resolution_);
    msg.origin_x = map_origin_x_, msg.origin_y = map_origin_y_;
    msg.origin_qx = origin_q_[0], msg.origin_qy = origin_q_[1], msg.origin_qz = origin_q_[2], msg.origin_qw = origin_q_[3];
#endif
    msg.data.resize(n);
    check(mpros_map_layer_msg(ctx_, layer, msg.data.data(), n, &n), "mpros_map_layer_msg");
    return msg;
}
OccupancyGrid NavMap::getStaticMapMsg() { return layerMsg(MPROS_LAYER_STATIC, false); }
OccupancyGrid NavMap::getGoalMapMsg() { return layerMsg(MPROS_LAYER_GOAL, false); }
OccupancyGrid NavMap::getVoroMapMsg() { return layerMsg(MPROS_LAYER_VORONOI, false); }
OccupancyGrid NavMap::getStaticOrigMapMsg() { return layerMsg(MPROS_LAYER_STATIC_ORIG, true); }
OccupancyGrid NavMap::getInflationOrigMapMsg() { return layerMsg(MPROS_LAYER_INFLATE_ORIG, true); }

void NavMap::updateGoal(const double &x, const double &y)
{
    check(mpros_map_update_goal(ctx_, x, y), "mpros_map_update_goal");
    goal_cache_ok_ = false;
    refresh();
}

static bool point_lookup(mpros_ctx *ctx, int layer, int i, int j)
{
    int32_t ij[2] = {i, j};
    uint8_t out = 1;
    check(mpros_map_point_in_collision(ctx, layer, ij, 1, &out), "mpros_map_point_in_collision");
    return out != 0;
}
bool NavMap::pointInCollision(const int &i, const int &j) { return point_lookup(ctx_, MPROS_LAYER_STATIC, i, j); }
bool NavMap::pointInCollisionOrig(const int &i, const int &j) { return point_lookup(ctx_, MPROS_LAYER_STATIC_ORIG, i, j); }
bool NavMap::pointInCollisionINFLOrig(const int &i, const int &j) { return point_lookup(ctx_, MPROS_LAYER_INFLATE_ORIG, i, j); }

template <typename T>
std::vector<T> NavMap::download(int layer, size_t n) const
{
    std::vector<T> v(n);
    check(mpros_map_download(ctx_, layer, v.data(), n * sizeof(T)), "mpros_map_download");
    return v;
}
std::vector<int> NavMap::getGoalMap()
{
    if (!goal_cache_ok_)
    {
        goal_cache_ = download<int>(MPROS_LAYER_GOAL, (size_t)map_height_ * map_width_);
        goal_cache_ok_ = true;
    }
    return goal_cache_;
}
std::vector<float> NavMap::getVoronoiMap() const { return download<float>(MPROS_LAYER_VORONOI, (size_t)map_height_ * map_width_); }
std::vector<int8_t> NavMap::getStaticMap() const { return download<int8_t>(MPROS_LAYER_STATIC, (size_t)map_height_ * map_width_); }
std::vector<int8_t> NavMap::getStaticOrigMap() const
{
    return download<int8_t>(MPROS_LAYER_STATIC_ORIG, (size_t)map_height_orig_ * map_width_orig_);
}
std::vector<int8_t> NavMap::getInflationOrigMap() const
{
    return download<int8_t>(MPROS_LAYER_INFLATE_ORIG, (size_t)map_height_orig_ * map_width_orig_);
}
double NavMap::getGoalCost(const int &i, const int &j)
{
    getGoalMap();
    return goal_cache_[(size_t)i * map_width_ + j] * map_resolution_;
}
double NavMap::getGoalCost(const double &x, const double &y)
{
    int i, j;
    posToIndex(x, y, i, j);
    return getGoalCost(i, j);
}
double NavMap::getVoronoiCost(const int &i, const int &j)
{
    if (!voro_cache_ok_)
    {
        voro_cache_ = getVoronoiMap();
        voro_cache_ok_ = true;
    }
    return voro_cache_[(size_t)i * map_width_ + j];
}
double NavMap::getVoronoiCost(const double &x, const double &y)
{
    int i, j;
    posToIndex(x, y, i, j);
    return getVoronoiCost(i, j);
}

// ---- RS::RSCurve ----------------------------------------------------------------------------------------------------
namespace RS
{
void RSCurve::FindOptimalPath()
{
    double segs[15], cost = 0;
    int32_t nseg = 0, npts = 0;
    const int cap = 256;
    std::vector<double> traj((size_t)cap * 5);
    check(mpros_rs_solve_batch(ctx_, &p_, start_, end_, 1, segs, &nseg, &cost, traj.data(), cap, &npts), "mpros_rs_solve_batch");
    optimal_len_ = cost;
    optimal_path_.clear();
    for (int k = 0; k < nseg; ++k)
        optimal_path_.push_back(PathSeg{segs[3 * k], (Steer)(int)segs[3 * k + 1], (Direc)(int)segs[3 * k + 2]});
    traj_.clear();
    for (int k = 0; k < npts && k < cap; ++k)
        traj_.push_back(std::vector<double>(traj.begin() + 5 * k, traj.begin() + 5 * k + 5));
}
std::vector<std::vector<double>> RSCurve::GetTraj() { return traj_; }
std::vector<PathPoint> RSCurve::getTrajectory()
{
    std::vector<PathPoint> out;
    for (auto &t : traj_)
    {
        PathPoint p(t[0], t[1], t[2]);
        p.curv_ = p_.steering_angle != 0 ? (1 / p_.radius) * (t[3] / p_.steering_angle) : 0.0; // 1/min_radius * steer
        p.direc_ = t[4];
        out.push_back(p);
    }
    return out;
}
namespace PathFunction
{
static thread_local mpros_ctx *g_ctx = nullptr;
void bind(mpros_ctx *ctx) { g_ctx = ctx; }
Path word(int k, const Point &pt)
{
    if (!g_ctx)
        throw MprosFailure("RS::PathFunction: no context bound (RS::PathFunction::bind)");
    const double in[3] = {pt.x, pt.y, pt.phi};
    double segs[15];
    int32_t n = 0;
    check(mpros_rs_word_batch(g_ctx, k, in, 1, segs, &n), "mpros_rs_word_batch");
    Path out;
    for (int t = 0; t < n; ++t)
        out.push_back(PathSeg{segs[3 * t], (Steer)(int)segs[3 * t + 1], (Direc)(int)segs[3 * t + 2]});
    return out;
}
} // namespace PathFunction
} // namespace RS

// ---- HybridAstar ------------------------------------------------------------------------------------------------------
HybridAstar::HybridAstar(const std::vector<double> &start_pt, const std::vector<double> &goal_pt, NavMap &map, const ParamServer &nh)
    : map_(&map)
{
    config(nh);
    initialize(start_pt, goal_pt, map);
}
HybridAstar::HybridAstar(const std::vector<double> &start_pt, const std::vector<double> &goal_pt, NavMap &map, Visualizer &vis,
                         const ParamServer &nh)
    : map_(&map), vis_(&vis)
{
    config(nh);
    initialize(start_pt, goal_pt, map);
}

void HybridAstar::setFootprint()
{
    // footprint_m_, footprint_longi_ and num_checking_step are derived from the vehicle parameters and map_resolution_orig_
    // inside mpros_planner_config; calling it again re-derives them for the current map
    if (!map_)
        throw MprosFailure("HybridAstar::setFootprint: no map bound");
    check(mpros_planner_config(map_->ctx(), &pp_), "mpros_planner_config");
}

bool HybridAstar::genRSPath(const double &x, const double &y, const double &yaw, int direction, double steer)
{
    if (!configured_ || goal_.size() < 3)
        return false;
    const double dx = x - goal_[0], dy = y - goal_[1];
    if (!(dx * dx + dy * dy < pp_.analytic_solution_range * pp_.analytic_solution_range)) // hybrid_a_star.cpp:869
        return false;
    const double s5[5] = {x, y, yaw, (double)direction, steer};
    double segs[15], cost = 0;
    int32_t nseg = 0, npts = 0;
    uint8_t hit = 1;
    check(mpros_rs_shot_batch(map_->ctx(), s5, goal_.data(), 1, segs, &nseg, &cost, nullptr, 0, &npts, &hit), "mpros_rs_shot_batch");
    return hit == 0;
}

void HybridAstar::config(const ParamServer &nh)
{
    if (!nh.ok())
        throw std::runtime_error("[Smoother]: Node died while configuring"); // hybrid_a_star.cpp:122-125
    mpros_planner_params_default(&pp_);
    const std::string v = "/mpros/vehicle/", p = "/mpros/planner/";
    nh.param<double>(v + "max_curvature", pp_.max_curvature, 0.4);
    nh.param<double>(v + "width", pp_.width, 2.0);
    nh.param<double>(v + "wheelbase", pp_.wheelbase, 3.0);
    nh.param<double>(v + "length", pp_.length, 4.2);
    nh.param<double>(v + "center_to_rotate_center_longitudinal", pp_.steer_offset, 0.0);
    nh.param<double>(v + "max_steering_angle", pp_.max_steering_angle, 0.35);
    nh.param<double>(p + "gear_penalty", pp_.gear_penalty, 10.0);
    nh.param<double>(p + "reverse_penalty", pp_.reverse_penalty, 0);
    nh.param<double>(p + "steering_penalty", pp_.steering_penalty, 0.1);
    nh.param<double>(p + "steering_change_penalty", pp_.steering_change_penalty, 10.0);
    nh.param<double>(p + "analytic_solution_range", pp_.analytic_solution_range, 10.0);
    nh.param<int>(p + "yaw_discretization", pp_.yaw_discretization, 36);
    nh.param<double>(p + "step_size", pp_.step_size, 1.1);
    nh.param<double>(p + "path_point_distance", pp_.path_point_distance, 0.2);
    nh.param<int>(p + "steering_discretization", pp_.steering_discretization, 3);
    nh.param<int>(p + "max_iteration", pp_.max_iteration, 10000);
    bool optimal = false, inter = true; // bool parameters in the reference (hybrid_a_star.cpp:158-161)
    nh.param<bool>(p + "optimal_path", optimal, false);
    nh.param<bool>(p + "interpolation_enable", inter, true);
    pp_.optimal_path = optimal ? 1 : 0;
    pp_.interpolation_enable = inter ? 1 : 0;
    yaw_resol_ = 2 * M_PI / pp_.yaw_discretization;
    if (!map_)
        throw MprosFailure("HybridAstar::config: no map bound (setFootprint reads map_->map_resolution_orig_)");
    check(mpros_planner_config(map_->ctx(), &pp_), "mpros_planner_config");
    configured_ = true;
}

void HybridAstar::initialize(const std::vector<double> &start_pt, const std::vector<double> &goal_pt, NavMap &map)
{
    if (!configured_)
        return; // "Planner not properly configured, stop initializing!"
    map_ = &map;
    start_ = start_pt;
    goal_ = goal_pt;
    // start / goal collision and goal-map availability are re-checked on the device by Plan(); mirror the early
    // returns here so that initialized_ has the reference's meaning (hybrid_a_star.cpp:71-100)
    if (footPrintInCollision4(start_pt[0], start_pt[1], start_pt[2]))
        return;
    if (footPrintInCollision4(goal_pt[0], goal_pt[1], goal_pt[2]))
        return;
    if (!map_->get_goal_)
        return;
    initialized_ = true;
}

void HybridAstar::reset()
{
    path_.clear();
    goal_cost_ = -1;
    std::memset(stats_, 0, sizeof(stats_));
}

bool HybridAstar::Plan()
{
    if (!initialized_)
        return false; // "Planner not properly initialized!"
    int cap = 4096;
    std::vector<double> paths((size_t)cap * 5);
    int32_t status = 0, len = 0;
    int64_t st[6] = {0, 0, 0, 0, 0, 0};
    double cost = -1;
    check(mpros_plan_batch(map_->ctx(), start_.data(), goal_.data(), 1, &status, &cost, &len, paths.data(), cap, st), "mpros_plan_batch");
    if (status == MPROS_PLAN_FOUND && len > cap)
    {
        // path_ never comes back truncated: the search is deterministic, run it again with room for the whole path
        cap = len;
        paths.assign((size_t)cap * 5, 0.0);
        check(mpros_plan_batch(map_->ctx(), start_.data(), goal_.data(), 1, &status, &cost, &len, paths.data(), cap, st), "mpros_plan_batch");
        if (len > cap)
            throw std::runtime_error("mpros_b200::HybridAstar::Plan: path longer than the buffer after the re-run");
    }
    for (int k = 0; k < 6; ++k)
        stats_[k] = st[k];
    goal_cost_ = cost;
    path_.clear();
    if (status != MPROS_PLAN_FOUND)
        return false;
    for (int k = 0; k < len && k < cap; ++k)
        path_.push_back(std::vector<double>(paths.begin() + 5 * k, paths.begin() + 5 * k + 5));
    return true;
}

std::vector<TreeSegment> HybridAstar::getSearchTree()
{
    std::vector<TreeSegment> tree;
    if (!initialized_)
        return tree;
    int32_t status = 0;
    double cost = -1;
    size_t n = 0;
    check(mpros_plan_search_tree(map_->ctx(), start_.data(), goal_.data(), &status, &cost, nullptr, 0, &n), "mpros_plan_search_tree");
    static_assert(sizeof(TreeSegment) == 4 * sizeof(double), "a search-tree segment is four packed doubles");
    tree.resize(n);
    if (n)
        check(mpros_plan_search_tree(map_->ctx(), start_.data(), goal_.data(), &status, &cost, reinterpret_cast<double *>(tree.data()), n, &n),
              "mpros_plan_search_tree");
    return tree;
}

std::vector<PathPoint> HybridAstar::getNewPath()
{
    // new_path_ as setPath builds it (hybrid_a_star.cpp:814-838): path length, cusps, curvature and, with
    // interpolation_enable, the collision-checked Bezier up-sampling of upsampleAndPopulatePath
    std::vector<PathPoint> out;
    if (path_.empty())
        return out;
    std::vector<double> raw;
    raw.reserve(path_.size() * 5);
    for (auto &p : path_)
        raw.insert(raw.end(), p.begin(), p.begin() + 5);
    int64_t offs[2] = {0, (int64_t)path_.size()}, out_offs[2] = {0, 0};
    std::vector<double> buf((size_t)8 * 64 * path_.size());
    int rc = mpros_path_postprocess_batch(map_->ctx(), raw.data(), offs, 1, buf.data(), buf.size() / 8, out_offs);
    if (rc == MPROS_ERR_CAPACITY)
    {
        buf.resize((size_t)8 * out_offs[1]);
        rc = mpros_path_postprocess_batch(map_->ctx(), raw.data(), offs, 1, buf.data(), buf.size() / 8, out_offs);
    }
    check(rc, "mpros_path_postprocess_batch");
    for (int64_t k = 0; k < out_offs[1]; ++k)
    {
        const double *o = buf.data() + 8 * k;
        PathPoint pt(o[0], o[1], o[2]);
        pt.curv_ = o[3], pt.direc_ = o[4], pt.steering_ = o[5], pt.is_cusp_ = o[6] != 0.0, pt.dist_ = o[7];
        out.push_back(pt);
    }
    return out;
}

bool HybridAstar::footPrintInCollision4(const double &x, const double &y, const double &yaw)
{
    double pose[3] = {x, y, yaw};
    uint8_t v = 1;
    check(mpros_collision_check_poses(map_->ctx(), pose, 1, &v), "mpros_collision_check_poses");
    return v != 0;
}

static bool footprint_method(mpros_ctx *ctx, int method, double x, double y, double yaw)
{
    double pose[3] = {x, y, yaw};
    uint8_t v = 1;
    if (mpros_collision_check_poses_method(ctx, method, pose, 1, &v) != MPROS_OK)
        throw MprosFailure(std::string("mpros_collision_check_poses_method: ") + mpros_last_error_string());
    return v != 0;
}
bool HybridAstar::footPrintInCollision(const double &x, const double &y, const double &yaw) { return footprint_method(map_->ctx(), 1, x, y, yaw); }
bool HybridAstar::footPrintInCollision2(const double &x, const double &y, const double &yaw) { return footprint_method(map_->ctx(), 2, x, y, yaw); }
bool HybridAstar::footPrintInCollision3(const double &x, const double &y, const double &yaw) { return footprint_method(map_->ctx(), 3, x, y, yaw); }

bool HybridAstar::pathInCollision(const std::vector<std::vector<double>> &path)
{
    if (path.empty())
        return false;
    std::vector<double> poses;
    poses.reserve(path.size() * 3);
    for (auto &p : path)
        poses.insert(poses.end(), p.begin(), p.begin() + 3);
    int64_t off[2] = {0, (int64_t)path.size()};
    uint8_t v = 1;
    check(mpros_collision_check_paths(map_->ctx(), poses.data(), off, 1, &v), "mpros_collision_check_paths");
    return v != 0;
}

bool HybridAstar::pathInCollision(const std::vector<PathPoint> &path)
{
    std::vector<std::vector<double>> p;
    for (auto &pt : path)
        p.push_back({pt.getX(), pt.getY(), pt.getYaw()});
    return pathInCollision(p);
}

int HybridAstar::calIndex(const int &idx_i, const int &idx_j, const int &idx_yaw, const double &direc)
{
    int idx_direc = direc > 0 ? 0 : 1;
    const int num_direc = 2;
    return idx_i * map_->map_width_ * pp_.yaw_discretization * num_direc + idx_j * pp_.yaw_discretization * num_direc +
           idx_yaw * num_direc + idx_direc;
}
int HybridAstar::yawToIdx(const double &yaw)
{
    double yaw_pos = yaw;
    if (yaw < 0)
        yaw_pos = yaw + 2 * M_PI;
    return int(yaw_pos / yaw_resol_);
}
void HybridAstar::regularYaw(double &yaw)
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

} // namespace mpros_b200
