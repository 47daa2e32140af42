// Host-side mirror of the reference's public C++ interfaces for the hot path, implemented on top of the C ABI
// (include/mpros_gpu.h): same class and method names, argument meaning and error behaviour as
//   NavMap            mpros_navigation_map/include/mpros_navigation_map/nav_map.h:17-547
//   HybridAstar       mpros_algorithm/include/mpros_algorithm/hybrid_a_star.h:26-404
//   RS::RSCurve       mpros_rs_path/include/mpros_rs_path/rs_curve.h:17-244
// so the mpros_planner node can swap the library in. Every method that computes forwards to the
// CUDA library; there is no CPU implementation here.
//
// Two build modes:
//   * default (headless): ROS is absent in this build environment, so the ROS types the interfaces touch are mirrored by small
//     PODs (OccupancyGridLite for nav_msgs::OccupancyGrid, ParamServer for ros::NodeHandle::param, a PathPoint with the
//     reference's member names, std::array<double, 4> for Eigen::Vector4d, an empty Visualizer);
//   * -DMPROS_FACADE_ROS (inside the catkin workspace): the REAL types are used - ros::NodeHandle, nav_msgs::OccupancyGrid,
//     the reference's own mpros_utils/pathpoint.h PathPoint and mpros_utils/visualizer.h Visualizer, Eigen::Vector4d - and
//     host/compat/{mpros_navigation_map/nav_map.h, mpros_algorithm/hybrid_a_star.h, mpros_rs_path/rs_curve.h} pull the classes
//     into the global namespace, so main_planner.cpp compiles unchanged (tests/test_facade_compile.py compiles the node's
//     main() against it).
#pragma once
#include <cmath>
#include <cstdint>
#include <array>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mpros_gpu.h"

#ifdef MPROS_FACADE_ROS
#include <Eigen/Core>

#include "nav_msgs/OccupancyGrid.h"
#include "ros/ros.h"
#include "mpros_utils/pathpoint.h" // the reference's own boundary record
class NavMap;                      // mpros_utils/visualizer.h names NavMap in publishCostmap (the compat headers alias it)
#ifndef MPROS_FACADE_NO_VISUALIZER
#include "mpros_utils/visualizer.h"
#endif
#endif

namespace mpros_b200
{

#ifdef MPROS_FACADE_ROS
typedef ros::NodeHandle ParamServer;
typedef nav_msgs::OccupancyGrid OccupancyGrid;
typedef ::PathPoint PathPoint;
typedef Eigen::Vector4d TreeSegment; // {next_x, next_y, curr_x, curr_y}, hybrid_a_star.cpp:1114-1164
typedef ::Visualizer Visualizer;
inline uint64_t grid_stamp_nsec(const OccupancyGrid &og) { return og.header.stamp.toNSec(); }
#else
// nav_msgs::OccupancyGrid, the fields NavMap::setOccupancyMap reads (nav_map.cpp:42-131) and vecToMsg writes (:654-691)
struct OccupancyGridLite
{
    typedef std::shared_ptr<const OccupancyGridLite> ConstPtr;
    uint64_t stamp_nsec = 1; // header.stamp; 0 = "initial map", ignored (nav_map.cpp:48-52)
    std::string frame_id;    // header.frame_id ("map" in every message NavMap builds)
    uint32_t width = 0, height = 0;
    float resolution = 0; // float32 in the message
    double origin_x = 0, origin_y = 0;
    double origin_qx = 0, origin_qy = 0, origin_qz = 0, origin_qw = 1; // origin.orientation
    std::vector<int8_t> data;
    void setOriginYaw(double yaw) // what map_server does: setRPY(0, 0, yaw)
    {
        origin_qx = origin_qy = 0;
        origin_qz = std::sin(yaw * 0.5);
        origin_qw = std::cos(yaw * 0.5);
    }
};
typedef OccupancyGridLite OccupancyGrid;
inline uint64_t grid_stamp_nsec(const OccupancyGrid &og) { return og.stamp_nsec; }

// ros::NodeHandle::param<T>(key, var, default) over a flat key -> double store
class ParamServer
{
public:
    void set(const std::string &key, double v) { store_[key] = v; }
    bool ok() const { return ok_; }
    void shutdown() { ok_ = false; }
    template <typename T>
    bool param(const std::string &key, T &var, const T &dflt) const
    {
        auto it = store_.find(key);
        if (it == store_.end())
        {
            var = dflt;
            return false;
        }
        var = static_cast<T>(it->second);
        return true;
    }

private:
    std::map<std::string, double> store_;
    bool ok_ = true;
};

// mpros_utils/pathpoint.h:7-35 (the boundary record of getNewPath / pathInCollision), same member names
struct PathPoint
{
    PathPoint() {}
    PathPoint(const double &x, const double &y, const double &yaw) : yaw_(yaw) { point_[0] = x, point_[1] = y; }
    double getX() const { return point_[0]; }
    double getY() const { return point_[1]; }
    double getYaw() const { return yaw_; }
    std::array<double, 2> point_ = {{0, 0}};
    double yaw_ = 0, curv_ = 0, vel_ = 0, accel_ = 0, jerk_ = 0;
    double direc_ = 1, steering_ = 0;
    bool fixed_ = false;
    bool is_cusp_ = false;
    double dist_ = 0; // distance from start to this point
    double time_ = 0;
};
typedef std::array<double, 4> TreeSegment; // Eigen::Vector4d {next_x, next_y, curr_x, curr_y}
struct Visualizer // mpros_utils/visualizer.h: only a reference is taken (hybrid_a_star.cpp:11-20, 66-68)
{
};
#endif

class NavMap
{
public:
    NavMap(int device = 0); // nav_map.cpp:10-12 (plus the CUDA device the layers live on)
    // nav_map.cpp:14-19: keeps the handle and, with the real ros::NodeHandle, subscribes getOccupancyMap to "map"
    explicit NavMap(ParamServer &nh, int device = 0);
    ~NavMap();
    NavMap(const NavMap &) = delete;
    NavMap &operator=(const NavMap &) = delete;

    void config(const ParamServer &nh);                            // nav_map.cpp:25-35
    void getOccupancyMap(const OccupancyGrid::ConstPtr &map_msg);  // nav_map.cpp:37-40 (the "map" subscriber callback)
    void setOccupancyMap(const OccupancyGrid &og);                 // nav_map.cpp:42-162
    void updateGoal(const double &x, const double &y);      // nav_map.h:204-210
    bool pointInCollision(const int &i, const int &j);      // nav_map.cpp:249-257
    bool pointInCollisionOrig(const int &i, const int &j);  // nav_map.cpp:259-267
    bool pointInCollisionINFLOrig(const int &i, const int &j); // nav_map.cpp:269-277
    // nav_map.h:86-199 (pure host arithmetic on the public members, identical expressions)
    inline int xToCol(const double &x) { return int((x - map_origin_x_) / map_resolution_); }
    inline int yToRow(const double &y) { return int((y - map_origin_y_) / map_resolution_); }
    inline int xToColOrig(const double &x) { return int((x - map_origin_x_) / map_resolution_orig_); }
    inline int yToRowOrig(const double &y) { return int((y - map_origin_y_) / map_resolution_orig_); }
    inline double colToX(const int &i) { return ((double)i + 0.5) * map_resolution_ + map_origin_x_; }
    inline double rowToY(const int &j) { return ((double)j + 0.5) * map_resolution_ + map_origin_y_; }
    inline double colToXOrig(const int &i) { return ((double)i + 0.5) * map_resolution_orig_ + map_origin_x_; }
    inline double rowToYOrig(const int &j) { return ((double)j + 0.5) * map_resolution_orig_ + map_origin_y_; }
    inline void posToIndex(const double &x, const double &y, int &i, int &j)
    {
        double dx = x - map_origin_x_, dy = y - map_origin_y_;
        j = static_cast<int>((dx * map_origin_cos_ + dy * map_origin_sin_) / map_resolution_);
        i = static_cast<int>((dx * -map_origin_sin_ + dy * map_origin_cos_) / map_resolution_);
    }
    inline void posToIndexOrig(const double &x, const double &y, int &i, int &j)
    {
        double dx = x - map_origin_x_, dy = y - map_origin_y_;
        j = static_cast<int>((dx * map_origin_cos_ + dy * map_origin_sin_) / map_resolution_orig_);
        i = static_cast<int>((dx * -map_origin_sin_ + dy * map_origin_cos_) / map_resolution_orig_);
    }
    inline void indexToPos(const int &i, const int &j, double &x, double &y)
    {
        double dx = ((double)j + 0.5) * map_resolution_, dy = ((double)i + 0.5) * map_resolution_;
        x = (dx * map_origin_cos_ - dy * map_origin_sin_) + map_origin_x_;
        y = (dx * map_origin_sin_ + dy * map_origin_cos_) + map_origin_y_;
    }
    inline void indexToPosOrig(const int &i, const int &j, double &x, double &y)
    {
        double dx = ((double)j + 0.5) * map_resolution_orig_, dy = ((double)i + 0.5) * map_resolution_orig_;
        x = (dx * map_origin_cos_ - dy * map_origin_sin_) + map_origin_x_;
        y = (dx * map_origin_sin_ + dy * map_origin_cos_) + map_origin_y_;
    }
    double getGoalCost(const double &x, const double &y); // nav_map.h:218-223
    double getGoalCost(const int &i, const int &j);       // nav_map.h:231-234
    double getVoronoiCost(const double &x, const double &y);
    double getVoronoiCost(const int &i, const int &j);
    std::vector<int> getGoalMap();           // nav_map.h:261-264
    std::vector<float> getVoronoiMap() const; // nav_map.h:270-273
    // the private layers, for the map messages of nav_map.cpp:693-716
    std::vector<int8_t> getStaticMap() const;
    std::vector<int8_t> getStaticOrigMap() const;
    std::vector<int8_t> getInflationOrigMap() const;
    // nav_map.cpp:693-716 -> vecToMsg (:654-691): the messages main_planner.cpp:216-220 publishes; data scaled on the device
    // (mpros_map_layer_msg), header.frame_id "map", geometry of the original or the down-sampled grid, origin orientation
    OccupancyGrid getStaticMapMsg();
    OccupancyGrid getGoalMapMsg();
    OccupancyGrid getVoroMapMsg();
    OccupancyGrid getStaticOrigMapMsg();
    OccupancyGrid getInflationOrigMapMsg();

    mpros_ctx *ctx() const { return ctx_; }

    // public data members of the reference (nav_map.h:469-487)
    bool get_map_ = false;
    bool get_goal_ = false;
    int map_height_ = 0, map_width_ = 0;
    int map_height_orig_ = 0, map_width_orig_ = 0;
    double map_resolution_ = 0, map_resolution_orig_ = 0;
    double map_origin_x_ = 0, map_origin_y_ = 0, map_origin_yaw_ = 0;
    double map_origin_sin_ = 0, map_origin_cos_ = 1;
    double x_max_ = 0, x_min_ = 0, y_max_ = 0, y_min_ = 0;

private:
    void refresh();
    template <typename T>
    std::vector<T> download(int layer, size_t n) const;
    OccupancyGrid layerMsg(int layer, bool is_original);
    mpros_ctx *ctx_ = nullptr;
    double origin_q_[4] = {0, 0, 0, 1}; // map_origin_q_ (x, y, z, w) as the last occupancy grid carried it
#ifdef MPROS_FACADE_ROS
    ros::NodeHandle nh_;
    ros::Subscriber map_sub;
#endif
    mpros_map_params params_;
    // host mirrors for the single-cell getters
    std::vector<int> goal_cache_;
    std::vector<float> voro_cache_;
    bool goal_cache_ok_ = false, voro_cache_ok_ = false;
};

namespace RS
{
// mpros_rs_path/include/mpros_rs_path/rs_curve_basic.h
enum Steer { LEFT = 1, STRAIGHT = 0, RIGHT = -1 };
enum Direc { FORWARD = 1, BACKWARD = -1 };
struct Point
{
    Point() : x(.0), y(.0), phi(.0) {}
    Point(double a, double b, double c) : x(a), y(b), phi(c) {}
    double x, y, phi;
};
struct PathSeg
{
    double value;
    Steer steer;
    Direc direc;
    double GetLength() const { return value; }
    Steer GetSteer() const { return steer; }
    Direc GetDirec() const { return direc; }
    // rs_curve_basic.h:94-96
    std::string GetManuever() const
    {
        return "Steer: " + std::to_string(GetSteer()) + "; Direction: " + std::to_string(GetDirec()) + "; Value: " + std::to_string(GetLength());
    }
};
typedef std::vector<PathSeg> Path;

class RSCurve
{
public:
    explicit RSCurve(mpros_ctx *ctx) : ctx_(ctx) {}
    void SetRadius(const double &r) { p_.radius = r; }
    void setSteeringAngle(const double &a) { p_.steering_angle = a; }
    void setStepSize(const double &s) { p_.step_size = s; }
    // declared order of rs_curve.h:97-100
    void setPenalty(double gear, double backward, double steer_change, double steer_angle)
    {
        p_.pen_gear = gear, p_.pen_backward = backward, p_.pen_steer_change = steer_change, p_.pen_steer_angle = steer_angle;
    }
    void SetStart(const double &x, const double &y, const double &yaw, int direction, double steer)
    {
        start_[0] = x, start_[1] = y, start_[2] = yaw, start_[3] = direction, start_[4] = steer;
    }
    void SetEnd(const double &x, const double &y, const double &yaw) { end_[0] = x, end_[1] = y, end_[2] = yaw; }
    void SetPoints(const Point &s, const Point &e)
    {
        start_[0] = s.x, start_[1] = s.y, start_[2] = s.phi;
        end_[0] = e.x, end_[1] = e.y, end_[2] = e.phi;
    }
    void FindOptimalPath();                            // rs_curve.cpp:111-163
    double GetLength() const { return optimal_len_; } // rs_curve.cpp:275-278
    std::vector<std::vector<double>> GetTraj();        // rs_curve.cpp:280-284 {x, y, yaw, steer*steer_angle, direc}
    std::vector<PathPoint> getTrajectory();
    Path getOptimalPath() const { return optimal_path_; }
    std::string GetManuever(const Path &path) // rs_curve.cpp:286-294
    {
        std::string manuever;
        for (auto &seg : path)
            manuever += seg.GetManuever() + "\n";
        return manuever;
    }

private:
    mpros_ctx *ctx_;
    mpros_rs_params p_{1.0, 0.0, 0.2, 0.0, 0.0, 0.0, 0.0};
    double start_[5] = {0, 0, 0, 0, 0}, end_[3] = {0, 0, 0};
    double optimal_len_ = 0;
    Path optimal_path_;
    std::vector<std::vector<double>> traj_;
};
// RS::PathFunction::path1..12 (rs_curve.h:248-334, rs_curve.cpp:353-599) evaluated on the device (mpros_rs_word_batch).
// The reference's free functions take only the point; here the context travels in a thread-local set by bind().
namespace PathFunction
{
void bind(mpros_ctx *ctx);
Path word(int k, const Point &pt); // k = 1..12
inline Path path1(const Point &pt) { return word(1, pt); }
inline Path path2(const Point &pt) { return word(2, pt); }
inline Path path3(const Point &pt) { return word(3, pt); }
inline Path path4(const Point &pt) { return word(4, pt); }
inline Path path5(const Point &pt) { return word(5, pt); }
inline Path path6(const Point &pt) { return word(6, pt); }
inline Path path7(const Point &pt) { return word(7, pt); }
inline Path path8(const Point &pt) { return word(8, pt); }
inline Path path9(const Point &pt) { return word(9, pt); }
inline Path path10(const Point &pt) { return word(10, pt); }
inline Path path11(const Point &pt) { return word(11, pt); }
inline Path path12(const Point &pt) { return word(12, pt); }
} // namespace PathFunction
} // namespace RS

class HybridAstar
{
public:
    HybridAstar() = default;
    // compact constructor: config, then initialize (hybrid_a_star.cpp:11-20); the Visualizer is only kept, as in the reference
    HybridAstar(const std::vector<double> &start_pt, const std::vector<double> &goal_pt, NavMap &map, Visualizer &vis, const ParamServer &nh);
    HybridAstar(const std::vector<double> &start_pt, const std::vector<double> &goal_pt, NavMap &map, const ParamServer &nh);
    void config(const ParamServer &nh);                                                                  // :120-188
    void initialize(const std::vector<double> &start_pt, const std::vector<double> &goal_pt, NavMap &map); // :52-118
    void reset();                                                                                        // :32-50
    bool Plan();                                                                                         // :682-777
    std::vector<std::vector<double>> getPath() { return path_; } // {x, y, yaw, steer, direction}
    // hybrid_a_star.cpp:1114-1164: leaf segments {next_x, next_y, curr_x, curr_y} of every evaluated primitive (re-plans
    // the query on the device with the expansion trace switched on)
    std::vector<TreeSegment> getSearchTree();
    std::vector<PathPoint> getNewPath();                          // hybrid_a_star.cpp:848-851: setPath's up-sampled path (:814-838)
    bool pathInCollision(const std::vector<std::vector<double>> &path); // :404-416
    bool pathInCollision(const std::vector<PathPoint> &path);           // :418-430
    bool footPrintInCollision4(const double &x, const double &y, const double &yaw); // :617-680
    // hybrid_a_star.cpp:432-615: the three methods no caller of the reference uses (mpros_collision_check_poses_method)
    bool footPrintInCollision(const double &x, const double &y, const double &yaw);
    bool footPrintInCollision2(const double &x, const double &y, const double &yaw);
    bool footPrintInCollision3(const double &x, const double &y, const double &yaw);
    void setFootprint(); // :190-206: footprint_m_ / footprint_longi_ from the configured vehicle (part of mpros_planner_config)
    // :853-900 for a free-standing pose: Reeds-Shepp shot to the goal if it lies within analytic_solution_range; true = a
    // collision-free path exists (the planner's own shots run inside Plan() on the device)
    bool genRSPath(const double &x, const double &y, const double &yaw, int direction = 1, double steer = 0.0);
    int calIndex(const int &idx_i, const int &idx_j, const int &idx_yaw, const double &direc); // :281-290
    int yawToIdx(const double &yaw);                                                          // :292-302
    void regularYaw(double &yaw);                                                             // :266-279
    // what Plan() found
    double goalCost() const { return goal_cost_; }
    long long iterations() const { return stats_[0]; }
    long long primitivesEvaluated() const { return stats_[1]; }

    void setMap(NavMap &map) { map_ = &map; }

private:
    NavMap *map_ = nullptr;
    Visualizer *vis_ = nullptr;
    mpros_planner_params pp_;
    bool configured_ = false, initialized_ = false;
    std::vector<double> start_, goal_;
    std::vector<std::vector<double>> path_;
    double goal_cost_ = -1;
    long long stats_[6] = {0, 0, 0, 0, 0, 0};
    double yaw_resol_ = 0;
};

// thrown only where the reference throws (config with a dead node handle) or when the CUDA library reports a failure
struct MprosFailure : public std::runtime_error
{
    explicit MprosFailure(const std::string &m) : std::runtime_error(m) {}
};

} // namespace mpros_b200
