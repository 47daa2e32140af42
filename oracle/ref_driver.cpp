// TEST INFRASTRUCTURE ONLY — CPU oracle. Never included, linked or executed by the product path.
//
// Headless C-ABI driver around the reference's OWN, UNMODIFIED sources, compiled where they lie under
// /root/reference (never copied into this repo) against the stub headers in oracle/stubs/.
// Build: `make -C oracle ref` -> oracle/_ref/libmpros_ref.so. The three reference translation units are
// pulled in as one unity TU because planner_utils.h defines non-inline functions in a header.
//
// What it wraps (reference file:line):
//   NavMap::config / setOccupancyMap / updateGoal          nav_map.cpp:25-35, 42-162, nav_map.h:204-210
//   NavMap::pointInCollision{,Orig,INFLOrig}               nav_map.cpp:249-277
//   HybridAstar ctor / config / initialize / Plan          hybrid_a_star.cpp:11-20, 52-188, 682-777
//   HybridAstar::footPrintInCollision4 / expandNode        hybrid_a_star.cpp:617-680, 304-402
//   RS::RSCurve::FindOptimalPath / GetTraj / GetLength     rs_curve.cpp:111-163, 280-284, 275-278
//   RSStateSpace::getCost (OMPL stub, PARITY UNPINNED)     rs_statespace.h:50-54
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "mpros_navigation_map/src/nav_map.cpp"
#include "mpros_rs_path/src/rs_curve.cpp"
#include "mpros_algorithm/src/hybrid_a_star.cpp"

namespace
{
    ros::NodeHandle g_nh;
    Visualizer g_vis;
    double now_s()
    {
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
}

extern "C"
{
    // ---- parameter server -------------------------------------------------------------------------
    void ref_param_set(const char *key, double v) { ros::param_store()[key] = v; }
    void ref_param_clear() { ros::param_store().clear(); }

    // ---- NavMap -----------------------------------------------------------------------------------
    void *ref_map_create()
    {
        NavMap *m = new NavMap();
        m->config(g_nh);
        return m;
    }
    void ref_map_destroy(void *h) { delete static_cast<NavMap *>(h); }

    // data: row-major int8 occupancy (-1/0/100), row 0 = bottom (ROS convention). res is float32 like
    // nav_msgs/MapMetaData. Orientation is built with setRPY(0,0,yaw) as map_server does.
    // Returns the wall time of setOccupancyMap in seconds.
    double ref_map_set_occupancy(void *h, const int8_t *data, int width, int height, float res,
                                 double ox, double oy, double oyaw)
    {
        NavMap *m = static_cast<NavMap *>(h);
        nav_msgs::OccupancyGrid og;
        og.header.stamp = ros::Time(1, 0);
        og.info.width = width;
        og.info.height = height;
        og.info.resolution = res;
        og.info.origin.position.x = ox;
        og.info.origin.position.y = oy;
        tf2::Quaternion q;
        q.setRPY(0, 0, oyaw);
        og.info.origin.orientation = tf2::toMsg(q);
        og.data.assign(data, data + (size_t)width * height);
        double t0 = now_s();
        m->setOccupancyMap(og);
        return now_s() - t0;
    }
    double ref_map_update_goal(void *h, double gx, double gy)
    {
        double t0 = now_s();
        static_cast<NavMap *>(h)->updateGoal(gx, gy);
        return now_s() - t0;
    }
    // out_i: {H, W, H_orig, W_orig, get_map, get_goal}; out_d: {res, res_orig, ox, oy, oyaw, sin, cos}
    void ref_map_info(void *h, int *out_i, double *out_d)
    {
        NavMap *m = static_cast<NavMap *>(h);
        out_i[0] = m->map_height_;
        out_i[1] = m->map_width_;
        out_i[2] = m->map_height_orig_;
        out_i[3] = m->map_width_orig_;
        out_i[4] = m->get_map_;
        out_i[5] = m->get_goal_;
        out_d[0] = m->map_resolution_;
        out_d[1] = m->map_resolution_orig_;
        out_d[2] = m->map_origin_x_;
        out_d[3] = m->map_origin_y_;
        out_d[4] = m->map_origin_yaw_;
        out_d[5] = m->map_origin_sin_;
        out_d[6] = m->map_origin_cos_;
    }
    // batch of cells through NavMap::pointInCollision (mpros_layer 1) / pointInCollisionOrig (0) / pointInCollisionINFLOrig (2), nav_map.cpp:249-277
    void ref_map_points_in_collision(void *h, int layer, const int *ij, long n, uint8_t *out)
    {
        NavMap *m = static_cast<NavMap *>(h);
        for (long k = 0; k < n; ++k)
        {
            const int i = ij[2 * k], j = ij[2 * k + 1];
            out[k] = layer == 1 ? m->pointInCollision(i, j) : layer == 0 ? m->pointInCollisionOrig(i, j) : m->pointInCollisionINFLOrig(i, j);
        }
    }
    // layer ids: 0 static_map_orig_ (i8, orig) 1 static_map_ (i8) 2 inflate_map_orig_ (i8, orig)
    // 3 goal_cost_map_ (i32) 4 obs_dist_ (i32) 5 region_ (i32) 6 isedge_ (u8) 7 edge_dist_ (i32)
    // 8 voronoi_value_map_ (f32). Returns number of elements written (or that would be, if dst==NULL).
    long ref_map_layer(void *h, int layer, void *dst)
    {
        NavMap *m = static_cast<NavMap *>(h);
        switch (layer)
        {
        case 0:
            if (dst)
                std::memcpy(dst, m->static_map_orig_.data(), m->static_map_orig_.size());
            return m->static_map_orig_.size();
        case 1:
            if (dst)
                std::memcpy(dst, m->static_map_.data(), m->static_map_.size());
            return m->static_map_.size();
        case 2:
            if (dst)
                std::memcpy(dst, m->inflate_map_orig_.data(), m->inflate_map_orig_.size());
            return m->inflate_map_orig_.size();
        case 3:
            if (dst)
                std::memcpy(dst, m->goal_cost_map_.data(), m->goal_cost_map_.size() * sizeof(int));
            return m->goal_cost_map_.size();
        case 8:
            if (dst)
                std::memcpy(dst, m->voronoi_value_map_.data(), m->voronoi_value_map_.size() * sizeof(float));
            return m->voronoi_value_map_.size();
        default:
            break;
        }
        long n = m->voronoi_map_.size();
        if (!dst)
            return n;
        for (long k = 0; k < n; ++k)
        {
            const VoroNode &v = m->voronoi_map_[k];
            switch (layer)
            {
            case 4:
                static_cast<int *>(dst)[k] = v.obs_dist_;
                break;
            case 5:
                static_cast<int *>(dst)[k] = v.region_;
                break;
            case 6:
                static_cast<uint8_t *>(dst)[k] = v.isedge_;
                break;
            case 7:
                static_cast<int *>(dst)[k] = v.edge_dist_;
                break;
            default:
                return -1;
            }
        }
        return n;
    }
    // The map-layer messages the node publishes (nav_map.cpp:654-716: getStaticMapMsg / getGoalMapMsg / getVoroMapMsg /
    // getStaticOrigMapMsg / getInflationOrigMapMsg -> vecToMsg): layer ids as ref_map_layer (1, 3, 8, 0, 2).
    // dst: int8[msg.data.size()]; meta3: {height, width, resolution}. Returns msg.data.size().
    long ref_map_layer_msg(void *h, int layer, int8_t *dst, double *meta3)
    {
        NavMap *m = static_cast<NavMap *>(h);
        nav_msgs::OccupancyGrid msg;
        switch (layer)
        {
        case 1:
            msg = m->getStaticMapMsg();
            break;
        case 3:
            msg = m->getGoalMapMsg();
            break;
        case 8:
            msg = m->getVoroMapMsg();
            break;
        case 0:
            msg = m->getStaticOrigMapMsg();
            break;
        case 2:
            msg = m->getInflationOrigMapMsg();
            break;
        default:
            return -1;
        }
        if (meta3)
            meta3[0] = msg.info.height, meta3[1] = msg.info.width, meta3[2] = msg.info.resolution;
        if (dst)
            std::memcpy(dst, msg.data.data(), msg.data.size());
        return (long)msg.data.size();
    }
    // NavMap point lookups (nav_map.cpp:249-277), layer: 0 = downsampled, 1 = orig, 2 = inflated orig
    int ref_map_point_in_collision(void *h, int layer, int i, int j)
    {
        NavMap *m = static_cast<NavMap *>(h);
        if (layer == 0)
            return m->pointInCollision(i, j);
        if (layer == 1)
            return m->pointInCollisionOrig(i, j);
        return m->pointInCollisionINFLOrig(i, j);
    }

    // ---- HybridAstar ------------------------------------------------------------------------------
    // A configured (not initialised) planner bound to `map`: enough for collision checks.
    void *ref_planner_create(void *map)
    {
        HybridAstar *p = new HybridAstar();
        p->map_ = static_cast<NavMap *>(map);
        p->config(g_nh);
        return p;
    }
    void ref_planner_destroy(void *p) { delete static_cast<HybridAstar *>(p); }

    // xyyaw: n x 3 doubles; out: n verdict bytes (1 = in collision). Returns wall seconds.
    double ref_planner_collision4(void *h, const double *xyyaw, long n, uint8_t *out)
    {
        HybridAstar *p = static_cast<HybridAstar *>(h);
        double t0 = now_s();
        for (long k = 0; k < n; ++k)
            out[k] = p->footPrintInCollision4(xyyaw[3 * k], xyyaw[3 * k + 1], xyyaw[3 * k + 2]);
        return now_s() - t0;
    }

    // footPrintInCollision (1), footPrintInCollision2 (2), footPrintInCollision3 (3): the unused methods of the interface
    void ref_planner_collision_method(void *h, int method, const double *xyyaw, long n, uint8_t *out)
    {
        HybridAstar *p = static_cast<HybridAstar *>(h);
        for (long k = 0; k < n; ++k)
        {
            const double x = xyyaw[3 * k], y = xyyaw[3 * k + 1], yaw = xyyaw[3 * k + 2];
            out[k] = method == 1 ? p->footPrintInCollision(x, y, yaw) : method == 2 ? p->footPrintInCollision2(x, y, yaw) : p->footPrintInCollision3(x, y, yaw);
        }
    }

    // initialize() for a query; returns initialized_ (0 when start/goal collide or no goal map)
    int ref_planner_initialize(void *h, void *map, const double *start3, const double *goal3)
    {
        HybridAstar *p = static_cast<HybridAstar *>(h);
        std::vector<double> s(start3, start3 + 3), g(goal3, goal3 + 3);
        p->initialize(s, g, *static_cast<NavMap *>(map), g_vis);
        return p->initialized_;
    }

    // Expand one parent {x,y,yaw,steer,dir,g} on an initialised planner whose closed set is fresh.
    // prim_out: P x 7 {x,y,yaw,steer,dir,collision,0} for every primitive in evaluation order (tree_).
    // child_out: up to P x 10 {x,y,yaw,steer,dir,g,h,f,idx_i,idx_j} for the children that were inserted
    // into the open list, in insertion order. Returns the number of inserted children.
    int ref_planner_expand(void *h, const double *parent6, double *prim_out, double *child_out)
    {
        HybridAstar *p = static_cast<HybridAstar *>(h);
        NodePtr grand = std::make_shared<HybridAstarNode>(parent6[0], parent6[1], parent6[2], 0.0, 1.0, nullptr, p->map_);
        NodePtr node = std::make_shared<HybridAstarNode>(parent6[0], parent6[1], parent6[2], parent6[3], parent6[4], grand, p->map_);
        node->setCost(parent6[5]);
        size_t tree0 = p->tree_.size();
        // remember what is already in the open list
        std::vector<NodePtr> before;
        for (auto &kv : p->open_map_)
            before.push_back(kv.second);
        p->expandNode(node);
        int np = 0;
        for (size_t k = tree0; k < p->tree_.size(); ++k, ++np)
        {
            auto &b = p->tree_[k];
            prim_out[7 * np + 0] = b(3);
            prim_out[7 * np + 1] = b(4);
            prim_out[7 * np + 2] = b(5);
            prim_out[7 * np + 3] = b(7);
            prim_out[7 * np + 4] = b(6);
            prim_out[7 * np + 5] = p->footPrintInCollision4(b(3), b(4), b(5));
            prim_out[7 * np + 6] = 0;
        }
        // collect the new children; order them as the primitives were generated
        std::vector<NodePtr> fresh;
        for (auto &kv : p->open_map_)
        {
            bool old = false;
            for (auto &b : before)
                if (b == kv.second)
                    old = true;
            if (!old && kv.second->parent_ == node)
                fresh.push_back(kv.second);
        }
        int nc = 0;
        for (int k = 0; k < np; ++k)
            for (auto &c : fresh)
                if (c->x_ == prim_out[7 * k] && c->y_ == prim_out[7 * k + 1] && c->yaw_ == prim_out[7 * k + 2] &&
                    c->steer_ == prim_out[7 * k + 3] && c->direction_ == prim_out[7 * k + 4])
                {
                    double *o = child_out + 10 * nc++;
                    o[0] = c->x_;
                    o[1] = c->y_;
                    o[2] = c->yaw_;
                    o[3] = c->steer_;
                    o[4] = c->direction_;
                    o[5] = c->gvalue_;
                    o[6] = c->hvalue_;
                    o[7] = c->fvalue_;
                    o[8] = c->idx_i_;
                    o[9] = c->idx_j_;
                    break;
                }
        return nc;
    }

    // setNodeHeuristic of a free-standing node (hybrid_a_star.cpp:248-264) on an initialised planner
    double ref_planner_heuristic(void *h, double x, double y, double yaw)
    {
        HybridAstar *p = static_cast<HybridAstar *>(h);
        NodePtr node = std::make_shared<HybridAstarNode>(x, y, yaw, 0.0, 1.0, nullptr, p->map_);
        p->setNodeHeuristic(node);
        return node->hvalue_;
    }

    // Whole query as main_planner.cpp:225-229 does it: updateGoal -> ctor (config + initialize) -> Plan().
    // stats_out (12 doubles): {plan_ok, initialized, goal_g, n_primitives (tree_), num_nodes_, path_len,
    //   goal_path_len, t_update_goal, t_ctor, t_plan, t_rs_shot, t_colli}
    // path_out: up to path_cap x 5 {x,y,yaw,steer,dir} of the raw path_ (getPath()).
    int ref_plan_query(void *map, const double *start3, const double *goal3, int do_update_goal,
                       double *stats_out, double *path_out, int path_cap)
    {
        NavMap *m = static_cast<NavMap *>(map);
        std::vector<double> s(start3, start3 + 3), g(goal3, goal3 + 3);
        double t0 = now_s();
        if (do_update_goal)
            m->updateGoal(g[0], g[1]);
        double t1 = now_s();
        HybridAstar planner(s, g, *m, g_vis, g_nh);
        double t2 = now_s();
        bool ok = planner.Plan();
        double t3 = now_s();
        std::vector<std::vector<double>> path = planner.getPath();
        stats_out[0] = ok;
        stats_out[1] = planner.initialized_;
        stats_out[2] = (planner.goal_node_ != nullptr) ? planner.goal_node_->gvalue_ : -1.0;
        stats_out[3] = (double)planner.tree_.size();
        stats_out[4] = planner.num_nodes_;
        stats_out[5] = (double)path.size();
        stats_out[6] = (double)planner.goal_path_.size();
        stats_out[7] = t1 - t0;
        stats_out[8] = t2 - t1;
        stats_out[9] = t3 - t2;
        stats_out[10] = planner.time_rs_shot_.count();
        stats_out[11] = planner.time_colli_.count();
        // why Plan() returned (hybrid_a_star.cpp:756-776), in enum mpros_plan_status codes: 1 goal found, 0 iteration limit
        // (the loop broke with entries left in open_map_), -1 open set empty, -2 planner not initialised
        stats_out[12] = !planner.initialized_ ? -2.0 : ok ? 1.0 : planner.open_map_.empty() ? -1.0 : 0.0;
        int n = (int)path.size() < path_cap ? (int)path.size() : path_cap;
        for (int k = 0; k < n; ++k)
            for (int c = 0; c < 5; ++c)
                path_out[5 * k + c] = path[k][c];
        return ok;
    }

    // setPath's PathPoint post-processing (hybrid_a_star.cpp:814-838) applied to a given raw path_ {x,y,yaw,steer,dir}:
    // updatePathLength / setCusp / updateCurvatureNew, then (interpolation_enable) upsampleAndPopulatePath +
    // updateCurvatureNew + updatePathLength — the reference's own functions, called in the reference's order on a planner
    // constructed for the path's end poses. newpath_out: cap x 8 {x, y, yaw, curv, direc, steering, is_cusp, dist} = what
    // getNewPath() returns. Returns new_path_.size().
    int ref_upsample_path(void *map, const double *path5, int n, double *newpath_out, int cap)
    {
        NavMap *m = static_cast<NavMap *>(map);
        std::vector<double> s(path5, path5 + 3), g(path5 + 5 * (n - 1), path5 + 5 * (n - 1) + 3);
        HybridAstar planner(s, g, *m, g_vis, g_nh);
        std::vector<PathPoint> &np_ = planner.new_path_;
        np_.clear();
        for (int i = 0; i < n; ++i)
        {
            PathPoint pp(path5[5 * i], path5[5 * i + 1], path5[5 * i + 2]);
            pp.steering_ = path5[5 * i + 3];
            pp.direc_ = path5[5 * i + 4];
            np_.push_back(pp);
        }
        mpros_utils::updatePathLength(np_);
        mpros_utils::setCusp(np_);
        mpros_utils::updateCurvatureNew(np_);
        if (planner.inter_enable_)
        {
            planner.upsampleAndPopulatePath(np_);
            mpros_utils::updateCurvatureNew(np_);
            mpros_utils::updatePathLength(np_);
        }
        int k = 0;
        for (; k < (int)np_.size() && k < cap; ++k)
        {
            const PathPoint &q = np_[k];
            double *o = newpath_out + 8 * k;
            o[0] = q.getX(), o[1] = q.getY(), o[2] = q.getYaw(), o[3] = q.curv_, o[4] = q.direc_, o[5] = q.steering_;
            o[6] = q.is_cusp_ ? 1.0 : 0.0, o[7] = q.dist_;
        }
        return (int)np_.size();
    }

    // getNewPath() of a whole query (main_planner.cpp:225-233): Plan() then the up-sampled path. Returns its size (0 when
    // the plan failed).
    int ref_plan_newpath(void *map, const double *start3, const double *goal3, int do_update_goal, double *newpath_out, int cap)
    {
        NavMap *m = static_cast<NavMap *>(map);
        std::vector<double> s(start3, start3 + 3), g(goal3, goal3 + 3);
        if (do_update_goal)
            m->updateGoal(g[0], g[1]);
        HybridAstar planner(s, g, *m, g_vis, g_nh);
        if (!planner.Plan())
            return 0;
        std::vector<PathPoint> np_ = planner.getNewPath();
        for (int k = 0; k < (int)np_.size() && k < cap; ++k)
        {
            const PathPoint &q = np_[k];
            double *o = newpath_out + 8 * k;
            o[0] = q.getX(), o[1] = q.getY(), o[2] = q.getYaw(), o[3] = q.curv_, o[4] = q.direc_, o[5] = q.steering_;
            o[6] = q.is_cusp_ ? 1.0 : 0.0, o[7] = q.dist_;
        }
        return (int)np_.size();
    }

    // getSearchTree() of a whole query (main_planner.cpp:260): Plan() then the visualisation tree. out4: cap x 4
    // {next_x, next_y, curr_x, curr_y}; returns the number of segments.
    long ref_plan_search_tree(void *map, const double *start3, const double *goal3, int do_update_goal, double *out4, long cap)
    {
        NavMap *m = static_cast<NavMap *>(map);
        std::vector<double> s(start3, start3 + 3), g(goal3, goal3 + 3);
        if (do_update_goal)
            m->updateGoal(g[0], g[1]);
        HybridAstar planner(s, g, *m, g_vis, g_nh);
        planner.Plan();
        std::vector<Eigen::Vector4d> tree = planner.getSearchTree();
        for (long k = 0; k < (long)tree.size() && k < cap; ++k)
            for (int c = 0; c < 4; ++c)
                out4[4 * k + c] = tree[k](c);
        return (long)tree.size();
    }

    // ---- RS::RSCurve ------------------------------------------------------------------------------
    // cfg: {radius, max_steer, step, pen_gear, pen_back, pen_a3, pen_a4} with pen_a3/pen_a4 passed in the
    // POSITIONS the planner uses (hybrid_a_star.cpp:170: setPenalty(gear, back, STEER_ANGLE, STEER_CHANGE)).
    // start5: {x,y,yaw,dir,steer}; goal3. seg_out: 5 x 3 {len, steer, dir}; traj_out: cap x 5.
    // Returns number of trajectory points; *nseg, *cost are set.
    int ref_rs_solve(const double *cfg, const double *start5, const double *goal3, double *seg_out, int *nseg,
                     double *cost, double *traj_out, int traj_cap)
    {
        RS::RSCurve gen;
        gen.SetRadius(cfg[0]);
        gen.setSteeringAngle(cfg[1]);
        gen.setStepSize(cfg[2]);
        gen.setPenalty(cfg[3], cfg[4], cfg[5], cfg[6]);
        gen.SetStart(start5[0], start5[1], start5[2], (int)start5[3], start5[4]);
        gen.SetEnd(goal3[0], goal3[1], goal3[2]);
        gen.FindOptimalPath();
        std::vector<std::vector<double>> traj = gen.GetTraj();
        *cost = gen.GetLength();
        const RS::Path &path = gen.optimal_path_;
        *nseg = (int)path.size();
        for (int k = 0; k < (int)path.size() && k < 5; ++k)
        {
            seg_out[3 * k + 0] = path[k].GetLength();
            seg_out[3 * k + 1] = path[k].GetSteer();
            seg_out[3 * k + 2] = path[k].GetDirec();
        }
        int n = (int)traj.size() < traj_cap ? (int)traj.size() : traj_cap;
        for (int k = 0; k < n; ++k)
            for (int c = 0; c < 5; ++c)
                traj_out[5 * k + c] = traj[k][c];
        return (int)traj.size();
    }

    // RS::PathFunction::path1..12 (rs_curve.cpp:353-599) on n normalised goals {x, y, phi}: seg_out n x 5 x 3 {len, steer, dir}
    void ref_rs_word(int word, const double *xyphi, long n, double *seg_out, int *nseg)
    {
        typedef RS::Path (*Fn)(const RS::Point &);
        static const Fn fns[12] = {RS::PathFunction::path1, RS::PathFunction::path2, RS::PathFunction::path3, RS::PathFunction::path4,
                                   RS::PathFunction::path5, RS::PathFunction::path6, RS::PathFunction::path7, RS::PathFunction::path8,
                                   RS::PathFunction::path9, RS::PathFunction::path10, RS::PathFunction::path11, RS::PathFunction::path12};
        for (long k = 0; k < n; ++k)
        {
            RS::Path path = fns[word - 1](RS::Point(xyphi[3 * k], xyphi[3 * k + 1], xyphi[3 * k + 2]));
            nseg[k] = (int)path.size();
            for (int t = 0; t < 5; ++t)
            {
                const bool have = t < (int)path.size();
                seg_out[15 * k + 3 * t + 0] = have ? path[t].GetLength() : 0.0;
                seg_out[15 * k + 3 * t + 1] = have ? (double)path[t].GetSteer() : 0.0;
                seg_out[15 * k + 3 * t + 2] = have ? (double)path[t].GetDirec() : 0.0;
            }
        }
    }

    // RSStateSpace(rho).getCost() — OMPL stand-in, PARITY UNPINNED (oracle/port/ompl_rs_distance.h)
    double ref_rs_distance(double rho, const double *a3, const double *b3)
    {
        RSStateSpace ss(rho);
        ss.setStart(a3[0], a3[1], a3[2]);
        ss.setEnd(b3[0], b3[1], b3[2]);
        return ss.getCost();
    }
}
