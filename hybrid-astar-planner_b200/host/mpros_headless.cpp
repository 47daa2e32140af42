// Headless single-query driver: what mpros_planner/src/main_planner.cpp:179-229 does, without a ROS master.
//   mpros_headless <map.yaml> <sx> <sy> <syaw> <gx> <gy> <gyaw> [key=value ...]
// Loads the PGM named by the map YAML with ROS map_server's trinary semantics, feeds NavMap,
// updates the goal, constructs HybridAstar, calls Plan() and prints one JSON line with the result.
// Extra key=value pairs override parameter-server entries (e.g. /mpros/planner/steering_discretization=5).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "mpros_facade.h"

using namespace mpros_b200;

static std::string dirname_of(const std::string &p)
{
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? "." : p.substr(0, k);
}

struct MapYaml
{
    std::string image;
    double resolution = 0.05, origin[3] = {0, 0, 0}, occupied_thresh = 0.65, free_thresh = 0.196;
    int negate = 0;
};

static MapYaml read_yaml(const std::string &path)
{
    MapYaml y;
    std::ifstream f(path);
    if (!f)
        throw std::runtime_error("cannot open " + path);
    std::string line;
    while (std::getline(f, line))
    {
        size_t h = line.find('#');
        if (h != std::string::npos)
            line = line.substr(0, h);
        size_t c = line.find(':');
        if (c == std::string::npos)
            continue;
        std::string k = line.substr(0, c), v = line.substr(c + 1);
        auto trim = [](std::string s) {
            size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
            return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
        };
        k = trim(k), v = trim(v);
        if (k == "image")
            y.image = v;
        else if (k == "resolution")
            y.resolution = atof(v.c_str());
        else if (k == "negate")
            y.negate = atoi(v.c_str());
        else if (k == "occupied_thresh")
            y.occupied_thresh = atof(v.c_str());
        else if (k == "free_thresh")
            y.free_thresh = atof(v.c_str());
        else if (k == "origin")
        {
            for (char &ch : v)
                if (ch == '[' || ch == ']' || ch == ',')
                    ch = ' ';
            std::istringstream ss(v);
            ss >> y.origin[0] >> y.origin[1] >> y.origin[2];
        }
    }
    return y;
}

// binary (P5) or ASCII (P2) PGM, 8-bit
static std::vector<int> read_pgm(const std::string &path, int &w, int &h, int &maxval)
{
    std::ifstream f(path, std::ios::binary);
    if (!f)
        throw std::runtime_error("cannot open " + path);
    std::string magic;
    f >> magic;
    auto next_int = [&]() {
        while (true)
        {
            f >> std::ws;
            if (f.peek() == '#')
            {
                std::string skip;
                std::getline(f, skip);
                continue;
            }
            int v;
            f >> v;
            return v;
        }
    };
    w = next_int(), h = next_int(), maxval = next_int();
    std::vector<int> px((size_t)w * h);
    if (magic == "P5")
    {
        f.get(); // single whitespace
        std::vector<unsigned char> raw((size_t)w * h);
        f.read(reinterpret_cast<char *>(raw.data()), raw.size());
        for (size_t k = 0; k < raw.size(); ++k)
            px[k] = raw[k];
    }
    else if (magic == "P2")
        for (auto &v : px)
            f >> v;
    else
        throw std::runtime_error("unsupported PGM magic " + magic);
    return px;
}

int main(int argc, char **argv)
{
    if (argc < 8)
    {
        std::fprintf(stderr, "usage: %s <map.yaml> <sx> <sy> <syaw> <gx> <gy> <gyaw> [key=value ...]\n", argv[0]);
        return 2;
    }
    try
    {
        MapYaml y = read_yaml(argv[1]);
        std::string img = y.image;
        if (!img.empty() && img[0] != '/')
            img = dirname_of(argv[1]) + "/" + img;
        int w, h, maxval;
        std::vector<int> px = read_pgm(img, w, h, maxval);
        OccupancyGridLite og;
        og.width = w, og.height = h, og.resolution = (float)y.resolution;
        og.origin_x = y.origin[0], og.origin_y = y.origin[1];
        og.setOriginYaw(y.origin[2]);
        og.data.resize((size_t)w * h);
        for (int r = 0; r < h; ++r)
            for (int c = 0; c < w; ++c)
            {
                double color = px[(size_t)r * w + c] * (255.0 / maxval);
                double p = y.negate ? color / 255.0 : (255.0 - color) / 255.0;
                int8_t v = -1;
                if (p > y.occupied_thresh)
                    v = 100;
                else if (p < y.free_thresh)
                    v = 0;
                og.data[(size_t)(h - 1 - r) * w + c] = v; // image row 0 is the top, grid row 0 the bottom
            }
        // shipped parameters (planner_config.yaml, vehicle_config.yaml) + the map YAML under /mpros/map (vehicle.launch:18)
        ParamServer nh;
        nh.set("/mpros/planner/step_size", 2.0), nh.set("/mpros/planner/max_iteration", 100000);
        nh.set("/mpros/planner/gear_penalty", 10.0), nh.set("/mpros/planner/reverse_penalty", 10.0);
        nh.set("/mpros/planner/steering_penalty", 0.5), nh.set("/mpros/planner/steering_change_penalty", 0.5);
        nh.set("/mpros/planner/analytic_solution_range", 10.0), nh.set("/mpros/planner/yaw_discretization", 36);
        nh.set("/mpros/planner/path_point_distance", 0.1), nh.set("/mpros/map/downsampling_factor", 1);
        nh.set("/mpros/map/obstacle_inflation_radius", 1.0), nh.set("/mpros/map/free_thresh", y.free_thresh);
        nh.set("/mpros/map/occupied_thresh", y.occupied_thresh), nh.set("/mpros/vehicle/max_steering_angle", 0.35);
        nh.set("/mpros/vehicle/max_curvature", 0.125), nh.set("/mpros/vehicle/width", 2.0), nh.set("/mpros/vehicle/wheelbase", 3.0);
        nh.set("/mpros/vehicle/length", 4.2), nh.set("/mpros/vehicle/center_to_rotate_center_longitudinal", 0.0);
        for (int k = 8; k < argc; ++k)
        {
            std::string kv = argv[k];
            size_t e = kv.find('=');
            if (e != std::string::npos)
                nh.set(kv.substr(0, e), atof(kv.substr(e + 1).c_str()));
        }
        std::vector<double> start = {atof(argv[2]), atof(argv[3]), atof(argv[4])}, goal = {atof(argv[5]), atof(argv[6]), atof(argv[7])};

        // main_planner.cpp:179-229: NavMap(nh) + config, the "map" callback, the five published layers, updateGoal, the
        // five-argument planner constructor, Plan()
        NavMap map(nh);
        Visualizer vis;
        map.config(nh);
        map.getOccupancyMap(std::make_shared<const OccupancyGridLite>(og));
        map.updateGoal(goal[0], goal[1]);
        struct NamedMsg
        {
            const char *name;
            OccupancyGridLite msg;
        } msgs[5] = {{"static", map.getStaticMapMsg()}, {"voronoi", map.getVoroMapMsg()}, {"static_orig", map.getStaticOrigMapMsg()},
                     {"inflate_orig", map.getInflationOrigMapMsg()}, {"goal", map.getGoalMapMsg()}};
        HybridAstar planner(start, goal, map, vis, nh);
        bool ok = planner.Plan();
        std::vector<std::vector<double>> path = planner.getPath();
        bool collides = ok ? planner.pathInCollision(path) : false;
        std::printf("{\"plan_ok\": %d, \"goal_g\": %.17g, \"iterations\": %lld, \"n_primitives\": %lld, \"path_len\": %zu, \"path_in_collision\": %d, "
                    "\"map\": [%d, %d], \"path\": [",
                    ok ? 1 : 0, planner.goalCost(), planner.iterations(), planner.primitivesEvaluated(), path.size(), collides ? 1 : 0,
                    map.map_height_, map.map_width_);
        for (size_t k = 0; k < path.size(); ++k)
            std::printf("%s[%.17g, %.17g, %.17g, %.17g, %.17g]", k ? ", " : "", path[k][0], path[k][1], path[k][2], path[k][3], path[k][4]);
        std::printf("], \"new_path\": [");
        std::vector<PathPoint> np = ok ? planner.getNewPath() : std::vector<PathPoint>();
        for (size_t k = 0; k < np.size(); ++k)
            std::printf("%s[%.17g, %.17g, %.17g, %.17g, %.17g, %.17g, %d, %.17g]", k ? ", " : "", np[k].getX(), np[k].getY(), np[k].getYaw(), np[k].curv_, np[k].direc_,
                        np[k].steering_, np[k].is_cusp_ ? 1 : 0, np[k].dist_);
        std::printf("], \"msgs\": {");
        for (int k = 0; k < 5; ++k)
        {
            // FNV-1a-64 of msg.data, with the geometry vecToMsg writes (nav_map.cpp:654-691)
            unsigned long long hsh = 1469598103934665603ull;
            for (int8_t v : msgs[k].msg.data)
                hsh = (hsh ^ (unsigned char)v) * 1099511628211ull;
            std::printf("%s\"%s\": {\"frame_id\": \"%s\", \"height\": %u, \"width\": %u, \"resolution\": %.9g, \"origin\": [%.17g, %.17g], \"n\": %zu, "
                        "\"fnv\": \"%016llx\"}",
                        k ? ", " : "", msgs[k].name, msgs[k].msg.frame_id.c_str(), msgs[k].msg.height, msgs[k].msg.width, (double)msgs[k].msg.resolution,
                        msgs[k].msg.origin_x, msgs[k].msg.origin_y, msgs[k].msg.data.size(), hsh);
        }
        // main_planner.cpp:259: the searched tree handed to the visualiser; :243 pathInCollision on the PathPoint path
        std::printf("}, \"search_tree_segments\": %zu, \"new_path_in_collision\": %d, \"rs_shot_from_start\": %d}\n", planner.getSearchTree().size(),
                    (ok && planner.pathInCollision(np)) ? 1 : 0, planner.genRSPath(start[0], start[1], start[2]) ? 1 : 0);
        return 0;
    }
    catch (const std::exception &e)
    {
        std::fprintf(stderr, "mpros_headless: %s\n", e.what());
        return 1;
    }
}
