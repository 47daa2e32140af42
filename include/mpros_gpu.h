/*
 * mpros_gpu.h — C ABI of the B200-native (sm_100a) data-parallel core of the mpros Hybrid A* planner.
 *
 * This is the drop-in boundary: plain pointers and sizes, opaque context handle, int status returns,
 * caller-owned buffers, one CUDA stream per context, a context is not thread-safe. There is NO CPU
 * fallback behind any entry point: without a CUDA device every call fails with MPROS_ERR_CUDA.
 *
 * Each entry point names the reference interface it replaces (file:line in Runningchauncey/Hybrid-Astar-Planner).
 * Index conventions are the reference's: i = row (y), j = col (x), linear index i*W + j, grid row 0 is the
 * bottom row of the map (nav_map.h:355,363).
 *
 * Entry points whose name ends in _dev take DEVICE pointers (for callers that keep data resident in HBM and
 * for device-timed benchmarking); all others take HOST pointers and do the copies themselves.
 */
#ifndef MPROS_GPU_H
#define MPROS_GPU_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define MPROS_API __attribute__((visibility("default")))
#else
#define MPROS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mpros_ctx mpros_ctx;

enum mpros_status
{
    MPROS_OK = 0,
    MPROS_ERR_INVALID = 1,   /* bad argument / call order (e.g. no map yet) */
    MPROS_ERR_CUDA = 2,      /* CUDA runtime failure, see mpros_last_error_string() */
    MPROS_ERR_CAPACITY = 3   /* caller-provided buffer too small */
};

/* Map layers, in the reference's member names (nav_map.h:494-507, node.h:29-44). */
enum mpros_layer
{
    MPROS_LAYER_STATIC_ORIG = 0,  /* static_map_orig_      int8   [H_orig*W_orig] 0/1           */
    MPROS_LAYER_STATIC = 1,       /* static_map_           int8   [H*W]           0/1           */
    MPROS_LAYER_INFLATE_ORIG = 2, /* inflate_map_orig_     int8   [H_orig*W_orig] 0/1           */
    MPROS_LAYER_GOAL = 3,         /* goal_cost_map_        int32  [H*W]           0..1000       */
    MPROS_LAYER_OBS_DIST = 4,     /* VoroNode::obs_dist_   int32  [H*W]           0..1000       */
    MPROS_LAYER_REGION = 5,       /* VoroNode::region_     int32  [H*W]           -1 = none     */
    MPROS_LAYER_ISEDGE = 6,       /* VoroNode::isedge_     uint8  [H*W]                         */
    MPROS_LAYER_EDGE_DIST = 7,    /* VoroNode::edge_dist_  int32  [H*W]           0..1000       */
    MPROS_LAYER_VORONOI = 8       /* voronoi_value_map_    float  [H*W]                         */
};

/* Parameters NavMap::config reads from /mpros/map/ (nav_map.cpp:25-35). */
typedef struct mpros_map_params
{
    double occupied_thresh;           /* occupied_thresh            default 0.68 (unused by the reference too) */
    double free_thresh;               /* free_thresh                default 0.2  */
    float voro_fallout_rate;          /* voro_fallout_rate (alpha)  default 10.0 */
    float voro_max_region;            /* voro_max_region (d_o_max)  default 50.0 */
    int downsampling_factor;          /* downsampling_factor        default 1    */
    double obstacle_inflation_radius; /* obstacle_inflation_radius  default 0.0  */
} mpros_map_params;

/* Parameters HybridAstar::config reads from /mpros/vehicle/ and /mpros/planner/ (hybrid_a_star.cpp:120-188). */
typedef struct mpros_planner_params
{
    double max_curvature;       /* default 0.4  */
    double width;               /* default 2.0  */
    double wheelbase;           /* default 3.0  */
    double length;              /* default 4.2  */
    double steer_offset;        /* center_to_rotate_center_longitudinal, default 0.0 */
    double max_steering_angle;  /* default 0.35 */
    double gear_penalty;            /* default 10.0 */
    double reverse_penalty;         /* default 0.0  */
    double steering_penalty;        /* default 0.1  */
    double steering_change_penalty; /* default 10.0 */
    double analytic_solution_range; /* default 10.0 */
    int yaw_discretization;         /* default 36   */
    double step_size;               /* default 1.1  */
    double path_point_distance;     /* default 0.2  */
    int steering_discretization;    /* default 3    */
    int max_iteration;              /* default 10000 */
    int optimal_path;               /* default 0 (false) */
    int interpolation_enable;       /* default 1 (true)  */
} mpros_planner_params;

/* Geometry of the current map: the public data members of NavMap (nav_map.h:469-487). */
typedef struct mpros_map_info
{
    int get_map, get_goal;
    int map_height, map_width;           /* downsampled grid */
    int map_height_orig, map_width_orig; /* original grid    */
    double map_resolution, map_resolution_orig;
    double map_origin_x, map_origin_y, map_origin_yaw, map_origin_sin, map_origin_cos;
    double x_max, x_min, y_max, y_min;
    int num_regions; /* number of obstacle components found by markObs (nav_map.cpp:384-403) */
} mpros_map_info;

/* ---- context ------------------------------------------------------------------------------------- */
MPROS_API int mpros_ctx_create(int device, mpros_ctx **out);
MPROS_API int mpros_ctx_destroy(mpros_ctx *ctx);
/* Last error message of the calling thread (never NULL). */
MPROS_API const char *mpros_last_error_string(void);
/* The context's cudaStream_t, for callers that time with CUDA events or chain their own work. */
MPROS_API void *mpros_ctx_stream(mpros_ctx *ctx);
MPROS_API int mpros_ctx_synchronize(mpros_ctx *ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
MPROS_API uint64_t mpros_ctx_launch_count(mpros_ctx *ctx);

/* ---- configuration ------------------------------------------------------------------------------- */
MPROS_API void mpros_map_params_default(mpros_map_params *p);         /* nav_map.cpp:27-34 defaults       */
MPROS_API void mpros_planner_params_default(mpros_planner_params *p); /* hybrid_a_star.cpp:128-161 defaults */
/* NavMap::config (nav_map.cpp:25-35). */
MPROS_API int mpros_map_config(mpros_ctx *ctx, const mpros_map_params *p);
/* HybridAstar::config + setFootprint (hybrid_a_star.cpp:120-206). Needs a map (uses map_resolution_orig_). */
MPROS_API int mpros_planner_config(mpros_ctx *ctx, const mpros_planner_params *p);

/* ---- NavMap ---------------------------------------------------------------------------------------- */
/* NavMap::setOccupancyMap (nav_map.cpp:42-162): threshold (N1), downsampleMap (N2, :592-652),
 * inflateObstacle (N3, :164-214), generateGoalMap if a goal is set (N4), generateVoronoiMap (N5-N9, :541-590).
 * data: host int8[height*width] occupancy (-1/0/100). resolution is the value NavMap stores in
 * map_resolution_orig_ (i.e. the float32 of the message widened to double); origin_yaw is what tf2::getYaw
 * returned. */
MPROS_API int mpros_map_set_occupancy(mpros_ctx *ctx, const int8_t *data, int width, int height, double resolution,
                            double origin_x, double origin_y, double origin_yaw);
/* Same, data already resident in HBM. */
MPROS_API int mpros_map_set_occupancy_dev(mpros_ctx *ctx, const int8_t *d_data, int width, int height, double resolution,
                                double origin_x, double origin_y, double origin_yaw);
/* NavMap::updateGoal -> generateGoalMap (nav_map.h:204-210, nav_map.cpp:279-349). Goal outside the map
 * clears get_goal (nav_map.cpp:297-302) and returns MPROS_OK. */
MPROS_API int mpros_map_update_goal(mpros_ctx *ctx, double goal_x, double goal_y);
MPROS_API int mpros_map_get_info(mpros_ctx *ctx, mpros_map_info *out);
/* getGoalMap / getVoronoiMap / the private layers (nav_map.h:261-273): copies one layer to host memory in the
 * reference's element type (see enum mpros_layer). dst_bytes must be >= the layer size. */
MPROS_API int mpros_map_download(mpros_ctx *ctx, int layer, void *dst, size_t dst_bytes);
/* NavMap::pointInCollision / pointInCollisionOrig / pointInCollisionINFLOrig (nav_map.cpp:249-277) for a batch of
 * cells; layer is MPROS_LAYER_STATIC, _STATIC_ORIG or _INFLATE_ORIG; ij: host int32[2*n] {i,j}; out: host uint8[n]. */
MPROS_API int mpros_map_point_in_collision(mpros_ctx *ctx, int layer, const int32_t *ij, size_t n, uint8_t *out);

/* ---- collision (C1/C2) ----------------------------------------------------------------------------- */
/* HybridAstar::footPrintInCollision4 (hybrid_a_star.cpp:617-680) for n poses. xyyaw: double[3*n] {x,y,yaw};
 * verdicts: uint8[n], 1 = in collision. */
MPROS_API int mpros_collision_check_poses(mpros_ctx *ctx, const double *xyyaw, size_t n, uint8_t *verdicts);
MPROS_API int mpros_collision_check_poses_dev(mpros_ctx *ctx, const double *d_xyyaw, size_t n, uint8_t *d_verdicts);
/* HybridAstar::pathInCollision (hybrid_a_star.cpp:404-430) for a batch of paths: path p owns poses
 * [offsets[p], offsets[p+1]); verdicts: uint8[n_paths]. */
MPROS_API int mpros_collision_check_paths(mpros_ctx *ctx, const double *xyyaw, const int64_t *offsets, size_t n_paths,
                                uint8_t *verdicts);

#ifdef __cplusplus
}
#endif
#endif /* MPROS_GPU_H */
