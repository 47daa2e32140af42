/*
 * mpros_gpu.h — C ABI of the B200-native (sm_100a) data-parallel core of the mpros Hybrid A* planner.
 *
 * This is the drop-in boundary: plain pointers and sizes, opaque context handle, int status returns,
 * caller-owned buffers, one CUDA stream per context (plus one per extra plan lane), a context is not
 * thread-safe. Everything a context allocates (map layers, staging areas, the planner's per-query workspaces of
 * every lane) belongs to that context and is freed by mpros_ctx_destroy; two contexts on one device share
 * nothing and may run concurrently from two host threads. There is NO CPU
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

/* Developer / tuning switches of one context. The MPROS_* environment variables give the defaults and are read ONCE, in
 * mpros_ctx_create; this call changes a switch afterwards (value NULL or "" = default). Keys and values:
 *   "c1_mode"   "2phase" (default) | "2phase_ldg" | "walk" footprint kernel: pre-classified two-phase form with the pose stream
 *                                                          staged by cp.async.bulk / the same without staging / plain probe walk
 *   "goal_mode" "cluster" (default) | "cluster8" | "grid"  N4 wavefront: 16-CTA cluster when granted / 8-CTA cluster / cooperative grid
 *   "plan_mode" "solo" (default) | "team" | "cluster" | "duo"  planner: warp-per-query pass, then the team kernel for the queries
 *                                                          that outgrow its workspaces / team kernel only (one CTA per query) /
 *                                                          a CTA pair on the two SMs of a TPC / two teams per CTA
 *   "solo_nodes" node-pool size of a small (warp-per-query) workspace (default 32768)
 *   "solo_min_queries" batches with fewer queries go straight to the team kernel (default 8192)
 *   "plan_grow" "1" (default) | "0"                       a query that outgrows its small workspace moves into one four times the
 *                                                          size and stays with its warp / goes to the team pass at once
 *   "plan_policy" "auto" (default) | "latency" | "throughput"  what the squad pass does with the queries of a thinned-out squad once
 *                                                          the tickets have run out: to the team pass at once (shortest time for ONE
 *                                                          batch) / onto a list from which idle warps of well-filled squads adopt them
 *                                                          (lowest cost per query); auto = throughput when another batch is in flight
 *                                                          at submission. "plan_age_limit": iterations after which a query goes to the
 *                                                          team pass for good under the throughput policy (default 30000)
 *   "debug"     "0" (default) | "1"                        per-pass timing on stderr
 * Results never depend on these switches (tests/ hold every mode to the same outputs). No reference counterpart. */
MPROS_API int mpros_ctx_set_option(mpros_ctx *ctx, const char *key, const char *value);

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
/* ---- the same, tiled by ROW BANDS across the GPUs of one box (very large maps) ----------------------
 * NavMap::setOccupancyMap's original-resolution stages split per band; every rank (one context per GPU) calls the same
 * sequence and the caller moves the bit-packed rows between ranks (NCCL over NVLink; hybrid-astar-planner_b200/sharding.py:
 * set_occupancy_banded):
 *   1. mpros_map_band_begin      geometry + allocation (nav_map.cpp:54-112); *halo_rows = the structuring element's
 *                                radius in pixels = rows of static_map_orig_ a band needs from each neighbour for N3
 *   2. mpros_map_band_threshold  N1 (nav_map.cpp:114-131) for the rank's rows [row0, row0 + nrows)
 *   3. halo exchange             rows [row0 - halo, row0) and [row0 + nrows, row0 + nrows + halo) of MPROS_LAYER_STATIC_ORIG
 *                                from the neighbouring ranks into the SAME rows of this rank's layer (mpros_map_band_layer_ptr)
 *   4. mpros_map_band_inflate    N3 (inflateObstacle, nav_map.cpp:164-214) for the rank's rows
 *   5. all-gather                every rank's rows of both layers into every rank's layers
 *   6. mpros_map_band_finish     N2, N4 (if a goal is set), N5-N9 on the complete layers (at the down-sampled resolution)
 * The result is bit-identical to mpros_map_set_occupancy on one GPU. Layers are bit-packed: bit (j & 31) of 32-bit word
 * [i * row_pitch_bytes / 4 + (j >> 5)]. */
MPROS_API int mpros_map_band_begin(mpros_ctx *ctx, int width, int height, double resolution, double origin_x, double origin_y,
                                   double origin_yaw, int *halo_rows);
/* rows: host int8[nrows * width], the rank's rows of the occupancy grid. */
MPROS_API int mpros_map_band_threshold(mpros_ctx *ctx, const int8_t *rows, int row0, int nrows);
MPROS_API int mpros_map_band_threshold_dev(mpros_ctx *ctx, const int8_t *d_rows, int row0, int nrows);
/* Device pointer and row pitch of MPROS_LAYER_STATIC_ORIG / MPROS_LAYER_INFLATE_ORIG in their bit-packed form. */
MPROS_API int mpros_map_band_layer_ptr(mpros_ctx *ctx, int layer, void **d_ptr, size_t *row_pitch_bytes);
MPROS_API int mpros_map_band_inflate(mpros_ctx *ctx, int row0, int nrows);
MPROS_API int mpros_map_band_finish(mpros_ctx *ctx);
/* NavMap::updateGoal -> generateGoalMap (nav_map.h:204-210, nav_map.cpp:279-349). Goal outside the map
 * clears get_goal (nav_map.cpp:297-302) and returns MPROS_OK. */
MPROS_API int mpros_map_update_goal(mpros_ctx *ctx, double goal_x, double goal_y);
MPROS_API int mpros_map_get_info(mpros_ctx *ctx, mpros_map_info *out);
/* getGoalMap / getVoronoiMap / the private layers (nav_map.h:261-273): copies one layer to host memory in the
 * reference's element type (see enum mpros_layer). dst_bytes must be >= the layer size. */
MPROS_API int mpros_map_download(mpros_ctx *ctx, int layer, void *dst, size_t dst_bytes);
/* NavMap::getStaticMapMsg / getGoalMapMsg / getVoroMapMsg / getStaticOrigMapMsg / getInflationOrigMapMsg -> vecToMsg
 * (nav_map.cpp:654-716): the int8 `data` of the nav_msgs/OccupancyGrid the node publishes for one layer,
 * floor(float(v) * 100 / min(100, max)) with the reference's int8 wrap-around (goal-map values above 127).
 * layer: MPROS_LAYER_STATIC, _GOAL, _VORONOI, _STATIC_ORIG or _INFLATE_ORIG. *n_out = msg.data.size() (0 while the layer
 * does not exist, like the reference's empty message); dst == NULL only queries the size. Geometry of the message:
 * mpros_map_get_info (height/width/resolution of the original or the down-sampled grid, origin). */
MPROS_API int mpros_map_layer_msg(mpros_ctx *ctx, int layer, int8_t *dst, size_t dst_bytes, size_t *n_out);
/* Overwrite the Voronoi-field layer (float[H*W], host) with caller-provided values, e.g. a field computed elsewhere.
 * Only MPROS_LAYER_VORONOI can be uploaded; the next mpros_map_set_occupancy recomputes it. */
MPROS_API int mpros_map_upload(mpros_ctx *ctx, int layer, const void *src, size_t src_bytes);
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

/* The other footprint tests of the public HybridAstar interface (no caller in the reference):
 * method 1 footPrintInCollision (hybrid_a_star.cpp:432-483, Bresenham lines between the corner cells, raw map),
 * method 2 footPrintInCollision2 (:485-575, winding-angle sum of the obstacle cells in the circumscribed square),
 * method 3 footPrintInCollision3 (:577-615, rasterised footprint kernel per yaw bin, mpros_utils/footprint.h),
 * method 4 = mpros_collision_check_poses. Same arguments as mpros_collision_check_poses. */
MPROS_API int mpros_collision_check_poses_method(mpros_ctx *ctx, int method, const double *xyyaw, size_t n, uint8_t *verdicts);

/* ---- Reeds-Shepp (R1-R5) ----------------------------------------------------------------------------- */
/* RS::RSCurve configuration: SetRadius / setSteeringAngle / setStepSize / setPenalty (rs_curve.h:58-100).
 * The four penalties are in RSCurve::setPenalty's DECLARED order (gear, backward, steer_change, steer_angle);
 * note the planner passes its steering_penalty / steering_change_penalty swapped (hybrid_a_star.cpp:170) —
 * mpros_rs_default_config returns the values as the planner's RSCurve actually holds them. */
typedef struct mpros_rs_params
{
    double radius;          /* min_radius_      */
    double steering_angle;  /* steer_angle_     */
    double step_size;       /* step_size_       */
    double pen_gear, pen_backward, pen_steer_change, pen_steer_angle;
} mpros_rs_params;
MPROS_API int mpros_rs_default_config(mpros_ctx *ctx, mpros_rs_params *out);
/* RSCurve::SetStart/SetEnd/FindOptimalPath/GetLength/GetTraj (rs_curve.cpp:111-222, 275-284) for n shots.
 * starts5: double[5n] {x, y, yaw, direction, steer} (SetStart's arguments); goals3: double[3n].
 * segs: double[15n] = 5 x {length, steer(-1/0/1), direction(+-1)} of optimal_path_; nseg: int32[n];
 * cost: double[n] = GetLength(); traj: double[n * traj_cap * 5] {x, y, yaw, steer*steer_angle, direction}
 * (may be NULL when traj_cap == 0); npts: int32[n] = GetTraj().size() (poses beyond traj_cap are not stored;
 * the device generates at most 256 poses per shot). rs == NULL uses the planner's configuration. */
MPROS_API int mpros_rs_solve_batch(mpros_ctx *ctx, const mpros_rs_params *rs, const double *starts5, const double *goals3,
                                   size_t n, double *segs, int32_t *nseg, double *cost, double *traj, int traj_cap,
                                   int32_t *npts);
/* HybridAstar::genRSPath without the range test (hybrid_a_star.cpp:872-880): solve + sampled-path collision.
 * verdicts: uint8[n], 1 = pathInCollision(rs_path). */
MPROS_API int mpros_rs_shot_batch(mpros_ctx *ctx, const double *starts5, const double *goals3, size_t n, double *segs,
                                  int32_t *nseg, double *cost, double *traj, int traj_cap, int32_t *npts, uint8_t *verdicts);
/* The same shot with every buffer in device memory (no trajectory output); asynchronous on the context's stream. */
MPROS_API int mpros_rs_shot_batch_dev(mpros_ctx *ctx, const double *d_starts5, const double *d_goals3, size_t n, double *d_segs,
                                      int32_t *d_nseg, double *d_cost, int32_t *d_npts, uint8_t *d_verdicts);
/* RS::PathFunction::path1 .. path12 (rs_curve.h:248-334, rs_curve.cpp:353-599): the closed-form word `word` (1..12) for n
 * goals already normalised into the start frame and divided by the radius. xyphi: double[3n]; segs: double[15n] = 5 x
 * {|value|, steer, direction} as the PathSeg constructor stores them (rs_curve_basic.h:84-88), zero-filled beyond nseg[k]. */
MPROS_API int mpros_rs_word_batch(mpros_ctx *ctx, int word, const double *xyphi, size_t n, double *segs, int32_t *nseg);
/* RSStateSpace(rho).setStart/setEnd/getCost (rs_statespace.h:37-54): OMPL-equivalent Reeds-Shepp distance. */
MPROS_API int mpros_rs_distance_batch(mpros_ctx *ctx, double rho, const double *a3, const double *b3, size_t n, double *out);

/* ---- expansion (E1-E5) ---------------------------------------------------------------------------------- */
/* HybridAstar::expandNode (hybrid_a_star.cpp:304-402) for n parents right after initialize(start, goal): the closed
 * set holds only the start (g = 0) and goal (10000) seeds (start3 may be NULL: no seeds). Needs the goal map
 * (mpros_map_update_goal). parents6: double[6n] {x, y, yaw, steer, direction, g}. P = 2 * steer_set_.size().
 * prims: double[n*P*7] {x, y, yaw, steer, direction, collision(0/1), cell_in_map(0/1)} for every primitive in
 * evaluation order (tree_); childs: double[n*P*10] {x, y, yaw, steer, direction, g, h, f, idx_i, idx_j} of the
 * children inserted into the open list, in insertion order; nchild: int32[n]. */
MPROS_API int mpros_expand_batch(mpros_ctx *ctx, const double *start3, const double *goal3, const double *parents6, size_t n,
                                 double *prims, double *childs, int32_t *nchild);
/* Same with the batch in device memory (start3/goal3 stay host pointers); asynchronous on the context's stream. */
MPROS_API int mpros_expand_batch_dev(mpros_ctx *ctx, const double *start3, const double *goal3, const double *d_parents6, size_t n,
                                     double *d_prims, double *d_childs, int32_t *d_nchild);

/* Compact output of the same expansion: only the children expandNode inserts into the open list (hybrid_a_star.cpp:369-398),
 * as 80-byte records appended in batches per parent - the children of one parent are contiguous and in insertion order
 * (`order`), parents appear in no particular order. About two records per expansion travel instead of the 816 bytes of
 * fixed-size arrays above (P = 6), which is what bounds the host-buffer form. *n_records = records produced (may exceed
 * cap_records: then the call returns MPROS_ERR_CAPACITY after storing the first cap_records); nchild: int32[n], may be NULL. */
typedef struct mpros_child_record
{
    uint32_t parent;  /* index of the parent in parents6                          */
    uint16_t prim;    /* motion primitive: 2 * steer index + direction index       */
    int16_t dir;      /* +1 forward, -1 backward                                   */
    int32_t i, j;     /* cell of the child on the down-sampled grid (calIndex)     */
    double x, y, yaw, steer, g, h, f;
    uint32_t order;   /* position among the parent's inserted children             */
    uint32_t reserved;
} mpros_child_record;
MPROS_API int mpros_expand_batch_compact(mpros_ctx *ctx, const double *start3, const double *goal3, const double *parents6, size_t n,
                                         mpros_child_record *records, size_t cap_records, size_t *n_records, int32_t *nchild);
/* Device buffers; d_n_records: one uint64 in device memory; asynchronous on the context's stream. */
MPROS_API int mpros_expand_batch_compact_dev(mpros_ctx *ctx, const double *start3, const double *goal3, const double *d_parents6, size_t n,
                                             mpros_child_record *d_records, size_t cap_records, unsigned long long *d_n_records,
                                             int32_t *d_nchild);

/* ---- planner ------------------------------------------------------------------------------------------------ */
/* Per-query result of mpros_plan_batch: what HybridAstar::Plan() returns and why. */
enum mpros_plan_status
{
    MPROS_PLAN_FOUND = 1,        /* Plan() == true ("Goal found")                                  */
    MPROS_PLAN_LIMIT = 0,        /* "Iteration limit reached! Plan failed!"                        */
    MPROS_PLAN_OPEN_EMPTY = -1,  /* "Open set empty! Plan failed!"                                 */
    MPROS_PLAN_NOT_INIT = -2,    /* initialize() bailed out: start/goal in collision or outside    */
    MPROS_PLAN_OVERFLOW = -3     /* internal node pool exhausted (never returned after the retries) */
};
/* For nq independent (start, goal) queries on the current map: NavMap::updateGoal + HybridAstar ctor/initialize +
 * Plan() + the raw path_ of setPath (main_planner.cpp:225-229; hybrid_a_star.cpp:52-118, 682-811; getPath()).
 * The per-query goal map is computed on demand on the device (values identical to generateGoalMap's).
 * starts3/goals3: double[3nq]; status: int32[nq] (enum mpros_plan_status); cost: double[nq] = goal_node_->gvalue_
 * (-1 when not found); path_len: int32[nq] = path_.size(); paths: double[nq*path_cap*5] {x, y, yaw, steer, direction}
 * (first path_cap points; may be NULL with path_cap == 0); stats (optional): int64[6nq] {iterations, primitives
 * evaluated (tree_.size()), RS shots, footPrintInCollision4 calls, goal-map cells computed, nodes inserted}. */
MPROS_API int mpros_plan_batch(mpros_ctx *ctx, const double *starts3, const double *goals3, size_t nq, int32_t *status,
                               double *cost, int32_t *path_len, double *paths, int path_cap, int64_t *stats);
MPROS_API int mpros_plan_batch_dev(mpros_ctx *ctx, const double *d_starts3, const double *d_goals3, size_t nq,
                                   int32_t *d_status, double *d_cost, int32_t *d_path_len, double *d_paths, int path_cap,
                                   int64_t *d_stats);

/* Batches in flight ("lanes"). The reference plans one query at a time in the ROS spin thread (main_planner.cpp:225-229,
 * planner_ros.cpp:44-51); a caller that streams query batches (a planning service) keeps up to MPROS_PLAN_LANES of them in
 * flight: mpros_plan_batch_submit[_dev] enqueues the batch on the lane's own stream (lane 0 = the context's stream) with
 * the lane's own planner workspaces and returns; mpros_plan_batch_wait(ctx, lane) blocks until that lane's batch and its
 * result copies are complete (lane < 0: all lanes). The iteration-limit tail of one batch - a handful of queries, each a
 * dependent chain of up to max_iteration expansions - then overlaps the next batch. Results are identical to
 * mpros_plan_batch (which is submit on lane 0 + wait). Host buffers of mpros_plan_batch_submit must stay valid until the
 * wait and should be page-locked (pageable buffers make the copies synchronous). The planner workspaces are POOLS of the
 * context that all lanes share - one workspace per query the device can keep resident, whatever the number of lanes: 296
 * team workspaces (~33 GB at the 100 000-iteration limit on a B200) and, once a batch of 8192 queries or more has been planned,
 * 4884 small ones of the warp-per-query pass (~43 GB) and 1184 grown ones (~33 GB). A team / warp takes a free workspace when
 * it starts and returns it when it exits, so the number of lanes costs streams and staging areas, not workspaces. With other
 * batches in flight when a batch is submitted it is planned for throughput rather than for its own latency ("plan_policy"
 * above; per-query results are the same either way). The map layers are shared by all lanes too:
 * do not call mpros_map_set_occupancy / mpros_map_upload / mpros_planner_config while a batch is in flight. A lane takes
 * one batch at a time (submitting to a busy lane queues behind it on the lane's stream). */
#define MPROS_PLAN_LANES 8
MPROS_API int mpros_plan_batch_submit(mpros_ctx *ctx, int lane, const double *starts3, const double *goals3, size_t nq,
                                      int32_t *status, double *cost, int32_t *path_len, double *paths, int path_cap,
                                      int64_t *stats);
MPROS_API int mpros_plan_batch_submit_dev(mpros_ctx *ctx, int lane, const double *d_starts3, const double *d_goals3, size_t nq,
                                          int32_t *d_status, double *d_cost, int32_t *d_path_len, double *d_paths,
                                          int path_cap, int64_t *d_stats);
MPROS_API int mpros_plan_batch_wait(mpros_ctx *ctx, int lane);
/* Frees the planner workspace pools of this context (node pools, open lists, closed sets, goal-wavefront windows: up to
 * ~109 GB on a B200 after 16 384-query batches, ~33 GB with "plan_mode" "team") after waiting for the batches in flight. They
 * are owned by the context, re-created by the next mpros_plan_batch* call and freed by mpros_ctx_destroy. */
MPROS_API int mpros_plan_release_workspaces(mpros_ctx *ctx);
/* cudaStream_t of a lane (NULL before its first submit), e.g. to record events or order other device work behind a batch */
MPROS_API void *mpros_plan_lane_stream(mpros_ctx *ctx, int lane);

/* HybridAstar::getSearchTree (hybrid_a_star.cpp:1114-1164) of ONE query: plans (start, goal) exactly like
 * mpros_plan_batch with nq = 1 while recording tree_ (the pose of every expanded node; hybrid_a_star.cpp:352 stores it once
 * per primitive) and turns every (expanded node, primitive) into ceil(step_size / path_point_distance) leaf segments
 * {next_x, next_y, curr_x, curr_y} (the reference's Eigen::Vector4d), in the reference's order. tree4: double[4 *
 * cap_segments]; *n_segments = number of segments; tree4 == NULL with cap_segments == 0 only queries the size (the plan is
 * run); MPROS_ERR_CAPACITY when the buffer is too small. */
MPROS_API int mpros_plan_search_tree(mpros_ctx *ctx, const double *start3, const double *goal3, int32_t *status, double *cost,
                                     double *tree4, size_t cap_segments, size_t *n_segments);

/* ---- several GPUs of one box --------------------------------------------------------------------------
 * One rank = one process (or host thread) + one context per GPU. The reference is a single-threaded ROS node
 * (main_planner.cpp:166-268) and has no counterpart; these entry points are what a multi-GPU planning service built on
 * NavMap / HybridAstar binds. All exchanges run inside the library: NCCL over NVLink / NVSwitch, enqueued on the
 * context's stream (libnccl.so.2 is resolved at run time; without it the mpros_comm_*_nccl calls fail with MPROS_ERR_CUDA
 * and everything single-GPU keeps working). */
typedef struct mpros_comm mpros_comm;
/* 128-byte NCCL unique id (ncclGetUniqueId). Rank 0 creates it; the launcher hands it to the other ranks (MPI, its own rendezvous,
 * a file). */
MPROS_API int mpros_comm_unique_id(void *id128);
/* ncclCommInitRank on the context's device; collective over all ranks. */
MPROS_API int mpros_comm_create_nccl(mpros_ctx *ctx, int world, int rank, const void *id128, mpros_comm **out);
/* Wrap an ncclComm_t the caller already owns (it is not destroyed with the handle). */
MPROS_API int mpros_comm_adopt_nccl(int world, int rank, void *nccl_comm, mpros_comm **out);
/* `world` ranks inside ONE process (out[0..world)): one host thread and one context per rank, parts exchanged with
 * cudaMemcpyAsync behind a host barrier. Every rank must make the same calls, each from its own thread. Lets the banded
 * algorithms run on a single GPU (tests) and serves boxes driven by one process. */
MPROS_API int mpros_comm_create_local(int world, mpros_comm **out);
MPROS_API int mpros_comm_destroy(mpros_comm *comm);
/* world / rank, payload bytes this rank contributed to exchanges so far and the number of exchange steps (any may be NULL). */
MPROS_API int mpros_comm_info(const mpros_comm *comm, int *world, int *rank, uint64_t *bytes_sent, uint64_t *exchanges);

/* NavMap::setOccupancyMap (nav_map.cpp:42-162) of a very large map with EVERY stage tiled by row bands across the ranks:
 * N1 threshold (:114-131) and N3 inflateObstacle (:164-214) on the rank's rows of the original grid, N5-N9
 * generateVoronoiMap (:541-590: markObs/floodFillObs :384-432, calObsDist :434-473, findEdge :475-507, calEdgeDist :509-539,
 * cost loop :565-582) on the rank's rows of the down-sampled grid. Exchanges: the finished bands of static_map_orig_ and
 * inflate_map_orig_ (the first doubles as N3's halo), for markObs the labels of each band's first and last row (seam
 * union) and the region ids of those rows, for each distance transform ONE carry exchange (two aggregate rows per band
 * serve the forward and the backward sweep), and the finished bands of obs_dist_ / region_ / isedge_ / edge_dist_ / the
 * Voronoi cost. N2 and the probe layouts are one cheap pass over the gathered bits on every rank; N4 never needs tiling
 * (values are capped at 999). Afterwards every rank holds the complete layers, bit-identical to mpros_map_set_occupancy.
 * rows: this rank's rows [row0, row0 + nrows) of the int8 occupancy grid (mpros_map_band_rows; a rank may own none: pass
 * any non-NULL pointer); bands are cut on the down-sampled grid, ceil(H / world) rows per rank. Collective. */
MPROS_API int mpros_map_band_rows(mpros_ctx *ctx, const mpros_comm *comm, int height, int *row0, int *nrows);
MPROS_API int mpros_map_set_occupancy_banded(mpros_ctx *ctx, mpros_comm *comm, const int8_t *rows, int width, int height,
                                             double resolution, double origin_x, double origin_y, double origin_yaw);
MPROS_API int mpros_map_set_occupancy_banded_dev(mpros_ctx *ctx, mpros_comm *comm, const int8_t *d_rows, int width, int height,
                                                 double resolution, double origin_x, double origin_y, double origin_yaw);

/* mpros_plan_batch for a GLOBAL set of nq queries sharded across the ranks (contiguous blocks, mpros_plan_shard_range):
 * every rank passes the same starts3 / goals3 and plans its block on its own replica of the map layers - no
 * communication during the search - then status / cost / path_len / stats of all blocks are gathered into every rank's
 * arrays (the only exchange; the paths too when gather_paths != 0, otherwise only the rank's own block of `paths` is
 * written). Arrays are sized for all nq queries on every rank. comm == NULL plans everything locally. Collective. */
MPROS_API int mpros_plan_shard_range(const mpros_comm *comm, size_t nq, size_t *lo, size_t *hi);
MPROS_API int mpros_plan_batch_sharded(mpros_ctx *ctx, mpros_comm *comm, const double *starts3, const double *goals3, size_t nq,
                                       int32_t *status, double *cost, int32_t *path_len, double *paths, int path_cap, int64_t *stats,
                                       int gather_paths);

/* Dynamic sharding over peer memory: ALL ranks draw query tickets from ONE counter in rank 0's memory (system-scope atomics
 * through the NVLink mapping: CUDA IPC between processes, the pointer itself between ranks of one process) and store every
 * query's results - status, cost, path length, statistics, path - straight into rank 0's arrays. The planner kernel balances
 * the load (a query that runs into the iteration limit costs ~50 mean queries; with fixed blocks the rank that holds more of
 * them finishes last) and performs the result gather itself; the communicator only carries the IPC handle once and three
 * 4-byte barriers per batch. mpros_plan_share_create is collective (rank 0 allocates the arrays for max_queries queries and
 * path_cap points each); mpros_plan_batch_shared is collective: every rank passes the same starts3 / goals3; result pointers
 * may be NULL on ranks that do not want them (non-NULL ones are filled from rank 0's arrays on any rank). Results per query
 * are identical to mpros_plan_batch. */
typedef struct mpros_plan_share mpros_plan_share;
MPROS_API int mpros_plan_share_create(mpros_ctx *ctx, mpros_comm *comm, size_t max_queries, int path_cap, mpros_plan_share **out);
MPROS_API int mpros_plan_share_destroy(mpros_ctx *ctx, mpros_plan_share *share);
MPROS_API int mpros_plan_batch_shared(mpros_ctx *ctx, mpros_plan_share *share, const double *starts3, const double *goals3, size_t nq,
                                      int32_t *status, double *cost, int32_t *path_len, double *paths, int path_cap, int64_t *stats);

/* ---- path post-processing -------------------------------------------------------------------- */
/* What HybridAstar::setPath does to the raw path_ before getNewPath() returns it (hybrid_a_star.cpp:814-838):
 * updatePathLength, setCusp, updateCurvatureNew (planner_utils.h:35-61, 251-309) and, when interpolation_enable is set,
 * upsampleAndPopulatePath (hybrid_a_star.cpp:902-1112: collision-checked cubic Bezier arcs) followed by
 * updateCurvatureNew + updatePathLength again. For a batch of raw paths: path q owns points [offsets[q], offsets[q+1])
 * of paths5 = double[5 * offsets[n_paths]] {x, y, yaw, steer, direction} (getPath()'s rows, e.g. mpros_plan_batch's).
 * out8: double[8 * out_cap_points] {x, y, yaw, curv_, direc_, steering_, is_cusp_, dist_} = the PathPoints of
 * getNewPath(), concatenated in path order; out_offsets: int64[n_paths + 1]. When out_cap_points is too small the call
 * returns MPROS_ERR_CAPACITY with out_offsets holding the required sizes. All candidate arcs of one segment of every path
 * are footprint-tested in ONE launch of the batched pathInCollision kernel. */
MPROS_API int mpros_path_postprocess_batch(mpros_ctx *ctx, const double *paths5, const int64_t *offsets, size_t n_paths, double *out8,
                                           size_t out_cap_points, int64_t *out_offsets);

#ifdef __cplusplus
}
#endif
#endif /* MPROS_GPU_H */
