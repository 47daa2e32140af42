// Internal definitions shared by the CUDA translation units of libmpros_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <stdint.h>
#include <string>
#include <vector>

#include "mpros_gpu.h"

namespace mpros
{

constexpr int kHugeInt = 1000; // HUGE_INT, mpros_navigation_map/node.h:4
// probe layout: words of a 16 x 16-cell block share a sector (1) or plain row-major tile words (0); 1 measured +5 %
#ifndef MPROS_TW_BLOCK16
#define MPROS_TW_BLOCK16 1
#endif
constexpr int kTwMargin = 256;  // cells of "outside the map" padding around the tile-word collision layers

// ---- HBM layout of one map (all device pointers) -------------------------------------------------------
// Occupancy layers are BIT-PACKED: bit (j & 31) of word [i * pitch + (j >> 5)] is cell (i, j); bits past
// the row end are zero. 8192^2 -> 8 MiB per layer, so the two collision layers stay resident in the
// 126 MB L2 together with everything else. Row pitch is a multiple of 4 words (16 B) so rows can be
// moved with 128-bit accesses. Byte/int layers in the reference's types exist only transiently, when a
// caller asks for them through mpros_map_download.
struct MapView
{
    const uint32_t *raw_bits;  // static_map_orig_      [H0 x pitch0]
    const uint32_t *infl_bits; // inflate_map_orig_     [H0 x pitch0]
    const uint32_t *ds_bits;   // static_map_           [H  x pitch ]
    // the two collision layers again as padded TILE WORDS for the fixed-point footprint walk (map_device.cuh): cell
    // (i, j) is bit ((pi & 3) << 3) | (pj & 7) of a 32-bit word holding 4 rows x 8 columns, (pi, pj) = (i, j) + kTwMargin;
    // the eight words of a 16 x 16-cell block (4 word rows x 2 word columns) are contiguous = one 32-byte sector:
    // word index (((pi >> 4) * tw_pitch + (pj >> 4)) << 3) | ((pi >> 1) & 6) | ((pj >> 3) & 1), tw_pitch = blocks per row
    const uint32_t *raw_tw, *infl_tw;
    int tw_pitch;
    const float *voronoi;      // voronoi_value_map_    [H*W]
    const uint16_t *goal;      // goal_cost_map_ as u16 [H*W] (single-goal context map)
    int H0, W0, pitch0;
    int H, W, pitch;
    double ox, oy, osin, ocos; // map_origin_x_/y_/sin_/cos_
    double res0, res;          // map_resolution_orig_, map_resolution_
    // fixed-point fast path of the footprint test (map_device.cuh: footprint_in_collision4): RN(1 / res0) and the
    // largest |x| + |y| of a pose for which the error bound of that path holds (2^22 cells minus the origin offset)
    double inv_res0, fx_span;
};

// ---- per-planner constants (HybridAstar::config / setFootprint, hybrid_a_star.cpp:120-206) -----------
constexpr int kMaxSteer = 16; // steer_set_ entries
struct PlannerView
{
    // footprint_longi_ x-coordinates (y = 0) and footprint_m_ corners, hybrid_a_star.cpp:193-199
    double longi_x0, longi_x1;
    double corner_x[4], corner_y[4];
    int n_check; // num_checking_step = ceil((L - W) / res_orig), hybrid_a_star.cpp:633
    // fixed-point fast path: 1 / n_check and the footprint's reach (subtracted from MapView::fx_span); fx_extent is
    // +inf when the path may not be used (n_check > 512)
    double inv_n_check, fx_extent;
    int fx_slack; // cells an axis end point must keep from the border of the padded layers (covers the corners)
    // expansion
    int n_steer; // steer_set_.size()
    double steer_set[kMaxSteer];
    double step_size, arc_radius_num; // arc radius numerator: wheelbase/2 + offset (hybrid_a_star.cpp:339)
    // per primitive (index 2 * steer_index + direction_index), evaluated ONCE on the host with the same libm the
    // reference calls for every expansion (hybrid_a_star.cpp:325-348): displacement in the vehicle frame and yaw change
    double prim_dx[2 * kMaxSteer], prim_dy[2 * kMaxSteer], prim_dyaw[2 * kMaxSteer];
    // the same for one leaf of getSearchTree (hybrid_a_star.cpp:1114-1164): path_granularity_ instead of step_size_
    double leaf_dx[2 * kMaxSteer], leaf_dy[2 * kMaxSteer], leaf_dyaw[2 * kMaxSteer];
    int leaf_num; // ceil(step_size_ / path_granularity_)
    double pen_gear, pen_back, pen_steer, pen_steer_change;
    double min_r, max_steer, rs_range;
    int num_yaw;
    double yaw_resol;
    int max_iteration;
};

// Device view of a workspace pool (plan_team.cuh: pool_acquire / pool_release)
struct WorkspacePool
{
    char *base;    // n x stride bytes
    int *owner;    // n flags: 0 free, 1 taken
    uint32_t *gen; // n generation counters (< 2^30; the host clears the pool before they could wrap)
    int n;
    // solo pool only: workspaces that travel with a handed-over query are "in custody" of the team pass until it has copied
    // the state out. At most custody_max at a time, and the pool holds custody_max workspaces more than the device has
    // resident warps, so a warp that needs a workspace always finds one (plan_squad.cuh).
    int *custody;
    int custody_max;
};
// Per-query planner workspaces of one pass (node-pool size class): node pools, open lists, closed sets and goal-wavefront
// windows, one per resident team of the DEVICE, shared by every lane of the context (plan.cu). Owned by the context, freed
// by mpros_ctx_destroy / mpros_plan_release_workspaces.
struct PlanPool
{
    char *base = nullptr;
    size_t bytes = 0, stride = 0;
    int n = 0;
    int *owner = nullptr;
    uint32_t *gen = nullptr;
    int meta_cap = 0;                    // entries of owner / gen
    unsigned long long issued = 0;       // upper bound of the generations handed out since the last clear
    unsigned char last_layout[128] = {}; // WarpWorkspaceLayout the bytes were last interpreted with (stale tags could alias)
};
// Work distribution of one (lane, pass): ticket counter and the list of queries to re-run
struct PlanScratch
{
    unsigned int *counter = nullptr; // 64 words: [0] ticket counter of the team pass, [16] of the solo pass, [32] length of the hand-off list
    int32_t *todo = nullptr;
    size_t todo_cap = 0;
    int32_t *handoff = nullptr; // queries the solo pass hands to the team pass (device-side list, plan_solo.cuh)
    size_t handoff_cap = 0;
    int32_t *mig = nullptr;     // migration list of the squad pass under the throughput policy (plan_squad.cuh)
    size_t mig_cap = 0;
};
constexpr int kPlanPasses = 3; // node-pool size classes

// Developer / tuning switches. Defaults come from the MPROS_* environment, read ONCE in mpros_ctx_create;
// mpros_ctx_set_option changes them per context afterwards.
struct Options
{
    bool c1_walk = false;      // MPROS_C1_MODE=walk: single-phase footprint kernel
    bool c1_ldg = false;       // MPROS_C1_MODE=2phase_ldg: the round-1 two-phase kernel (two probes per pose, byte stores)
    int c1_stages = -1;        // MPROS_C1_MODE=2phase_s0 / _s1 / _s2: shared-memory stages of the pose stream (-1 = default)
    int goal_mode = 0;         // MPROS_GOAL_MODE: 0 cluster (16 CTAs when granted), 1 "cluster8", 2 "grid"
    bool plan_cluster = false; // MPROS_PLAN_MODE=cluster: a team is a CTA pair
    bool plan_duo = false;     // MPROS_PLAN_MODE=duo: two teams per CTA keeping their iterations in step (plan_team.cuh: DuoSync)
    bool plan_solo = true;     // default (MPROS_PLAN_MODE unset or "solo"): every query starts in the warp-per-query kernel, the team
                               // kernel takes over the ones that outgrow its workspaces; "team" = team kernel only
    int solo_nodes = 32768;    // MPROS_SOLO_NODES: node-pool size of a solo workspace
    int solo_min_queries = 8192; // MPROS_SOLO_MIN_QUERIES: smaller batches go straight to the team kernel (a few thousand queries do
                                 // not fill the squads; the team kernel alone finishes them sooner: 2 048 queries 0.74 s against 0.95 s)
    int plan_policy = 0;       // MPROS_PLAN_POLICY: 0 auto (throughput when another batch is in flight at submission), 1 latency, 2 throughput
    int plan_age_limit = 30000; // MPROS_PLAN_AGE_LIMIT: throughput policy, iterations after which a query goes to the team pass for good
    bool plan_grow = true;     // MPROS_PLAN_GROW=0: queries that outgrow a small workspace go to the team pass at once
    bool plan_nosquad = false; // MPROS_PLAN_MODE=solo_warp: independent warps (plan_solo_kernel) instead of squads (plan_squad_kernel)
    bool debug = false;        // MPROS_DEBUG: per-pass timing (and the phase profile of the -DMPROS_PROFILE build) on stderr
};

struct Ctx
{
    int device = 0;
    int n_sm = 0; // cudaDevAttrMultiProcessorCount, queried once in mpros_ctx_create; every grid is a multiple of it
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr; // second stream of the host-buffer entry points that pipeline their copies (created on first use)
    cudaEvent_t aux_event = nullptr;
    uint64_t launches = 0;
    Options opt;
    PlanScratch plan_scratch[MPROS_PLAN_LANES][kPlanPasses];
    PlanPool plan_pool[kPlanPasses];
    PlanPool solo_pool;    // small workspaces of the warp-per-query pass, one per resident warp of the device
    PlanPool mid_pool;     // grown workspaces of the squad pass (four times the nodes), eight per SM
    int solo_resident = 0, squad_resident = 0; // resident CTAs per SM of the two solo-pass kernels (occupancy query, cached)
    int plan_resident[3] = {0, 0, 0}; // resident teams per SM: [0] CTA teams, [1] CTA-pair teams (occupancy query, cached), [2] two teams per CTA

    mpros_map_params mp;
    mpros_planner_params pp;
    bool planner_configured = false;

    // NavMap public state
    bool get_map = false, get_goal = false;
    // row-band session (mpros_map_band_begin .. mpros_map_band_finish)
    bool band_open = false;
    int band_have_old = 0, band_smax = -1;
    int H0 = 0, W0 = 0, pitch0 = 0, H = 0, W = 0, pitch = 0;
    double res0 = 0, res = 0, ox = 0, oy = 0, oyaw = 0, osin = 0, ocos = 1;
    double goal_x = 0, goal_y = 0;
    int num_regions = 0;

    // device buffers
    uint32_t *raw_bits = nullptr, *infl_bits = nullptr, *ds_bits = nullptr;
    uint32_t *raw_tw = nullptr, *infl_tw = nullptr; // derived from raw_bits / infl_bits (build_tilewords_kernel)
    int tw_pitch = 0;
    size_t cap_tw = 0;
    // "not surely free" layer of the batched footprint test (collision.cu), same tile-word layout; built on demand from the
    // two collision layers and the planner's footprint (map_build_unsafe), invalidated when either changes
    uint32_t *unsafe_tw = nullptr;
    uint2 *cls_tw = nullptr; // {unsafe word, inflated word} interleaved: one 8-byte probe classifies a pose (collision.cu)
    size_t cap_unsafe_tw = 0;
    bool unsafe_valid = false;
    bool c1_attr_set = false; // dynamic shared-memory limit of collision_poses_tma_kernel raised
    int unsafe_center_probe = 0; // the axis line has a probe exactly at the pose (n_check even, symmetric axis)
    uint16_t *goal = nullptr, *obs_dist = nullptr, *edge_dist = nullptr;
    int32_t *region = nullptr;
    uint32_t *edge_bits = nullptr;
    float *voronoi = nullptr;
    size_t cap0_words = 0, cap_words = 0, cap_cells = 0;
    // generic scratch
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    void *stage = nullptr; // device staging for host-pointer entry points
    size_t stage_bytes = 0;
    void *dist_buf = nullptr; // per-query arrays of ALL ranks' queries (mpros_plan_batch_sharded, dist.cu)
    size_t dist_bytes = 0;
    // plan lanes 1 .. MPROS_PLAN_LANES-1 (batches in flight beside the context's own stream, plan.cu); created on first use
    cudaStream_t lane_stream[MPROS_PLAN_LANES] = {};
    cudaEvent_t lane_event[MPROS_PLAN_LANES] = {};
    void *lane_stage[MPROS_PLAN_LANES] = {};
    size_t lane_stage_bytes[MPROS_PLAN_LANES] = {};

    PlannerView pv;

    MapView view() const
    {
        MapView v;
        v.raw_bits = raw_bits;
        v.infl_bits = infl_bits;
        v.ds_bits = ds_bits;
        v.raw_tw = raw_tw;
        v.infl_tw = infl_tw;
        v.tw_pitch = tw_pitch;
        v.voronoi = voronoi;
        v.goal = goal;
        v.H0 = H0;
        v.W0 = W0;
        v.pitch0 = pitch0;
        v.H = H;
        v.W = W;
        v.pitch = pitch;
        v.ox = ox;
        v.oy = oy;
        v.osin = osin;
        v.ocos = ocos;
        v.res0 = res0;
        v.res = res;
        v.inv_res0 = 1.0 / res0;
        v.fx_span = 4194304.0 * res0 - (std::fabs(ox) + std::fabs(oy));
        if ((unsigned long long)(H0 + 2 * kTwMargin) * (unsigned long long)(W0 + 2 * kTwMargin) >= (1ull << 36)) // 32-bit word index
            v.fx_span = -1.0;
        return v;
    }
};

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
int ensure_scratch(Ctx *c, size_t bytes);
int ensure_stage(Ctx *c, size_t bytes);

#define MPROS_CUDA(expr)                                                   \
    do                                                                     \
    {                                                                      \
        cudaError_t e_ = (expr);                                           \
        if (e_ != cudaSuccess)                                             \
            return ::mpros::cuda_fail(e_, #expr, __FILE__, __LINE__);      \
    } while (0)

#define MPROS_LAUNCH_CHECK(c)                                              \
    do                                                                     \
    {                                                                      \
        (c)->launches++;                                                   \
        cudaError_t e_ = cudaGetLastError();                               \
        if (e_ != cudaSuccess)                                             \
            return ::mpros::cuda_fail(e_, "kernel launch", __FILE__, __LINE__); \
    } while (0)

inline int pitch_words(int width) { return ((width + 31) / 32 + 3) & ~3; }
inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
// grid of a grid-stride kernel: enough CTAs for the work, capped at 16 resident CTAs on every SM of the device
inline int grid_for(const Ctx *c, size_t work_items, int block = 256)
{
    size_t g = (work_items + block - 1) / block;
    const size_t cap = (size_t)(c->n_sm > 0 ? c->n_sm : 1) * 16;
    if (g > cap)
        g = cap;
    return g < 1 ? 1 : (int)g;
}

// stage entry points implemented in the other translation units
int map_build_voronoi(Ctx *c);          // N5-N9, voronoi.cu
int map_build_goal(Ctx *c);             // N4, goal_map.cu
int map_build_unsafe(Ctx *c);           // pre-classification layer of the batched footprint test, map_preprocess.cu

} // namespace mpros

struct mpros_ctx : public mpros::Ctx
{
};
