// Context lifetime, error reporting and configuration (NavMap::config, HybridAstar::config/setFootprint).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "mpros_internal.cuh"

namespace mpros
{
static thread_local std::string g_last_error = "";

void set_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    char buf[512];
    std::snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    set_error(buf);
    return MPROS_ERR_CUDA;
}

static int grow(void **p, size_t *cap, size_t bytes)
{
    if (bytes <= *cap)
        return MPROS_OK;
    if (*p)
        MPROS_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    MPROS_CUDA(cudaMalloc(p, want));
    *cap = want;
    return MPROS_OK;
}
int ensure_scratch(Ctx *c, size_t bytes) { return grow(&c->scratch, &c->scratch_bytes, bytes); }
int ensure_stage(Ctx *c, size_t bytes) { return grow(&c->stage, &c->stage_bytes, bytes); }

// one switch of Options; value == nullptr or "" restores the default
static bool apply_option(Options &o, const char *key, const char *value)
{
    const std::string k = key ? key : "", v = value ? value : "";
    if (k == "c1_mode")
    {
        if (v != "" && v != "2phase" && v != "walk" && v != "2phase_ldg" && v != "2phase_s0" && v != "2phase_s1" && v != "2phase_s2")
            return false;
        o.c1_walk = v == "walk";
        o.c1_ldg = v == "2phase_ldg";
        o.c1_stages = v == "2phase_s0" ? 0 : v == "2phase_s1" ? 1 : v == "2phase_s2" ? 2 : -1;
    }
    else if (k == "goal_mode")
    {
        if (v != "" && v != "cluster" && v != "cluster8" && v != "grid")
            return false;
        o.goal_mode = v == "grid" ? 2 : v == "cluster8" ? 1 : 0;
    }
    else if (k == "plan_mode")
    {
        if (v != "" && v != "team" && v != "cluster" && v != "duo" && v != "solo" && v != "solo_warp")
            return false;
        o.plan_cluster = v == "cluster";
        o.plan_duo = v == "duo";
        o.plan_solo = v == "" || v == "solo" || v == "solo_warp";
        o.plan_nosquad = v == "solo_warp";
    }
    else if (k == "solo_min_queries")
    {
        const long n = v == "" ? 8192 : std::strtol(v.c_str(), nullptr, 10);
        if (n < 0 || n > (1l << 30))
            return false;
        o.solo_min_queries = (int)n;
    }
    else if (k == "plan_policy")
    {
        if (v != "" && v != "auto" && v != "latency" && v != "throughput")
            return false;
        o.plan_policy = v == "latency" ? 1 : v == "throughput" ? 2 : 0;
    }
    else if (k == "plan_age_limit")
    {
        const long n = v == "" ? 30000 : std::strtol(v.c_str(), nullptr, 10);
        if (n < 64 || n > (1l << 30))
            return false;
        o.plan_age_limit = (int)n;
    }
    else if (k == "plan_grow")
        o.plan_grow = !(v == "0");
    else if (k == "solo_nodes")
    {
        const long n = v == "" ? 32768 : std::strtol(v.c_str(), nullptr, 10);
        if (n < 64 || n > (1l << 24))
            return false;
        o.solo_nodes = (int)n;
    }
    else if (k == "debug")
        o.debug = !(v == "" || v == "0");
    else
        return false;
    return true;
}

static void free_plan_scratch(Ctx *c)
{
    for (int l = 0; l < MPROS_PLAN_LANES; ++l)
        for (int k = 0; k < kPlanPasses; ++k)
        {
            PlanScratch &S = c->plan_scratch[l][k];
            if (S.counter)
                cudaFree(S.counter);
            if (S.todo)
                cudaFree(S.todo);
            if (S.handoff)
                cudaFree(S.handoff);
            if (S.mig)
                cudaFree(S.mig);
            S = PlanScratch();
        }
    for (int k = 0; k < kPlanPasses; ++k)
    {
        PlanPool &P = c->plan_pool[k];
        if (P.base)
            cudaFree(P.base);
        if (P.owner)
            cudaFree(P.owner);
        if (P.gen)
            cudaFree(P.gen);
        P = PlanPool();
    }
    for (PlanPool *pp : {&c->solo_pool, &c->mid_pool})
    {
        PlanPool &P = *pp;
        if (P.base)
            cudaFree(P.base);
        if (P.owner)
            cudaFree(P.owner);
        if (P.gen)
            cudaFree(P.gen);
        P = PlanPool();
    }
}
} // namespace mpros

using namespace mpros;

extern "C" const char *mpros_last_error_string(void) { return g_last_error.c_str(); }

extern "C" int mpros_ctx_create(int device, mpros_ctx **out)
{
    if (!out)
    {
        set_error("mpros_ctx_create: out is NULL");
        return MPROS_ERR_INVALID;
    }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
    {
        // no CPU fallback: the product path requires a CUDA device
        cuda_fail(e == cudaSuccess ? cudaErrorNoDevice : e, "cudaGetDeviceCount (no CUDA device, no CPU fallback)", __FILE__,
                  __LINE__);
        return MPROS_ERR_CUDA;
    }
    if (device < 0 || device >= n)
    {
        set_error("mpros_ctx_create: device index out of range");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(device));
    mpros_ctx *c = new (std::nothrow) mpros_ctx();
    if (!c)
    {
        set_error("out of host memory");
        return MPROS_ERR_INVALID;
    }
    c->device = device;
    cudaError_t ae = cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (ae != cudaSuccess || c->n_sm < 1)
    {
        delete c;
        return cuda_fail(ae == cudaSuccess ? cudaErrorInvalidDevice : ae, "cudaDevAttrMultiProcessorCount", __FILE__, __LINE__);
    }
    // the MPROS_* environment is read here and nowhere else; unknown values keep the default
    apply_option(c->opt, "c1_mode", std::getenv("MPROS_C1_MODE"));
    apply_option(c->opt, "goal_mode", std::getenv("MPROS_GOAL_MODE"));
    apply_option(c->opt, "plan_mode", std::getenv("MPROS_PLAN_MODE"));
    apply_option(c->opt, "solo_nodes", std::getenv("MPROS_SOLO_NODES"));
    apply_option(c->opt, "plan_grow", std::getenv("MPROS_PLAN_GROW"));
    apply_option(c->opt, "plan_policy", std::getenv("MPROS_PLAN_POLICY"));
    apply_option(c->opt, "plan_age_limit", std::getenv("MPROS_PLAN_AGE_LIMIT"));
    apply_option(c->opt, "solo_min_queries", std::getenv("MPROS_SOLO_MIN_QUERIES"));
    apply_option(c->opt, "debug", std::getenv("MPROS_DEBUG"));
    mpros_map_params_default(&c->mp);
    mpros_planner_params_default(&c->pp);
    std::memset(&c->pv, 0, sizeof(c->pv));
    cudaError_t se = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess)
    {
        delete c;
        return cuda_fail(se, "cudaStreamCreate", __FILE__, __LINE__);
    }
    *out = c;
    return MPROS_OK;
}

extern "C" int mpros_ctx_destroy(mpros_ctx *c)
{
    if (!c)
        return MPROS_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int l = 1; l < MPROS_PLAN_LANES; ++l)
        if (c->lane_stream[l])
        {
            cudaStreamSynchronize(c->lane_stream[l]);
            cudaStreamDestroy(c->lane_stream[l]);
            cudaEventDestroy(c->lane_event[l]);
            if (c->lane_stage[l])
                cudaFree(c->lane_stage[l]);
        }
    if (c->aux_stream)
    {
        cudaStreamSynchronize(c->aux_stream);
        cudaStreamDestroy(c->aux_stream);
        cudaEventDestroy(c->aux_event);
    }
    free_plan_scratch(c); // node pools, open lists, closed sets of every lane (tens of GB at full size)
    void *bufs[] = {c->raw_bits, c->infl_bits, c->ds_bits, c->goal, c->obs_dist, c->edge_dist, c->region,
                    c->edge_bits, c->voronoi, c->scratch, c->stage, c->dist_buf, c->raw_tw, c->infl_tw, c->unsafe_tw, c->cls_tw};
    for (void *b : bufs)
        if (b)
            cudaFree(b);
    cudaStreamDestroy(c->stream);
    delete c;
    return MPROS_OK;
}

extern "C" int mpros_ctx_set_option(mpros_ctx *c, const char *key, const char *value)
{
    if (!c || !key)
    {
        set_error("mpros_ctx_set_option: NULL argument");
        return MPROS_ERR_INVALID;
    }
    if (!apply_option(c->opt, key, value))
    {
        set_error(std::string("mpros_ctx_set_option: unknown option or value: ") + key + "=" + (value ? value : ""));
        return MPROS_ERR_INVALID;
    }
    return MPROS_OK;
}

extern "C" int mpros_plan_release_workspaces(mpros_ctx *c)
{
    if (!c)
    {
        set_error("mpros_plan_release_workspaces: NULL context");
        return MPROS_ERR_INVALID;
    }
    MPROS_CUDA(cudaSetDevice(c->device));
    MPROS_CUDA(cudaStreamSynchronize(c->stream));
    for (int l = 1; l < MPROS_PLAN_LANES; ++l)
        if (c->lane_stream[l])
            MPROS_CUDA(cudaStreamSynchronize(c->lane_stream[l]));
    free_plan_scratch(c);
    return MPROS_OK;
}

extern "C" void *mpros_ctx_stream(mpros_ctx *c) { return c ? (void *)c->stream : nullptr; }

extern "C" int mpros_ctx_synchronize(mpros_ctx *c)
{
    if (!c)
        return MPROS_ERR_INVALID;
    MPROS_CUDA(cudaStreamSynchronize(c->stream));
    return MPROS_OK;
}

extern "C" uint64_t mpros_ctx_launch_count(mpros_ctx *c) { return c ? c->launches : 0; }

// nav_map.cpp:27-34
extern "C" void mpros_map_params_default(mpros_map_params *p)
{
    p->occupied_thresh = 0.68;
    p->free_thresh = 0.2;
    p->voro_fallout_rate = 10.0f;
    p->voro_max_region = 50.0f;
    p->downsampling_factor = 1;
    p->obstacle_inflation_radius = 0.0;
}

// hybrid_a_star.cpp:128-161
extern "C" void mpros_planner_params_default(mpros_planner_params *p)
{
    p->max_curvature = 0.4;
    p->width = 2.0;
    p->wheelbase = 3.0;
    p->length = 4.2;
    p->steer_offset = 0.0;
    p->max_steering_angle = 0.35;
    p->gear_penalty = 10.0;
    p->reverse_penalty = 0.0;
    p->steering_penalty = 0.1;
    p->steering_change_penalty = 10.0;
    p->analytic_solution_range = 10.0;
    p->yaw_discretization = 36;
    p->step_size = 1.1;
    p->path_point_distance = 0.2;
    p->steering_discretization = 3;
    p->max_iteration = 10000;
    p->optimal_path = 0;
    p->interpolation_enable = 1;
}

extern "C" int mpros_map_config(mpros_ctx *c, const mpros_map_params *p)
{
    if (!c || !p)
    {
        set_error("mpros_map_config: NULL argument");
        return MPROS_ERR_INVALID;
    }
    if (p->downsampling_factor < 1)
    {
        set_error("mpros_map_config: downsampling_factor must be >= 1");
        return MPROS_ERR_INVALID;
    }
    c->mp = *p;
    return MPROS_OK;
}

// HybridAstar::config + setFootprint (hybrid_a_star.cpp:120-206)
extern "C" int mpros_planner_config(mpros_ctx *c, const mpros_planner_params *p)
{
    if (!c || !p)
    {
        set_error("mpros_planner_config: NULL argument");
        return MPROS_ERR_INVALID;
    }
    if (!c->get_map)
    {
        // setFootprint reads map_->map_resolution_orig_ (hybrid_a_star.cpp:204)
        set_error("mpros_planner_config: no map set (setFootprint needs map_resolution_orig_)");
        return MPROS_ERR_INVALID;
    }
    int num_pos_steer = p->steering_discretization / 2;
    if (2 * num_pos_steer + 1 > kMaxSteer || p->yaw_discretization < 1)
    {
        set_error("mpros_planner_config: steering_discretization / yaw_discretization out of range");
        return MPROS_ERR_INVALID;
    }
    // Everything is validated and built in a local PlannerView; the context is only touched on success, so a failed call
    // leaves the previous configuration (or "not configured") intact.
    auto finite_pos = [](double v) { return std::isfinite(v) && v > 0; };
    if (!finite_pos(p->step_size) || !finite_pos(p->path_point_distance) || !finite_pos(p->max_curvature) ||
        !finite_pos(p->max_steering_angle) || !finite_pos(p->length) || !finite_pos(p->width) || !std::isfinite(p->wheelbase) ||
        !std::isfinite(p->steer_offset) || !std::isfinite(p->analytic_solution_range) || p->max_iteration < 0 ||
        !std::isfinite(p->gear_penalty) || !std::isfinite(p->reverse_penalty) || !std::isfinite(p->steering_penalty) ||
        !std::isfinite(p->steering_change_penalty))
    {
        set_error("mpros_planner_config: step_size, path_point_distance, max_curvature, max_steering_angle, length and width must "
                  "be finite and positive, the other parameters finite");
        return MPROS_ERR_INVALID;
    }
    if (!(p->step_size / p->path_point_distance < 1e6))
    {
        set_error("mpros_planner_config: step_size / path_point_distance too large");
        return MPROS_ERR_INVALID;
    }
    PlannerView v;
    std::memset(&v, 0, sizeof(v));
    const double L = p->length, Wd = p->width, off = p->steer_offset;
    // footprint_m_ corners, clockwise (hybrid_a_star.cpp:193-196)
    v.corner_x[0] = L / 2 + off;
    v.corner_y[0] = Wd / 2;
    v.corner_x[1] = L / 2 + off;
    v.corner_y[1] = -Wd / 2;
    v.corner_x[2] = -L / 2 + off;
    v.corner_y[2] = -Wd / 2;
    v.corner_x[3] = -L / 2 + off;
    v.corner_y[3] = Wd / 2;
    // footprint_longi_ (hybrid_a_star.cpp:198-199)
    v.longi_x0 = L / 2 + off - 1.0 / 2 * Wd;
    v.longi_x1 = -L / 2 + off + 1.0 / 2 * Wd;
    // num_checking_step (hybrid_a_star.cpp:632-633)
    double footprint_len = L - Wd;
    const double n_check_d = std::ceil(footprint_len / c->res0);
    if (!(n_check_d >= 1) || !(n_check_d < 1e8))
    {
        set_error("mpros_planner_config: length - width must be positive (num_checking_step >= 1)");
        return MPROS_ERR_INVALID;
    }
    v.n_check = (int)n_check_d;
    v.inv_n_check = 1.0 / (double)v.n_check;
    v.fx_extent = v.n_check <= 512 ? std::fabs(L) + std::fabs(Wd) + std::fabs(off) : INFINITY;
    // corners lie within |W| (> sqrt(2) * W / 2) of an axis end point
    v.fx_slack = (int)std::ceil(std::fabs(Wd) / c->res0) + 2;
    if (!(v.fx_slack < kTwMargin - 2) || !(Wd > 0) || !(L > 0))
        v.fx_extent = INFINITY;
    // steer_set_ (hybrid_a_star.cpp:176-184)
    v.n_steer = 0;
    for (int i = 1; i <= num_pos_steer; ++i)
    {
        v.steer_set[v.n_steer++] = p->max_steering_angle / num_pos_steer * i;
        v.steer_set[v.n_steer++] = -p->max_steering_angle / num_pos_steer * i;
    }
    v.steer_set[v.n_steer++] = 0.0;
    v.step_size = p->step_size;
    v.arc_radius_num = p->wheelbase / 2 + off;
    // motion primitives, hybrid_a_star.cpp:320-350 (steer outer loop, direction {+1, -1} inner loop)
    for (int si = 0; si < v.n_steer; ++si)
        for (int di = 0; di < 2; ++di)
        {
            const double steer = v.steer_set[si], direc = di ? -1.0 : 1.0;
            double dx, dy, dyaw_signed;
            if (steer == 0)
            {
                dx = direc * p->step_size;
                dy = 0;
                dyaw_signed = 0;
            }
            else
            {
                double radius = (p->wheelbase / 2 + off) / std::tan(std::abs(steer));
                double steer_sign = steer / std::abs(steer);
                double dyaw = p->step_size / radius;
                dx = std::sin(dyaw) * radius * direc;
                dy = (1 - std::cos(dyaw)) * radius * steer_sign;
                dyaw_signed = dyaw * steer_sign * direc;
            }
            v.prim_dx[2 * si + di] = dx;
            v.prim_dy[2 * si + di] = dy;
            v.prim_dyaw[2 * si + di] = dyaw_signed;
            // one leaf of the search-tree visualisation (hybrid_a_star.cpp:1133-1153)
            if (steer == 0)
            {
                dx = direc * p->path_point_distance;
                dy = 0;
                dyaw_signed = 0;
            }
            else
            {
                double radius = (p->wheelbase / 2 + off) / std::tan(std::abs(steer));
                double steer_sign = steer / std::abs(steer);
                double dyaw = p->path_point_distance / radius;
                dx = std::sin(dyaw) * radius * direc;
                dy = (1 - std::cos(dyaw)) * radius * steer_sign;
                dyaw_signed = dyaw * steer_sign * direc;
            }
            v.leaf_dx[2 * si + di] = dx;
            v.leaf_dy[2 * si + di] = dy;
            v.leaf_dyaw[2 * si + di] = dyaw_signed;
        }
    v.leaf_num = (int)std::ceil(p->step_size / p->path_point_distance);
    v.pen_gear = p->gear_penalty;
    v.pen_back = p->reverse_penalty;
    v.pen_steer = p->steering_penalty;
    v.pen_steer_change = p->steering_change_penalty;
    v.min_r = 1.0 / p->max_curvature;
    v.max_steer = p->max_steering_angle;
    v.rs_range = p->analytic_solution_range;
    v.num_yaw = p->yaw_discretization;
    v.yaw_resol = 2 * M_PI / p->yaw_discretization;
    v.max_iteration = p->max_iteration;
    // commit
    c->pp = *p;
    c->pv = v;
    c->unsafe_valid = false; // depends on the footprint
    c->planner_configured = true;
    return MPROS_OK;
}
