#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native mpros hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload plan|collision|expand|preprocess|rs|arena]

Workload "plan" (default; BASELINE.json configs[4]): synthetic 8192x8192 occupancy grid @0.1 m
(seed 20261019: 1-m border wall + 1100 rectangles), inflation 1.0 m, down-sampling 8, shipped planner parameters
(3 steering angles, 36 yaw bins, step 2.0 m); 16384 independent seeded start/goal queries PER GPU (goal 20-80 m from
the start, both collision-free and connected). One STEP = one mpros_plan_batch over the rank's 16384 queries =
per query: obstacle-aware goal heuristic (on demand) + Hybrid A* with batched footprint checks + Reeds-Shepp shots.
metric = plans/s; the same run also reports node expansions/s, footprint checks/s inside planning and the standalone
batched collision kernel (2^24 poses). Multi-GPU: queries sharded across ranks, map replicated, the only exchange
is the gather of the per-query results (scaling "weak": 16384 queries per GPU).

Workload "collision": the C1 kernel alone, 2^24 seeded uniform poses per GPU per step.

One JSON line on stdout (rank 0). The oracle (oracle/) is used only as the CPU baseline / reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))

import ctypes

C_DOUBLE = ctypes.c_double
import synthetic  # noqa: E402  (workload generator: numpy only)

SYN_SIZE, SYN_RES, SYN_DS = synthetic.SYN_SIZE, synthetic.SYN_RES, synthetic.SYN_DS
EXTENT = SYN_SIZE * SYN_RES
QUERIES_PER_GPU = 16384
PATH_CAP = 128
POSES_PER_STEP = 1 << 24
POSE_SEED = 1
# algorithmic bytes per unit of work, SURVEY.md §8(d)
B_CHECK = 24 + 1 + (22 + 1 + 4)  # C1 @0.1 m: pose in + verdict out + (n+1+4) one-byte map cells
B_EXPANSION = 450                # E1 with the case-1 survivor / insert ratios
B_SHOT = 440                     # R4 without the trajectory
EXPANSIONS_PER_STEP = 1 << 20
RS_SHOTS_PER_STEP = 1 << 20
ARENA_QUERIES = 2048
ARENA_CPU_QUERIES = 4
B_GOAL_CELL = 5                  # N4: 1 B in, 4 B out per cell of the goal map


def shipped_params(ds=SYN_DS):
    """planner_config.yaml + vehicle_config.yaml as the shipped launch loads them (the YAML's misspelled
    steering_dsicretization leaves the code default of 3 steering angles)."""
    return {
        "/mpros/planner/optimal_path": 0, "/mpros/planner/interpolation_enable": 1, "/mpros/planner/step_size": 2.0,
        "/mpros/planner/max_iteration": 100000, "/mpros/planner/gear_penalty": 10.0, "/mpros/planner/reverse_penalty": 10.0,
        "/mpros/planner/steering_penalty": 0.5, "/mpros/planner/steering_change_penalty": 0.5,
        "/mpros/planner/analytic_solution_range": 10.0, "/mpros/planner/yaw_discretization": 36,
        "/mpros/planner/path_point_distance": 0.1, "/mpros/map/downsampling_factor": ds,
        "/mpros/map/obstacle_inflation_radius": 1.0, "/mpros/map/free_thresh": 0.2, "/mpros/map/occupied_thresh": 0.68,
        "/mpros/vehicle/max_steering_angle": 0.35, "/mpros/vehicle/max_curvature": 0.125, "/mpros/vehicle/width": 2.0,
        "/mpros/vehicle/wheelbase": 3.0, "/mpros/vehicle/length": 4.2,
        "/mpros/vehicle/center_to_rotate_center_longitudinal": 0.0,
    }


def ncu_traffic(kernel):
    """DRAM bytes (read + write) per launch of `kernel` from the newest committed ncu capture of this very command
    (profiles/rNN_traffic.json, written by tools/make_traffic.py from the `ncu --set full` captures); None if absent."""
    for tag in ("r02", "r01"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", tag + "_traffic.json")))[kernel]["dram_bytes_per_launch"]
        except Exception:
            continue
    return None


def plan_traffic():
    """DRAM bytes of one batch of the plan workload = the two launches of a batch (squad pass + team pass)."""
    a, b = ncu_traffic("plan_squad_kernel"), ncu_traffic("plan_team_kernel")
    return None if a is None or b is None else a + b


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu_index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        def num(v):
            try:
                return float(v)
            except ValueError:
                return None
        sm = [num(s[0]) for s in self.samples if s and num(s[0]) is not None]
        mx = [num(s[1]) for s in self.samples if len(s) > 1 and num(s[1]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 2 + k and s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def gen_poses(n, seed, extent=EXTENT):
    rng = np.random.default_rng(seed)
    p = np.empty((n, 3), dtype=np.float64)
    p[:, 0] = rng.uniform(0.0, extent, n)
    p[:, 1] = rng.uniform(0.0, extent, n)
    p[:, 2] = rng.uniform(-np.pi, np.pi, n)
    return p


def gen_queries(ctx, n, seed):
    """The §8(d) query recipe; footprint checks for the rejection sampling run on the GPU context."""
    static_ds = ctx.layer("static")
    return synthetic.synthetic_queries(n, ctx.collision_check, static_ds, SYN_RES * SYN_DS, EXTENT, seed)


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref) — one process per core, one NavMap copy each, because the
# reference is single-threaded and NavMap::updateGoal mutates shared state.
_W = {}


def _cpu_worker_init(so_path, params, shm_name, shape, res):
    from multiprocessing import shared_memory

    from oracle import refapi

    shm = shared_memory.SharedMemory(name=shm_name)
    occ = np.ndarray(shape, dtype=np.int8, buffer=shm.buf)
    lib = refapi.RefLib(so_path)
    rm = refapi.RefMap(lib, params)
    t0 = time.perf_counter()
    rm.set_occupancy(occ, res, [0.0, 0.0, 0.0])
    _W["t_pre"] = time.perf_counter() - t0
    _W["rp"] = refapi.RefPlanner(rm)
    _W["rm"] = rm
    shm.close()


def _cpu_worker_collision(poses):
    t0 = time.perf_counter()
    v = _W["rp"].collision4(poses)
    return time.perf_counter() - t0, int(v.sum()), len(poses)


def _cpu_worker_rs(job):
    from oracle import refapi

    st5, g3, cfg = job  # cfg: radius, steering angle, step, penalties as HybridAstar::config hands them to RSCurve
    rp, lib = _W["rp"], _W["rp"].lib
    t0 = time.perf_counter()
    for s, g in zip(st5, g3):
        seg, cost, traj = refapi.ref_rs_solve(lib, cfg, s, g)
        if len(traj):
            rp.collision4(traj[:, :3])
    dt = time.perf_counter() - t0
    z = np.zeros(3)
    t0 = time.perf_counter()
    for _ in range(2000):
        lib.L.ref_rs_distance(C_DOUBLE(8.0), z.ctypes.data, z.ctypes.data)
    overhead = (time.perf_counter() - t0) / 2000 * 2  # two library calls per shot
    return dt, len(st5), overhead


def _cpu_worker_pre(goal):
    t0 = time.perf_counter()
    _W["rm"].update_goal(goal[0], goal[1])
    return _W["t_pre"], time.perf_counter() - t0


def _cpu_worker_plan(job):
    from oracle import refapi

    starts, goals = job
    t0 = time.perf_counter()
    ok = prims = 0
    per_query = []
    for s, g in zip(starts, goals):
        r = refapi.ref_plan_query(_W["rm"], s, g)  # updateGoal + ctor (dense state) + Plan(), main_planner.cpp:225-229
        ok += int(r["plan_ok"])
        prims += int(r["n_primitives"])
        per_query.append((int(r["plan_ok"]), int(r["n_primitives"]), float(r["goal_g"]), int(r["status"])))
    return time.perf_counter() - t0, ok, prims, len(starts), _W["t_pre"], per_query


class CpuArm:
    def __init__(self, occ, params, cores, res=None):
        from multiprocessing import get_context, shared_memory

        from oracle import refapi

        if not os.path.exists(refapi.REF_SO):
            raise RuntimeError("oracle/_ref/libmpros_ref.so missing")
        self.cores = cores
        self.shm = shared_memory.SharedMemory(create=True, size=occ.nbytes)
        np.ndarray(occ.shape, dtype=np.int8, buffer=self.shm.buf)[:] = occ
        self.pool = get_context("spawn").Pool(cores, initializer=_cpu_worker_init,
                                              initargs=(refapi.REF_SO, params, self.shm.name, occ.shape, SYN_RES if res is None else res))

    def collision_rate(self, n_poses, seed, extent=None):
        poses = gen_poses(n_poses, seed) if extent is None else gen_poses(n_poses, seed, extent)
        res = self.pool.map(_cpu_worker_collision, np.array_split(poses, self.cores))
        # workers run concurrently: throughput = total checks / slowest worker's footPrintInCollision4 loop
        return sum(r[2] for r in res) / max(r[0] for r in res)

    def plan_rate(self, starts, goals):
        jobs = list(zip(np.array_split(starts, self.cores), np.array_split(goals, self.cores)))
        res = self.pool.map(_cpu_worker_plan, jobs)
        t = max(r[0] for r in res)
        n = sum(r[3] for r in res)
        return {"plans_per_sec": n / t, "expansions_per_sec": sum(r[2] for r in res) / 6.0 / t, "solved": sum(r[1] for r in res),
                "n": n, "preprocess_s": max(r[4] for r in res), "per_query": [q for r in res for q in r[5]]}

    def rs_rate(self, st5, g3, cfg):
        res = self.pool.map(_cpu_worker_rs, [(a, b, cfg) for a, b in zip(np.array_split(st5, self.cores), np.array_split(g3, self.cores))])
        return sum(r[1] for r in res) / max(r[0] for r in res), max(r[2] for r in res)

    def preprocess_time(self, gx, gy):
        res = self.pool.map(_cpu_worker_pre, [(gx, gy)] * self.cores)
        return max(r[0] for r in res), max(r[1] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()
        self.shm.close()
        self.shm.unlink()


def usable_cores(per_process_bytes=3.5e9):
    """(processes to run, nproc, processes the memory allows): all host cores, limited by memory - the reference allocates
    a dense H'*W'*36*2 state per query (~1.8 GB at 1024^2) on top of its map copy."""
    nproc = os.cpu_count() or 1
    by_mem = nproc
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
        by_mem = max(1, int(avail * 0.8 / per_process_bytes))
    except Exception:
        pass
    return max(1, min(nproc, by_mem)), nproc, by_mem


def effective_lanes(args, world):
    """Query batches in flight of the plan workload: --lanes, else 8 (the library's MPROS_PLAN_LANES; with other batches in flight
    the squad pass keeps long queries in squad form - its throughput policy - and a batch alone takes ~1.7 s, so eight batches
    cover each other's tails: 25.0 k plans/s against 23.1 k with six and 20.2 k with four)."""
    want = args.lanes if args.lanes > 0 else 8
    return max(1, min(want, 8))


def plan_config(lanes, scaling="weak", world=1):
    """config of the plan workload, identical for the GPU arm and the reference arm (the reference arm times a seeded
    sample of the same query set; what the sample was is stated in its cpu_baseline.sample)."""
    cfg = {"workload": workload_text("plan"), "path_cap": PATH_CAP, "batches_in_flight": lanes,
           "l2_policy": "per-query working sets (node pool, open list, closed set, goal wavefront) live in per-query HBM workspaces: 3 552 resident "
                        "warp-per-query searches x 8.8 MB, grown ones 27.9 MB, 296 team workspaces x 110 MB (>> L2 in aggregate); map layers are "
                        "L2-resident by design"}
    if scaling == "strong":
        cfg["queries_total"] = QUERIES_PER_GPU
    else:
        cfg["queries_per_gpu"] = QUERIES_PER_GPU
    return cfg


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    occ = synthetic.synthetic_occupancy()
    params = shipped_params()
    cores, nproc, by_mem = usable_cores()
    arm = CpuArm(occ, params, cores)
    vals, extra = [], {}
    t0 = time.perf_counter()
    try:
        if args.workload == "collision":
            n = (1 << 20) * cores
            for s in range(args.warmup + args.steps):
                v = arm.collision_rate(n, POSE_SEED + s)
                if s >= args.warmup:
                    vals.append(v)
                if time.perf_counter() - t0 > 200 and vals:
                    break
            metric, unit = "collision_checks_per_sec", "checks/s"
            sample = "%d seeded uniform poses per step on the synthetic 8192^2 map, %d processes x 1 thread; timed region = footPrintInCollision4 loop" % (n, cores)
            per_step = n
        else:
            # the query set is generated with the reference's own collision test so that this arm needs no GPU
            # at least 64 queries per process over the run (warm-up steps included)
            per_proc = max(2, -(-64 // max(1, args.steps + args.warmup)))
            lib_poses = _ref_queries(occ, params, per_proc * cores * max(1, args.steps + args.warmup))
            per_step = per_proc * cores
            for s in range(args.warmup + args.steps):
                sl = slice(s * per_step, (s + 1) * per_step)
                r = arm.plan_rate(lib_poses[0][sl], lib_poses[1][sl])
                if s >= args.warmup:
                    vals.append(r["expansions_per_sec"] if args.workload == "expand" else r["plans_per_sec"])
                    extra = r
                if time.perf_counter() - t0 > 420 and vals:
                    break
            metric, unit = ("node_expansions_per_sec", "expansions/s") if args.workload == "expand" else ("plans_per_sec", "plans/s")
            sample = ("%d seeded queries per step (%d per process, %d per process over the run) of the synthetic 8192^2 query set (ds=8), %d processes x 1 thread "
                      "(nproc %d, memory allows %d: 3.5 GB per process); timed region per query = NavMap::updateGoal + HybridAstar ctor + Plan() "
                      "(map preprocessing %.1f s excluded)" % (per_step, per_proc, per_proc * (args.steps + args.warmup), cores, nproc, by_mem,
                                                              extra.get("preprocess_s", 0)))
    finally:
        arm.close()
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": args.warmup, "ms_per_step": (1e3 * per_step / extra["plans_per_sec"]) if args.workload == "expand" else 1e3 * per_step / value, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": plan_config(effective_lanes(args, args.gpus), args.scaling) if args.workload == "plan" else {"workload": workload_text(args.workload)},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "nproc": nproc, "processes_memory_allows": by_mem, "kind": "reference",
                         "sample": sample, "units_per_step": per_step},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if extra:
        line["extra"] = {"node_expansions_per_sec": extra["expansions_per_sec"], "solved": extra["solved"], "queries": extra["n"]}
    print(json.dumps(line))
    return 0


def _ref_queries(occ, params, n):
    """§8(d) query recipe evaluated with the reference's own footprint test (CPU-only path for --impl reference)."""
    from oracle import refapi

    lib = refapi.RefLib()
    # collision checks need only the raw + inflated layers: down-sampling 64 keeps the Voronoi/CCL part of the
    # reference's preprocessing negligible here; the free-component test uses the ds=8 static layer built with numpy
    p2 = dict(params)
    rm = refapi.RefMap(lib, p2)
    rm.set_occupancy(occ, SYN_RES, [0.0, 0.0, 0.0])
    rp = refapi.RefPlanner(rm)
    static_ds = rm.layer("static")
    out = synthetic.synthetic_queries(n, rp.collision4, static_ds, SYN_RES * SYN_DS, EXTENT, synthetic.SYN_QUERY_SEED)
    rp.close()
    rm.close()
    return out


def verify_against_reference(ctx, occ, params, starts, goals, status, cost, stats, k):
    """SURVEY 9.4 on the headline query set, live: K evenly spaced queries plus every query the GPU did not solve, re-planned
    by the unmodified reference (oracle/_ref) on all host cores. Mode (ii): everything GPU-computed against the reference with
    its own Voronoi field -> distribution of the cost difference. Mode (i): the reference's Voronoi field injected into the
    GPU planner -> status, tree_.size() and goal cost must be equal per query."""
    from oracle import refapi

    nq = len(starts)
    idx = np.union1d(np.arange(0, nq, max(1, nq // max(1, k)))[:k], np.nonzero(status != 1)[0])
    cores, nproc, by_mem = usable_cores()
    arm = CpuArm(occ, params, cores)
    try:
        ref = arm.plan_rate(starts[idx], goals[idx])["per_query"]
    finally:
        arm.close()
    r_status = np.array([q[3] for q in ref])
    r_prims = np.array([q[1] for q in ref])
    r_cost = np.array([q[2] for q in ref])
    g_status, g_prims, g_cost = status[idx], stats[idx, 1], cost[idx]
    both = (g_status == 1) & (r_status == 1)
    d = (g_cost[both] - r_cost[both]) / r_cost[both]
    edges = [-1, -1e-1, -1e-2, -1e-3, -1e-4, 1e-4, 1e-3, 1e-2, 1e-1, 10]
    out = {"queries": int(len(idx)), "non_found_included": int((status != 1).sum()), "reference_processes": cores,
           "mode_ii_full_gpu": {"status_mismatches": int((g_status != r_status).sum()), "same_primitive_count": int((g_prims == r_prims).sum()),
                                "rel_cost_difference": {"median": float(np.median(d)), "mean": float(d.mean()), "min": float(d.min()), "max": float(d.max()),
                                                        "bin_edges": edges, "histogram": np.histogram(d, bins=edges)[0].tolist(),
                                                        "other_homotopy_class_gt_5pct": int((np.abs(d) > 0.05).sum())}}}
    # mode (i): the reference's order-dependent Voronoi field on the GPU
    rm = refapi.RefMap(refapi.RefLib(), params)
    rm.set_occupancy(occ, SYN_RES, [0.0, 0.0, 0.0])
    canon = ctx.layer("voronoi").copy()
    ref_v = rm.layer("voronoi")
    rm.close()
    try:
        ctx.upload_voronoi(ref_v)
        g = ctx.plan_batch(starts[idx], goals[idx], path_cap=PATH_CAP, truncate_ok=True)
    finally:
        ctx.upload_voronoi(canon)
    ok = (g["status"] == 1) & (r_status == 1)
    rel = np.abs(g["cost"][ok] - r_cost[ok]) / np.abs(r_cost[ok])
    out["mode_i_reference_voronoi_injected"] = {
        "status_mismatches": int((g["status"] != r_status).sum()), "primitive_count_mismatches": int((g["primitives"] != r_prims).sum()),
        "max_rel_cost_difference": float(rel.max()) if len(rel) else None, "cost_beyond_1e-9": int((rel > 1e-9).sum()),
        "voronoi_cells_differing_from_canonical": int((ref_v.view(np.uint32) != canon.view(np.uint32)).sum()),
        "mismatching_queries": [int(q) for q in idx[(g["status"] != r_status) | (g["primitives"] != r_prims)]][:16]}
    return out


def workload_text(w):
    if w == "collision":
        return "collision: synthetic 8192x8192 @0.1 m (seed 20261019), inflation 1.0 m, batched footPrintInCollision4 over seeded uniform poses"
    if w == "rs":
        return ("rs: synthetic 8192x8192 @0.1 m (seed 20261019), inflation 1.0 m, shipped planner config (rho 8 m, step 2 m); batched "
                "HybridAstar::genRSPath shots: RSCurve::FindOptimalPath (48 candidate words) + GetTraj + pathInCollision of the sampled "
                "path, over seeded collision-free start poses and goals within the 10 m analytic range")
    if w == "preprocess":
        return ("preprocess: synthetic 8192x8192 @0.1 m (seed 20261019), ds=8, inflation 1.0 m; NavMap::setOccupancyMap (threshold, inflation, "
                "down-sampling, obstacle labelling, L1 distance maps, Voronoi field) + NavMap::updateGoal (holonomic goal map) per step")
    if w == "expand":
        return ("expand: synthetic 8192x8192 @0.1 m (seed 20261019), ds=8, inflation 1.0 m, shipped planner config; batched "
                "HybridAstar::expandNode (6 primitives, footprint test, node cost, goal-map + Reeds-Shepp heuristic, closed-set prune) over "
                "seeded collision-free parent poses of one query")
    return ("plan: synthetic 8192x8192 @0.1 m (seed 20261019), ds=8, inflation 1.0 m, shipped planner config (3 steers, 36 yaw bins, step 2 m); "
            "independent seeded start/goal queries 20-80 m apart; per query goal heuristic + Hybrid A* + Reeds-Shepp shots")


def run_b200(args):
    import torch
    import torch.distributed as dist

    import mpros_b200

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    occ = synthetic.synthetic_occupancy()
    params = shipped_params()
    ctx = mpros_b200.Context(local_rank)
    ctx.map_config(params)
    occ_pinned = torch.from_numpy(occ).pin_memory()
    t0 = time.perf_counter()
    ctx.set_occupancy(occ_pinned.numpy(), SYN_RES, [0.0, 0.0, 0.0])
    t_pre = time.perf_counter() - t0
    ctx.planner_config(params)
    stream = torch.cuda.ExternalStream(ctx.stream())
    warm = max(3, args.warmup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_device(step_fn, join=None):
        """W warm-up steps, then K steps timed with CUDA events on the context's stream (per step + whole region).
        join(): makes the context's stream wait for the other streams the steps ran on (plan lanes)."""
        for _ in range(warm):
            step_fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(stream)
        for k in range(args.steps):
            ev[k][0].record(stream)
            step_fn()
            ev[k][1].record(stream)
        if join:
            join()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), [a.elapsed_time(b) for a, b in ev], ctx.launch_count() - l0

    def timed_host(step_fn, finish=None):
        for _ in range(2):
            step_fn()
        if finish:
            finish()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_fn()
        if finish:
            finish()
        barrier()
        return (time.perf_counter() - t0) * 1e3

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    # ---- collision kernel (always measured: headline for --workload collision, extra otherwise) ----
    n = POSES_PER_STEP
    col_steps = args.steps if args.workload == "collision" else min(args.steps, 5)
    poses_h = torch.from_numpy(gen_poses(n, POSE_SEED + 1000 * rank)).pin_memory()
    poses_d = poses_h.cuda()
    verd_d = torch.zeros(n, dtype=torch.uint8, device="cuda")
    verd_h = torch.zeros(n, dtype=torch.uint8).pin_memory()
    sampler = ClockSampler(local_rank)
    sampler.start()
    keep_steps = args.steps
    args.steps = col_steps
    col_total_ms, col_ms, col_launches = timed_device(lambda: ctx.collision_check_dev(poses_d.data_ptr(), n, verd_d.data_ptr()))
    col_e2e_ms = timed_host(lambda: ctx._check(ctx.L.mpros_collision_check_poses(ctx.h, poses_h.data_ptr(), n, verd_h.data_ptr())))
    args.steps = keep_steps
    assert torch.equal(verd_h, verd_d.cpu()), "host-API and device-API verdicts differ"
    hit_rate = float(verd_h.float().mean())
    col_total_ms, col_e2e_ms = max_over_ranks([col_total_ms, col_e2e_ms])
    col_value = world * n * col_steps / (col_total_ms * 1e-3)
    col_e2e = world * n * col_steps / (col_e2e_ms * 1e-3)
    peak, peak_kind = measured_peak_gbs()
    c1_mode = os.environ.get("MPROS_C1_MODE", "")
    col_kernel = ("collision_poses_kernel" if c1_mode == "walk" else "collision_poses_2phase_kernel" if c1_mode == "2phase_ldg"
                  else "collision_poses_tma_kernel")
    col_roof = {"bound": "hbm", "achieved": n * B_CHECK / (float(np.mean(col_ms)) * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "traffic": ncu_traffic(col_kernel), "kernel": col_kernel, "kernel_ms": float(np.mean(col_ms)), "peak_source": peak_kind,
                "algorithmic_bytes_per_unit": B_CHECK}
    col_roof["frac"] = col_roof["achieved"] / peak
    del poses_d, verd_d

    line = None
    if args.workload == "collision":
        sampler.stop_flag = True
        sampler.join(timeout=2)
        if rank == 0:
            line = {"metric": "collision_checks_per_sec", "value": col_value, "unit": "checks/s", "ms_per_step": col_total_ms / col_steps,
                    "e2e": {"value": col_e2e, "unit": "checks/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n},
                    "gpu_launches": int(col_launches), "roofline": col_roof,
                    "config": {"workload": workload_text("collision"), "poses_per_step_per_gpu": n, "hit_rate": hit_rate,
                               "l2_policy": "inputs larger than L2 (403 MB of poses per step)", "preprocess_s_first_call": t_pre}}
    elif args.workload == "rs":
        # R1-R4: batched Reeds-Shepp word evaluation + sampled-path collision
        ns = RS_SHOTS_PER_STEP
        rng = np.random.default_rng(POSE_SEED + 99 + rank)
        cand = gen_poses(3 * ns, POSE_SEED + 60 + 1000 * rank)
        cand = cand[ctx.collision_check(cand) == 0][:ns]
        assert len(cand) == ns
        st5 = np.concatenate([cand, rng.choice([-1.0, 1.0], ns)[:, None],
                              (params["/mpros/vehicle/max_steering_angle"] * rng.integers(-1, 2, ns))[:, None]], axis=1)
        dd, aa = rng.uniform(0.5, 10.0, ns), rng.uniform(-np.pi, np.pi, ns)
        g3 = np.stack([cand[:, 0] + dd * np.cos(aa), cand[:, 1] + dd * np.sin(aa), rng.uniform(-np.pi, np.pi, ns)], axis=1)
        s_h, g_h = torch.from_numpy(np.ascontiguousarray(st5)).pin_memory(), torch.from_numpy(np.ascontiguousarray(g3)).pin_memory()
        s_d, g_d = s_h.cuda(), g_h.cuda()
        mk_d = lambda shape, dt: torch.zeros(shape, dtype=dt, device="cuda")
        mk_h = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
        seg_d, nseg_d, cost_d, npts_d, ver_d = mk_d(ns * 15, torch.float64), mk_d(ns, torch.int32), mk_d(ns, torch.float64), mk_d(ns, torch.int32), mk_d(ns, torch.uint8)
        seg_h, nseg_h, cost_h, npts_h, ver_h = mk_h(ns * 15, torch.float64), mk_h(ns, torch.int32), mk_h(ns, torch.float64), mk_h(ns, torch.int32), mk_h(ns, torch.uint8)

        def step_dev():
            ctx._check(ctx.L.mpros_rs_shot_batch_dev(ctx.h, s_d.data_ptr(), g_d.data_ptr(), ns, seg_d.data_ptr(), nseg_d.data_ptr(), cost_d.data_ptr(),
                                                     npts_d.data_ptr(), ver_d.data_ptr()))

        def step_host():
            ctx._check(ctx.L.mpros_rs_shot_batch(ctx.h, s_h.data_ptr(), g_h.data_ptr(), ns, seg_h.data_ptr(), nseg_h.data_ptr(), cost_h.data_ptr(),
                                                 None, 0, npts_h.data_ptr(), ver_h.data_ptr()))

        total_ms, step_ms, launches = timed_device(step_dev)
        e2e_ms = timed_host(step_host)
        sampler.stop_flag = True
        sampler.join(timeout=2)
        assert torch.equal(ver_h, ver_d.cpu()) and torch.equal(cost_h, cost_d.cpu()), "host-API and device-API shots differ"
        total_ms, e2e_ms = max_over_ranks([total_ms, e2e_ms])
        if rank == 0:
            k_ms = float(np.mean(step_ms))
            mean_pts = float(npts_h.float().mean())
            n_probe = B_CHECK - 25
            algo_unit = 64 + 16 + mean_pts * (n_probe + 5 - 4)  # 64 B in + 16 B out + n_pts x (n + 5) B probes (n = 22)
            achieved = ns * algo_unit / (k_ms * 1e-3) / 1e9
            line = {"metric": "rs_shots_per_sec", "value": world * ns * args.steps / (total_ms * 1e-3), "unit": "shots/s",
                    "ms_per_step": total_ms / args.steps,
                    "e2e": {"value": world * ns * args.steps / (e2e_ms * 1e-3), "unit": "shots/s", "h2d_bytes_per_step": ns * 64,
                            "d2h_bytes_per_step": ns * (120 + 4 + 8 + 4 + 1)},
                    "gpu_launches": int(launches),
                    "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                 "traffic": (lambda a, b: a + b if a is not None and b is not None else None)(ncu_traffic("rs_solve_kernel"), ncu_traffic("rs_traj_kernel")),
                                 "kernel": "rs_solve_kernel + rs_traj_kernel (one shot batch = the two launches, timed together)", "kernel_ms": k_ms,
                                 "peak_source": peak_kind, "algorithmic_bytes_per_unit": algo_unit,
                                 "note": "FP64-issue / instruction-fetch bound (48 Reeds-Shepp candidate words with sin/cos/atan2/acos per shot, sequential trajectory integration), not HBM bound"},
                    "extra": {"mean_trajectory_poses": mean_pts, "shots_in_collision": float(ver_h.float().mean()),
                              "trajectory_poses_checked_per_sec": world * ns * mean_pts * args.steps / (total_ms * 1e-3),
                              "collision_kernel_checks_per_sec": col_value},
                    "config": {"workload": workload_text("rs"), "shots_per_step_per_gpu": ns,
                               "l2_policy": "inputs + outputs 207 MB per step (larger than L2)", "preprocess_s_first_call": t_pre}}
    elif args.workload == "preprocess":
        # N1-N9 + N4: every map layer from the int8 occupancy grid, then the goal map of one goal
        H0, W0 = occ.shape
        occ_d = occ_pinned.cuda()
        info = ctx.info()
        cells0, cells = H0 * W0, int(info["map_height"]) * int(info["map_width"])
        gx, gy = EXTENT * 0.5 + 3.3, EXTENT * 0.5 - 1.7
        vor_h = torch.zeros(cells, dtype=torch.float32).pin_memory()
        goal_h = torch.zeros(cells, dtype=torch.int32).pin_memory()

        def step_dev():
            ctx.set_occupancy_dev(occ_d.data_ptr(), W0, H0, SYN_RES, [0.0, 0.0, 0.0])
            ctx.update_goal(gx, gy)

        def step_host():
            ctx._check(ctx.L.mpros_map_set_occupancy(ctx.h, occ_pinned.data_ptr(), W0, H0, SYN_RES, 0.0, 0.0, 0.0))
            ctx.update_goal(gx, gy)
            ctx._check(ctx.L.mpros_map_download(ctx.h, mpros_b200.LAYERS["voronoi"][0], vor_h.data_ptr(), vor_h.numel() * 4))
            ctx._check(ctx.L.mpros_map_download(ctx.h, mpros_b200.LAYERS["goal"][0], goal_h.data_ptr(), goal_h.numel() * 4))

        total_ms, step_ms, launches = timed_device(step_dev)
        e2e_ms = timed_host(step_host)
        # split of one step: the map layers (N1-N3, N5-N9) and the goal map (N4)
        split = []
        for _ in range(5):
            ea, eb, ec = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ea.record(stream)
            ctx.set_occupancy_dev(occ_d.data_ptr(), W0, H0, SYN_RES, [0.0, 0.0, 0.0])
            eb.record(stream)
            ctx.update_goal(gx, gy)
            ec.record(stream)
            torch.cuda.synchronize()
            split.append((ea.elapsed_time(eb), eb.elapsed_time(ec)))
        occ_ms, goal_ms = [float(np.median(v)) for v in zip(*split)]
        goal_layer = ctx.layer("goal")
        sampler.stop_flag = True
        sampler.join(timeout=2)
        total_ms, e2e_ms = max_over_ranks([total_ms, e2e_ms])
        if rank == 0:
            k_ms = float(np.mean(step_ms))
            ds = SYN_DS
            algo = cells0 * (2 + (1 + 1.0 / (ds * ds)) + 2) + cells * (18 + 2 * B_GOAL_CELL)
            achieved = algo / (k_ms * 1e-3) / 1e9
            line = {"metric": "map_cells_per_sec", "value": world * cells0 * args.steps / (total_ms * 1e-3), "unit": "cells/s",
                    "ms_per_step": total_ms / args.steps,
                    "e2e": {"value": world * cells0 * args.steps / (e2e_ms * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": cells0,
                            "d2h_bytes_per_step": cells * 8},
                    "gpu_launches": int(launches),
                    "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                                 "kernel": "whole step (%d launches: N1 threshold_pack_vec, N3 inflate, N2 downsample, tile words, N5 ccl_*, N6/N8 dt_row/dt_col, N7 edge, N9 voronoi_cost, N4 goal_bfs_cluster x2)" % (launches // max(1, args.steps)),
                                 "kernel_ms": k_ms, "peak_source": peak_kind, "algorithmic_bytes_per_launch": algo,
                                 "note": "per original cell N1 2 B + N2 1+1/ds^2 B + N3 2 B, per down-sampled cell N5-N9 18 B + N4 2 x 5 B (SURVEY 8d; setOccupancyMap regenerates the goal map once a goal is known, nav_map.cpp:150-156, updateGoal does it again); the step is a chain of ~40 short kernels on a 1024^2 grid plus <= 999 wavefront levels, launch/sync latency bound rather than HBM bound"},
                    "extra": {"set_occupancy_ms": occ_ms, "update_goal_ms": goal_ms,
                              "goal_map_levels": int(goal_layer[goal_layer < 1000].max()), "goal_map_cells_reached": int((goal_layer < 1000).sum()),
                              "collision_kernel_checks_per_sec": col_value},
                    "config": {"workload": workload_text("preprocess"), "cells_per_step_per_gpu": cells0, "downsampled_cells": cells,
                               "l2_policy": "64 MiB int8 grid read once per step; the bit-packed layers (8 MiB each) stay in L2 by design",
                               "preprocess_s_first_call": t_pre}}
    elif args.workload == "expand":
        # E1-E5: the frontier-parallel expansion kernel on a node batch of ONE query (goal fixed, goal map resident)
        ne = EXPANSIONS_PER_STEP
        P = ctx.n_primitives()
        rng = np.random.default_rng(POSE_SEED + 77 + rank)
        cand = gen_poses(3 * ne, POSE_SEED + 50 + 1000 * rank)
        cand = cand[ctx.collision_check(cand) == 0][:ne]
        assert len(cand) == ne
        goal = np.array([EXTENT * 0.5 + 3.3, EXTENT * 0.5 - 1.7, 0.6])
        while ctx.collision_check(goal[None, :])[0]:
            goal[:2] += 0.9
        ctx.update_goal(goal[0], goal[1])
        steer = params["/mpros/vehicle/max_steering_angle"] * rng.integers(-1, 2, ne)
        par = np.concatenate([cand, steer[:, None], rng.choice([-1.0, 1.0], ne)[:, None], rng.uniform(0, 100, ne)[:, None]], axis=1)
        par_h = torch.from_numpy(np.ascontiguousarray(par)).pin_memory()
        par_d = par_h.cuda()
        pr_d = torch.zeros(ne * P * 7, dtype=torch.float64, device="cuda")
        ch_d = torch.zeros(ne * P * 10, dtype=torch.float64, device="cuda")
        nc_d = torch.zeros(ne, dtype=torch.int32, device="cuda")
        pr_h = torch.zeros(ne * P * 7, dtype=torch.float64).pin_memory()
        ch_h = torch.zeros(ne * P * 10, dtype=torch.float64).pin_memory()
        nc_h = torch.zeros(ne, dtype=torch.int32).pin_memory()
        g_c = goal.ctypes.data

        def step_dev():
            ctx._check(ctx.L.mpros_expand_batch_dev(ctx.h, None, g_c, par_d.data_ptr(), ne, pr_d.data_ptr(), ch_d.data_ptr(), nc_d.data_ptr()))

        def step_host_ref_format():
            ctx._check(ctx.L.mpros_expand_batch(ctx.h, None, g_c, par_h.data_ptr(), ne, pr_h.data_ptr(), ch_h.data_ptr(), nc_h.data_ptr()))

        # the call a user makes for the open-list inserts: compact child records (80 B each, ~2 per expansion) instead of
        # the 816 B of fixed-size reference-format arrays per expansion
        import ctypes as C

        rec_h = torch.zeros(ne * P * 80, dtype=torch.uint8).pin_memory()
        n_rec = C.c_size_t(0)

        def step_host():
            ctx._check(ctx.L.mpros_expand_batch_compact(ctx.h, None, g_c, par_h.data_ptr(), ne, rec_h.data_ptr(), ne * P, C.byref(n_rec),
                                                        nc_h.data_ptr()))

        total_ms, step_ms, launches = timed_device(step_dev)
        e2e_ref_ms = timed_host(step_host_ref_format)
        nc_ref = nc_h.clone()
        e2e_ms = timed_host(step_host)
        assert torch.equal(nc_ref, nc_h) and n_rec.value == int(nc_h.sum()), "compact and reference-format expansions differ"
        step_host_ref_format()
        sampler.stop_flag = True
        sampler.join(timeout=2)
        assert torch.equal(nc_h, nc_d.cpu()) and torch.equal(ch_h, ch_d.cpu()), "host-API and device-API expansions differ"
        total_ms, e2e_ms, e2e_ref_ms = max_over_ranks([total_ms, e2e_ms, e2e_ref_ms])
        if rank == 0:
            k_ms = float(np.mean(step_ms))
            achieved = ne * B_EXPANSION / (k_ms * 1e-3) / 1e9
            survivors = float((pr_h.view(ne, P, 7)[:, :, 5] == 0).float().mean())
            line = {"metric": "node_expansions_per_sec", "value": world * ne * args.steps / (total_ms * 1e-3), "unit": "expansions/s",
                    "ms_per_step": total_ms / args.steps,
                    "e2e": {"value": world * ne * args.steps / (e2e_ms * 1e-3), "unit": "expansions/s", "h2d_bytes_per_step": ne * 48,
                            "d2h_bytes_per_step": int(n_rec.value) * 80 + ne * 4 + 8, "api": "mpros_expand_batch_compact",
                            "reference_format": {"value": world * ne * args.steps / (e2e_ref_ms * 1e-3), "api": "mpros_expand_batch",
                                                 "d2h_bytes_per_step": ne * (P * 17 * 8 + 4)}},
                    "gpu_launches": int(launches),
                    "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                 "traffic": ncu_traffic("expand_batch_kernel"), "kernel": "expand_batch_kernel", "kernel_ms": k_ms,
                                 "peak_source": peak_kind, "algorithmic_bytes_per_unit": B_EXPANSION,
                                 "note": "the kernel also writes the reference-format records (P x 7 + P x 10 doubles = 816 B per expansion), above the 450 B algorithmic figure"},
                    "extra": {"primitives_per_sec": world * ne * P * args.steps / (total_ms * 1e-3), "primitive_survival_rate": survivors,
                              "children_inserted_per_expansion": float(nc_h.float().mean()),
                              "collision_kernel_checks_per_sec": col_value, "collision_kernel_roofline": col_roof},
                    "config": {"workload": workload_text("expand"), "expansions_per_step_per_gpu": ne, "primitives_per_expansion": P,
                               "l2_policy": "inputs + outputs larger than L2 (50 MB of parents in, 860 MB of records out per step)",
                               "preprocess_s_first_call": t_pre}}
    else:
        if args.scaling == "strong":
            # BASELINE.json configs[4] read literally: 16 384 queries IN TOTAL, contiguous shards (sharding.shard_range)
            import sharding

            all_s, all_g = gen_queries(ctx, QUERIES_PER_GPU, synthetic.SYN_QUERY_SEED)
            lo, hi = sharding.shard_range(QUERIES_PER_GPU, rank, world)
            starts, goals = np.ascontiguousarray(all_s[lo:hi]), np.ascontiguousarray(all_g[lo:hi])
        else:
            starts, goals = gen_queries(ctx, QUERIES_PER_GPU, synthetic.SYN_QUERY_SEED + rank)
        nq = len(starts)
        s_h, g_h = torch.from_numpy(starts).pin_memory(), torch.from_numpy(goals).pin_memory()
        s_d, g_d = s_h.cuda(), g_h.cuda()
        # Batches in flight: step k runs on lane k % LANES (own stream + planner workspaces, include/mpros_gpu.h), so the
        # iteration-limit tail of one batch (a few queries, each a dependent chain of 100 001 expansions) overlaps the next
        # batch. Every step still plans all nq queries and writes all results; --lanes 1 gives one batch at a time.
        # Strong scaling shrinks a rank's batch (16 384 / world queries) but not its tail (a limit query is 100 001 dependent
        # expansions whatever the batch size), so more batches must be in flight to cover it: 8 lanes unless --lanes is given.
        # Lanes share the context's workspace pool, so this costs streams and staging areas only.
        LANES = min(effective_lanes(args, world), mpros_b200.PLAN_LANES)

        def dev_bufs():
            return {"status": torch.zeros(nq, dtype=torch.int32, device="cuda"), "cost": torch.zeros(nq, dtype=torch.float64, device="cuda"),
                    "len": torch.zeros(nq, dtype=torch.int32, device="cuda"),
                    "paths": torch.zeros(nq * PATH_CAP * 5, dtype=torch.float64, device="cuda"),
                    "stats": torch.zeros(nq * 6, dtype=torch.int64, device="cuda")}

        def host_bufs():
            return {"status": torch.zeros(nq, dtype=torch.int32).pin_memory(), "cost": torch.zeros(nq, dtype=torch.float64).pin_memory(),
                    "len": torch.zeros(nq, dtype=torch.int32).pin_memory(),
                    "paths": torch.zeros(nq * PATH_CAP * 5, dtype=torch.float64).pin_memory(),
                    "stats": torch.zeros(nq * 6, dtype=torch.int64).pin_memory()}

        dbuf = [dev_bufs() for _ in range(LANES)]
        hbuf = [host_bufs() for _ in range(LANES)]
        status_d, cost_d, stats_d = dbuf[0]["status"], dbuf[0]["cost"], dbuf[0]["stats"]
        status_h, cost_h = hbuf[0]["status"], hbuf[0]["cost"]
        # result gather (world > 1): per-step cost buffers in a ring, so that a lane's next batch never waits for a gather
        RING = 16
        cost_ring = [torch.zeros(nq, dtype=torch.float64, device="cuda") for _ in range(RING)] if world > 1 else None
        gather = [[torch.zeros(nq, dtype=torch.float64, device="cuda") for _ in range(world)] for _ in range(RING)] if world > 1 else None
        gathered = [None] * RING
        lane_ext = {}
        counter = {"dev": 0, "host": 0}

        def lane_stream(lane):
            if lane not in lane_ext:
                lane_ext[lane] = torch.cuda.ExternalStream(ctx.plan_lane_stream(lane))
            return lane_ext[lane]

        def step_dev():
            k = counter["dev"]
            lane = k % LANES
            counter["dev"] += 1
            b = dbuf[lane]
            cost_k = b["cost"]
            if world > 1:
                j = k % RING
                cost_k = cost_ring[j]
                if gathered[j] is not None and ctx.plan_lane_stream(lane):
                    lane_stream(lane).wait_event(gathered[j])  # the gather of step k - RING has read this buffer
            ctx.plan_batch_submit_dev(lane, s_d.data_ptr(), g_d.data_ptr(), nq, b["status"].data_ptr(), cost_k.data_ptr(),
                                      b["len"].data_ptr(), b["paths"].data_ptr(), PATH_CAP, b["stats"].data_ptr())
            if world > 1:  # the path's only exchange step: result gather over NCCL/NVLink, ordered behind the lane's batch
                cur = torch.cuda.current_stream()
                cur.wait_stream(lane_stream(lane))
                dist.all_gather(gather[j], cost_k)
                gathered[j] = torch.cuda.Event()
                gathered[j].record(cur)
                if lane == 0:
                    b["cost"].copy_(cost_k, non_blocking=True)  # lane 0's buffers are the ones checked after the run

        def step_host():
            lane = counter["host"] % LANES
            counter["host"] += 1
            b = hbuf[lane]
            ctx.plan_batch_wait(lane)  # the lane's previous results have been consumed
            ctx.plan_batch_submit(lane, s_h.data_ptr(), g_h.data_ptr(), nq, b["status"].data_ptr(), b["cost"].data_ptr(),
                                  b["len"].data_ptr(), b["paths"].data_ptr(), PATH_CAP, b["stats"].data_ptr())

        def join_lanes():
            """everything issued on the lanes precedes what follows on the context's stream (where the events are)"""
            for lane in range(1, LANES):
                if ctx.plan_lane_stream(lane):
                    stream.wait_stream(lane_stream(lane))
            if world > 1:
                stream.wait_stream(torch.cuda.current_stream())  # the result gathers

        total_ms, step_ms, launches = timed_device(step_dev, join_lanes)
        rank_ms = total_ms  # this rank's own timed region (the headline uses the max over ranks)
        e2e_ms = timed_host(step_host, lambda: ctx.plan_batch_wait(-1))
        # one batch alone (no overlap): its latency and the planner kernel's own duration
        lone = []
        for _ in range(2):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a0.record(stream)
            b = dbuf[0]
            ctx.plan_batch_submit_dev(0, s_d.data_ptr(), g_d.data_ptr(), nq, b["status"].data_ptr(), b["cost"].data_ptr(),
                                      b["len"].data_ptr(), b["paths"].data_ptr(), PATH_CAP, b["stats"].data_ptr())
            a1.record(stream)
            barrier()
            lone.append(a0.elapsed_time(a1))
        lone_ms = float(min(lone))
        sampler.stop_flag = True
        sampler.join(timeout=2)
        st = stats_d.cpu().numpy().reshape(nq, 6)
        status = status_d.cpu().numpy()
        assert np.array_equal(status, status_h.numpy()) and torch.equal(cost_h, cost_d.cpu()), "host-API and device-API plans differ"
        iters, prims, shots, checks, cells = [int(st[:, k].sum()) for k in range(5)]
        P = ctx.n_primitives()
        expansions = prims // P
        total_ms, e2e_ms = max_over_ranks([total_ms, e2e_ms])
        sums = torch.tensor([expansions, checks, shots, cells, int((status == 1).sum()), int((status == 0).sum()), int((status < 0).sum())],
                            dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(sums)
        expansions_all, checks_all, shots_all, cells_all, n_found, n_limit, n_fail = [int(v) for v in sums.tolist()]
        # attribution of the scaling loss: every rank's own step time, iteration total and single-batch latency
        mine = torch.tensor([rank_ms / args.steps, float(iters), lone_ms, float(nq), float((status == 0).sum())], dtype=torch.float64, device="cuda")
        per_rank = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(per_rank, mine)
        else:
            per_rank = [mine]
        per_rank = [t.tolist() for t in per_rank]
        nq_all = int(sum(r[3] for r in per_rank))
        verify = None
        if rank == 0 and args.verify > 0:
            verify = verify_against_reference(ctx, occ, params, starts, goals, status, cost_d.cpu().numpy(), st, args.verify)
        if rank == 0:
            secs = total_ms * 1e-3 / args.steps
            # launches overlap when LANES > 1: the kernel's share of the timed region is region / launches (CUDA events on the
            # context's stream around the whole region); one launch alone takes lone_ms
            k_ms = float(np.mean(step_ms)) if LANES == 1 else total_ms / args.steps
            algo_bytes = B_EXPANSION * expansions + B_SHOT * shots + B_GOAL_CELL * cells
            achieved = algo_bytes / (k_ms * 1e-3) / 1e9
            line = {"metric": "plans_per_sec", "value": nq_all / secs, "unit": "plans/s", "ms_per_step": total_ms / args.steps,
                    "e2e": {"value": nq_all * args.steps / (e2e_ms * 1e-3), "unit": "plans/s", "h2d_bytes_per_step": nq * 48,
                            "d2h_bytes_per_step": nq * (4 + 8 + 4 + PATH_CAP * 40 + 48)},
                    "gpu_launches": int(launches),
                    "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                 "traffic": plan_traffic(),
                                 "kernel": "plan_squad_kernel + plan_team_kernel", "kernel_ms": k_ms, "kernel_ms_alone": lone_ms, "peak_source": peak_kind,
                                 "algorithmic_bytes_per_launch": algo_bytes,
                                 "note": "a batch is two launches on one stream: plan_squad_kernel (every query, one warp each, 24 per CTA) and "
                                         "plan_team_kernel (the queries still running when the squads thin out, one CTA each); achieved / traffic are "
                                         "for the pair. Latency / FP64-issue bound (A* is a dependent chain per query, FP64 transcendentals), "
                                         "not HBM bound; see DESIGN.md"},
                    "extra": {"node_expansions_per_sec": expansions_all / secs, "footprint_checks_per_sec_in_planning": checks_all / secs,
                              "rs_shots_per_sec": shots_all / secs, "goal_map_cells_per_sec": cells_all / secs,
                              "collision_kernel_checks_per_sec": col_value, "collision_kernel_e2e_checks_per_sec": col_e2e,
                              "collision_kernel_roofline": col_roof, "collision_hit_rate": hit_rate,
                              "plans_found": n_found, "plans_iteration_limit": n_limit, "plans_failed_other": n_fail,
                              "mean_expansions_per_query": expansions_all / nq_all,
                              "single_batch_ms": lone_ms, "single_batch_plans_per_sec": nq / (lone_ms * 1e-3),
                              "note_value": "value is the streaming rate with %d batches in flight per GPU; one batch alone takes single_batch_ms" % LANES,
                              "preprocess_s_first_call": t_pre,
                              "per_rank": {"step_ms": [r[0] for r in per_rank], "iterations_per_step": [int(r[1]) for r in per_rank],
                                           "single_batch_ms": [r[2] for r in per_rank], "queries": [int(r[3]) for r in per_rank],
                                           "iteration_limit_queries": [int(r[4]) for r in per_rank],
                                           "note": "step_ms = each rank's own timed region / steps (the headline takes the max); a limit query is a "
                                                   "dependent chain of 100 001 expansions (0.65-1.0 s), so a single batch can never finish faster than that"}},
                    "config": plan_config(LANES, args.scaling, world)}
            line["config"]["single_batch_ms"] = lone_ms
            if verify is not None:
                line["extra"]["verify_vs_reference"] = verify

    if rank == 0 and line is not None:
        cpu = None
        if args.gpus == 1 and not args.no_cpu:
            try:
                arm = CpuArm(occ, params, 1)
                try:
                    if args.workload == "collision":
                        arm.collision_rate(1 << 18, POSE_SEED + 5)
                        v = arm.collision_rate(1 << 22, POSE_SEED + 6)
                        cpu = {"value": v, "unit": "checks/s", "cores": 1, "kind": "reference",
                               "sample": "4194304 seeded uniform poses on the synthetic 8192^2 map, 1 process x 1 thread, unmodified reference footPrintInCollision4 loop (preprocessing excluded)"}
                    elif args.workload == "rs":
                        rc = ctx.rs_default_config()
                        v = arm.rs_rate(st5[:20000], g3[:20000], [rc.radius, rc.steering_angle, rc.step_size, rc.pen_gear, rc.pen_backward,
                                                                    rc.pen_steer_change, rc.pen_steer_angle])
                        cpu = {"value": v[0], "unit": "shots/s", "cores": 1, "kind": "reference",
                               "sample": "the first 20000 shots of the step, 1 process x 1 thread, unmodified reference RSCurve::SetStart/SetEnd/FindOptimalPath/GetTraj + pathInCollision of the trajectory, one ctypes call per shot (call overhead measured on an empty call: %.1f us of %.1f us per shot)" % (v[1] * 1e6, 1e6 / v[0])}
                    elif args.workload == "preprocess":
                        r = arm.preprocess_time(EXTENT * 0.5 + 3.3, EXTENT * 0.5 - 1.7)
                        cpu = {"value": occ.size / (r[0] + r[1]), "unit": "cells/s", "cores": 1, "kind": "reference",
                               "sample": "one pass: unmodified reference NavMap::setOccupancyMap (%.2f s) + updateGoal (%.2f s) on the same 8192^2 map, 1 process x 1 thread" % r}
                    elif args.workload == "expand":
                        qs, qg = gen_queries(ctx, QUERIES_PER_GPU, synthetic.SYN_QUERY_SEED)  # the plan workload's first 8 queries
                        r = arm.plan_rate(qs[:8], qg[:8])
                        cpu = {"value": r["expansions_per_sec"], "unit": "expansions/s", "cores": 1, "kind": "reference",
                               "sample": "expandNode calls inside the unmodified reference's Plan() on 8 seeded queries of the plan workload, 1 process x 1 thread (primitives evaluated / 6 / time of updateGoal + ctor + Plan())"}
                    else:
                        r = arm.plan_rate(starts[:8], goals[:8])
                        # the 8 reference results are compared, not thrown away: status and tree_.size() must equal the GPU's
                        # wherever the search does not cross a tie-broken Voronoi cell; the goal cost is reported as a difference
                        same = [int(a_[0] == (1 if status[k] == 1 else 0)) for k, a_ in enumerate(r["per_query"])]
                        same_tree = [int(a_[1] == int(st[k, 1])) for k, a_ in enumerate(r["per_query"])]
                        dcost = [abs(a_[2] - float(cost_d[k])) / abs(a_[2]) for k, a_ in enumerate(r["per_query"]) if a_[0] and status[k] == 1]
                        cpu = {"value": r["plans_per_sec"], "unit": "plans/s", "cores": 1, "kind": "reference",
                               "sample": "first 8 queries of rank 0's set, 1 process x 1 thread, unmodified reference: NavMap::updateGoal + HybridAstar ctor + Plan() per query (map preprocessing %.1f s excluded); %d/8 solved" % (r["preprocess_s"], r["solved"]),
                               "node_expansions_per_sec": r["expansions_per_sec"],
                               "agreement_with_gpu": {"same_status": sum(same), "same_primitive_count": sum(same_tree), "of": len(same),
                                                      "max_rel_cost_difference": max(dcost) if dcost else None,
                                                      "note": "fully GPU (canonical Voronoi tie-breaks) against the reference's own field; tests/ pin all 16384 queries"}}
                        assert sum(same) == len(same), "the reference disagrees with the GPU on the status of a sampled query"
                finally:
                    arm.close()
            except Exception as e:  # oracle absent: say so, do not fail the GPU bench
                cpu = {"value": None, "unit": line["unit"], "cores": 0, "kind": "unavailable", "sample": str(e)}
        if args.workload == "plan" and world == 1 and not args.no_arena:
            # BASELINE.json configs[2] (case 3, the configuration the north star's 50x target is quoted on) in the default line
            try:
                ctx.release_workspaces()
                torch.cuda.empty_cache()
                ar = arena_measure(args, 2, not args.no_cpu)
                line["extra"]["case3_arena"] = {"plans_per_sec": ar["value"], "plans_per_sec_e2e": ar["e2e"]["value"], "ms_per_step": ar["ms_per_step"],
                                                "queries_per_step": ar["config"]["queries_per_gpu"], "steps": 2,
                                                "workload": ar["config"]["workload"], "gpu": ar["extra"], "cpu_reference": ar["cpu_baseline"]}
            except Exception as e:
                line["extra"]["case3_arena"] = {"error": str(e)}
        line.update({"n_gpus": world, "steps": args.steps, "warmup": warm, "higher_is_better": True, "scaling": args.scaling if args.workload == "plan" else "weak",
                     "vs_baseline": None, "dtype": "f64", "data": "synthetic", "clocks": sampler.summary(), "cpu_baseline": cpu})
        order = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                 "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "extra"]
        print(json.dumps({k: line[k] for k in order if k in line}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_arena(args):
    print(json.dumps(arena_measure(args, args.steps, not args.no_cpu)))
    return 0


def arena_measure(args, steps, with_cpu):
    """BASELINE.json configs[2], "case 3 Arena whole map" (the case the north star's 50x target names): the 2000 x 2000 @0.04 m
    stand-in of SURVEY.md 8(d) (Arena2036's pgm is missing from the reference: ARENA_part's image read with Arena2036.yaml's
    thresholds, nearest-neighbour x4; committed input fixture tests/golden/maps/arena_standin.npz), down-sampling 4. One line:
    plans/s (headline), footprint checks/s and map preprocessing on this map, each beside the unmodified reference."""
    import torch

    import mpros_b200

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise RuntimeError("--workload arena is a single-GPU workload")
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    z = np.load(os.path.join(ROOT, "tests", "golden", "maps", "arena_standin.npz"))
    occ = np.ascontiguousarray(np.repeat(np.repeat(z["occupancy"], 4, axis=0), 4, axis=1))
    res, ds = 0.04, 4
    extent = occ.shape[0] * res
    params = shipped_params(ds)
    params["/mpros/map/free_thresh"], params["/mpros/map/occupied_thresh"] = float(z["free_thresh"]), float(z["occupied_thresh"])
    ctx = mpros_b200.Context(0)
    ctx.map_config(params)
    occ_pinned = torch.from_numpy(occ).pin_memory()
    ctx.set_occupancy(occ_pinned.numpy(), res, [0.0, 0.0, 0.0])
    ctx.planner_config(params)
    stream = torch.cuda.ExternalStream(ctx.stream())
    warm, K = max(3, args.warmup), steps
    sampler = ClockSampler(0)
    sampler.start()

    def timed(fn, steps, join=None):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(stream)
        for _ in range(steps):
            fn()
        if join:
            join()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), ctx.launch_count() - l0

    # map preprocessing (host buffer -> all layers), median of 5
    t_pre = []
    for _ in range(5):
        t0 = time.perf_counter()
        ctx.set_occupancy(occ_pinned.numpy(), res, [0.0, 0.0, 0.0])
        t_pre.append(time.perf_counter() - t0)
    t_pre = float(np.median(t_pre))
    # footprint checks: 2^24 uniform poses, n_check = 55 probes at this resolution -> 24 + 1 + 60 = 85 B per check
    n = POSES_PER_STEP
    poses_d = torch.from_numpy(gen_poses(n, POSE_SEED, extent)).cuda()
    verd_d = torch.zeros(n, dtype=torch.uint8, device="cuda")
    col_ms, _ = timed(lambda: ctx.collision_check_dev(poses_d.data_ptr(), n, verd_d.data_ptr()), 10)
    col_rate = n * 10 / (col_ms * 1e-3)
    hit_rate = float(verd_d.float().mean())
    del poses_d, verd_d
    # plans: seeded independent queries 20-80 m apart (the synthetic recipe on this map), batches in flight as in run_b200
    nq = ARENA_QUERIES
    static_ds = ctx.layer("static")
    starts, goals = synthetic.synthetic_queries(nq, ctx.collision_check, static_ds, res * ds, extent, synthetic.SYN_QUERY_SEED)
    s_h, g_h = torch.from_numpy(starts).pin_memory(), torch.from_numpy(goals).pin_memory()
    s_d, g_d = s_h.cuda(), g_h.cuda()
    LANES = min(effective_lanes(args, 1), mpros_b200.PLAN_LANES)
    mk = lambda pin: {k: (torch.zeros(sz, dtype=dt).pin_memory() if pin else torch.zeros(sz, dtype=dt, device="cuda"))
                      for k, sz, dt in (("status", nq, torch.int32), ("cost", nq, torch.float64), ("len", nq, torch.int32),
                                        ("paths", nq * PATH_CAP * 5, torch.float64), ("stats", nq * 6, torch.int64))}
    dbuf, hbuf = [mk(False) for _ in range(LANES)], [mk(True) for _ in range(LANES)]
    cnt = {"d": 0, "h": 0}

    def step_dev():
        lane = cnt["d"] % LANES
        cnt["d"] += 1
        b = dbuf[lane]
        ctx.plan_batch_submit_dev(lane, s_d.data_ptr(), g_d.data_ptr(), nq, b["status"].data_ptr(), b["cost"].data_ptr(), b["len"].data_ptr(),
                                  b["paths"].data_ptr(), PATH_CAP, b["stats"].data_ptr())

    def join():
        for lane in range(1, LANES):
            if ctx.plan_lane_stream(lane):
                stream.wait_stream(torch.cuda.ExternalStream(ctx.plan_lane_stream(lane)))

    plan_ms, launches = timed(step_dev, K, join)

    def step_host():
        lane = cnt["h"] % LANES
        cnt["h"] += 1
        b = hbuf[lane]
        ctx.plan_batch_wait(lane)
        ctx.plan_batch_submit(lane, s_h.data_ptr(), g_h.data_ptr(), nq, b["status"].data_ptr(), b["cost"].data_ptr(), b["len"].data_ptr(),
                              b["paths"].data_ptr(), PATH_CAP, b["stats"].data_ptr())

    for _ in range(2):
        step_host()
    ctx.plan_batch_wait(-1)
    t0 = time.perf_counter()
    for _ in range(K):
        step_host()
    ctx.plan_batch_wait(-1)
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    st = dbuf[0]["stats"].cpu().numpy().reshape(nq, 6)
    status = dbuf[0]["status"].cpu().numpy()
    assert np.array_equal(status, hbuf[0]["status"].numpy()), "host-API and device-API plans differ"
    P = ctx.n_primitives()
    expansions, shots, cells = int(st[:, 1].sum()) // P, int(st[:, 2].sum()), int(st[:, 4].sum())
    secs = plan_ms * 1e-3 / K
    peak, peak_kind = measured_peak_gbs()
    algo = B_EXPANSION * expansions + B_SHOT * shots + B_GOAL_CELL * cells
    line = {"metric": "plans_per_sec", "value": nq / secs, "unit": "plans/s", "n_gpus": 1, "steps": K, "warmup": warm, "ms_per_step": plan_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "arena: case 3 'Arena whole map' stand-in, 2000x2000 @0.04 m (ARENA_part x4 with Arena2036.yaml's thresholds), ds=4, inflation 1.0 m, "
                                   "shipped planner config; seeded independent start/goal queries 20-80 m apart; per query goal heuristic + Hybrid A* + Reeds-Shepp shots",
                       "queries_per_gpu": nq, "path_cap": PATH_CAP, "batches_in_flight": LANES,
                       "l2_policy": "per-query working sets in per-team HBM workspaces (>> L2 in aggregate); map layers L2-resident by design"},
            "e2e": {"value": nq * K / e2e_s, "unit": "plans/s", "h2d_bytes_per_step": nq * 48, "d2h_bytes_per_step": nq * (4 + 8 + 4 + PATH_CAP * 40 + 48)},
            "gpu_launches": int(launches), "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": algo / secs / 1e9, "peak": peak, "unit": "GB/s", "frac": algo / secs / 1e9 / peak, "traffic": None,
                         "kernel": "plan_squad_kernel + plan_team_kernel", "kernel_ms": plan_ms / K, "peak_source": peak_kind, "algorithmic_bytes_per_launch": algo,
                         "note": "latency/FP64-issue bound (sequential A* per query), not HBM bound; see DESIGN.md"},
            "extra": {"node_expansions_per_sec": expansions / secs, "plans_found": int((status == 1).sum()), "plans_iteration_limit": int((status == 0).sum()),
                      "plans_failed_other": int((status < 0).sum()), "mean_expansions_per_query": expansions / nq,
                      "collision_checks_per_sec": col_rate, "collision_hit_rate": hit_rate,
                      "collision_roofline_frac": col_rate * (24 + 1 + ctx_n_check(ctx) + 1 + 4) / 1e9 / peak,
                      "preprocess_s_host_buffer": t_pre}}
    ctx.close()
    cpu = None
    if with_cpu:
        try:
            arm = CpuArm(occ, params, 1, res)
            try:
                arm.collision_rate(1 << 16, POSE_SEED + 5, extent)
                cv = arm.collision_rate(1 << 20, POSE_SEED + 6, extent)
                r = arm.plan_rate(starts[:ARENA_CPU_QUERIES], goals[:ARENA_CPU_QUERIES])
                cpu = {"value": r["plans_per_sec"], "unit": "plans/s", "cores": 1, "kind": "reference",
                       "sample": "first %d queries of the step, 1 process x 1 thread, unmodified reference: NavMap::updateGoal + HybridAstar ctor + Plan() per query; %d/%d solved; "
                                 "map preprocessing (setOccupancyMap) %.2f s; footPrintInCollision4 on 1048576 uniform poses" % (ARENA_CPU_QUERIES, r["solved"], ARENA_CPU_QUERIES, r["preprocess_s"]),
                       "collision_checks_per_sec": cv, "preprocess_s": r["preprocess_s"], "node_expansions_per_sec": r["expansions_per_sec"]}
            finally:
                arm.close()
            # the whole host: one process per core (0.5 GB each at 500 x 500 cells), 2 queries per process
            cores, nproc, by_mem = usable_cores(0.8e9)
            arm = CpuArm(occ, params, cores, res)
            try:
                m = min(nq, 2 * cores)
                ra = arm.plan_rate(starts[:m], goals[:m])
                cva = arm.collision_rate((1 << 17) * cores, POSE_SEED + 7, extent)
                cpu["all_cores"] = {"plans_per_sec": ra["plans_per_sec"], "collision_checks_per_sec": cva, "processes": cores, "nproc": nproc,
                                    "processes_memory_allows": by_mem, "queries": m, "solved": ra["solved"],
                                    "node_expansions_per_sec": ra["expansions_per_sec"]}
            finally:
                arm.close()
        except Exception as e:
            if cpu is None:
                cpu = {"value": None, "unit": "plans/s", "cores": 0, "kind": "unavailable", "sample": str(e)}
            else:
                cpu["all_cores"] = {"error": str(e)}
    line["cpu_baseline"] = cpu
    if cpu and cpu.get("value"):
        line["extra"]["vs_reference_one_core"] = {"plans": line["value"] / cpu["value"], "collision_checks": col_rate / cpu["collision_checks_per_sec"]}
        if "all_cores" in cpu and "plans_per_sec" in cpu["all_cores"]:
            line["extra"]["vs_reference_all_cores"] = {"plans": line["value"] / cpu["all_cores"]["plans_per_sec"], "plans_e2e": line["e2e"]["value"] / cpu["all_cores"]["plans_per_sec"],
                                                      "collision_checks": col_rate / cpu["all_cores"]["collision_checks_per_sec"],
                                                      "processes": cpu["all_cores"]["processes"], "nproc": cpu["all_cores"]["nproc"],
                                                      "note": "north-star target: >= 50x the reference CPU planner on this case (collision checks/s and planning throughput)"}
    return line


def ctx_n_check(ctx):
    """num_checking_step of the configured planner = ceil((L - W) / map_resolution_orig_) (hybrid_a_star.cpp:632-633)."""
    import math

    inf = ctx.info()
    return int(math.ceil((4.2 - 2.0) / inf["map_resolution_orig"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="plan", choices=["plan", "collision", "expand", "preprocess", "rs", "arena"])
    ap.add_argument("--lanes", type=int, default=0,
                    help="plan workload: query batches in flight (1 = one batch at a time; default 8)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="plan workload: weak = 16384 queries per GPU (default), strong = 16384 queries in total, sharded across the GPUs")
    ap.add_argument("--verify", type=int, default=0,
                    help="plan workload, rank 0: re-plan K seeded queries plus every non-FOUND one with the unmodified reference (oracle/_ref, "
                         "~2.5 s each on one core, run on all cores) and report status / expansion-count / cost agreement per SURVEY 9.4")
    ap.add_argument("--no-arena", action="store_true", help="plan workload at N=1: skip the case-3 (Arena stand-in) figures in extra")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "arena":
        return run_arena(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
