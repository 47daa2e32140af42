"""Developer probe for compute-sanitizer: a small batch on a small map with workspaces so small that every path between the
planner kernels is taken (squad pass -> grown workspace -> team pass with state, hand-over without growth, restart), checked
against the team kernel alone.   compute-sanitizer --tool memcheck python tools/sanitize_plan.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))
import mpros_b200
from oracle import refapi
m = refapi.load_map("reverse_parking")
params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"])
ctx = mpros_b200.Context(0)
ctx.map_config(params)
ctx.set_occupancy(m["occupancy"], m["resolution"], m["origin"])
ctx.planner_config(params)
rng = np.random.default_rng(5)
base = np.array([[17.0, 8.0, 3.14], [3.0, 8.0, 0.0], [16.0, 9.0, 1.0], [5.0, 9.0, -1.0]])
s = np.repeat(base, 24, axis=0) + np.c_[rng.uniform(-0.3, 0.3, (96, 2)), rng.uniform(-0.2, 0.2, 96)]
g = np.array([[11.5, 2.1, 1.57]] * 96) + np.c_[rng.uniform(-0.2, 0.2, (96, 2)), np.zeros(96)]
keys = ["status", "cost", "path_len", "iterations", "primitives", "rs_shots", "collision_checks", "goal_map_cells", "nodes_inserted"]
ctx.set_option("solo_min_queries", 0)
ctx.set_option("plan_mode", "team")
ref = ctx.plan_batch(s, g, path_cap=64, truncate_ok=True)
print("team:", dict(zip(*np.unique(ref["status"], return_counts=True))), "iterations", ref["iterations"].min(), ref["iterations"].max())
ok = True
for mode, nodes, grow in (("solo", 64, "1"), ("solo", 64, "0"), ("solo", 256, "1"), ("solo_warp", 64, "1"), ("throughput", 64, "1")):
    ctx.set_option("plan_policy", "throughput" if mode == "throughput" else None)
    ctx.set_option("plan_age_limit", 100 if mode == "throughput" else None)
    mode = "solo" if mode == "throughput" else mode
    ctx.set_option("plan_mode", mode)
    ctx.set_option("solo_nodes", nodes)
    ctx.set_option("plan_grow", grow)
    r = ctx.plan_batch(s, g, path_cap=64, truncate_ok=True)
    bad = [k for k in keys if not np.array_equal(r[k], ref[k])]
    valid = np.arange(64)[None, :] < np.minimum(ref["path_len"], 64)[:, None]
    if not np.array_equal(r["paths"][valid], ref["paths"][valid]):
        bad.append("paths")
    print(mode, nodes, grow, "OK" if not bad else "DIFFERS: %s" % bad)
    ok = ok and not bad
ctx.close()
sys.exit(0 if ok else 1)
