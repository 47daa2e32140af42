"""Developer probe: one small plan batch (for compute-sanitizer runs)."""
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
ctx.update_goal(11.5, 2.1)
s = np.array([[17.0, 8.0, 3.14], [3.0, 8.0, 0.0], [16.0, 9.0, 1.0], [5.0, 9.0, -1.0]] * 4)
g = np.array([[11.5, 2.1, 1.57]] * 16)
r = ctx.plan_batch(s, g, path_cap=64)
print(r["status"], r["iterations"], r["cost"])
print(ctx.collision_check(s).tolist())
