"""Developer probe: GPU plan_batch vs the CPU restatement on synthetic queries (first N, or explicit indices)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))
import mpros_b200, synthetic, bench
from oracle import portapi
ctx = mpros_b200.Context(0)
params = bench.shipped_params()
ctx.map_config(params)
occ = synthetic.synthetic_occupancy()
ctx.set_occupancy(occ, synthetic.SYN_RES, [0, 0, 0])
ctx.planner_config(params)
s, g = bench.gen_queries(ctx, 4096, synthetic.SYN_QUERY_SEED)
idx = [int(v) for v in sys.argv[1:]] or list(range(24))
r = ctx.plan_batch(s[idx], g[idx], path_cap=16)
plib = portapi.PortLib()
pm = portapi.PortMap(plib, params)
t0 = time.time(); pm.set_occupancy(occ, synthetic.SYN_RES, [0, 0, 0]); print("port preprocessing %.1f s" % (time.time() - t0))
print("voronoi equal:", np.array_equal(pm.layer("voronoi"), ctx.layer("voronoi")), "inflate equal:", np.array_equal(pm.layer("inflate_orig"), ctx.layer("inflate_orig")))
pp = portapi.PortPlanner(pm)
for k, q in enumerate(idx):
    pm.update_goal(g[q][0], g[q][1])
    o = pp.plan(s[q], g[q])
    print(q, "gpu it %d st %d cost %.9f | port it %d st %d cost %.9f %s" % (r["iterations"][k], r["status"][k], r["cost"][k], o["iterations"], o["status"], o["goal_g"],
          "OK" if (r["iterations"][k] == o["iterations"] and abs(r["cost"][k] - o["goal_g"]) < 1e-6 * max(1, abs(o["goal_g"]))) else "DIFF"))
