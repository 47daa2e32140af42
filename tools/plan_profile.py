"""Developer probe (not a test): per-pass timing and per-query work distribution of the synthetic plan workload."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))
import mpros_b200, synthetic, bench
os.environ["MPROS_DEBUG"] = "1"
ctx = mpros_b200.Context(0)
params = bench.shipped_params()
if os.environ.get("MPROS_MAXIT"):
    params["/mpros/planner/max_iteration"] = int(os.environ["MPROS_MAXIT"])
ctx.map_config(params)
ctx.set_occupancy(synthetic.synthetic_occupancy(), synthetic.SYN_RES, [0, 0, 0])
ctx.planner_config(params)
nq = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
s, g = bench.gen_queries(ctx, nq, synthetic.SYN_QUERY_SEED)
for rep in range(2):
    t0 = time.perf_counter()
    r = ctx.plan_batch(s, g, path_cap=128)
    print("plan_batch wall %.3f s" % (time.perf_counter() - t0))
it = r["iterations"]
if os.environ.get("MPROS_SAVE"):
    np.savez_compressed(os.environ["MPROS_SAVE"], starts=s, goals=g, **{k: v for k, v in r.items() if isinstance(v, np.ndarray) and v.ndim == 1})
print("status counts", {int(k): int((r["status"] == k).sum()) for k in np.unique(r["status"])})
print("iterations: mean %.1f median %.1f p90 %.0f p99 %.0f max %d" % (it.mean(), np.median(it), np.percentile(it, 90), np.percentile(it, 99), it.max()))
print("nodes inserted: mean %.1f p90 %.0f p99 %.0f max %d" % (r["nodes_inserted"].mean(), np.percentile(r["nodes_inserted"], 90), np.percentile(r["nodes_inserted"], 99), r["nodes_inserted"].max()))
for cap in [4096, 16384, 65536, 131072]:
    print("queries needing >", cap, "nodes:", int((r["nodes_inserted"] >= cap).sum()))
print("sum iterations", it.sum(), "excluding limit-hitters", it[r["status"] != 0].sum())
print("goal cells mean", r["goal_map_cells"].mean(), "rs shots mean", r["rs_shots"].mean())
# lone-team latency: only the queries that ran into the iteration limit
hard = np.nonzero(r["status"] == 0)[0]
if len(hard):
    t0 = time.perf_counter()
    r2 = ctx.plan_batch(s[hard], g[hard], path_cap=16)
    dt = time.perf_counter() - t0
    print("limit-hitters alone: %d queries, %.3f s -> %.2f us per iteration (lone team)" % (len(hard), dt, 1e6 * dt / r2["iterations"].max()))
