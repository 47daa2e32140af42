"""Developer probe: compare plan results of two builds of the library on the synthetic query set."""
import os, sys, subprocess, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))
    import mpros_b200, synthetic, bench
    ctx = mpros_b200.Context(0)
    params = bench.shipped_params()
    ctx.map_config(params)
    ctx.set_occupancy(synthetic.synthetic_occupancy(), synthetic.SYN_RES, [0, 0, 0])
    ctx.planner_config(params)
    nq = 4096
    s, g = bench.gen_queries(ctx, nq, synthetic.SYN_QUERY_SEED)
    r = ctx.plan_batch(s, g, path_cap=16)
    np.savez(sys.argv[2], status=r["status"], cost=r["cost"], it=r["iterations"], nodes=r["nodes_inserted"], s=s, g=g)
    sys.exit(0)
outs = []
for k, lib in enumerate(sys.argv[1:3]):
    env = dict(os.environ, MPROS_LIB=lib)
    out = "/tmp/cmp_%d.npz" % k
    subprocess.check_call([sys.executable, __file__, "child", out], env=env)
    outs.append(np.load(out))
a, b = outs
d = np.nonzero((a["it"] != b["it"]) | (a["status"] != b["status"]))[0]
print("queries differing:", len(d), "of", len(a["it"]))
for q in d[:10]:
    print(q, "it", a["it"][q], b["it"][q], "status", a["status"][q], b["status"][q], "cost", a["cost"][q], b["cost"][q], "start", a["s"][q], "goal", a["g"][q])
print("max |cost diff| among same-status", np.nanmax(np.abs(a["cost"] - b["cost"])[(a["status"] == b["status"])]))
