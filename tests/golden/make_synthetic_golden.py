#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — golden vectors of the headline configuration (BASELINE.json configs[4]: synthetic 8192^2 grid,
ds = 8, the 16 384 seeded queries of bench.py) recorded from the UNMODIFIED reference (oracle/_ref) in this container.

    python tests/golden/make_synthetic_golden.py [--procs 8] [--every 32]

Writes tests/golden/synthetic_plans.npz:
  port_*      every one of the 16 384 queries through the CPU restatement (oracle/port; canonical Voronoi field = the field
              the GPU computes): status, iterations, primitives, RS shots, cost, path length
  ref_idx     the queries run through the reference: every `--every`-th query (512 at the default) plus ALL queries whose
              restated status is not FOUND (iteration limit / open set empty)
  ref_*       HybridAstar::Plan() of the reference for those (its own, order-dependent Voronoi field): status
              (enum mpros_plan_status), tree_.size() (primitives), goal g, path_, seconds per phase
  refv_*      the restatement on the same queries with the REFERENCE's Voronoi field injected (search parity in isolation)
  voro_idx/voro_val   the cells where the reference's voronoi_value_map_ differs from the canonical field, and its values
                      there (so that a test can rebuild the reference's field from the GPU's without running the reference)
The reference needs ~2.2 s and 1.8 GB per query (updateGoal + dense cost_set_/node_set_ fill + Plan), hence the subset."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))

PATH_CAP = 128
_W = {}


def _setup(kind, ref_voronoi=None):
    import bench
    import synthetic
    from oracle import portapi, refapi

    occ = synthetic.synthetic_occupancy()
    params = bench.shipped_params()
    if kind == "ref":
        rm = refapi.RefMap(refapi.RefLib(), params)
        rm.set_occupancy(occ, synthetic.SYN_RES, [0.0, 0.0, 0.0])
        _W["rm"] = rm
    else:
        pm = portapi.PortMap(portapi.PortLib(), params)
        pm.set_occupancy(occ, synthetic.SYN_RES, [0.0, 0.0, 0.0])
        if ref_voronoi is not None:
            pm.set_voronoi(np.load(ref_voronoi))
        _W["pm"], _W["pp"] = pm, portapi.PortPlanner(pm)


def _port_job(job):
    out = []
    for q, s, g in job:
        _W["pm"].update_goal(g[0], g[1])
        o = _W["pp"].plan(s, g, path_cap=PATH_CAP)
        out.append((q, int(o["status"]), int(o["iterations"]), int(o["n_primitives"]), int(o["n_rs_shots"]), float(o["goal_g"]),
                    int(o["path_len"]), o["path"]))
    return out


def _ref_job(job):
    from oracle import refapi

    out = []
    for q, s, g in job:
        r = refapi.ref_plan_query(_W["rm"], s, g, path_cap=PATH_CAP)
        out.append((q, int(r["status"]), int(r["n_primitives"]), float(r["goal_g"]), int(r["path_len"]), r["path"],
                    float(r["t_update_goal"]), float(r["t_ctor"]), float(r["t_plan"])))
    return out


def _pool(procs, kind, ref_voronoi=None):
    from multiprocessing import get_context

    return get_context("spawn").Pool(procs, initializer=_setup, initargs=(kind, ref_voronoi))


def _jobs(idx, starts, goals, procs, chunk=8):
    items = [(int(q), starts[q].copy(), goals[q].copy()) for q in idx]
    return [items[k:k + chunk] for k in range(0, len(items), chunk)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--every", type=int, default=32)
    ap.add_argument("--nq", type=int, default=16384)
    args = ap.parse_args()
    import bench
    import synthetic
    from oracle import portapi, refapi

    t0 = time.time()
    _setup("port")
    pm, pp = _W["pm"], _W["pp"]
    starts, goals = synthetic.synthetic_queries(args.nq, pp.collision4, pm.layer("static"), synthetic.SYN_RES * synthetic.SYN_DS,
                                                bench.EXTENT, synthetic.SYN_QUERY_SEED)
    canon = pm.layer("voronoi")
    # 1. the whole set through the restatement
    with _pool(args.procs, "port") as pool:
        res = [r for chunk in pool.imap_unordered(_port_job, _jobs(range(args.nq), starts, goals, args.procs, 64)) for r in chunk]
    res.sort(key=lambda r: r[0])
    port = {"status": np.array([r[1] for r in res], np.int8), "iterations": np.array([r[2] for r in res], np.int32),
            "primitives": np.array([r[3] for r in res], np.int32), "shots": np.array([r[4] for r in res], np.int32),
            "cost": np.array([r[5] for r in res], np.float64), "path_len": np.array([r[6] for r in res], np.int32)}
    print("restatement, %d queries: %.0f s; status counts %s" % (args.nq, time.time() - t0, dict(zip(*np.unique(port["status"], return_counts=True)))))
    # 2. the reference on the subset
    idx = np.union1d(np.arange(0, args.nq, args.every), np.nonzero(port["status"] != 1)[0]).astype(np.int32)
    t1 = time.time()
    with _pool(args.procs, "ref") as pool:
        rres = [r for chunk in pool.imap_unordered(_ref_job, _jobs(idx, starts, goals, args.procs, 4)) for r in chunk]
    rres.sort(key=lambda r: r[0])
    print("reference, %d queries: %.0f s" % (len(idx), time.time() - t1))
    ref_paths = np.zeros((len(idx), PATH_CAP, 5))
    for k, r in enumerate(rres):
        ref_paths[k, : len(r[5])] = r[5]
    # the reference's own Voronoi field (one more pass; cheap next to the above)
    _setup("ref")
    ref_v = _W["rm"].layer("voronoi")
    diff = np.nonzero(ref_v.ravel().view(np.uint32) != canon.ravel().view(np.uint32))[0]
    tmp = "/tmp/_ref_voronoi.npy"
    np.save(tmp, ref_v)
    # 3. the restatement with the reference's field injected, same subset
    with _pool(args.procs, "port", tmp) as pool:
        vres = [r for chunk in pool.imap_unordered(_port_job, _jobs(idx, starts, goals, args.procs, 8)) for r in chunk]
    vres.sort(key=lambda r: r[0])
    out = os.path.join(ROOT, "tests", "golden", "synthetic_plans.npz")
    np.savez_compressed(
        out, nq=args.nq, seed=synthetic.SYN_QUERY_SEED, starts_sum=starts.sum(), goals_sum=goals.sum(),
        **{"port_" + k: v for k, v in port.items()},
        ref_idx=idx, ref_status=np.array([r[1] for r in rres], np.int8), ref_primitives=np.array([r[2] for r in rres], np.int32),
        ref_cost=np.array([r[3] for r in rres]), ref_path_len=np.array([r[4] for r in rres], np.int32),
        ref_paths=ref_paths[:, : max(1, int(max(r[4] for r in rres)))],
        ref_seconds=np.array([[r[6], r[7], r[8]] for r in rres], np.float32),
        refv_status=np.array([r[1] for r in vres], np.int8), refv_iterations=np.array([r[2] for r in vres], np.int32),
        refv_primitives=np.array([r[3] for r in vres], np.int32), refv_cost=np.array([r[5] for r in vres]),
        voro_idx=diff.astype(np.int32), voro_val=ref_v.ravel()[diff])
    print("wrote %s (%.1f KB); Voronoi cells that differ from the canonical field: %d of %d" % (out, os.path.getsize(out) / 1024, len(diff), ref_v.size))


if __name__ == "__main__":
    main()
