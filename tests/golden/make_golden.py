"""Generate tests/golden/*.npz from the reference's own code compiled unmodified (oracle/_ref).

Run HERE (needs /root/reference to build oracle/_ref); the fixtures are committed and pin both the CPU restatement
(oracle/port, `-m "not gpu"` tests) and the CUDA path (`-m gpu` tests) to the reference without the reference being
present at test time. Inputs are regenerated from seeds by tests/golden_inputs.py (shared with the tests)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_inputs as gi  # noqa: E402
from oracle import refapi  # noqa: E402


def main():
    lib = refapi.RefLib()
    for case in gi.MAP_CASES:
        name, ds, goal, num_steer = case
        m = refapi.load_map(name)
        params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"], ds=ds, num_steer=num_steer)
        rm = refapi.RefMap(lib, params)
        rm.set_occupancy(m["occupancy"], m["resolution"], m["origin"])
        rm.update_goal(goal[0], goal[1])
        inf = rm.info()
        rp = refapi.RefPlanner(rm)
        poses = gi.collision_poses(inf)
        out = {
            "info_i": np.array([inf[k] for k in ["H", "W", "H_orig", "W_orig", "get_map", "get_goal"]], dtype=np.int32),
            "info_d": np.array([inf[k] for k in ["res", "res_orig", "ox", "oy", "oyaw", "sin", "cos"]], dtype=np.float64),
            "static_orig": np.packbits(rm.layer("static_orig").astype(bool)),
            "static": np.packbits(rm.layer("static").astype(bool)),
            "inflate_orig": np.packbits(rm.layer("inflate_orig").astype(bool)),
            "goal": rm.layer("goal").astype(np.int16),
            "obs_dist": rm.layer("obs_dist").astype(np.int16),
            "region": rm.layer("region").astype(np.int32),
            "isedge": np.packbits(rm.layer("isedge").astype(bool)),
            "edge_dist": rm.layer("edge_dist").astype(np.int16),
            "voronoi": rm.layer("voronoi"),
            "verdicts": np.packbits(rp.collision4(poses).astype(bool)),
        }
        np.savez_compressed(os.path.join(HERE, "layers_%s_ds%d.npz" % (name, ds)), **out)
        print("layers", name, ds, {k: v.nbytes for k, v in out.items() if v.nbytes > 20000})
        rp.close()
        rm.close()

    # Reeds-Shepp generator and the (stand-in) OMPL distance
    cfg = gi.RS_CFG
    s, g = gi.rs_pairs()
    segs = np.zeros((len(s), 5, 3))
    nseg = np.zeros(len(s), dtype=np.int32)
    cost = np.zeros(len(s))
    npts = np.zeros(len(s), dtype=np.int32)
    endpose = np.zeros((len(s), 5))
    traj_sum = np.zeros((len(s), 5))
    for k in range(len(s)):
        sg, c, tr = refapi.ref_rs_solve(lib, cfg, s[k], g[k])
        nseg[k], cost[k], npts[k] = len(sg), c, len(tr)
        segs[k, : len(sg)] = sg
        if len(tr):
            endpose[k] = tr[-1]
            traj_sum[k] = tr.sum(axis=0)
    a, b = gi.distance_pairs()
    dist = np.array([refapi.ref_rs_distance(lib, 8.0, a[k], b[k]) for k in range(len(a))])
    np.savez_compressed(os.path.join(HERE, "rs.npz"), segs=segs, nseg=nseg, cost=cost, npts=npts, endpose=endpose, traj_sum=traj_sum, dist=dist)
    print("rs", len(s), "distance", len(a))

    # expansions and whole plans
    plans = {}
    for ci, (name, ds, num_steer, start, goal) in enumerate(gi.PLAN_CASES):
        m = refapi.load_map(name)
        params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"], ds=ds, num_steer=num_steer)
        rm = refapi.RefMap(lib, params)
        rm.set_occupancy(m["occupancy"], m["resolution"], m["origin"])
        r = refapi.ref_plan_query(rm, start, goal)
        plans["stats_%d" % ci] = np.array([r["plan_ok"], r["goal_g"], r["n_primitives"], r["num_nodes"], r["path_len"]])
        plans["path_%d" % ci] = r["path"]
        # getNewPath(): setPath's post-processing of that raw path incl. Bezier up-sampling (hybrid_a_star.cpp:814-838)
        plans["newpath_%d" % ci] = refapi.ref_upsample_path(rm, r["path"])
        P = 2 * (2 * ((num_steer or 3) // 2) + 1)
        parents = gi.expansion_parents(rm, refapi.RefPlanner(rm), num_steer)
        prims = np.zeros((len(parents), P, 7))
        childs = np.zeros((len(parents), P, 10))
        nchild = np.zeros(len(parents), dtype=np.int32)
        for k in range(len(parents)):
            rp2 = refapi.RefPlanner(rm)
            assert rp2.initialize(start, goal)
            pr, ch = rp2.expand(parents[k], P)
            rp2.close()
            prims[k] = pr
            childs[k, : len(ch)] = ch
            nchild[k] = len(ch)
        plans["parents_%d" % ci] = parents
        plans["prims_%d" % ci] = prims
        plans["childs_%d" % ci] = childs
        plans["nchild_%d" % ci] = nchild
        print("plan", name, start, goal, plans["stats_%d" % ci])
        rm.close()
    np.savez_compressed(os.path.join(HERE, "plans.npz"), **plans)


if __name__ == "__main__":
    main()
