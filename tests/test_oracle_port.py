"""CPU tests (`-m "not gpu"`): the CPU restatement (oracle/port) against the golden vectors recorded from the reference's
own code compiled unmodified (tests/golden/make_golden.py), and — where oracle/_ref is present — against that build live."""
import os

import numpy as np
import pytest

import golden_inputs as gi
from oracle import portapi, refapi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def plib():
    if not os.path.exists(refapi.PORT_SO):
        pytest.skip("oracle/libmpros_oracle.so not built (make -C oracle port)")
    return portapi.PortLib()


def port_map_for(plib, name, ds, goal, num_steer, gold_info):
    m = refapi.load_map(name)
    params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"], ds=ds, num_steer=num_steer)
    pm = portapi.PortMap(plib, params)
    d = gold_info
    # geometry exactly as NavMap stored it (float32 resolution, tf2 yaw)
    pm.set_occupancy(m["occupancy"], d[1], [d[2], d[3], d[4]], exact_geometry=True)
    if goal is not None:
        pm.update_goal(goal[0], goal[1])
    return pm, params


def unpack(bits, shape):
    return np.unpackbits(bits)[: shape[0] * shape[1]].reshape(shape)


@pytest.mark.parametrize("name,ds,goal,num_steer", gi.MAP_CASES)
def test_layers_and_verdicts_match_golden(plib, name, ds, goal, num_steer):
    z = np.load(os.path.join(GOLD, "layers_%s_ds%d.npz" % (name, ds)))
    pm, params = port_map_for(plib, name, ds, goal, num_steer, z["info_d"])
    inf = pm.info()
    H, W, H0, W0 = [int(v) for v in z["info_i"][:4]]
    assert (inf["H"], inf["W"], inf["H_orig"], inf["W_orig"]) == (H, W, H0, W0)
    assert inf["res"] == z["info_d"][0] and inf["sin"] == z["info_d"][5] and inf["cos"] == z["info_d"][6]
    # bit-exact integer layers
    assert np.array_equal(pm.layer("static_orig"), unpack(z["static_orig"], (H0, W0)))
    assert np.array_equal(pm.layer("static"), unpack(z["static"], (H, W)))
    assert np.array_equal(pm.layer("inflate_orig"), unpack(z["inflate_orig"], (H0, W0)))
    assert np.array_equal(pm.layer("goal"), z["goal"].astype(np.int32))
    assert np.array_equal(pm.layer("obs_dist"), z["obs_dist"].astype(np.int32))
    # Voronoi contract: obstacle labels exact; free-cell region is a nearest region (canonical tie-break), agreement reported
    occ = pm.layer("static").astype(bool)
    assert np.array_equal(pm.layer("region")[occ], z["region"][occ])
    agree = float((pm.layer("region")[~occ] == z["region"][~occ]).mean()) if (~occ).any() else 1.0
    assert agree > 0.90
    same_in = (pm.layer("edge_dist") == z["edge_dist"].astype(np.int32)) & (pm.layer("isedge") == unpack(z["isedge"], (H, W)))
    assert np.array_equal(pm.layer("voronoi")[same_in], z["voronoi"][same_in])  # float cost bitwise given equal inputs
    # C1 verdicts
    poses = gi.collision_poses({"H_orig": H0, "W_orig": W0, "res_orig": z["info_d"][1], "ox": z["info_d"][2], "oy": z["info_d"][3],
                                "cos": z["info_d"][6], "sin": z["info_d"][5]})
    want = np.unpackbits(z["verdicts"])[: len(poses)]
    got = portapi.PortPlanner(pm, params).collision4(poses)
    assert np.array_equal(got, want)
    pm.close()


def test_rs_generator_matches_golden(plib):
    z = np.load(os.path.join(GOLD, "rs.npz"))
    s, g = gi.rs_pairs()
    for k in range(len(s)):
        seg, cost, traj = portapi.rs_solve(plib, gi.RS_CFG, s[k], g[k])
        assert len(seg) == z["nseg"][k] and np.array_equal(seg, z["segs"][k, : len(seg)])
        assert cost == z["cost"][k] and len(traj) == z["npts"][k]
        if len(traj):
            assert np.array_equal(traj[-1], z["endpose"][k]) and np.array_equal(traj.sum(axis=0), z["traj_sum"][k])
    a, b = gi.distance_pairs()
    d = np.array([portapi.rs_distance(plib, 8.0, a[k], b[k]) for k in range(len(a))])
    assert np.array_equal(d, z["dist"])


def test_ompl_standin_against_reference_generator(plib):
    """The only available pin of the un-vendored OMPL distance: with zero penalties the
    reference's own RS generator must never be shorter than the shortest RS path, and equal it wherever its word
    actually reaches the goal."""
    rng = np.random.default_rng(9)
    n = 4000
    s, g = gi.rs_pairs(n, seed=17, reach=25.0)
    cfg = [8.0, 0.35, 2.0, 0.0, 0.0, 0.0, 0.0]
    agree = reached = 0
    for k in range(n):
        seg, cost, traj = portapi.rs_solve(plib, cfg, s[k], g[k])
        d = portapi.rs_distance(plib, 8.0, s[k, :3], g[k])
        length = cost / 1.001
        end = traj[-1]
        ok = abs(end[0] - g[k, 0]) < 1e-6 and abs(end[1] - g[k, 1]) < 1e-6 and abs(np.angle(np.exp(1j * (end[2] - g[k, 2])))) < 1e-6
        if ok:
            reached += 1
            assert length >= d - 1e-7, "generator found a path shorter than the claimed shortest length"
            agree += int(abs(length - d) < 1e-7 * max(1.0, d))
    assert reached > 0.9 * n
    assert agree > 0.95 * reached


@pytest.mark.parametrize("ci", range(len(gi.PLAN_CASES)))
def test_expansion_and_plan_match_golden(plib, ci):
    name, ds, num_steer, start, goal = gi.PLAN_CASES[ci]
    z = np.load(os.path.join(GOLD, "plans.npz"))
    lay = np.load(os.path.join(GOLD, "layers_%s_ds%d.npz" % (name, ds)))
    pm, params = port_map_for(plib, name, ds, goal[:2], num_steer, lay["info_d"])
    pm.set_voronoi(lay["voronoi"])  # the reference's (order-dependent) field: isolates search parity
    pp = portapi.PortPlanner(pm, params)
    parents = z["parents_%d" % ci]
    P = z["prims_%d" % ci].shape[1]
    for k in range(len(parents)):
        prim, child, nc = pp.expand(start, goal, parents[k], P)
        assert np.array_equal(prim, z["prims_%d" % ci][k])
        assert nc == z["nchild_%d" % ci][k] and np.array_equal(child, z["childs_%d" % ci][k, :nc])
    if ci == 2 and os.environ.get("MPROS_FAST_TESTS"):
        return
    q = pp.plan(start, goal)
    ok, g_ref, prims_ref, nodes_ref, plen = z["stats_%d" % ci]
    assert (q["status"] == 1) == bool(ok)
    assert q["goal_g"] == g_ref and q["n_primitives"] == prims_ref and q["num_nodes"] == nodes_ref
    assert np.array_equal(q["path"], z["path_%d" % ci])
    pm.close()


def test_port_matches_live_reference(plib):
    """Where the compiled reference is present: a map not covered by the golden set, compared live."""
    if not os.path.exists(refapi.REF_SO):
        pytest.skip("oracle/_ref not built")
    lib = refapi.RefLib()
    m = refapi.load_map("map_slalom")
    params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"])
    rm = refapi.RefMap(lib, params)
    rm.set_occupancy(m["occupancy"], m["resolution"], m["origin"])
    rm.update_goal(50.0, 5.0)
    inf = rm.info()
    pm = portapi.PortMap(plib, params)
    pm.set_occupancy(m["occupancy"], inf["res_orig"], [inf["ox"], inf["oy"], inf["oyaw"]], exact_geometry=True)
    pm.update_goal(50.0, 5.0)
    for layer in ["static_orig", "static", "inflate_orig", "goal", "obs_dist"]:
        assert np.array_equal(rm.layer(layer), pm.layer(layer)), layer
    poses = gi.collision_poses(inf, n=20000, seed=4)
    assert np.array_equal(refapi.RefPlanner(rm).collision4(poses), portapi.PortPlanner(pm, params).collision4(poses))


def test_footprint_preclassification_rules_hold_for_the_reference_arithmetic():
    """The two rules behind the GPU's two-phase footprint test (collision.cu, map_device.cuh: footprint_preclass), checked on
    the CPU against the restatement of footPrintInCollision4 (itself bit-identical to the reference):
      surely free: no inflated cell within Chebyshev distance floor(reach_axis/res)+1 and no raw cell within
                   floor(reach_corner/res)+1 of the pose's cell, that far inside the map  =>  verdict 0;
      sure hit:    n_check even, symmetric axis, the pose's own cell occupied in the inflated layer and the pose at least
                   2^-22 cells away from every cell edge  =>  verdict 1."""
    import sys

    from scipy import ndimage

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hybrid-astar-planner_b200"))
    import synthetic

    res = 0.1
    occ = synthetic.synthetic_occupancy(size=2048, n_rect=70)
    params = refapi.shipped_params(0.2, 0.68, ds=8)
    pm = portapi.PortMap(portapi.PortLib(), params)
    pm.set_occupancy(occ, res, [0.0, 0.0, 0.0])
    raw, infl = pm.layer("static_orig") != 0, pm.layer("inflate_orig") != 0
    L, W = 4.2, 2.0
    r_axis = int(np.floor((L - W) / 2 / res + 1e-6) + 1)
    r_corner = int(np.floor(np.hypot(L / 2, W / 2) / res + 1e-6) + 1)
    assert (r_axis, r_corner) == (12, 24) and int(np.ceil((L - W) / res)) % 2 == 0
    unsafe = ndimage.maximum_filter(infl, size=2 * r_axis + 1) | ndimage.maximum_filter(raw, size=2 * r_corner + 1)
    unsafe[:r_corner, :] = unsafe[-r_corner:, :] = True
    unsafe[:, :r_corner] = unsafe[:, -r_corner:] = True
    rng = np.random.default_rng(17)
    n = 1 << 20
    ext = occ.shape[0] * res
    poses = np.stack([rng.uniform(0, ext, n), rng.uniform(0, ext, n), rng.uniform(-np.pi, np.pi, n)], axis=1)
    k = n // 4
    poses[:k, :2] = np.round(poses[:k, :2] / res) * res                      # on cell edges: neither rule may fire wrongly
    poses[k:2 * k, :2] = (np.floor(poses[k:2 * k, :2] / res) + 0.5) * res    # cell centres
    want = portapi.PortPlanner(pm).collision4(poses)
    q = poses[:, :2] / res
    cell = np.floor(q).astype(np.int64)
    inside = (cell >= 0).all(1) & (cell < occ.shape[0]).all(1)
    cj, ci = np.clip(cell[:, 0], 0, occ.shape[1] - 1), np.clip(cell[:, 1], 0, occ.shape[0] - 1)
    frac = q - np.floor(q)
    centred = (np.minimum(frac, 1 - frac) >= 2.0 ** -22).all(1)
    free = inside & ~unsafe[ci, cj]
    hit = inside & unsafe[ci, cj] & infl[ci, cj] & centred
    assert free.mean() > 0.3 and hit.mean() > 0.1
    assert not want[free].any(), "%d 'surely free' poses collide for the reference" % int(want[free].sum())
    assert want[hit].all(), "%d 'sure hit' poses are free for the reference" % int((want[hit] == 0).sum())
    pm.close()


def test_restatement_is_pinned_to_the_reference_on_the_headline_config():
    """BASELINE.json configs[4] (synthetic 8192^2 grid, ds = 8, bench.py's seeded queries): with the reference's own Voronoi
    field (rebuilt from the canonical one + the recorded differences) the restatement must reproduce HybridAstar::Plan() of
    the unmodified reference (golden: tests/golden/synthetic_plans.npz, made by make_synthetic_golden.py from oracle/_ref)
    on a sample that includes an iteration-limit query and an open-set-empty one; and the file's own refv_* records (the
    same comparison on all 544 recorded queries, evaluated when the file was made) must show no difference."""
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "hybrid-astar-planner_b200"))
    import bench
    import synthetic

    z = np.load(os.path.join(GOLD, "synthetic_plans.npz"))
    assert np.array_equal(z["refv_status"], z["ref_status"]) and np.array_equal(z["refv_primitives"], z["ref_primitives"])
    found = z["ref_status"] == 1
    assert np.array_equal(z["refv_cost"][found], z["ref_cost"][found])
    assert len(z["ref_idx"]) >= 512 and set(np.nonzero(z["port_status"] != 1)[0]) <= set(z["ref_idx"])
    occ = synthetic.synthetic_occupancy()
    params = bench.shipped_params()
    pm = portapi.PortMap(portapi.PortLib(), params)
    pm.set_occupancy(occ, synthetic.SYN_RES, [0.0, 0.0, 0.0])
    pp = portapi.PortPlanner(pm)
    starts, goals = synthetic.synthetic_queries(int(z["nq"]), pp.collision4, pm.layer("static"), synthetic.SYN_RES * synthetic.SYN_DS,
                                                bench.EXTENT, synthetic.SYN_QUERY_SEED)
    assert float(starts.sum()) == float(z["starts_sum"]) and float(goals.sum()) == float(z["goals_sum"])
    v = pm.layer("voronoi").copy().ravel()
    v[z["voro_idx"]] = z["voro_val"]
    pm.set_voronoi(v)
    idx = z["ref_idx"]
    cheap = np.nonzero(z["ref_primitives"] < 60000)[0]
    sample = list(cheap[:: max(1, len(cheap) // 12)][:12])
    sample.append(int(np.nonzero(z["ref_status"] == 0)[0][0]))   # runs into max_iteration
    sample.append(int(np.nonzero(z["ref_status"] == -1)[0][0]))  # open set empty
    for k in sample:
        q = int(idx[k])
        pm.update_goal(goals[q][0], goals[q][1])
        o = pp.plan(starts[q], goals[q], path_cap=128)
        assert int(o["status"]) == int(z["ref_status"][k]) and int(o["n_primitives"]) == int(z["ref_primitives"][k]), "query %d" % q
        if o["status"] == 1:
            assert o["goal_g"] == z["ref_cost"][k]
            n = int(z["ref_path_len"][k])
            assert int(o["path_len"]) == n and np.array_equal(o["path"][:n], z["ref_paths"][k, :n])
    pm.close()
