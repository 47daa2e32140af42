"""BASELINE.json configs[4] at full size: the synthetic 8192 x 8192 grid (ds = 8) and the 16 384-query batch of bench.py.
Exact comparison with the CPU restatement where it finishes in seconds (layers, footprint verdicts, a sample of whole
plans) and size-independent properties on the full batch (determinism, order independence, collision-free paths that join
start and goal, costs bounded below by the straight-line distance)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))

pytestmark = pytest.mark.gpu
NQ = 16384


@pytest.fixture(scope="module")
def syn(gpu_ctx_factory):
    import bench
    import synthetic

    ctx = gpu_ctx_factory()
    params = bench.shipped_params()
    ctx.map_config(params)
    occ = synthetic.synthetic_occupancy()
    ctx.set_occupancy(occ, synthetic.SYN_RES, [0.0, 0.0, 0.0])
    ctx.planner_config(params)
    starts, goals = bench.gen_queries(ctx, NQ, synthetic.SYN_QUERY_SEED)
    return ctx, params, occ, starts, goals


@pytest.fixture(scope="module")
def port(syn):
    from oracle import portapi

    if not os.path.exists(os.path.join(ROOT, "oracle", "libmpros_oracle.so")):
        pytest.skip("oracle/libmpros_oracle.so not built")
    import synthetic

    ctx, params, occ, starts, goals = syn
    pm = portapi.PortMap(portapi.PortLib(), params)
    pm.set_occupancy(occ, synthetic.SYN_RES, [0.0, 0.0, 0.0])
    return pm


def test_layers_bit_exact_8192(syn, port):
    ctx = syn[0]
    for layer in ["static_orig", "inflate_orig", "static", "obs_dist", "region", "isedge", "edge_dist", "voronoi"]:
        a, b = port.layer(layer), ctx.layer(layer)
        assert a.shape == b.shape
        assert np.array_equal(a, b), "%s: %d mismatches" % (layer, int((a != b).sum()))


def test_collision_verdicts_bit_exact_4m_poses(syn, port):
    from oracle import portapi

    import bench

    ctx = syn[0]
    poses = bench.gen_poses(1 << 22, 7)
    k = len(poses) // 4
    poses[:k, :2] = np.round(poses[:k, :2] / 0.1) * 0.1
    # cell centres, cell edges (where the two-phase kernel must hand over to the exact path) and non-finite poses
    poses[k:k + 1000, :2] = (np.floor(poses[k:k + 1000, :2] / 0.1) + 0.5) * 0.1
    poses[k + 1000:k + 1008, 2] = [np.nan, np.inf, -np.inf, 1e300, -1e300, 0.0, np.pi, -np.pi]
    poses[k + 1008:k + 1012, 0] = [np.nan, np.inf, -1e300, 1e12]
    want = portapi.PortPlanner(port).collision4(poses)
    got = ctx.collision_check(poses)
    assert 0.2 < want.mean() < 0.4
    assert np.array_equal(want, got), "%d verdict mismatches of %d" % (int((want != got).sum()), len(poses))
    # the single-phase kernel (every pose through the probe walk) gives the same verdicts
    ctx.set_option("c1_mode", "walk")
    try:
        walk = ctx.collision_check(poses)
        ctx.set_option("c1_mode", "2phase_ldg")  # the two-phase kernel without the bulk-copy staging of the pose stream
        ldg = ctx.collision_check(poses)
    finally:
        ctx.set_option("c1_mode", None)
    assert np.array_equal(want, walk) and np.array_equal(want, ldg)
    # device buffers the bulk copy cannot take: poses 8- but not 16-byte aligned (lanes fill the stages), a verdict buffer
    # that is not 4-byte aligned (falls back to byte stores), ragged sizes around the 128-pose tile
    import torch

    pd = torch.from_numpy(np.concatenate([np.zeros(1), poses.ravel()])).cuda()
    for n_sub, in_off, out_off in ((len(poses), 1, 0), (100003, 1, 1), (127, 4, 0), (128, 4, 0), (129, 4, 2), (1, 1, 0), (1 << 20, 1 + 3 * 77, 0)):
        out = torch.full((n_sub + 8,), 7, dtype=torch.uint8, device="cuda")
        first = (in_off - 1) // 3
        ctx.collision_check_dev(pd.data_ptr() + 8 * in_off, n_sub, out.data_ptr() + out_off)
        torch.cuda.synchronize()
        got_sub = out.cpu().numpy()
        assert np.array_equal(got_sub[out_off:out_off + n_sub], want[first:first + n_sub]), (n_sub, in_off, out_off)
        assert (got_sub[:out_off] == 7).all() and (got_sub[out_off + n_sub:] == 7).all(), "wrote outside the verdict buffer"


def test_sampled_plans_equal_restatement(syn, port):
    from oracle import portapi

    ctx, params, occ, starts, goals = syn
    idx = np.arange(0, NQ, NQ // 12)[:12]
    g = ctx.plan_batch(starts[idx], goals[idx], path_cap=256)
    pp = portapi.PortPlanner(port)
    for k, q in enumerate(idx):
        port.update_goal(goals[q][0], goals[q][1])
        o = pp.plan(starts[q], goals[q])
        assert o["status"] == g["status"][k] and int(o["iterations"]) == g["iterations"][k], "query %d" % q
        assert int(o["n_primitives"]) == g["primitives"][k] and int(o["n_rs_shots"]) == g["rs_shots"][k]
        if o["status"] == 1:
            assert abs(o["goal_g"] - g["cost"][k]) <= 1e-9 * abs(o["goal_g"])
            assert np.allclose(g["paths"][k, : g["path_len"][k]], o["path"], rtol=0, atol=1e-8)


GOLDEN = os.path.join(ROOT, "tests", "golden", "synthetic_plans.npz")


@pytest.fixture(scope="module")
def golden(syn):
    """tests/golden/synthetic_plans.npz (made by tests/golden/make_synthetic_golden.py in the container that holds the
    reference): the restatement on all 16 384 queries, the UNMODIFIED reference on 511 sampled queries plus every non-FOUND
    one, and the cells where the reference's order-dependent Voronoi field differs from the canonical one."""
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/synthetic_plans.npz missing")
    z = np.load(GOLDEN)
    ctx, params, occ, starts, goals = syn
    # the fixture belongs to exactly this query set
    assert int(z["nq"]) == NQ and float(z["starts_sum"]) == float(starts.sum()) and float(z["goals_sum"]) == float(goals.sum())
    return z


def test_all_16384_plans_equal_restatement(syn, golden):
    """EVERY query of the headline batch against the CPU restatement (which the golden file pins to the reference on 544 of
    them): status, iteration count, primitives, Reeds-Shepp shots, path length exact, cost <= 1e-9 relative."""
    ctx, params, occ, starts, goals = syn
    g = ctx.plan_batch(starts, goals, path_cap=128)
    z = golden
    same = (g["status"] == z["port_status"]) & (g["iterations"] == z["port_iterations"]) & (g["primitives"] == z["port_primitives"]) & \
           (g["rs_shots"] == z["port_shots"]) & (g["path_len"] == z["port_path_len"])
    found = (g["status"] == 1) & (z["port_status"] == 1)
    rel = np.zeros(NQ)
    rel[found] = np.abs(g["cost"][found] - z["port_cost"][found]) / np.abs(z["port_cost"][found])
    same &= rel <= 1e-9
    bad = np.nonzero(~same)[0]
    print("headline batch vs restatement: %d of %d queries differ %s; status counts GPU %s" % (
        len(bad), NQ, bad[:16].tolist(), dict(zip(*np.unique(g["status"], return_counts=True)))))
    for q in bad[:8]:
        print("  query %d: status %d/%d iterations %d/%d primitives %d/%d shots %d/%d cost %.9g/%.9g" % (
            q, g["status"][q], z["port_status"][q], g["iterations"][q], z["port_iterations"][q], g["primitives"][q], z["port_primitives"][q],
            g["rs_shots"][q], z["port_shots"][q], g["cost"][q], z["port_cost"][q]))
    # libm vs libdevice: a Reeds-Shepp candidate tie decided by the last ulp picks the other (equal-cost) word, whose sampled
    # path may collide where the first does not; measured on this batch: 2 of 16 384 queries differ, one of them in its status
    # (query 7612: the GPU's shot succeeds at iteration 38 576, the CPU arithmetic runs into the limit) - DESIGN.md 5
    n_status = int((g["status"] != z["port_status"]).sum())
    assert n_status <= 2, "status differs for queries %s" % np.nonzero(g["status"] != z["port_status"])[0][:16]
    assert len(bad) <= 4, "%d queries differ from the restatement" % len(bad)


def _reference_voronoi(ctx, z):
    v = ctx.layer("voronoi").copy().ravel()
    v[z["voro_idx"]] = z["voro_val"]
    return v.reshape(ctx.layer("voronoi").shape)


def test_headline_queries_equal_the_reference_with_its_voronoi_field(syn, golden):
    """SURVEY 9.4 mode (i): the reference's voronoi_value_map_ injected, every sampled query (511 + all 33 non-FOUND ones)
    must reproduce HybridAstar::Plan() of the unmodified reference: status, tree_.size(), goal g, raw path."""
    ctx, params, occ, starts, goals = syn
    z = golden
    canon = ctx.layer("voronoi").copy()
    idx = z["ref_idx"]
    try:
        ctx.upload_voronoi(_reference_voronoi(ctx, z))
        g = ctx.plan_batch(starts[idx], goals[idx], path_cap=128)
    finally:
        ctx.upload_voronoi(canon)
    n_bad = 0
    for k, q in enumerate(idx):
        ok = g["status"][k] == z["ref_status"][k] and g["primitives"][k] == z["ref_primitives"][k]
        if ok and z["ref_status"][k] == 1:
            ok = abs(g["cost"][k] - z["ref_cost"][k]) <= 1e-9 * abs(z["ref_cost"][k]) and g["path_len"][k] == z["ref_path_len"][k]
            n = int(min(z["ref_path_len"][k], z["ref_paths"].shape[1], 128))
            ok = ok and np.allclose(g["paths"][k, :n], z["ref_paths"][k, :n], rtol=0, atol=1e-8)
        if not ok:
            n_bad += 1
            print("  query %d: status %d/%d primitives %d/%d cost %.9g/%.9g" % (q, g["status"][k], z["ref_status"][k], g["primitives"][k],
                                                                                  z["ref_primitives"][k], g["cost"][k], z["ref_cost"][k]))
    print("mode (i), reference Voronoi injected: %d of %d queries differ from the reference (%d FOUND, %d iteration limit, %d open set empty)" % (
        n_bad, len(idx), int((z["ref_status"] == 1).sum()), int((z["ref_status"] == 0).sum()), int((z["ref_status"] == -1).sum())))
    assert int((g["status"] != z["ref_status"]).sum()) <= 1 and n_bad <= 2  # measured: query 7612 only (Reeds-Shepp ulp tie, see above)


def test_headline_queries_full_gpu_against_the_reference(syn, golden):
    """SURVEY 9.4 mode (ii): everything computed on the GPU (canonical Voronoi field: 23 % of this map's free cells sit in
    L1-distance tie areas where the reference's region_ depends on its heap order). Reported as a distribution."""
    ctx, params, occ, starts, goals = syn
    z = golden
    idx = z["ref_idx"]
    g = ctx.plan_batch(starts[idx], goals[idx], path_cap=128)
    both = (g["status"] == 1) & (z["ref_status"] == 1)
    d = (g["cost"][both] - z["ref_cost"][both]) / z["ref_cost"][both]
    edges = [-1, -1e-1, -1e-2, -1e-3, -1e-4, 1e-4, 1e-3, 1e-2, 1e-1, 10]
    hist = np.histogram(d, bins=edges)[0]
    n_status = int((g["status"] != z["ref_status"]).sum())
    print("mode (ii), fully GPU: status mismatches %d of %d; relative cost difference (GPU - reference) / reference: median %.3g, "
          "mean %.3g, min %.3g, max %.3g; histogram over %s: %s; |d| > 5 %% (other homotopy class): %d" % (
              n_status, len(idx), np.median(d), d.mean(), d.min(), d.max(), edges, hist.tolist(), int((np.abs(d) > 0.05).sum())))
    assert n_status <= 3            # 1 measured with the restatement: query 1946 runs into the limit on the canonical field
    assert abs(np.median(d)) < 1e-3 and abs(d.mean()) < 5e-3  # no bias: an equally valid Voronoi field, not a worse planner
    assert (np.abs(d) <= 1e-4).sum() >= 0.15 * len(d)
    # every path the GPU returns is collision-free for the reference's own footprint test (bit-exact verdicts)
    ok = np.nonzero(g["status"] == 1)[0]
    poses = np.concatenate([g["paths"][k, : min(g["path_len"][k], 128), :3] for k in ok])
    assert not ctx.collision_check(poses).any()


def test_full_batch_properties(syn):
    ctx, params, occ, starts, goals = syn
    a = ctx.plan_batch(starts, goals, path_cap=128)
    b = ctx.plan_batch(starts, goals, path_cap=128)
    # determinism: team scheduling is dynamic, results must not depend on it
    for key in ["status", "cost", "path_len", "iterations", "primitives", "rs_shots"]:
        assert np.array_equal(a[key], b[key]), key
    found = a["status"] == 1
    assert found.mean() > 0.99 and (a["status"] == 0).sum() < 64
    # order independence: a shuffled subset planned alone gives the same per-query results
    rng = np.random.default_rng(1)
    sub = rng.permutation(NQ)[:1024]
    c = ctx.plan_batch(starts[sub], goals[sub], path_cap=128)
    assert np.array_equal(c["status"], a["status"][sub]) and np.array_equal(c["cost"], a["cost"][sub])
    assert np.array_equal(c["iterations"], a["iterations"][sub])
    # returned paths: start at the start, end at the goal (the Reeds-Shepp tail lands on it), collision-free for the
    # (bit-exact) footprint test, cost at least the straight-line distance, every limit-hitter ran max_iteration + 1 pops
    assert (a["iterations"][a["status"] == 0] == params["/mpros/planner/max_iteration"] + 1).all()
    ok = np.nonzero(found & (a["path_len"] <= 128))[0]
    first = a["paths"][ok, 0, :3]
    last = a["paths"][ok, a["path_len"][ok] - 1, :3]
    assert np.allclose(first, starts[ok])
    # the reference never checks a Reeds-Shepp candidate against the goal, so the tail lands on the goal
    # only for the words whose formula is right: typical error ~1e-13, a few words miss by metres (never more than the
    # 10 m range a shot is tried from)
    err = np.hypot(last[:, 0] - goals[ok, 0], last[:, 1] - goals[ok, 1])
    print("end-point error: median %.3g, 99th percentile %.3g, max %.3g m" % (np.median(err), np.percentile(err, 99), err.max()))
    assert np.median(err) < 1e-6 and err.max() < 10.0
    d = np.hypot(goals[ok, 0] - starts[ok, 0], goals[ok, 1] - starts[ok, 1])
    assert (a["cost"][ok] >= d - 1e-9).all()
    poses = np.concatenate([a["paths"][q, : a["path_len"][q], :3] for q in ok[:4096]])
    assert not ctx.collision_check(poses).any(), "a returned path collides"


def test_solo_and_team_passes_give_identical_plans(syn):
    """Default planner = warp-per-query pass (squads of 24 queries per CTA whose heuristics the whole CTA evaluates) + team
    pass for the queries that outgrow the small workspaces. Every
    output (status, cost, path, all six statistics) equals the team kernel's alone, bit for bit - with the default solo
    workspaces and with tiny ones that hand most of the batch over."""
    ctx, params, occ, starts, goals = syn
    keys = ["status", "cost", "path_len", "iterations", "primitives", "rs_shots", "collision_checks", "goal_map_cells", "nodes_inserted"]

    def same(a, b, what):
        for key in keys:
            assert np.array_equal(a[key], b[key]), "%s: %s differs for queries %s" % (what, key, np.nonzero(a[key] != b[key])[0][:8])
        valid = np.arange(a["paths"].shape[1])[None, :] < np.minimum(a["path_len"], a["paths"].shape[1])[:, None]
        assert np.array_equal(a["paths"][valid], b["paths"][valid]), what

    n = 4096
    try:
        ctx.set_option("plan_mode", "team")
        team = ctx.plan_batch(starts, goals, path_cap=128)
        ctx.set_option("plan_mode", None)
        solo = ctx.plan_batch(starts, goals, path_cap=128)
        same(team, solo, "default solo workspaces")
        ctx.set_option("plan_policy", "throughput")  # what a batch gets when others are in flight (auto)
        thr = ctx.plan_batch(starts, goals, path_cap=128)
        same(team, thr, "default solo workspaces, throughput policy")
        ctx.set_option("plan_policy", None)
        ctx.set_option("solo_nodes", 512)
        ctx.set_option("solo_min_queries", 0)  # batches below 8192 queries go to the team kernel alone by default
        tiny = ctx.plan_batch(starts[:n], goals[:n], path_cap=128)
        same({k: v[:n] for k, v in team.items()}, tiny, "solo workspaces of 512 nodes")
        ctx.set_option("plan_policy", "throughput")  # thinned-out squads' queries are adopted by idle warps of well-filled squads
        ctx.set_option("plan_age_limit", 2000)       # ... and queries older than this go to the team pass for good
        thr = ctx.plan_batch(starts[:n], goals[:n], path_cap=128)
        same({k: v[:n] for k, v in team.items()}, thr, "solo workspaces of 512 nodes, throughput policy")
        ctx.set_option("plan_policy", None)
        ctx.set_option("plan_age_limit", None)
        ctx.set_option("plan_grow", "0")  # no grown workspaces: every query that outgrows 512 nodes goes to the team pass
        nogrow = ctx.plan_batch(starts[:n], goals[:n], path_cap=128)
        same({k: v[:n] for k, v in team.items()}, nogrow, "solo workspaces of 512 nodes, no grown ones")
        ctx.set_option("plan_grow", None)
        ctx.set_option("solo_nodes", None)
        ctx.set_option("plan_mode", "solo_warp")  # independent warps instead of squads (the form for more than 8 primitives)
        warp = ctx.plan_batch(starts[:n], goals[:n], path_cap=128)
        same({k: v[:n] for k, v in team.items()}, warp, "plan_solo_kernel")
    finally:
        ctx.set_option("plan_mode", None)
        ctx.set_option("solo_nodes", None)
        ctx.set_option("plan_grow", None)
        ctx.set_option("solo_min_queries", None)
        ctx.set_option("plan_policy", None)
        ctx.set_option("plan_age_limit", None)
    it = team["iterations"]
    print("nodes inserted per iteration: mean %.2f, 99th percentile %.2f; queries with more than 32768 nodes: %d" % (
        team["nodes_inserted"].sum() / max(1, it.sum()), np.percentile(team["nodes_inserted"] / np.maximum(it, 1), 99),
        int((team["nodes_inserted"] > 32768).sum())))


def test_batches_in_flight_match_the_synchronous_call(syn):
    """mpros_plan_batch_submit[_dev] / mpros_plan_batch_wait: batches in flight on all eight lanes (own streams and staging
    areas, ONE workspace pool of the context: a team takes a free workspace when its CTA starts) give exactly the per-query
    results of mpros_plan_batch, in any interleaving; eight lanes need no more workspace memory than one."""
    import torch

    import mpros_b200

    ctx, params, occ, starts, goals = syn
    n = 2048
    cap = 64
    assert mpros_b200.PLAN_LANES == 8
    ref = [ctx.plan_batch(starts[k * n:(k + 1) * n], goals[k * n:(k + 1) * n], path_cap=cap, truncate_ok=True) for k in range(8)]
    lanes = list(range(mpros_b200.PLAN_LANES))
    free_after_one_lane = torch.cuda.mem_get_info()[0]

    def bufs(pin):
        mk = (lambda *a, **k: torch.zeros(*a, **k).pin_memory()) if pin else (lambda *a, **k: torch.zeros(*a, device="cuda", **k))
        return {"status": mk(n, dtype=torch.int32), "cost": mk(n, dtype=torch.float64), "len": mk(n, dtype=torch.int32),
                "paths": mk(n * cap * 5, dtype=torch.float64), "stats": mk(n * 6, dtype=torch.int64)}

    # host buffers (page-locked): all lanes submitted before the first wait
    hb = [bufs(True) for _ in lanes]
    hs = [torch.from_numpy(starts[k * n:(k + 1) * n].copy()).pin_memory() for k in range(len(lanes))]
    hg = [torch.from_numpy(goals[k * n:(k + 1) * n].copy()).pin_memory() for k in range(len(lanes))]
    for k, lane in enumerate(lanes):
        b = hb[k]
        ctx.plan_batch_submit(lane, hs[k].data_ptr(), hg[k].data_ptr(), n, b["status"].data_ptr(), b["cost"].data_ptr(),
                              b["len"].data_ptr(), b["paths"].data_ptr(), cap, b["stats"].data_ptr())
    ctx.plan_batch_wait(-1)
    for k in range(len(lanes)):
        b, r = hb[k], ref[k]
        assert np.array_equal(b["status"].numpy(), r["status"]) and np.array_equal(b["cost"].numpy(), r["cost"]), "lane %d" % k
        assert np.array_equal(b["len"].numpy(), r["path_len"])
        got = b["paths"].numpy().reshape(n, cap, 5)
        valid = np.arange(cap)[None, :] < np.minimum(r["path_len"], cap)[:, None]  # rows past path_len are not written
        assert np.array_equal(got[valid], r["paths"][valid])
        assert np.array_equal(b["stats"].numpy().reshape(n, 6)[:, 0], r["iterations"])
    # device buffers, two rounds per lane (the second queues behind the first on the lane's stream), waited per lane
    ds = [t.cuda() for t in hs]
    dg = [t.cuda() for t in hg]
    db = [bufs(False) for _ in lanes]
    torch.cuda.synchronize()
    for rnd in range(2):
        for k, lane in enumerate(lanes):
            b = db[k]
            ctx.plan_batch_submit_dev(lane, ds[k].data_ptr(), dg[k].data_ptr(), n, b["status"].data_ptr(), b["cost"].data_ptr(),
                                      b["len"].data_ptr(), b["paths"].data_ptr(), cap, b["stats"].data_ptr())
    for k, lane in reversed(list(enumerate(lanes))):
        ctx.plan_batch_wait(lane)
        b, r = db[k], ref[k]
        assert np.array_equal(b["status"].cpu().numpy(), r["status"]) and np.array_equal(b["cost"].cpu().numpy(), r["cost"])
        assert np.array_equal(b["stats"].cpu().numpy().reshape(n, 6)[:, 0], r["iterations"])
    assert ctx.plan_lane_stream(1) and ctx.plan_lane_stream(0) == ctx.stream()
    # the lanes added staging areas and the test its buffers (< 1 GB), not workspaces (33 GB per set)
    assert free_after_one_lane - torch.cuda.mem_get_info()[0] < 2 << 30
    with pytest.raises(Exception):
        ctx.plan_batch_wait(mpros_b200.PLAN_LANES)


def test_big_batches_in_flight_match_the_team_kernel(syn):
    """Batches of 8 192 queries on three lanes at once: the first finds the device idle (latency policy of the squad pass),
    the others find batches in flight (throughput policy: queries of thinned-out squads are adopted by other squads, also
    while the other batches' kernels compete for the SMs). Per-query results equal the team kernel's alone."""
    import torch

    ctx, params, occ, starts, goals = syn
    n, cap = 8192, 64
    ctx.set_option("plan_mode", "team")
    try:
        ref = [ctx.plan_batch(starts[k * n:(k + 1) * n], goals[k * n:(k + 1) * n], path_cap=cap, truncate_ok=True) for k in range(2)]
    finally:
        ctx.set_option("plan_mode", None)
    ds = [torch.from_numpy(starts[k * n:(k + 1) * n].copy()).cuda() for k in range(2)]
    dg = [torch.from_numpy(goals[k * n:(k + 1) * n].copy()).cuda() for k in range(2)]
    jobs = [(1, 0), (2, 1), (3, 0)]  # (lane, batch)
    bufs = [{"status": torch.zeros(n, dtype=torch.int32, device="cuda"), "cost": torch.zeros(n, dtype=torch.float64, device="cuda"),
             "len": torch.zeros(n, dtype=torch.int32, device="cuda"), "paths": torch.zeros(n * cap * 5, dtype=torch.float64, device="cuda"),
             "stats": torch.zeros(n * 6, dtype=torch.int64, device="cuda")} for _ in jobs]
    torch.cuda.synchronize()
    for (lane, k), b in zip(jobs, bufs):
        ctx.plan_batch_submit_dev(lane, ds[k].data_ptr(), dg[k].data_ptr(), n, b["status"].data_ptr(), b["cost"].data_ptr(),
                                  b["len"].data_ptr(), b["paths"].data_ptr(), cap, b["stats"].data_ptr())
    ctx.plan_batch_wait(-1)
    for (lane, k), b in zip(jobs, bufs):
        r = ref[k]
        st = b["stats"].cpu().numpy().reshape(n, 6)
        assert np.array_equal(b["status"].cpu().numpy(), r["status"]) and np.array_equal(b["cost"].cpu().numpy(), r["cost"]), "lane %d" % lane
        assert np.array_equal(b["len"].cpu().numpy(), r["path_len"]) and np.array_equal(st[:, 0], r["iterations"]), "lane %d" % lane
        assert np.array_equal(st[:, 5], r["nodes_inserted"]) and np.array_equal(st[:, 3], r["collision_checks"]), "lane %d" % lane
        got = b["paths"].cpu().numpy().reshape(n, cap, 5)
        valid = np.arange(cap)[None, :] < np.minimum(r["path_len"], cap)[:, None]
        assert np.array_equal(got[valid], r["paths"][valid]), "lane %d" % lane


def test_goal_map_cluster_and_grid_kernels_agree(syn, port):
    """N4 at 1024 x 1024: the cluster/DSMEM wavefront (default) and the cooperative grid kernel (goal_mode=grid) both
    reproduce generateGoalMap bit for bit, for goals in the middle, in a corner (window clipped) and on an obstacle."""
    ctx = syn[0]
    static = ctx.layer("static")
    occ_cells = np.argwhere(static != 0)
    goals = [(409.6 + 3.3, 409.6 - 1.7), (1.3, 1.2), (817.9, 818.1), (12.0, 700.0),
             ((occ_cells[len(occ_cells) // 2][1] + 0.5) * 0.8, (occ_cells[len(occ_cells) // 2][0] + 0.5) * 0.8)]
    try:
        for gx, gy in goals:
            port.update_goal(gx, gy)
            want = port.layer("goal")
            ctx.set_option("goal_mode", "cluster")
            ctx.update_goal(gx, gy)
            a = ctx.layer("goal")
            ctx.set_option("goal_mode", "grid")
            ctx.update_goal(gx, gy)
            b = ctx.layer("goal")
            assert np.array_equal(a, want), "cluster kernel, goal (%.1f, %.1f): %d mismatches" % (gx, gy, int((a != want).sum()))
            assert np.array_equal(b, want), "grid kernel, goal (%.1f, %.1f)" % (gx, gy)
    finally:
        ctx.set_option("goal_mode", None)
