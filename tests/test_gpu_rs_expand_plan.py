"""GPU parity (through the C ABI) of the Reeds-Shepp generator, the OMPL-equivalent distance, node expansion and
whole-query planning against the oracles (unmodified reference = `ref`, CPU restatement = `port`)."""
import numpy as np
import pytest

from helpers import make_pair
from oracle import portapi, refapi

pytestmark = pytest.mark.gpu

REL = 1e-4  # north_star tolerance for heuristic values and path cost
TIGHT = 1e-9  # what FP64 libm-vs-libdevice differences actually allow


def rs_cfg(ctx):
    r = ctx.rs_default_config()
    return [r.radius, r.steering_angle, r.step_size, r.pen_gear, r.pen_backward, r.pen_steer_change, r.pen_steer_angle]


def random_rs_pairs(rng, n, reach=10.0):
    s = np.zeros((n, 5))
    s[:, 0] = rng.uniform(5, 15, n)
    s[:, 1] = rng.uniform(5, 15, n)
    s[:, 2] = rng.uniform(-np.pi, np.pi, n)
    s[:, 3] = rng.choice([1.0, -1.0], n)
    s[:, 4] = rng.choice([0.0, 0.35, -0.35, 0.175, -0.175], n)
    ang = rng.uniform(-np.pi, np.pi, n)
    d = rng.uniform(0.1, reach, n)
    g = np.stack([s[:, 0] + d * np.cos(ang), s[:, 1] + d * np.sin(ang), rng.uniform(-np.pi, np.pi, n)], axis=1)
    return s, g


def test_rs_solve_matches_oracles(reflib, gpu_ctx_factory):
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, "reverse_parking", goal=(11.5, 2.1))
    cfg = rs_cfg(ctx)
    # the planner hands RSCurve (gear, backward, STEER_ANGLE, STEER_CHANGE) -> positional slots (hybrid_a_star.cpp:170)
    assert cfg == [8.0, 0.35, 2.0, 10.0, 10.0, 0.5, 0.5]
    rng = np.random.default_rng(3)
    n = 20000
    s, g = random_rs_pairs(rng, n)
    out = ctx.rs_solve_batch(s, g, traj_cap=160)
    plib = portapi.PortLib()
    n_word_diff = 0
    n_bitwise = 0
    for k in range(n):
        seg_p, cost_p, traj_p = portapi.rs_solve(plib, cfg, s[k], g[k])
        if k < 2000:  # the reference itself on a subset (slower: multimap of vectors)
            seg_r, cost_r, traj_r = refapi.ref_rs_solve(reflib, cfg, s[k], g[k])
            assert np.array_equal(seg_r, seg_p) and cost_r == cost_p and np.array_equal(traj_r, traj_p)
        ns = out["nseg"][k]
        same = (ns == len(seg_p) and np.array_equal(out["segs"][k, :ns, 1:], seg_p[:, 1:])
                and np.allclose(out["segs"][k, :ns, 0], seg_p[:, 0], rtol=0, atol=1e-9))
        if not same:
            # another candidate may only win on a (near) tie of the penalised cost: the reference's multimap picks by the
            # last ulp of glibc's transcendentals there
            n_word_diff += 1
            assert abs(out["cost"][k] - cost_p) <= TIGHT * max(1.0, abs(cost_p))
            continue
        assert abs(out["cost"][k] - cost_p) <= 1e-12 * max(1.0, abs(cost_p))
        assert out["npts"][k] == len(traj_p)
        assert np.allclose(out["traj"][k, : len(traj_p)], traj_p, rtol=0, atol=1e-9)
        n_bitwise += int(out["cost"][k] == cost_p and np.array_equal(out["traj"][k, : len(traj_p)], traj_p))
    print("rs_solve: %d shots, %d bitwise identical, %d different word on cost tie" % (n, n_bitwise, n_word_diff))
    # exact-arithmetic ties between formulas (e.g. path3 vs path4) are decided by the last ulp of libm vs libdevice: measured 276 of
    # 20 000 shots (1.4 %), each with |dcost| <= 1e-9 (asserted above)
    assert n_word_diff <= 0.03 * n
    rm.close()


def test_rs_distance_matches_port(gpu_ctx_factory, reflib):
    ctx = gpu_ctx_factory()
    rng = np.random.default_rng(5)
    n = 100000
    a = np.stack([rng.uniform(-30, 30, n), rng.uniform(-30, 30, n), rng.uniform(-np.pi, np.pi, n)], 1)
    b = np.stack([rng.uniform(-30, 30, n), rng.uniform(-30, 30, n), rng.uniform(-np.pi, np.pi, n)], 1)
    b[: n // 10, :2] = a[: n // 10, :2] + rng.uniform(-3, 3, (n // 10, 2))  # short manoeuvres
    got = ctx.rs_distance_batch(8.0, a, b)
    plib = portapi.PortLib()
    want = np.array([portapi.rs_distance(plib, 8.0, a[k], b[k]) for k in range(n)])
    rel = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    print("rs_distance: max rel err %.3g, bitwise equal %.4f" % (rel.max(), float((got == want).mean())))
    assert rel.max() <= 1e-12
    # and the reference wrapper (RSStateSpace over the OMPL stand-in) agrees with the port by construction
    for k in range(200):
        assert refapi.ref_rs_distance(reflib, 8.0, a[k], b[k]) == want[k]


def test_rs_shot_verdicts(reflib, gpu_ctx_factory):
    ctx = gpu_ctx_factory()
    goal = [11.5, 2.1, 1.57]
    rm, params, m = make_pair(reflib, ctx, "reverse_parking", goal=goal[:2])
    rp = refapi.RefPlanner(rm)
    cfg = rs_cfg(ctx)
    rng = np.random.default_rng(7)
    cand = np.stack([rng.uniform(2, 19, 60000), rng.uniform(1, 14, 60000), rng.uniform(-np.pi, np.pi, 60000)], 1)
    cand = cand[(np.hypot(cand[:, 0] - goal[0], cand[:, 1] - goal[1]) < 10.0)]
    free = cand[rp.collision4(cand) == 0][:4000]
    n = len(free)
    assert n >= 1000
    s = np.zeros((n, 5))
    s[:, :3] = free
    s[:, 3] = rng.choice([1.0, -1.0], n)
    s[:, 4] = rng.choice([0.0, 0.35, -0.35], n)
    g = np.tile(np.array(goal), (n, 1))
    out = ctx.rs_solve_batch(s, g, traj_cap=160, shot=True)
    plib = portapi.PortLib()
    mism = 0
    compared = 0
    for k in range(n):
        seg_p, cost_p, traj_p = portapi.rs_solve(plib, cfg, s[k], g[k])
        want = int(rp.collision4(traj_p[:, :3]).any()) if len(traj_p) else 0
        same = (out["nseg"][k] == len(seg_p) and np.array_equal(out["segs"][k, : len(seg_p), 1:], seg_p[:, 1:])
                and np.allclose(out["segs"][k, : len(seg_p), 0], seg_p[:, 0], rtol=0, atol=1e-9))
        if same:
            compared += 1
            mism += int(want != out["verdicts"][k])
    print("rs_shot: %d shots, free fraction %.3f, compared %d, verdict mismatches %d" % (n, 1 - out["verdicts"].mean(), compared, mism))
    assert 0.0 < out["verdicts"].mean() < 1.0
    assert mism == 0 and compared > n // 2
    rp.close()
    rm.close()


@pytest.mark.parametrize("name,num_steer,start,goal", [
    ("reverse_parking", None, [17.0, 8.0, 3.14], [11.5, 2.1, 1.57]),
    ("parking4", 5, [-28.0, -1.5, 0.0], [10.0, -2.0, 1.57]),
])
def test_expand_matches_reference(reflib, gpu_ctx_factory, name, num_steer, start, goal):
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, name, num_steer=num_steer, goal=goal[:2])
    ctx.upload_voronoi(rm.layer("voronoi"))  # isolate E1-E5 from the order-dependent Voronoi field
    inf = rm.info()
    rp = refapi.RefPlanner(rm)
    P = ctx.n_primitives()
    rng = np.random.default_rng(11)
    # parents: collision-free poses
    cand = np.stack([inf["ox"] + rng.uniform(-30, 30, 20000), inf["oy"] + rng.uniform(-5, 50, 20000), rng.uniform(-np.pi, np.pi, 20000)], 1)
    if inf["oyaw"] == 0.0:
        cand[:, 0] = rng.uniform(0, inf["W_orig"] * inf["res_orig"], 20000)
        cand[:, 1] = rng.uniform(0, inf["H_orig"] * inf["res_orig"], 20000)
    free = cand[rp.collision4(cand) == 0][:300]
    assert len(free) >= 50
    steer_set = [0.0, 0.35, -0.35, 0.175, -0.175] if num_steer == 5 else [0.0, 0.35, -0.35]
    parents = np.zeros((len(free), 6))
    parents[:, :3] = free
    parents[:, 3] = rng.choice(steer_set, len(free))
    parents[:, 4] = rng.choice([1.0, -1.0], len(free))
    parents[:, 5] = rng.uniform(0, 50, len(free))
    prims, childs, nchild = ctx.expand_batch(start, goal, parents)
    for k in range(len(free)):
        # the reference planner must be re-initialised per parent to get a fresh closed set
        rp2 = refapi.RefPlanner(rm)
        assert rp2.initialize(start, goal)
        prim_r, child_r = rp2.expand(parents[k], P)
        rp2.close()
        assert np.allclose(prims[k, :, :5], prim_r[:, :5], rtol=0, atol=1e-12)
        assert np.array_equal(prims[k, :, 5], prim_r[:, 5]), "collision verdicts of the primitives differ"
        assert nchild[k] == len(child_r)
        c = childs[k, : nchild[k]]
        assert np.allclose(c[:, :5], child_r[:, :5], rtol=0, atol=1e-12)
        assert np.array_equal(c[:, 8:], child_r[:, 8:])
        assert np.allclose(c[:, 5:8], child_r[:, 5:8], rtol=1e-12, atol=0)  # g, h, f (tolerance budget: 1e-4)
    rp.close()
    rm.close()


PLAN_CASES = [
    # name, ds, num_steer, start, goal  (cases 0, 1, 5 and ARENA_part)
    ("reverse_parking", 1, None, [17.0, 8.0, 3.14], [11.5, 2.1, 1.57]),
    ("reverse_parking", 1, None, [3.0, 8.0, 0.0], [11.5, 2.1, 1.57]),
    ("parallel_parking3", 1, None, [3.0, 8.5, 0.0], [10.7, 4.0, 0.0]),
    ("parking4", 1, 5, [-28.0, -1.5, 0.0], [10.0, -2.0, 1.57]),
    ("parking4", 1, 5, [-28.0, -1.5, 0.0], [8.0, -2.0, -1.57]),
]


@pytest.mark.parametrize("name,ds,num_steer,start,goal", PLAN_CASES)
def test_plan_search_parity_with_reference_voronoi(reflib, gpu_ctx_factory, name, ds, num_steer, start, goal):
    """(i) of SURVEY §9.4: the reference's Voronoi field injected -> isolates search parity. Expansion count, cost and
    the raw path must match the unmodified reference."""
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, name, ds=ds, num_steer=num_steer, goal=goal[:2])
    ctx.upload_voronoi(rm.layer("voronoi"))
    r = refapi.ref_plan_query(rm, start, goal)
    g = ctx.plan_batch([start], [goal], path_cap=512)
    P = ctx.n_primitives()
    print("%s: ref ok=%d g=%.9f prims=%d | gpu status=%d g=%.9f prims=%d iters=%d shots=%d" % (
        name, r["plan_ok"], r["goal_g"], r["n_primitives"], g["status"][0], g["cost"][0], g["primitives"][0], g["iterations"][0], g["rs_shots"][0]))
    assert (g["status"][0] == 1) == bool(r["plan_ok"])
    assert g["primitives"][0] == int(r["n_primitives"]), "expansion count differs from the reference"
    assert abs(g["cost"][0] - r["goal_g"]) <= TIGHT * abs(r["goal_g"])
    assert g["path_len"][0] == len(r["path"])
    assert np.allclose(g["paths"][0, : len(r["path"])], r["path"], rtol=0, atol=1e-8)
    rm.close()


@pytest.mark.parametrize("name,ds,num_steer,start,goal", PLAN_CASES)
def test_plan_full_gpu_vs_reference(reflib, gpu_ctx_factory, name, ds, num_steer, start, goal):
    """(ii) of SURVEY §9.4: everything GPU-generated (canonical Voronoi tie-breaks). Path collision-free for the
    reference's own checker and cost no worse than the reference beyond 1e-4 relative."""
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, name, ds=ds, num_steer=num_steer, goal=goal[:2])
    r = refapi.ref_plan_query(rm, start, goal)
    g = ctx.plan_batch([start], [goal], path_cap=512)
    assert (g["status"][0] == 1) == bool(r["plan_ok"])
    rp = refapi.RefPlanner(rm)
    path = g["paths"][0, : g["path_len"][0]]
    assert not rp.collision4(path[:, :3]).any(), "returned path collides"
    assert np.allclose(path[0, :3], start) and np.allclose(path[-1, :3], goal, atol=0.6)
    rel = (g["cost"][0] - r["goal_g"]) / r["goal_g"]
    print("%s: ref g=%.6f gpu g=%.6f rel %.3g; prims ref %d gpu %d" % (name, r["goal_g"], g["cost"][0], rel, r["n_primitives"], g["primitives"][0]))
    assert rel <= REL
    # and against the restatement, which shares the canonical Voronoi contract: search must be identical
    plib = portapi.PortLib()
    pm = portapi.PortMap(plib, params)
    inf = rm.info()
    pm.set_occupancy(m["occupancy"], inf["res_orig"], [inf["ox"], inf["oy"], inf["oyaw"]], exact_geometry=True)
    pm.update_goal(goal[0], goal[1])
    assert np.array_equal(pm.layer("voronoi"), ctx.layer("voronoi")), "GPU Voronoi field differs from the canonical restatement"
    pp = portapi.PortPlanner(pm)
    q = pp.plan(start, goal)
    assert q["status"] == g["status"][0]
    assert int(q["n_primitives"]) == g["primitives"][0] and int(q["iterations"]) == g["iterations"][0]
    assert int(q["n_rs_shots"]) == g["rs_shots"][0]
    # footprint-check counts include the RS poses tested before the first hit; a shot that ties on cost may sample another
    # (equally blocked) word, so this count is not a strict parity quantity
    assert abs(int(q["n_collision"]) - g["collision_checks"][0]) <= 0.01 * q["n_collision"]
    assert abs(q["goal_g"] - g["cost"][0]) <= TIGHT * abs(q["goal_g"])
    assert np.allclose(path, q["path"], rtol=0, atol=1e-8)
    rp.close()
    rm.close()


def test_cluster_mode_gives_identical_plans(reflib, gpu_ctx_factory, monkeypatch):
    """plan_mode=cluster (mpros_ctx_set_option): a team is a pair of CTAs on the two SMs of a TPC (search CTA + heuristic CTA exchanging the
    children's symmetry images and heuristics through distributed shared memory). Same searches, bit for bit."""
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, "parking4", num_steer=5, goal=(10.0, -2.0))
    rng = np.random.default_rng(9)
    inf = rm.info()
    rp = refapi.RefPlanner(rm)
    c, s = inf["cos"], inf["sin"]
    u = rng.uniform(0, inf["W_orig"] * inf["res_orig"], 6000)
    v = rng.uniform(0, inf["H_orig"] * inf["res_orig"], 6000)
    cand = np.stack([inf["ox"] + u * c - v * s, inf["oy"] + u * s + v * c, rng.uniform(-np.pi, np.pi, 6000)], 1)
    free = cand[rp.collision4(cand) == 0]
    starts, goals = free[:200], free[200:400]
    ctx.set_option("plan_mode", "team")
    a = ctx.plan_batch(starts, goals, path_cap=256)
    ctx.set_option("plan_mode", "cluster")
    b = ctx.plan_batch(starts, goals, path_cap=256)
    # plan_mode=duo: two teams per CTA that keep their iterations in step (shared instruction-cache fills)
    ctx.set_option("plan_mode", "duo")
    d = ctx.plan_batch(starts, goals, path_cap=256)
    d1 = ctx.plan_batch(starts[:1], goals[:1], path_cap=256)  # a lone team: its partner never becomes active
    ctx.set_option("plan_mode", None)
    assert (a["status"] == 1).sum() > 50
    for key in ["status", "cost", "path_len", "iterations", "primitives", "rs_shots", "paths"]:
        assert np.array_equal(a[key], b[key]), key
        assert np.array_equal(a[key], d[key]), "duo: " + key
        if key != "paths":  # rows past path_len are not written
            assert np.array_equal(a[key][:1], d1[key]), "duo, one query: " + key
    n0 = int(a["path_len"][0])
    assert np.array_equal(a["paths"][0, :n0], d1["paths"][0, :n0])
    rp.close()
    rm.close()


def test_expand_batch_device_buffers_match_host_buffers(gpu_ctx_factory):
    """mpros_expand_batch_dev (batch resident in device memory, asynchronous) == mpros_expand_batch."""
    import torch

    import bench
    import synthetic

    ctx = gpu_ctx_factory()
    params = bench.shipped_params()
    ctx.map_config(params)
    ctx.set_occupancy(synthetic.synthetic_occupancy(size=2048, n_rect=70), synthetic.SYN_RES, [0.0, 0.0, 0.0])
    ctx.planner_config(params)
    rng = np.random.default_rng(5)
    poses = np.stack([rng.uniform(0, 204.8, 6000), rng.uniform(0, 204.8, 6000), rng.uniform(-np.pi, np.pi, 6000)], axis=1)
    poses = poses[ctx.collision_check(poses) == 0][:2000]
    goal = poses[0].copy()
    ctx.update_goal(goal[0], goal[1])
    n, P = len(poses), ctx.n_primitives()
    par = np.concatenate([poses, 0.35 * rng.integers(-1, 2, n)[:, None], rng.choice([-1.0, 1.0], n)[:, None], rng.uniform(0, 50, n)[:, None]], axis=1)
    prims, childs, nchild = ctx.expand_batch(poses[1], goal, par)
    par_d = torch.from_numpy(np.ascontiguousarray(par)).cuda()
    pr_d = torch.zeros(n * P * 7, dtype=torch.float64, device="cuda")
    ch_d = torch.full((n * P * 10,), 7.0, dtype=torch.float64, device="cuda")
    nc_d = torch.zeros(n, dtype=torch.int32, device="cuda")
    s3 = np.ascontiguousarray(poses[1])
    ctx._check(ctx.L.mpros_expand_batch_dev(ctx.h, s3.ctypes.data, goal.ctypes.data, par_d.data_ptr(), n, pr_d.data_ptr(), ch_d.data_ptr(),
                                            nc_d.data_ptr()))
    ctx.synchronize()
    assert np.array_equal(nc_d.cpu().numpy(), nchild) and nchild.sum() > n
    assert np.array_equal(pr_d.cpu().numpy().reshape(n, P, 7), prims)
    assert np.array_equal(ch_d.cpu().numpy().reshape(n, P, 10), childs)


def test_expand_batch_compact_records_equal_the_reference_format_children(gpu_ctx_factory):
    """mpros_expand_batch_compact: the inserted children as 80-byte records == the rows of mpros_expand_batch's childs array,
    bit for bit; a buffer that is too small reports the required size."""
    import bench
    import mpros_b200
    import synthetic

    ctx = gpu_ctx_factory()
    params = bench.shipped_params()
    ctx.map_config(params)
    ctx.set_occupancy(synthetic.synthetic_occupancy(size=2048, n_rect=70), synthetic.SYN_RES, [0.0, 0.0, 0.0])
    ctx.planner_config(params)
    rng = np.random.default_rng(6)
    poses = np.stack([rng.uniform(0, 204.8, 9000), rng.uniform(0, 204.8, 9000), rng.uniform(-np.pi, np.pi, 9000)], axis=1)
    poses = poses[ctx.collision_check(poses) == 0][:3001]
    goal = poses[0].copy()
    ctx.update_goal(goal[0], goal[1])
    n = len(poses)
    par = np.concatenate([poses, 0.35 * rng.integers(-1, 2, n)[:, None], rng.choice([-1.0, 1.0], n)[:, None], rng.uniform(0, 50, n)[:, None]], axis=1)
    prims, childs, nchild = ctx.expand_batch(poses[1], goal, par)
    rec, nchild_c = ctx.expand_batch_compact(poses[1], goal, par)
    assert np.array_equal(nchild, nchild_c) and len(rec) == int(nchild.sum()) > n
    at = 0
    for k in range(n):
        r = rec[at:at + nchild[k]]
        at += nchild[k]
        assert np.all(r["parent"] == k) and np.array_equal(r["order"], np.arange(nchild[k]))
        c = childs[k, : nchild[k]]
        for col, name in enumerate(["x", "y", "yaw", "steer"]):
            assert np.array_equal(r[name], c[:, col])
        assert np.array_equal(r["dir"].astype(np.float64), c[:, 4])
        for col, name in [(5, "g"), (6, "h"), (7, "f")]:
            assert np.array_equal(r[name], c[:, col])
        assert np.array_equal(r["i"], c[:, 8].astype(np.int32)) and np.array_equal(r["j"], c[:, 9].astype(np.int32))
        # the primitive a child came from: its end pose is that primitive's row
        assert np.array_equal(prims[k, r["prim"], 0], r["x"])
    with pytest.raises(mpros_b200.MprosError) as e:
        ctx.expand_batch_compact(poses[1], goal, par, cap=100)
    assert "too small" in str(e.value)


def test_rs_shot_batch_device_buffers_match_host_buffers(gpu_ctx_factory):
    """mpros_rs_shot_batch_dev (device buffers, asynchronous, no trajectory) == mpros_rs_shot_batch."""
    import torch

    import bench
    import synthetic

    ctx = gpu_ctx_factory()
    params = bench.shipped_params()
    ctx.map_config(params)
    ctx.set_occupancy(synthetic.synthetic_occupancy(size=2048, n_rect=70), synthetic.SYN_RES, [0.0, 0.0, 0.0])
    ctx.planner_config(params)
    rng = np.random.default_rng(9)
    poses = np.stack([rng.uniform(0, 204.8, 9000), rng.uniform(0, 204.8, 9000), rng.uniform(-np.pi, np.pi, 9000)], axis=1)
    poses = poses[ctx.collision_check(poses) == 0][:3000]
    n = len(poses)
    s5 = np.concatenate([poses, rng.choice([-1.0, 1.0], n)[:, None], 0.35 * rng.integers(-1, 2, n)[:, None]], axis=1)
    d, a = rng.uniform(0.5, 10, n), rng.uniform(-np.pi, np.pi, n)
    g3 = np.stack([poses[:, 0] + d * np.cos(a), poses[:, 1] + d * np.sin(a), rng.uniform(-np.pi, np.pi, n)], axis=1)
    want = ctx.rs_solve_batch(s5, g3, traj_cap=0, shot=True)
    s_d, g_d = torch.from_numpy(np.ascontiguousarray(s5)).cuda(), torch.from_numpy(np.ascontiguousarray(g3)).cuda()
    seg = torch.zeros(n * 15, dtype=torch.float64, device="cuda")
    nseg = torch.zeros(n, dtype=torch.int32, device="cuda")
    cost = torch.zeros(n, dtype=torch.float64, device="cuda")
    npts = torch.zeros(n, dtype=torch.int32, device="cuda")
    ver = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ctx._check(ctx.L.mpros_rs_shot_batch_dev(ctx.h, s_d.data_ptr(), g_d.data_ptr(), n, seg.data_ptr(), nseg.data_ptr(), cost.data_ptr(),
                                             npts.data_ptr(), ver.data_ptr()))
    ctx.synchronize()
    assert np.array_equal(ver.cpu().numpy(), want["verdicts"]) and 0.05 < want["verdicts"].mean() < 0.95
    assert np.array_equal(cost.cpu().numpy(), want["cost"]) and np.array_equal(npts.cpu().numpy(), want["npts"])
    assert np.array_equal(nseg.cpu().numpy(), want["nseg"])


def test_rs_words_equal_reference_path_functions(reflib, gpu_ctx_factory):
    """RS::PathFunction::path1..12 (rs_curve.cpp:353-599) through mpros_rs_word_batch: segment count, steer and direction
    exact, |value| <= 1e-12 relative (NaN where the reference yields NaN: candidates are filtered later, rs_curve.cpp:99-109)."""
    import ctypes as C

    ctx = gpu_ctx_factory()
    rng = np.random.default_rng(21)
    n = 4000
    pts = np.stack([rng.uniform(-6, 6, n), rng.uniform(-6, 6, n), rng.uniform(-np.pi, np.pi, n)], axis=1)
    pts[:50, 1] = 0.0   # degenerate geometry: goals on the axis, zero heading change
    pts[50:100, 2] = 0.0
    ctx.L.mpros_rs_word_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    for word in range(1, 13):
        want_seg, want_n = refapi.ref_rs_word(reflib, word, pts)
        segs = np.zeros((n, 5, 3))
        nseg = np.zeros(n, dtype=np.int32)
        ctx._check(ctx.L.mpros_rs_word_batch(ctx.h, word, pts.ctypes.data, n, segs.ctypes.data, nseg.ctypes.data))
        assert np.array_equal(nseg, want_n), "word %d: segment counts differ" % word
        assert np.array_equal(np.isnan(segs), np.isnan(want_seg)), "word %d: NaN pattern differs" % word
        fin = ~np.isnan(want_seg[:, :, 0])
        # steer / direction: exact unless the value is an exact-arithmetic zero decided by the last ulp (direction flips at 0)
        tiny = np.abs(want_seg[:, :, 0]) < 1e-9
        same_sd = (segs[:, :, 1:] == want_seg[:, :, 1:]).all(axis=2) | tiny | ~fin
        assert same_sd.all(), "word %d: %d steer/direction mismatches" % (word, int((~same_sd).sum()))
        a, b = segs[:, :, 0][fin], want_seg[:, :, 0][fin]
        assert np.allclose(a, b, rtol=1e-12, atol=1e-12), "word %d: max |dvalue| %.3g" % (word, float(np.abs(a - b).max()))
    with pytest.raises(Exception):
        ctx._check(ctx.L.mpros_rs_word_batch(ctx.h, 13, pts.ctypes.data, n, segs.ctypes.data, nseg.ctypes.data))
