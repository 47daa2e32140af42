"""C3 (NavMap::pointInCollision*, nav_map.cpp:249-277) through mpros_map_point_in_collision against the unmodified reference,
and the ownership rules of a context: two contexts on one device share nothing, a failed mpros_planner_config changes
nothing, released workspaces come back."""
import threading

import numpy as np
import pytest

from helpers import make_pair
from oracle import refapi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,ds", [("reverse_parking", 1), ("reverse_parking", 5), ("parking4", 1), ("ARENA_part", 4)])
def test_point_in_collision_three_layers(reflib, gpu_ctx_factory, name, ds):
    """Every cell of the map, a ring of cells outside it (outside the map = collision) and far-away / negative indices, on
    static_map_ (pointInCollision), static_map_orig_ (pointInCollisionOrig) and inflate_map_orig_ (pointInCollisionINFLOrig)."""
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, name, ds=ds)
    inf = rm.info()
    rng = np.random.default_rng(3)
    for layer, (H, W) in (("static", (inf["H"], inf["W"])), ("static_orig", (inf["H_orig"], inf["W_orig"])),
                          ("inflate_orig", (inf["H_orig"], inf["W_orig"]))):
        ii, jj = np.meshgrid(np.arange(-3, H + 3), np.arange(-3, W + 3), indexing="ij")
        ij = np.stack([ii.ravel(), jj.ravel()], axis=1)
        far = np.array([[-1, 0], [0, -1], [H, 0], [0, W], [H - 1, W], [H, W - 1], [-2 ** 31, 5], [5, -2 ** 31], [2 ** 31 - 1, 2 ** 31 - 1],
                        [H * W, 0], [0, H * W]])
        ij = np.concatenate([ij, far, rng.integers(-10 * max(H, W), 10 * max(H, W), (20000, 2))]).astype(np.int32)
        want = rm.points_in_collision(layer, ij)
        got = ctx.point_in_collision(layer, ij)
        assert np.array_equal(want, got), "%s: %d mismatches" % (layer, int((want != got).sum()))
        inside = (ij[:, 0] >= 0) & (ij[:, 0] < H) & (ij[:, 1] >= 0) & (ij[:, 1] < W)
        assert got[~inside].all(), "outside the map must read as collision"
        assert np.array_equal(got[inside], rm.layer(layer)[ij[inside, 0], ij[inside, 1]].astype(np.uint8))
    assert len(ctx.point_in_collision("static", np.zeros((0, 2), dtype=np.int32))) == 0
    with pytest.raises(Exception):
        ctx.point_in_collision("goal", [[0, 0]])  # not a collision layer
    rm.close()


def test_two_contexts_on_one_device_do_not_share_workspaces(reflib, gpu_ctx_factory):
    """Planner workspaces belong to the context: two contexts plan DIFFERENT maps concurrently from two host threads (each on
    its own stream) and both reproduce what they compute alone."""
    a, b = gpu_ctx_factory(), gpu_ctx_factory()
    rma, _, _ = make_pair(reflib, a, "parking4", num_steer=5, goal=(10.0, -2.0))
    rmb, _, _ = make_pair(reflib, b, "reverse_parking", goal=(11.5, 2.1))

    def free_poses(rm, n, seed):
        inf = rm.info()
        rng = np.random.default_rng(seed)
        c, s = inf["cos"], inf["sin"]
        u = rng.uniform(0, inf["W_orig"] * inf["res_orig"], 40 * n)
        v = rng.uniform(0, inf["H_orig"] * inf["res_orig"], 40 * n)
        cand = np.stack([inf["ox"] + u * c - v * s, inf["oy"] + u * s + v * c, rng.uniform(-np.pi, np.pi, 40 * n)], 1)
        rp = refapi.RefPlanner(rm)
        free = cand[rp.collision4(cand) == 0]
        rp.close()
        assert len(free) >= 2 * n
        return free[:n], free[n:2 * n]

    sa, ga = free_poses(rma, 300, 1)
    sb, gb = free_poses(rmb, 300, 2)
    alone_a, alone_b = a.plan_batch(sa, ga, path_cap=512), b.plan_batch(sb, gb, path_cap=512)
    assert (alone_a["status"] == 1).sum() > 30 and (alone_b["status"] == 1).sum() > 30
    out = {}

    def run(tag, ctx, s, g):
        out[tag] = [ctx.plan_batch(s, g, path_cap=512) for _ in range(3)]

    ta, tb = threading.Thread(target=run, args=("a", a, sa, ga)), threading.Thread(target=run, args=("b", b, sb, gb))
    ta.start(), tb.start()
    ta.join(), tb.join()
    for tag, alone in (("a", alone_a), ("b", alone_b)):
        for r in out[tag]:
            for key in ["status", "cost", "path_len", "iterations", "primitives", "rs_shots", "paths"]:
                assert np.array_equal(alone[key], r[key]), (tag, key)
    # released workspaces are re-created by the next call, with the same results
    a.release_workspaces()
    again = a.plan_batch(sa, ga, path_cap=512)
    for key in ["status", "cost", "path_len", "iterations", "paths"]:
        assert np.array_equal(alone_a[key], again[key]), key
    rma.close(), rmb.close()


def test_failed_planner_config_leaves_the_context_untouched(reflib, gpu_ctx_factory):
    """mpros_planner_config validates into a local view and commits on success only (a zero path_point_distance, a zero
    curvature or length <= width used to leave a half-written configuration behind)."""
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, "reverse_parking", goal=(11.5, 2.1))
    rng = np.random.default_rng(5)
    poses = np.stack([rng.uniform(0, 20, 5000), rng.uniform(0, 15, 5000), rng.uniform(-np.pi, np.pi, 5000)], axis=1)
    before = ctx.collision_check(poses)
    plan_before = ctx.plan_batch([[17.0, 8.0, 3.14]], [[11.5, 2.1, 1.57]], path_cap=512)
    for key, bad in (("/mpros/planner/path_point_distance", 0.0), ("/mpros/vehicle/max_curvature", 0.0), ("/mpros/vehicle/length", 1.5),
                     ("/mpros/planner/step_size", float("nan")), ("/mpros/vehicle/max_steering_angle", -0.1),
                     ("/mpros/planner/yaw_discretization", 0), ("/mpros/planner/steering_discretization", 99)):
        p = dict(params)
        p[key] = bad
        with pytest.raises(Exception):
            ctx.planner_config(p)
        assert np.array_equal(ctx.collision_check(poses), before), key
    plan_after = ctx.plan_batch([[17.0, 8.0, 3.14]], [[11.5, 2.1, 1.57]], path_cap=512)
    for key in ["status", "cost", "path_len", "iterations", "paths"]:
        assert np.array_equal(plan_before[key], plan_after[key]), key
    # options: unknown keys / values are rejected, known ones accepted
    with pytest.raises(Exception):
        ctx.set_option("no_such_option", "1")
    with pytest.raises(Exception):
        ctx.set_option("plan_mode", "bogus")
    ctx.set_option("goal_mode", "grid")
    ctx.set_option("goal_mode", None)
    rm.close()
