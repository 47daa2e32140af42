"""BASELINE.json configs[2] ("case 3 Arena whole map"): the 2000 x 2000 @0.04 m stand-in of SURVEY.md §8(d) (Arena2036's pgm
is missing from the reference; ARENA_part's image with Arena2036.yaml's thresholds, nearest-neighbour x4), down-sampling 4. Layers, footprint verdicts and whole
plans against the unmodified reference, including a query that runs into max_iteration (failure with the identical
iteration count is the parity target there)."""
import numpy as np
import pytest

from helpers import make_pair, random_poses
from oracle import refapi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def arena(reflib, gpu_ctx_factory):
    part = refapi.load_map("arena_standin")  # ARENA_part's image read with Arena2036.yaml's thresholds (make_maps.py)
    occ = np.ascontiguousarray(np.repeat(np.repeat(part["occupancy"], 4, axis=0), 4, axis=1))
    assert occ.shape == (2000, 2000)
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, "arena2000", ds=4, occupancy=occ, resolution=0.04, origin=[0.0, 0.0, 0.0])
    yield ctx, rm, params, m
    rm.close()


def test_case3_layers_and_verdicts(arena):
    ctx, rm, params, m = arena
    for layer in ["static_orig", "static", "inflate_orig", "obs_dist"]:
        a, b = rm.layer(layer), ctx.layer(layer)
        assert np.array_equal(a, b), "%s: %d mismatches" % (layer, int((a != b).sum()))
    inf = rm.info()
    rp = refapi.RefPlanner(rm)
    rng = np.random.default_rng(3)
    poses = random_poses(rng, 200000, inf)
    k = len(poses) // 4
    poses[:k, :2] = np.round(poses[:k, :2] / 0.04) * 0.04  # n_check = 55 probes per pose at this resolution
    want, got = rp.collision4(poses), ctx.collision_check(poses)
    assert 0.05 < want.mean() < 0.95
    assert np.array_equal(want, got), "%d verdict mismatches" % int((want != got).sum())
    rp.close()


@pytest.mark.parametrize("start,goal,solved", [
    ([5.0, 5.0, 0.0], [70.0, 21.0, 0.0], True),     # SURVEY: 58 151 expansions, Plan() 2.6 s on the CPU
    ([5.0, 75.0, 0.0], [70.0, 9.0, 0.0], False),    # SURVEY: runs into max_iteration = 100 000 (600 006 primitives)
])
def test_case3_plan_search_parity(arena, start, goal, solved):
    ctx, rm, params, m = arena
    rm.update_goal(goal[0], goal[1])
    ctx.update_goal(goal[0], goal[1])
    assert np.array_equal(rm.layer("goal"), ctx.layer("goal"))
    ctx.upload_voronoi(rm.layer("voronoi"))  # the reference's (order-dependent) field: isolates search parity
    r = refapi.ref_plan_query(rm, start, goal)
    g = ctx.plan_batch([start], [goal], path_cap=1024)
    print("case 3 %s -> %s: ref ok=%d g=%.9f prims=%d | gpu status=%d g=%.9f prims=%d iters=%d" % (
        start, goal, r["plan_ok"], r["goal_g"], r["n_primitives"], g["status"][0], g["cost"][0], g["primitives"][0], g["iterations"][0]))
    assert bool(r["plan_ok"]) == solved
    assert (g["status"][0] == 1) == solved
    assert g["primitives"][0] == int(r["n_primitives"]), "expansion count differs from the reference"
    if solved:
        assert abs(g["cost"][0] - r["goal_g"]) <= 1e-9 * abs(r["goal_g"])
        assert g["path_len"][0] == len(r["path"])
        assert np.allclose(g["paths"][0, : len(r["path"])], r["path"], rtol=0, atol=1e-8)
    else:
        assert g["status"][0] == 0 and g["iterations"][0] == 100001


def test_case3_goal_map_full_resolution(gpu_ctx_factory):
    """N4 on the 2000 x 2000 grid at ds = 1 (the largest window the cluster kernel takes: 1999 rows x 63 words, 189 KB of
    shared memory per CTA) against the CPU restatement (bit-identical to the reference, tests/test_oracle_port.py); the
    cooperative grid kernel must agree too."""
    import os

    from oracle import portapi

    part = refapi.load_map("arena_standin")
    occ = np.ascontiguousarray(np.repeat(np.repeat(part["occupancy"], 4, axis=0), 4, axis=1))
    params = refapi.shipped_params(part["free_thresh"], part["occupied_thresh"], ds=1)
    pm = portapi.PortMap(portapi.PortLib(), params)
    pm.set_occupancy(occ, 0.04, [0.0, 0.0, 0.0])
    ctx = gpu_ctx_factory()
    ctx.map_config(params)
    ctx.set_occupancy(occ, 0.04, [0.0, 0.0, 0.0])
    try:
        for gx, gy in [(40.0, 40.0), (5.0, 5.0), (70.0, 21.0), (79.5, 0.5)]:
            pm.update_goal(gx, gy)
            want = pm.layer("goal")
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
        pm.close()
