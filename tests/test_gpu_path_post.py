"""setPath's PathPoint post-processing incl. upsampleAndPopulatePath (hybrid_a_star.cpp:814-838,
902-1112) through mpros_path_postprocess_batch, against the unmodified reference's getNewPath() / its own functions."""
import numpy as np
import pytest

import golden_inputs as gi
from helpers import make_pair
from oracle import refapi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,ds,num_steer,start,goal", gi.PLAN_CASES)
def test_new_path_matches_reference(reflib, gpu_ctx_factory, name, ds, num_steer, start, goal):
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, name, ds=ds, num_steer=num_steer, goal=goal[:2])
    r = refapi.ref_plan_query(rm, start, goal)
    want = refapi.ref_plan_newpath(rm, start, goal)          # Plan() + getNewPath()
    assert np.array_equal(want, refapi.ref_upsample_path(rm, r["path"]))
    got = ctx.path_postprocess([r["path"]])[0]
    print("%s: raw %d points -> new path %d points (reference %d), %d cusps" % (name, len(r["path"]), len(got), len(want), int(want[:, 6].sum())))
    assert got.shape == want.shape
    assert np.array_equal(got[:, 4:7], want[:, 4:7])         # direc_, steering_, is_cusp_
    assert np.allclose(got, want, rtol=0, atol=1e-9)         # same libm, same operation order: equal to rounding
    rm.close()


def test_batch_of_gpu_plans_matches_reference(reflib, gpu_ctx_factory):
    """Many raw paths at once (one footprint launch per segment round): the GPU planner's own paths for seeded queries."""
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, "parking4", num_steer=5, goal=(10.0, -2.0))
    ctx.upload_voronoi(rm.layer("voronoi"))
    rng = np.random.default_rng(5)
    inf = rm.info()
    rp = refapi.RefPlanner(rm)
    c, s = inf["cos"], inf["sin"]
    u = rng.uniform(0, inf["W_orig"] * inf["res_orig"], 4000)
    v = rng.uniform(0, inf["H_orig"] * inf["res_orig"], 4000)
    cand = np.stack([inf["ox"] + u * c - v * s, inf["oy"] + u * s + v * c, rng.uniform(-np.pi, np.pi, 4000)], 1)
    free = cand[rp.collision4(cand) == 0]
    starts, goals = free[:24], free[24:48]
    res = ctx.plan_batch(starts, goals, path_cap=512)
    found = np.nonzero(res["status"] == 1)[0]
    assert len(found) >= 8
    raws = [res["paths"][k, : res["path_len"][k]] for k in found]
    l0 = ctx.launch_count()
    got = ctx.path_postprocess(raws)
    launches = ctx.launch_count() - l0
    n_seg = 0
    for k, raw in enumerate(raws):
        want = refapi.ref_upsample_path(rm, raw)
        assert got[k].shape == want.shape, "path %d: %d vs %d points" % (k, len(got[k]), len(want))
        assert np.allclose(got[k], want, rtol=0, atol=1e-9)
        assert not rp.collision4(got[k][:, :3]).any() or rp.collision4(want[:, :3]).any()
        n_seg += len(raw)
    print("%d paths, %d raw points -> %d points; %d footprint launches for the whole batch" % (len(raws), n_seg, sum(len(g) for g in got), launches))
    assert launches <= max(len(r) for r in raws)  # rounds, not paths x segments x tries
    # interpolation disabled: lengths, cusps and curvature only
    p2 = dict(params)
    p2["/mpros/planner/interpolation_enable"] = 0
    ctx.planner_config(p2)
    plain = ctx.path_postprocess(raws[:2])
    assert all(len(a) == len(b) for a, b in zip(plain, raws[:2]))
    assert np.allclose(plain[0][:, :3], raws[0][:, :3]) and plain[0][-1, 7] > 0
    ctx.planner_config(params)
    rp.close()
    rm.close()


def test_degenerate_inputs(gpu_ctx_factory, reflib):
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, "reverse_parking", goal=(11.5, 2.1))
    out = ctx.path_postprocess([])
    assert out == []
    one = ctx.path_postprocess([np.array([[3.0, 8.0, 0.0, 0.0, 1.0]])])
    assert one[0].shape == (1, 8)
    rm.close()


@pytest.mark.parametrize("ci", [0, 3])
def test_search_tree_matches_reference(reflib, gpu_ctx_factory, ci):
    """HybridAstar::getSearchTree (hybrid_a_star.cpp:1114-1164) through mpros_plan_search_tree: the same expansions in the
    same order, every (expanded node, primitive) as ceil(step / granularity) leaf segments."""
    name, ds, num_steer, start, goal = gi.PLAN_CASES[ci]
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, name, ds=ds, num_steer=num_steer, goal=goal[:2])
    ctx.upload_voronoi(rm.layer("voronoi"))
    want = refapi.ref_plan_search_tree(rm, start, goal)
    status, cost, got = ctx.plan_search_tree(start, goal)
    r = refapi.ref_plan_query(rm, start, goal)
    print("%s: %d segments (%d primitives x %d leaves)" % (name, len(got), int(r["n_primitives"]), len(got) // max(1, int(r["n_primitives"]))))
    assert status == 1 and abs(cost - r["goal_g"]) <= 1e-9 * r["goal_g"]
    assert got.shape == want.shape and len(got) == int(r["n_primitives"]) * 20
    assert np.allclose(got, want, rtol=0, atol=1e-9)
    rm.close()
