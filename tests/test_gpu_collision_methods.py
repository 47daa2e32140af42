"""footPrintInCollision / footPrintInCollision2 / footPrintInCollision3 (hybrid_a_star.cpp:432-615, no
caller in the reference) through mpros_collision_check_poses_method, against the unmodified reference."""
import numpy as np
import pytest

from helpers import make_pair, random_poses
from oracle import refapi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,ds", [("reverse_parking", 1), ("parking4", 1), ("ARENA_part", 2)])
def test_collision_methods_1_2_3(reflib, gpu_ctx_factory, name, ds):
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, name, ds=ds)
    inf = rm.info()
    rp = refapi.RefPlanner(rm)
    rng = np.random.default_rng(4)
    poses = random_poses(rng, 20000, inf)
    k = len(poses) // 4
    poses[:k, 0] = np.round(poses[:k, 0] / m["resolution"]) * m["resolution"]
    poses[:k, 1] = np.round(poses[:k, 1] / m["resolution"]) * m["resolution"]
    poses[: k // 2, 2] = rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi, 1.57, 3.14, -3.0, 6.0], k // 2)
    for method in (1, 3):  # integer / bit work behind the pose's cell index: bit-exact
        want = rp.collision_method(method, poses)
        got = ctx.collision_check_method(method, poses)
        assert 0.02 < want.mean() < 0.999
        bad = np.nonzero(want != got)[0]
        assert len(bad) == 0, "method %d: %d mismatches, first pose %s" % (method, len(bad), poses[bad[0]])
    # method 2: a threshold (pi) on a sum of winding angles that is a multiple of 2 pi up to rounding noise; device and host
    # libm differ in the last ulp, which cannot move the verdict unless an obstacle cell centre lies on a footprint edge
    sub = poses[:4000]
    want = rp.collision_method(2, sub)
    got = ctx.collision_check_method(2, sub)
    assert np.array_equal(want, got), "method 2: %d mismatches" % int((want != got).sum())
    assert np.array_equal(ctx.collision_check_method(4, sub), rp.collision4(sub))
    rp.close()
    rm.close()
