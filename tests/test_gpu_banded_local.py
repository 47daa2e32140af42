"""Every stage of NavMap::setOccupancyMap tiled by row bands (mpros_map_set_occupancy_banded), run with the
in-process communicator: `world` ranks = host threads with one context each on the one GPU of the test box. The banded
algorithms (seam union of the component labelling nav_map.cpp:384-432, carry rows of the two distance transforms
:434-473 / :509-539, two-row look-ahead of findEdge :475-507) must reproduce mpros_map_set_occupancy bit for bit; the
2-GPU NCCL form of the same calls is tests/mgpu_banded_check.py."""
import numpy as np
import pytest

import mpros_b200
from oracle import refapi

pytestmark = pytest.mark.gpu
LAYERS = ["static_orig", "inflate_orig", "static", "obs_dist", "region", "isedge", "edge_dist", "voronoi"]


def _params(ds, infl=1.0):
    p = refapi.shipped_params(0.2, 0.68, ds=ds)
    p["/mpros/map/obstacle_inflation_radius"] = infl
    return p


def _single(occ, res, origin, params, goal=None):
    ctx = mpros_b200.Context(0)
    ctx.map_config(params)
    ctx.set_occupancy(occ, res, origin)
    if goal is not None:
        ctx.update_goal(*goal)
    out = {k: ctx.layer(k) for k in LAYERS}
    out["num_regions"] = ctx.info()["num_regions"]
    ctx.close()
    return out


def _banded(occ, res, origin, params, world, twice=False):
    H, W = occ.shape
    comms = mpros_b200.Comm.local(world)

    def rank_fn(r):
        ctx = mpros_b200.Context(0)
        ctx.map_config(params)
        r0, n = ctx.band_of(comms[r], H)
        for _ in range(2 if twice else 1):
            ctx.set_occupancy_banded(comms[r], occ[r0:r0 + n], W, H, res, origin)
        out = {k: ctx.layer(k) for k in LAYERS}
        out["num_regions"] = ctx.info()["num_regions"]
        out["band"] = (r0, n)
        out["comm"] = comms[r].info()
        ctx.close()
        return out

    outs = mpros_b200.run_local_ranks(world, rank_fn)
    for c in comms:
        c.close()
    return outs


def _compare(want, outs, what):
    for r, got in enumerate(outs):
        for k in LAYERS:
            bad = int((want[k] != got[k]).sum())
            assert bad == 0, "%s: rank %d layer %s differs in %d cells" % (what, r, k, bad)
        assert got["num_regions"] == want["num_regions"], "%s: rank %d region count" % (what, r)


def _random_map(rng, H, W, density):
    occ = np.zeros((H, W), dtype=np.int8)
    n = int(density * H * W / 40)
    for _ in range(n):
        h, w = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        i, j = int(rng.integers(0, H)), int(rng.integers(0, W))
        occ[i:i + h, j:j + w] = 100
    occ[rng.random((H, W)) < 0.002] = 100       # isolated cells: many tiny components
    occ[rng.random((H, W)) < 0.01] = -1         # unknown cells keep the old state (free on a first map)
    return occ


@pytest.mark.parametrize("name,ds", [("reverse_parking", 1), ("parking4", 1), ("ARENA_part", 2), ("reverse_parking", 5), ("parkinglot", 1)])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_banded_layers_equal_single_gpu_on_the_reference_maps(name, ds, world):
    m = refapi.load_map(name)
    params = _params(ds)
    res = float(np.float32(m["resolution"]))
    want = _single(m["occupancy"], res, m["origin"], params)
    outs = _banded(m["occupancy"], res, m["origin"], params, world)
    _compare(want, outs, "%s ds=%d world=%d" % (name, ds, world))


@pytest.mark.parametrize("H,W,ds,world,density", [
    (700, 900, 1, 4, 0.25),     # components that cross several seams, long vertical carries
    (700, 900, 1, 7, 0.02),     # sparse: distances beyond the band height come from other bands
    (64, 2100, 1, 8, 0.10),     # eight-row bands, ragged last word
    (5, 300, 1, 8, 0.20),       # more ranks than rows: empty bands
    (333, 257, 3, 5, 0.15),     # ds does not divide the map
    (1, 64, 1, 2, 0.30),        # a single row
])
def test_banded_layers_equal_single_gpu_on_random_maps(H, W, ds, world, density):
    rng = np.random.default_rng(H * 1000 + W + world)
    occ = _random_map(rng, H, W, density)
    params = _params(ds, infl=0.45)
    want = _single(occ, 0.1, [0.0, 0.0, 0.0], params)
    outs = _banded(occ, 0.1, [0.0, 0.0, 0.0], params, world, twice=True)
    _compare(want, outs, "random %dx%d ds=%d world=%d" % (H, W, ds, world))
    # one band per rank, bands cover the grid exactly once and in order
    rows = [o["band"] for o in outs]
    assert rows[0][0] == 0 and sum(n for _, n in rows) == H
    assert all(rows[k][0] + rows[k][1] == rows[k + 1][0] or rows[k + 1][1] == 0 for k in range(world - 1))
    assert all(o["comm"]["exchanges"] > 0 for o in outs)


def test_one_component_snaking_through_every_band():
    """A serpentine wall visits every band many times and is ONE component; a second one lies inside its last loop."""
    H, W, world = 240, 200, 6
    occ = np.zeros((H, W), dtype=np.int8)
    for k, j in enumerate(range(4, W - 4, 8)):
        occ[4:H - 4, j] = 100
        if k % 2 == 0:
            occ[H - 5, j:j + 9] = 100
        else:
            occ[4, j:j + 9] = 100
    occ[100:104, 198:200] = 100
    params = _params(1, infl=0.0)
    want = _single(occ, 0.1, [0.0, 0.0, 0.0], params)
    assert want["num_regions"] == 2
    outs = _banded(occ, 0.1, [0.0, 0.0, 0.0], params, world)
    _compare(want, outs, "serpentine")


def test_sharded_plans_equal_the_single_context_plans():
    """mpros_plan_batch_sharded over 3 local ranks: every rank ends with the summaries of all queries; paths gathered."""
    m = refapi.load_map("parking4")
    params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"], num_steer=5)
    res = float(np.float32(m["resolution"]))
    rng = np.random.default_rng(5)
    starts = np.tile([-28.0, -1.5, 0.0], (10, 1)) + rng.uniform(-0.2, 0.2, (10, 3)) * [1, 1, 0.2]
    goals = np.tile([10.0, -2.0, 1.57], (10, 1)) + rng.uniform(-0.2, 0.2, (10, 3)) * [1, 1, 0.1]
    ctx = mpros_b200.Context(0)
    ctx.map_config(params)
    ctx.set_occupancy(m["occupancy"], res, m["origin"])
    ctx.planner_config(params)
    want = ctx.plan_batch(starts, goals, path_cap=512)
    ctx.close()
    world = 3
    comms = mpros_b200.Comm.local(world)
    H, W = m["occupancy"].shape

    def rank_fn(r):
        c = mpros_b200.Context(0)
        c.map_config(params)
        r0, n = c.band_of(comms[r], H)
        c.set_occupancy_banded(comms[r], m["occupancy"][r0:r0 + n], W, H, res, m["origin"])
        c.planner_config(params)
        out = c.plan_batch_sharded(comms[r], starts, goals, path_cap=512, gather_paths=True)  # collective: same arguments on all ranks
        own = c.plan_batch_sharded(comms[r], starts, goals, path_cap=512, gather_paths=False)
        c.close()
        out["own"] = own
        return out

    outs = mpros_b200.run_local_ranks(world, rank_fn)
    for c in comms:
        c.close()
    assert (want["status"] == 1).sum() >= 5
    for r, got in enumerate(outs):
        for k in ["status", "cost", "path_len", "iterations", "primitives"]:
            assert np.array_equal(want[k], got[k]), "rank %d: %s differs" % (r, k)
        lo, hi = got["owned"]
        assert np.array_equal(want["paths"], got["paths"]), "rank %d: gathered paths differ" % r
        own = got["own"]
        assert np.array_equal(want["paths"][lo:hi], own["paths"][lo:hi]) and np.array_equal(want["cost"], own["cost"])
        assert not own["paths"][:lo].any() and not own["paths"][hi:].any()  # without the gather only the own block is written
    assert [o["owned"] for o in outs] == [(0, 4), (4, 7), (7, 10)]


def test_shared_tickets_give_the_single_context_plans():
    """mpros_plan_batch_shared over 3 local ranks: all ranks draw query tickets from rank 0's counter and store into rank 0's
    arrays (peer memory between GPUs; here the ranks share one device); results per query equal mpros_plan_batch, two
    batches in a row (the counter is cleared behind a barrier), every rank can read them."""
    m = refapi.load_map("parking4")
    params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"], num_steer=5)
    res = float(np.float32(m["resolution"]))
    rng = np.random.default_rng(9)
    n = 40
    starts = np.tile([-28.0, -1.5, 0.0], (n, 1)) + rng.uniform(-0.3, 0.3, (n, 3)) * [1, 1, 0.2]
    goals = np.tile([10.0, -2.0, 1.57], (n, 1)) + rng.uniform(-0.3, 0.3, (n, 3)) * [1, 1, 0.1]
    ctx = mpros_b200.Context(0)
    ctx.map_config(params)
    ctx.set_occupancy(m["occupancy"], res, m["origin"])
    ctx.planner_config(params)
    want = ctx.plan_batch(starts, goals, path_cap=512)
    want2 = ctx.plan_batch(starts[::-1], goals[::-1], path_cap=512)
    ctx.close()
    world = 3
    comms = mpros_b200.Comm.local(world)

    def rank_fn(r):
        c = mpros_b200.Context(0)
        c.map_config(params)
        c.set_occupancy(m["occupancy"], res, m["origin"])
        c.planner_config(params)
        sh = c.plan_share_create(comms[r], 64, 512)
        a = c.plan_batch_shared(sh, starts, goals, 512)
        b = c.plan_batch_shared(sh, starts[::-1], goals[::-1], 512, want_results=(r != 2))
        c.plan_share_destroy(sh)
        c.close()
        return a, b

    outs = mpros_b200.run_local_ranks(world, rank_fn)
    for c in comms:
        c.close()
    assert (want["status"] == 1).sum() >= n // 2
    for r, (a, b) in enumerate(outs):
        for k in ["status", "cost", "path_len", "iterations", "primitives", "paths"]:
            assert np.array_equal(want[k], a[k]), "rank %d: %s differs" % (r, k)
            if b is not None:
                assert np.array_equal(want2[k], b[k]), "rank %d, second batch: %s differs" % (r, k)
    assert outs[2][1] is None
