"""CPU test of the N>1 host path: world_size-2 gloo, queries sharded across ranks, results all-gathered. The per-rank
planner is stood in by the CPU restatement (test infrastructure) so that only the sharding/gather logic is under test."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QUERIES = [([17.0, 8.0, 3.14], [11.5, 2.1, 1.57]), ([3.0, 8.0, 0.0], [11.5, 2.1, 1.57]), ([16.0, 9.0, 1.0], [11.5, 2.1, 1.57]),
           ([5.0, 9.0, -1.0], [11.5, 2.1, 1.57]), ([10.0, 9.0, 0.5], [11.5, 2.1, 1.57])]


class PortCtx:
    """plan_batch with the signature of mpros_b200.Context.plan_batch, backed by the CPU oracle."""

    def __init__(self):
        from oracle import portapi, refapi

        m = refapi.load_map("reverse_parking")
        params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"])
        self.pm = portapi.PortMap(portapi.PortLib(), params)
        self.pm.set_occupancy(m["occupancy"], m["resolution"], m["origin"])
        self.pp = portapi.PortPlanner(self.pm, params)

    def plan_batch(self, starts, goals, path_cap=64):
        n = len(starts)
        out = {"status": np.zeros(n, np.int32), "cost": np.zeros(n), "path_len": np.zeros(n, np.int32),
               "iterations": np.zeros(n, np.int64), "paths": np.zeros((n, path_cap, 5))}
        for k in range(n):
            self.pm.update_goal(goals[k][0], goals[k][1])
            q = self.pp.plan(starts[k], goals[k])
            out["status"][k], out["cost"][k] = q["status"], q["goal_g"] if q["status"] == 1 else -1.0
            out["path_len"][k], out["iterations"][k] = q["path_len"], q["iterations"]
            m = min(path_cap, len(q["path"]))
            out["paths"][k, :m] = q["path"][:m]
        return out


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))
    import torch.distributed as dist

    import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    starts = np.array([q[0] for q in QUERIES])
    goals = np.array([q[1] for q in QUERIES])
    r = sharding.plan_batch_sharded(PortCtx(), starts, goals, 64, dist=dist)
    ret[rank] = (r["status"].tolist(), r["cost"].tolist(), r["iterations"].tolist(), r["owned"])
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))
    import sharding

    for n in [0, 1, 5, 16384, 16385]:
        for world in [1, 2, 3, 8]:
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[k][1] == spans[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo_matches_single_process():
    if not os.path.exists(os.path.join(ROOT, "oracle", "libmpros_oracle.so")):
        pytest.skip("oracle/libmpros_oracle.so not built")
    sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))
    import sharding

    starts = np.array([q[0] for q in QUERIES])
    goals = np.array([q[1] for q in QUERIES])
    single = sharding.plan_batch_sharded(PortCtx(), starts, goals, 64)
    world = 2
    manager = mp.Manager()
    ret = manager.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        status, cost, iters, owned = ret[rank]
        assert status == single["status"].tolist()
        assert cost == single["cost"].tolist()
        assert iters == single["iterations"].tolist()
    assert ret[0][3] == (0, 3) and ret[1][3] == (3, 5)


# ---- very large maps: row-band tiling of N1 / N3 with halo exchange and all-gather (sharding.set_occupancy_banded) --------
class NumpyBandBackend:
    """The band_* methods of mpros_b200.Context on the CPU (numpy, one int32 per cell) so that the exchange logic of
    set_occupancy_banded (band ranges, halo rows, all-gather placement) is testable without a GPU. N1 / N3 follow
    nav_map.cpp:114-131 / :164-214 literally."""

    def __init__(self, free_thresh, inflation_radius):
        self.free_thresh, self.radius = free_thresh, inflation_radius

    def band_begin(self, W, H, resolution, origin):
        self.H, self.W = H, W
        self.layers = {"static_orig": np.zeros((H, W), np.int32), "inflate_orig": np.zeros((H, W), np.int32)}
        self.r = int(np.ceil(self.radius / float(np.float32(resolution))))
        a = np.arange(-self.r, self.r + 1)
        self.se = ((np.abs(a)[:, None] - 0.5) ** 2 + (np.abs(a)[None, :] - 0.5) ** 2) <= self.r * self.r
        self.finished = False
        return self.r

    def band_threshold(self, rows, row0):
        self.layers["static_orig"][row0:row0 + rows.shape[0]] = (rows > self.free_thresh) | (rows < 0)

    def band_rows(self, layer, row0, nrows):
        import torch

        return torch.from_numpy(self.layers[layer][row0:row0 + nrows])

    def band_inflate(self, row0, nrows):
        raw, out, r = self.layers["static_orig"], self.layers["inflate_orig"], self.r
        for i in range(row0, row0 + nrows):
            acc = np.zeros(self.W, bool)
            for di in range(-r, r + 1):
                if not 0 <= i + di < self.H:
                    continue
                src = raw[i + di].astype(bool)
                for dj in np.nonzero(self.se[di + r])[0] - r:
                    if dj >= 0:
                        acc[dj:] |= src[: self.W - dj]
                    else:
                        acc[:dj] |= src[-dj:]
            out[i] = acc

    def band_finish(self):
        self.finished = True


def _band_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))
    import torch.distributed as dist

    import sharding
    from oracle import refapi

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = refapi.load_map("reverse_parking")
    occ = m["occupancy"]
    H, W = occ.shape
    lo, hi = sharding.band_rows(H, rank, world)
    be = NumpyBandBackend(m["free_thresh"], 1.0)
    sharding.set_occupancy_banded(be, occ[lo:hi], H, W, m["resolution"], m["origin"], dist=dist)
    ret[rank] = (be.layers["static_orig"].astype(np.int8), be.layers["inflate_orig"].astype(np.int8), be.finished, (lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_band_partition_and_halo_plan():
    sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))
    import sharding

    for H in [1, 7, 150, 8192, 8193]:
        for world in [1, 2, 3, 8]:
            spans = [sharding.band_rows(H, r, world) for r in range(world)]
            assert spans[0][0] == 0 and max(b for _, b in spans) == H
            assert all(spans[k][1] == spans[k + 1][0] or spans[k + 1][0] == H for k in range(world - 1))
            for halo in [0, 3, 10, 200]:
                got = {}
                for src, dst, a, b in sharding.halo_transfers(H, world, halo):
                    slo, shi = spans[src]
                    assert slo <= a < b <= shi and src != dst
                    got.setdefault(dst, set()).update(range(a, b))
                for dst, (lo, hi) in enumerate(spans):
                    want = set() if hi <= lo else (set(range(max(0, lo - halo), lo)) | set(range(hi, min(H, hi + halo))))
                    assert got.get(dst, set()) == want


def test_world_size_2_gloo_banded_map_matches_oracle():
    if not os.path.exists(os.path.join(ROOT, "oracle", "libmpros_oracle.so")):
        pytest.skip("oracle/libmpros_oracle.so not built")
    from oracle import portapi, refapi

    m = refapi.load_map("reverse_parking")
    params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"])
    pm = portapi.PortMap(portapi.PortLib(), params)
    pm.set_occupancy(m["occupancy"], m["resolution"], m["origin"])
    want_raw, want_infl = pm.layer("static_orig"), pm.layer("inflate_orig")
    world = 2
    manager = mp.Manager()
    ret = manager.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_band_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        raw, infl, finished, span = ret[rank]
        assert finished and span == ((0, 75) if rank == 0 else (75, 150))
        assert np.array_equal(raw, want_raw), "rank %d: raw layer differs after the all-gather" % rank
        assert np.array_equal(infl, want_infl), "rank %d: inflated layer differs (halo exchange)" % rank
