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
