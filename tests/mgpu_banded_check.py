"""Multi-GPU check (run under torchrun, one rank per GPU): NavMap::setOccupancyMap tiled by row bands across the ranks
(sharding.set_occupancy_banded: N1 + halo exchange + N3 per band, all-gather, replicated finish) must give layers that are
bit-identical to the single-GPU mpros_map_set_occupancy, and the sharded planner must return the single-GPU results.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_banded_check.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))


def same_plans(ref, got, with_paths=True):
    """status / cost / iterations equal and, for every query, the path rows below its length (rows past path_len are not written)."""
    if not (np.array_equal(ref["status"], got["status"]) and np.array_equal(ref["cost"], got["cost"]) and
            np.array_equal(ref["iterations"], got["iterations"]) and np.array_equal(ref["path_len"], got["path_len"])):
        return False
    if not with_paths:
        return True
    cap = ref["paths"].shape[1]
    valid = np.arange(cap)[None, :] < np.minimum(ref["path_len"], cap)[:, None]
    return np.array_equal(ref["paths"][valid], got["paths"][valid])


def main():
    import torch
    import torch.distributed as dist

    import bench
    import mpros_b200
    import sharding
    import synthetic

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    occ = synthetic.synthetic_occupancy()[:size, :size].copy()
    occ[-10:, :] = 100  # close the cropped map with a wall again
    occ[:, -10:] = 100
    H, W = occ.shape
    params = bench.shipped_params()
    single = mpros_b200.Context(local)
    single.map_config(params)
    t0 = time.perf_counter()
    single.set_occupancy(occ, synthetic.SYN_RES, [0.0, 0.0, 0.0])
    t_single = time.perf_counter() - t0
    banded = mpros_b200.Context(local)
    banded.map_config(params)
    lo, hi = sharding.band_rows(H, rank, world)
    for rep in range(2):
        dist.barrier()
        t0 = time.perf_counter()
        sharding.set_occupancy_banded(banded, occ[lo:hi], H, W, synthetic.SYN_RES, [0.0, 0.0, 0.0], dist=dist)
        t_banded = time.perf_counter() - t0
    bad = 0
    for layer in ["static_orig", "inflate_orig", "static", "obs_dist", "region", "isedge", "edge_dist", "voronoi"]:
        a, b = single.layer(layer), banded.layer(layer)
        if not np.array_equal(a, b):
            bad += 1
            print("rank %d: layer %s differs in %d cells" % (rank, layer, int((a != b).sum())), flush=True)
    # queries sharded across the ranks on the banded map vs all of them on the single-GPU map
    single.planner_config(params)
    banded.planner_config(params)
    starts, goals = synthetic.synthetic_queries(64, single.collision_check, single.layer("static"), synthetic.SYN_RES * synthetic.SYN_DS,
                                                size * synthetic.SYN_RES, synthetic.SYN_QUERY_SEED)
    ref = single.plan_batch(starts, goals, path_cap=128)
    got = sharding.plan_batch_sharded(banded, starts, goals, 128, dist=dist, device="cuda")
    if not (np.array_equal(ref["status"], got["status"]) and np.array_equal(ref["cost"], got["cost"]) and
            np.array_equal(ref["iterations"], got["iterations"])):
        bad += 1
        print("rank %d: sharded plans differ from the single-GPU plans" % rank, flush=True)
    # the same two steps entirely behind the C ABI: mpros_comm over NCCL (ncclCommInitRank from an id broadcast once),
    # every stage of the map banded incl. N5-N9, result gather by the library
    comm = mpros_b200.Comm.from_torch_distributed(banded, dist, "cuda")
    cb = mpros_b200.Context(local)
    cb.map_config(params)
    r0, n = cb.band_of(comm, H)
    t_c = []
    for rep in range(3):
        dist.barrier()
        t0 = time.perf_counter()
        cb.set_occupancy_banded(comm, occ[r0:r0 + n], W, H, synthetic.SYN_RES, [0.0, 0.0, 0.0])
        t_c.append(time.perf_counter() - t0)
    for layer in ["static_orig", "inflate_orig", "static", "obs_dist", "region", "isedge", "edge_dist", "voronoi"]:
        a, b = single.layer(layer), cb.layer(layer)
        if not np.array_equal(a, b):
            bad += 1
            print("rank %d: C-ABI banded layer %s differs in %d cells" % (rank, layer, int((a != b).sum())), flush=True)
    if cb.info()["num_regions"] != single.info()["num_regions"]:
        bad += 1
        print("rank %d: region count differs" % rank, flush=True)
    # a ds = 1 map (N5-N9 at the full resolution: the case the banding is for), cropped so that it stays quick
    p1 = dict(params)
    p1["/mpros/map/downsampling_factor"] = 1
    s1, b1 = mpros_b200.Context(local), mpros_b200.Context(local)
    s1.map_config(p1)
    b1.map_config(p1)
    occ1 = occ[:1536, :1792]
    t0 = time.perf_counter()
    s1.set_occupancy(occ1, synthetic.SYN_RES, [0.0, 0.0, 0.0])
    t_s1 = time.perf_counter() - t0
    r1, n1 = b1.band_of(comm, occ1.shape[0])
    for rep in range(2):
        dist.barrier()
        t0 = time.perf_counter()
        b1.set_occupancy_banded(comm, occ1[r1:r1 + n1], occ1.shape[1], occ1.shape[0], synthetic.SYN_RES, [0.0, 0.0, 0.0])
        t_b1 = time.perf_counter() - t0
    for layer in ["static_orig", "inflate_orig", "static", "obs_dist", "region", "isedge", "edge_dist", "voronoi"]:
        a, b = s1.layer(layer), b1.layer(layer)
        if not np.array_equal(a, b):
            bad += 1
            print("rank %d: ds=1 banded layer %s differs in %d cells" % (rank, layer, int((a != b).sum())), flush=True)
    cb.planner_config(params)
    got2 = cb.plan_batch_sharded(comm, starts, goals, path_cap=128, gather_paths=True)
    if not same_plans(ref, got2):
        bad += 1
        print("rank %d: mpros_plan_batch_sharded differs from the single-GPU plans" % rank, flush=True)
    # dynamic sharding: tickets from rank 0's counter and results into rank 0's arrays over the NVLink mapping (CUDA IPC)
    n_big = 4096
    bs, bg = synthetic.synthetic_queries(n_big, single.collision_check, single.layer("static"), synthetic.SYN_RES * synthetic.SYN_DS,
                                         size * synthetic.SYN_RES, synthetic.SYN_QUERY_SEED + 3)
    ref_big = single.plan_batch(bs, bg, path_cap=128, truncate_ok=True)
    share = cb.plan_share_create(comm, n_big, 128)
    t_shared, t_static = [], []
    for rep in range(3):
        dist.barrier()
        t0 = time.perf_counter()
        got3 = cb.plan_batch_shared(share, bs, bg, 128)
        t_shared.append(time.perf_counter() - t0)
        dist.barrier()
        t0 = time.perf_counter()
        got4 = cb.plan_batch_sharded(comm, bs, bg, path_cap=128, gather_paths=True)
        t_static.append(time.perf_counter() - t0)
    for name, got_x in (("shared tickets", got3), ("fixed blocks", got4)):
        if not same_plans(ref_big, got_x):
            bad += 1
            print("rank %d: %s differ from the single-GPU plans" % (rank, name), flush=True)
    cb.plan_share_destroy(share)
    if rank == 0:
        print("%d queries on %d GPUs, host buffers: shared tickets + peer stores %.1f ms, fixed blocks + gather %.1f ms (best of 3); "
              "iterations per rank with fixed blocks %s" % (n_big, world, 1e3 * min(t_shared), 1e3 * min(t_static),
              [int(ref_big["iterations"][a:b].sum()) for a, b in [sharding.shard_range(n_big, r, world) for r in range(world)]]), flush=True)
    ci = comm.info()
    if rank == 0:
        print("C ABI over NCCL: banded map %.1f ms (best of 3; single %.1f ms), ds=1 %dx%d banded %.1f ms vs single %.1f ms, "
              "%d exchanges, %.1f MB sent by rank 0" % (1e3 * min(t_c), 1e3 * t_single, occ1.shape[0], occ1.shape[1], 1e3 * t_b1, 1e3 * t_s1,
                                                       ci["exchanges"], ci["bytes_sent"] / 1e6), flush=True)
    comm.close()
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("mgpu_banded_check: world %d, map %dx%d, halo exchange + all-gather; single %.1f ms, banded %.1f ms; %s" % (
            world, H, W, 1e3 * t_single, 1e3 * t_banded, "OK" if int(t.item()) == 0 else "MISMATCH"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(t.item()) == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
