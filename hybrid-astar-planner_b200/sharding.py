"""Multi-GPU host logic: independent (start, goal) queries are sharded across ranks (one process per GPU), every rank
holds a replica of the map layers, and the ONLY exchange step is the gather of the fixed-size per-query results
(NCCL over NVLink on the GPU box; gloo in the CPU tests). No communication happens during the search."""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous block of query indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def plan_batch_sharded(ctx, starts, goals, path_cap, dist=None, device="cpu"):
    """Plan the GLOBAL query set cooperatively: this rank plans its shard with ctx.plan_batch, then status / cost /
    path_len / iterations of all shards are all-gathered so every rank returns the full result arrays (paths stay
    on the owning rank: result["paths"] covers the local shard, result["owned"] = (lo, hi))."""
    import torch

    n = len(starts)
    if dist is None or not dist.is_initialized():
        rank, world = 0, 1
    else:
        rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(n, rank, world)
    local = ctx.plan_batch(starts[lo:hi], goals[lo:hi], path_cap=path_cap)
    out = {"owned": (lo, hi), "paths": local["paths"]}
    # fixed-size record per query: pad the local shard to the largest shard so all_gather sees equal shapes
    width = (n + world - 1) // world
    rec = np.zeros((width, 4), dtype=np.float64)
    rec[: hi - lo, 0] = local["status"]
    rec[: hi - lo, 1] = local["cost"]
    rec[: hi - lo, 2] = local["path_len"]
    rec[: hi - lo, 3] = local["iterations"]
    t = torch.from_numpy(rec).to(device)
    if world > 1:
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
    else:
        parts = [t]
    full = np.zeros((n, 4), dtype=np.float64)
    for r, p in enumerate(parts):
        a, b = shard_range(n, r, world)
        full[a:b] = p.cpu().numpy()[: b - a]
    out["status"] = full[:, 0].astype(np.int32)
    out["cost"] = full[:, 1]
    out["path_len"] = full[:, 2].astype(np.int32)
    out["iterations"] = full[:, 3].astype(np.int64)
    return out
