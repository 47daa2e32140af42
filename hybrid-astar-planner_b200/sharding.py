"""Multi-GPU host logic (one process per GPU).

* Queries: independent (start, goal) queries are sharded across ranks, every rank holds a replica of the map layers, and
  the ONLY exchange step is the gather of the fixed-size per-query results. No communication happens during the search.
* Very large maps: NavMap::setOccupancyMap's original-resolution stages (N1 threshold, N3 inflation = the thresholded
  exact distance transform of nav_map.cpp:164-214) are tiled by ROW BANDS; the exchange steps are the halo rows of the
  raw layer N3 needs from the neighbouring bands and the all-gather of the finished bands, after which every rank
  completes the down-sampled stages on its replica (set_occupancy_banded).
Collectives run over NCCL / NVLink on the GPU box and over gloo in the CPU tests."""
import numpy as np


def band_rows(H, rank, world):
    """Rows [lo, hi) of the original grid owned by `rank`: equal bands of ceil(H / world) rows (the last ones may be short
    or empty), so that the all-gather sees equal shapes."""
    per = (H + world - 1) // world
    lo = min(H, rank * per)
    return lo, min(H, lo + per)


def halo_transfers(H, world, halo):
    """Every (src, dst, row_lo, row_hi) the halo exchange needs: rank dst's band needs rows within `halo` of it that rank
    src owns. Bands thinner than the halo make a rank receive from several neighbours."""
    out = []
    for dst in range(world):
        lo, hi = band_rows(H, dst, world)
        if hi <= lo:
            continue
        for want_lo, want_hi in ((max(0, lo - halo), lo), (hi, min(H, hi + halo))):
            for src in range(world):
                if src == dst:
                    continue
                slo, shi = band_rows(H, src, world)
                a, b = max(want_lo, slo), min(want_hi, shi)
                if b > a:
                    out.append((src, dst, a, b))
    return out


def set_occupancy_banded(backend, occ_rows, H, W, resolution, origin, dist=None):
    """NavMap::setOccupancyMap with the original-resolution stages tiled by row bands across the ranks of `dist`.

    backend: the rank's map context — mpros_b200.Context on a GPU (tests use a numpy stand-in with the same methods:
    band_begin, band_threshold, band_rows, band_inflate, band_finish). occ_rows: int8 [hi - lo, W], this rank's rows
    band_rows(H, rank, world) of the occupancy grid (a rank never sees the rest of the grid). Returns (lo, hi)."""
    import torch

    if dist is None or not dist.is_initialized():
        rank, world = 0, 1
    else:
        rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = band_rows(H, rank, world)
    assert occ_rows.shape == (hi - lo, W), "occ_rows must be this rank's band"
    halo = backend.band_begin(W, H, resolution, origin)
    backend.band_threshold(occ_rows, lo)                        # N1 on the band
    if world > 1:
        # halo exchange: raw rows next to the band, straight between the layers' device memory
        ops, keep = [], []
        for src, dst, a, b in halo_transfers(H, world, halo):
            if src == rank:
                t = backend.band_rows("static_orig", a, b - a)
                ops.append(dist.P2POp(dist.isend, t, dst))
                keep.append(t)
            elif dst == rank:
                t = backend.band_rows("static_orig", a, b - a)
                ops.append(dist.P2POp(dist.irecv, t, src))
                keep.append(t)
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if torch.cuda.is_available() and keep and keep[0].is_cuda:
            torch.cuda.synchronize()
    backend.band_inflate(lo, hi - lo)                           # N3 on the band
    if world > 1:
        per = (H + world - 1) // world
        for layer in ("static_orig", "inflate_orig"):
            probe = backend.band_rows(layer, 0, 1)
            send = torch.zeros((per, probe.shape[1]), dtype=probe.dtype, device=probe.device)
            if hi > lo:
                send[: hi - lo] = backend.band_rows(layer, lo, hi - lo)
            parts = [torch.zeros_like(send) for _ in range(world)]
            dist.all_gather(parts, send)
            for r, part in enumerate(parts):
                a, b = band_rows(H, r, world)
                if r != rank and b > a:
                    backend.band_rows(layer, a, b - a).copy_(part[: b - a])
        if torch.cuda.is_available() and send.is_cuda:
            torch.cuda.synchronize()
    backend.band_finish()                                       # N2, N4-N9 on the complete layers
    return lo, hi


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
