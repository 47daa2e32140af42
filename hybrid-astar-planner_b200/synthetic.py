"""Synthetic workloads of SURVEY.md §8(d): the 8192x8192 occupancy grid and the independent start/goal query set.
Pure numpy; shared by bench.py, the tests and (as an input generator only) the CPU oracle arm."""
import numpy as np

SYN_SIZE = 8192
SYN_RES = 0.1
SYN_DS = 8
SYN_MAP_SEED = 20261019
SYN_QUERY_SEED = 20261020


def synthetic_occupancy(size=SYN_SIZE, n_rect=1100, seed=SYN_MAP_SEED, res=SYN_RES):
    """1-m border wall + n_rect axis-aligned rectangles, sides U[2,20] m, centres uniform, clipped; int8 in {0,100}.
    (numpy PCG64 instead of mt19937_64: the array itself is the fixture, every arm consumes the same one.)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    occ = np.zeros((size, size), dtype=np.int8)
    b = int(round(1.0 / res))
    occ[:b, :] = 100
    occ[-b:, :] = 100
    occ[:, :b] = 100
    occ[:, -b:] = 100
    ext = size * res
    cx = rng.uniform(0, ext, n_rect)
    cy = rng.uniform(0, ext, n_rect)
    sx = rng.uniform(2, 20, n_rect)
    sy = rng.uniform(2, 20, n_rect)
    for k in range(n_rect):
        j0 = max(0, int((cx[k] - sx[k] / 2) / res))
        j1 = min(size, int((cx[k] + sx[k] / 2) / res) + 1)
        i0 = max(0, int((cy[k] - sy[k] / 2) / res))
        i1 = min(size, int((cy[k] + sy[k] / 2) / res) + 1)
        occ[i0:i1, j0:j1] = 100
    return occ


def free_components(static_ds):
    """4-connected components of the free cells of the down-sampled static layer (for the reachability condition)."""
    from scipy import ndimage

    lab, _ = ndimage.label(static_ds == 0, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    return lab


def synthetic_queries(n, collision_fn, static_ds, res_ds, extent, seed=SYN_QUERY_SEED):
    """start uniform over poses with footPrintInCollision4 == false, yaw U(-pi, pi]; goal = start + U[20, 80] m in a
    U(-pi, pi] direction, yaw U(-pi, pi]; resampled until inside the map, collision-free and in the start's free
    component (goal-map cost at the start cell < 1000). collision_fn(poses[n,3]) -> uint8[n]."""
    rng = np.random.default_rng(seed)
    comp = free_components(static_ds)
    starts = np.zeros((0, 3))
    goals = np.zeros((0, 3))
    while len(starts) < n:
        m = max(4096, 3 * (n - len(starts)))
        s = np.stack([rng.uniform(0, extent, m), rng.uniform(0, extent, m), rng.uniform(-np.pi, np.pi, m)], 1)
        d = rng.uniform(20.0, 80.0, m)
        a = rng.uniform(-np.pi, np.pi, m)
        g = np.stack([s[:, 0] + d * np.cos(a), s[:, 1] + d * np.sin(a), rng.uniform(-np.pi, np.pi, m)], 1)
        inside = (g[:, 0] > 0) & (g[:, 0] < extent) & (g[:, 1] > 0) & (g[:, 1] < extent)
        s, g = s[inside], g[inside]
        ok = (collision_fn(s) == 0) & (collision_fn(g) == 0)
        s, g = s[ok], g[ok]
        ci = comp[(s[:, 1] / res_ds).astype(int), (s[:, 0] / res_ds).astype(int)]
        cg = comp[(g[:, 1] / res_ds).astype(int), (g[:, 0] / res_ds).astype(int)]
        same = (ci == cg) & (ci > 0)
        starts = np.concatenate([starts, s[same]])
        goals = np.concatenate([goals, g[same]])
    return np.ascontiguousarray(starts[:n]), np.ascontiguousarray(goals[:n])
