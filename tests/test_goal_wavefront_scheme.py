"""CPU model of the scheme behind goal_bfs_cluster_kernel (hybrid-astar-planner_b200/csrc/goal_map.cu): the window is cut into
row bands, every band carries K halo rows copied from its neighbours at an exchange, runs K levels on its own (the halo
decays from the outside, the band stays exact) and writes a distance only for its own rows. The model works on boolean
arrays instead of bit-packed words; it pins the ARGUMENT (exactness of the bands between exchanges, termination vote every
K levels, bands shorter than the halo, empty bands) against a plain breadth-first search with the reference's cap of 999
(nav_map.cpp:279-349). The kernel itself is compared with generateGoalMap on the GPU (tests/test_gpu_*.py)."""
import numpy as np
import pytest

HUGE = 1000


def plain_bfs(blocked, gi, gj):
    H, W = blocked.shape
    dist = np.full((H, W), HUGE, dtype=np.int32)
    seen = blocked.copy()
    front = np.zeros((H, W), dtype=bool)
    front[gi, gj] = True
    seen[gi, gj] = True
    dist[gi, gj] = 0  # the goal cell is 0 even if occupied
    for level in range(1, HUGE):
        nb = np.zeros_like(front)
        nb[1:, :] |= front[:-1, :]
        nb[:-1, :] |= front[1:, :]
        nb[:, 1:] |= front[:, :-1]
        nb[:, :-1] |= front[:, 1:]
        front = nb & ~seen
        if not front.any():
            break
        seen |= front
        dist[front] = level
    return dist


def banded_bfs(blocked, gi, gj, n_bands, halo_max=8):
    H, W = blocked.shape
    band = -(-H // n_bands)
    K = min(halo_max, band)
    ext = band + 2 * K
    dist = np.full((H, W), HUGE, dtype=np.int32)
    # per band: ext rows = K halo + band + K halo; rows that do not exist are blocked for good
    blk, fr = [], []
    for r in range(n_bands):
        b = np.ones((ext, W), dtype=bool)
        f = np.zeros((ext, W), dtype=bool)
        for e in range(ext):
            g = r * band - K + e
            if 0 <= g < H:
                b[e] = blocked[g]
        ge = gi - r * band + K
        if 0 <= ge < ext:
            f[ge, gj] = True
            b[ge, gj] = True
        blk.append(b)
        fr.append(f)
    dist[gi, gj] = 0
    found = [False] * n_bands
    for level in range(1, HUGE):
        for r in range(n_bands):
            f, b = fr[r], blk[r]
            nb = np.zeros_like(f)
            nb[1:, :] |= f[:-1, :]
            nb[:-1, :] |= f[1:, :]
            nb[:, 1:] |= f[:, :-1]
            nb[:, :-1] |= f[:, 1:]
            fresh = nb & ~b
            b |= fresh
            fr[r] = fresh
            my_rows = max(0, min(band, H - r * band))
            own = fresh[K:K + my_rows]
            if own.any():
                found[r] = True
                dist[r * band:r * band + my_rows][own] = level
        if level % K:
            continue
        if not any(found):
            break
        found = [False] * n_bands
        # exchange: the K halo rows on each side from the neighbours' OWN rows (which are exact)
        new_blk = [b.copy() for b in blk]
        new_fr = [f.copy() for f in fr]
        for r in range(n_bands):
            if r > 0:
                new_blk[r][:K] = blk[r - 1][band:band + K]
                new_fr[r][:K] = fr[r - 1][band:band + K]
            if r + 1 < n_bands:
                new_blk[r][K + band:] = blk[r + 1][K:2 * K]
                new_fr[r][K + band:] = fr[r + 1][K:2 * K]
        blk, fr = new_blk, new_fr
    return dist


@pytest.mark.parametrize("shape,n_bands,density", [((40, 37), 8, 0.25), ((5, 60), 8, 0.2), ((130, 90), 16, 0.3), ((64, 64), 8, 0.0),
                                                   ((33, 200), 16, 0.35), ((1, 1), 8, 0.0), ((300, 40), 8, 0.28)])
def test_banded_halo_wavefront_equals_plain_bfs(shape, n_bands, density):
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    H, W = shape
    for trial in range(4):
        blocked = rng.random((H, W)) < density
        gi, gj = int(rng.integers(0, H)), int(rng.integers(0, W))
        want = plain_bfs(blocked, gi, gj)
        got = banded_bfs(blocked, gi, gj, n_bands)
        assert np.array_equal(want, got), "shape %s trial %d: %d mismatches" % (shape, trial, int((want != got).sum()))


def test_cap_at_999_levels():
    # a 1 x 1200 corridor: cells further than 999 steps keep HUGE_INT = 1000 (nav_map.cpp: relax iff neighbour > current + 1)
    blocked = np.zeros((3, 1200), dtype=bool)
    blocked[0] = blocked[2] = True
    want = plain_bfs(blocked, 1, 0)
    got = banded_bfs(blocked, 1, 0, 8)
    assert np.array_equal(want, got) and want[1, 999] == 999 and want[1, 1000] == HUGE
