"""Seeded inputs shared by tests/golden/make_golden.py (which records the reference's outputs) and the tests."""
import numpy as np

# name, ds, goal (x, y), steering_discretization (None = shipped default 3)
MAP_CASES = [
    ("reverse_parking", 1, (11.5, 2.1), None),
    ("parallel_parking3", 1, (10.7, 4.0), None),
    ("parking4", 1, (10.0, -2.0), 5),
    ("ARENA_part", 1, (-20.0, -1.5), None),
    ("ARENA_part", 4, (5.0, 5.0), None),
    ("reverse_parking", 5, (11.5, 2.1), None),
    ("parkinglot", 1, (50.0, 50.0), None),
]

# name, ds, steering_discretization, start, goal  (cases 0, 1, 5)
PLAN_CASES = [
    ("reverse_parking", 1, None, [17.0, 8.0, 3.14], [11.5, 2.1, 1.57]),
    ("reverse_parking", 1, None, [3.0, 8.0, 0.0], [11.5, 2.1, 1.57]),
    ("parallel_parking3", 1, None, [3.0, 8.5, 0.0], [10.7, 4.0, 0.0]),
    ("parking4", 1, 5, [-28.0, -1.5, 0.0], [10.0, -2.0, 1.57]),
    ("parking4", 1, 5, [-28.0, -1.5, 0.0], [8.0, -2.0, -1.57]),
]

# RSCurve as the planner configures it: {radius, steer angle, step, gear, backward, slot3, slot4} (hybrid_a_star.cpp:142-170)
RS_CFG = [8.0, 0.35, 2.0, 10.0, 10.0, 0.5, 0.5]


def collision_poses(info, n=40000, seed=1, margin=3.0):
    """Uniform poses over the (possibly rotated) map rectangle plus a margin; a quarter snapped to the cell lattice,
    an eighth with axis-aligned yaws (cell-edge cases)."""
    rng = np.random.default_rng(seed)
    H, W, res = info["H_orig"], info["W_orig"], info["res_orig"]
    c, s = info["cos"], info["sin"]
    u = rng.uniform(-margin, W * res + margin, n)
    v = rng.uniform(-margin, H * res + margin, n)
    k = n // 4
    u[:k] = np.round(u[:k] / res) * res
    v[:k] = np.round(v[:k] / res) * res
    x = info["ox"] + u * c - v * s
    y = info["oy"] + u * s + v * c
    yaw = rng.uniform(-np.pi, np.pi, n)
    yaw[: k // 2] = rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi, 1.57, 3.14], k // 2)
    return np.stack([x, y, yaw], axis=1)


def rs_pairs(n=4000, seed=3, reach=10.0):
    rng = np.random.default_rng(seed)
    s = np.zeros((n, 5))
    s[:, 0] = rng.uniform(5, 15, n)
    s[:, 1] = rng.uniform(5, 15, n)
    s[:, 2] = rng.uniform(-np.pi, np.pi, n)
    s[:, 3] = rng.choice([1.0, -1.0], n)
    s[:, 4] = rng.choice([0.0, 0.35, -0.35, 0.175, -0.175], n)
    ang = rng.uniform(-np.pi, np.pi, n)
    d = rng.uniform(0.1, reach, n)
    g = np.stack([s[:, 0] + d * np.cos(ang), s[:, 1] + d * np.sin(ang), rng.uniform(-np.pi, np.pi, n)], axis=1)
    return s, g


def distance_pairs(n=4000, seed=5):
    rng = np.random.default_rng(seed)
    a = np.stack([rng.uniform(-30, 30, n), rng.uniform(-30, 30, n), rng.uniform(-np.pi, np.pi, n)], 1)
    b = np.stack([rng.uniform(-30, 30, n), rng.uniform(-30, 30, n), rng.uniform(-np.pi, np.pi, n)], 1)
    b[: n // 10, :2] = a[: n // 10, :2] + rng.uniform(-3, 3, (n // 10, 2))
    return a, b


def expansion_parents(rmap, planner, num_steer, n=40, seed=11):
    """Collision-free parent states {x, y, yaw, steer, dir, g} (the oracle's own footprint test picks them)."""
    inf = rmap.info()
    rng = np.random.default_rng(seed)
    m = 20000
    u = rng.uniform(0, inf["W_orig"] * inf["res_orig"], m)
    v = rng.uniform(0, inf["H_orig"] * inf["res_orig"], m)
    cand = np.stack([inf["ox"] + u * inf["cos"] - v * inf["sin"], inf["oy"] + u * inf["sin"] + v * inf["cos"], rng.uniform(-np.pi, np.pi, m)], 1)
    free = cand[planner.collision4(cand) == 0][:n]
    steer_set = [0.0, 0.35, -0.35, 0.175, -0.175] if num_steer == 5 else [0.0, 0.35, -0.35]
    parents = np.zeros((len(free), 6))
    parents[:, :3] = free
    parents[:, 3] = rng.choice(steer_set, len(free))
    parents[:, 4] = rng.choice([1.0, -1.0], len(free))
    parents[:, 5] = rng.uniform(0, 50, len(free))
    return parents
