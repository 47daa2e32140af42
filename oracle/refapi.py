"""TEST INFRASTRUCTURE ONLY — ctypes access to the CPU oracles. Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the product.

Two libraries:
  * oracle/_ref/libmpros_ref.so   — the reference's own unmodified sources (ref_driver.cpp), kind "reference"
  * oracle/libmpros_oracle.so     — the CPU restatement (oracle/port/), kind "port"
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libmpros_ref.so")
PORT_SO = os.path.join(HERE, "libmpros_oracle.so")
MAPS_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden", "maps")

LAYERS = {
    "static_orig": (0, np.int8, True),
    "static": (1, np.int8, False),
    "inflate_orig": (2, np.int8, True),
    "goal": (3, np.int32, False),
    "obs_dist": (4, np.int32, False),
    "region": (5, np.int32, False),
    "isedge": (6, np.uint8, False),
    "edge_dist": (7, np.int32, False),
    "voronoi": (8, np.float32, False),
}


def shipped_params(free_thresh=0.2, occupied_thresh=0.68, ds=1, num_steer=None):
    """Parameter-server content of the shipped launch: planner_config.yaml + vehicle_config.yaml + the map
    YAML loaded into /mpros/map (vehicle.launch:18). The YAML misspells `steering_dsicretization`, so the
    code default num_steer_=3 is what runs (hybrid_a_star.cpp:158); pass num_steer=5 for the full set."""
    p = {
        "/mpros/planner/optimal_path": 0,
        "/mpros/planner/interpolation_enable": 1,
        "/mpros/planner/step_size": 2.0,
        "/mpros/planner/max_iteration": 100000,
        "/mpros/planner/gear_penalty": 10.0,
        "/mpros/planner/reverse_penalty": 10.0,
        "/mpros/planner/steering_penalty": 0.5,
        "/mpros/planner/steering_change_penalty": 0.5,
        "/mpros/planner/analytic_solution_range": 10.0,
        "/mpros/planner/yaw_discretization": 36,
        "/mpros/planner/path_point_distance": 0.1,
        "/mpros/map/downsampling_factor": ds,
        "/mpros/map/obstacle_inflation_radius": 1.0,
        "/mpros/map/free_thresh": free_thresh,
        "/mpros/map/occupied_thresh": occupied_thresh,
        "/mpros/vehicle/max_steering_angle": 0.35,
        "/mpros/vehicle/max_curvature": 0.125,
        "/mpros/vehicle/width": 2.0,
        "/mpros/vehicle/wheelbase": 3.0,
        "/mpros/vehicle/length": 4.2,
        "/mpros/vehicle/center_to_rotate_center_longitudinal": 0.0,
    }
    if num_steer is not None:
        p["/mpros/planner/steering_discretization"] = num_steer
    return p


def load_map(name):
    """Committed occupancy fixture (tests/golden/make_maps.py): dict(occupancy int8 HxW, resolution,
    origin[3], free_thresh, occupied_thresh)."""
    z = np.load(os.path.join(MAPS_DIR, name + ".npz"))
    return {
        "occupancy": np.ascontiguousarray(z["occupancy"]),
        "resolution": float(z["resolution"]),
        "origin": [float(v) for v in z["origin"]],
        "free_thresh": float(z["free_thresh"]),
        "occupied_thresh": float(z["occupied_thresh"]),
    }


def synthetic_occupancy(size=8192, n_rect=1100, seed=20261019, res=0.1):
    """Synthetic map of SURVEY.md §8(d) — the generator lives with the workloads (input data, not oracle logic)."""
    import sys

    pkg = os.path.join(os.path.dirname(HERE), "hybrid-astar-planner_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    import synthetic

    return synthetic.synthetic_occupancy(size, n_rect, seed, res)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefLib:
    """oracle/_ref/libmpros_ref.so — the unmodified reference behind ref_driver.cpp."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(path)
        L.ref_param_set.argtypes = [C.c_char_p, C.c_double]
        L.ref_map_create.restype = C.c_void_p
        L.ref_map_destroy.argtypes = [C.c_void_p]
        L.ref_map_set_occupancy.restype = C.c_double
        L.ref_map_set_occupancy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double, C.c_double]
        L.ref_map_update_goal.restype = C.c_double
        L.ref_map_update_goal.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.ref_map_info.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_map_layer.restype = C.c_long
        L.ref_map_layer.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_map_point_in_collision.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_map_points_in_collision.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p]
        L.ref_planner_create.restype = C.c_void_p
        L.ref_planner_create.argtypes = [C.c_void_p]
        L.ref_planner_destroy.argtypes = [C.c_void_p]
        L.ref_planner_collision4.restype = C.c_double
        L.ref_planner_collision4.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.ref_planner_initialize.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_planner_expand.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_planner_heuristic.restype = C.c_double
        L.ref_planner_heuristic.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.ref_plan_query.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_rs_solve.argtypes = [C.c_void_p] * 4 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_rs_distance.restype = C.c_double
        L.ref_rs_distance.argtypes = [C.c_double, C.c_void_p, C.c_void_p]
        self.L = L

    def set_params(self, params):
        self.L.ref_param_clear()
        for k, v in params.items():
            self.L.ref_param_set(k.encode(), float(v))


class RefMap:
    def __init__(self, lib, params):
        self.lib = lib
        lib.set_params(params)
        self.h = lib.L.ref_map_create()
        self.t_set = None

    def close(self):
        if self.h:
            self.lib.L.ref_map_destroy(self.h)
            self.h = None

    def set_occupancy(self, occ, res, origin):
        occ = np.ascontiguousarray(occ, dtype=np.int8)
        H, W = occ.shape
        self.t_set = self.lib.L.ref_map_set_occupancy(self.h, occ.ctypes.data, W, H, res, origin[0], origin[1], origin[2])
        return self.t_set

    def update_goal(self, gx, gy):
        return self.lib.L.ref_map_update_goal(self.h, gx, gy)

    def info(self):
        oi = np.zeros(6, dtype=np.int32)
        od = np.zeros(7, dtype=np.float64)
        self.lib.L.ref_map_info(self.h, oi.ctypes.data, od.ctypes.data)
        keys_i = ["H", "W", "H_orig", "W_orig", "get_map", "get_goal"]
        keys_d = ["res", "res_orig", "ox", "oy", "oyaw", "sin", "cos"]
        d = {k: int(v) for k, v in zip(keys_i, oi)}
        d.update({k: float(v) for k, v in zip(keys_d, od)})
        return d

    def points_in_collision(self, name, ij):
        """NavMap::pointInCollision ('static') / pointInCollisionOrig ('static_orig') / pointInCollisionINFLOrig
        ('inflate_orig') for a batch of cells ij = [[i, j], ...] (nav_map.cpp:249-277)."""
        ij = np.ascontiguousarray(ij, dtype=np.int32).reshape(-1, 2)
        out = np.zeros(len(ij), dtype=np.uint8)
        self.lib.L.ref_map_points_in_collision(self.h, LAYERS[name][0], ij.ctypes.data, len(ij), out.ctypes.data)
        return out

    def layer(self, name):
        lid, dt, orig = LAYERS[name]
        n = self.lib.L.ref_map_layer(self.h, lid, None)
        out = np.zeros(n, dtype=dt)
        if n:
            self.lib.L.ref_map_layer(self.h, lid, out.ctypes.data)
        inf = self.info()
        shape = (inf["H_orig"], inf["W_orig"]) if orig else (inf["H"], inf["W"])
        return out.reshape(shape) if n == shape[0] * shape[1] else out


def ref_layer_msg(rmap, name):
    """msg.data (int8, flat) and (height, width, resolution) of NavMap::get*MapMsg (nav_map.cpp:654-716) for the layer
    'static', 'goal', 'voronoi', 'static_orig' or 'inflate_orig'."""
    lid = LAYERS[name][0]
    meta = np.zeros(3, dtype=np.float64)
    L = rmap.lib.L
    L.ref_map_layer_msg.restype = C.c_long
    n = L.ref_map_layer_msg(C.c_void_p(rmap.h), C.c_int(lid), None, C.c_void_p(meta.ctypes.data))
    out = np.zeros(max(n, 0), dtype=np.int8)
    if n > 0:
        L.ref_map_layer_msg(C.c_void_p(rmap.h), C.c_int(lid), C.c_void_p(out.ctypes.data), C.c_void_p(meta.ctypes.data))
    return out, (int(meta[0]), int(meta[1]), float(meta[2]))


class RefPlanner:
    """Configured planner bound to a RefMap (uses the parameter store as set when created)."""

    def __init__(self, rmap):
        self.map = rmap
        self.lib = rmap.lib
        self.h = self.lib.L.ref_planner_create(rmap.h)

    def close(self):
        if self.h:
            self.lib.L.ref_planner_destroy(self.h)
            self.h = None

    def collision4(self, poses):
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(poses), dtype=np.uint8)
        self.t_collision = self.lib.L.ref_planner_collision4(self.h, poses.ctypes.data, len(poses), out.ctypes.data)
        return out

    def collision_method(self, method, poses):
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(poses), dtype=np.uint8)
        self.lib.L.ref_planner_collision_method(C.c_void_p(self.h), C.c_int(method), C.c_void_p(poses.ctypes.data), C.c_long(len(poses)),
                                                C.c_void_p(out.ctypes.data))
        return out

    def initialize(self, start, goal):
        s = np.array(start, dtype=np.float64)
        g = np.array(goal, dtype=np.float64)
        return bool(self.lib.L.ref_planner_initialize(self.h, self.map.h, s.ctypes.data, g.ctypes.data))

    def expand(self, parent6, n_prim):
        p = np.array(parent6, dtype=np.float64)
        prim = np.zeros((n_prim, 7), dtype=np.float64)
        child = np.zeros((n_prim, 10), dtype=np.float64)
        nc = self.lib.L.ref_planner_expand(self.h, p.ctypes.data, prim.ctypes.data, child.ctypes.data)
        return prim, child[:nc]

    def heuristic(self, x, y, yaw):
        return self.lib.L.ref_planner_heuristic(self.h, x, y, yaw)


STAT_KEYS = ["plan_ok", "initialized", "goal_g", "n_primitives", "num_nodes", "path_len", "goal_path_len",
             "t_update_goal", "t_ctor", "t_plan", "t_rs_shot", "t_colli", "status"]


def ref_plan_query(rmap, start, goal, update_goal=True, path_cap=4096):
    s = np.array(start, dtype=np.float64)
    g = np.array(goal, dtype=np.float64)
    stats = np.zeros(len(STAT_KEYS), dtype=np.float64)
    path = np.zeros((path_cap, 5), dtype=np.float64)
    rmap.lib.L.ref_plan_query(rmap.h, s.ctypes.data, g.ctypes.data, int(update_goal), stats.ctypes.data, path.ctypes.data, path_cap)
    d = dict(zip(STAT_KEYS, stats.tolist()))
    d["path"] = path[: int(min(d["path_len"], path_cap))].copy()
    return d


def ref_upsample_path(rmap, path5, cap=8192):
    """setPath's post-processing (hybrid_a_star.cpp:814-838) of a raw path -> getNewPath() rows
    {x, y, yaw, curv, direc, steering, is_cusp, dist}."""
    p = np.ascontiguousarray(path5, dtype=np.float64)
    out = np.zeros((cap, 8), dtype=np.float64)
    rmap.lib.L.ref_upsample_path.restype = C.c_int
    n = rmap.lib.L.ref_upsample_path(C.c_void_p(rmap.h), C.c_void_p(p.ctypes.data), C.c_int(len(p)), C.c_void_p(out.ctypes.data), C.c_int(cap))
    assert n <= cap
    return out[:n].copy()


def ref_plan_newpath(rmap, start, goal, update_goal=True, cap=8192):
    s = np.array(start, dtype=np.float64)
    g = np.array(goal, dtype=np.float64)
    out = np.zeros((cap, 8), dtype=np.float64)
    rmap.lib.L.ref_plan_newpath.restype = C.c_int
    n = rmap.lib.L.ref_plan_newpath(C.c_void_p(rmap.h), C.c_void_p(s.ctypes.data), C.c_void_p(g.ctypes.data), C.c_int(int(update_goal)),
                                    C.c_void_p(out.ctypes.data), C.c_int(cap))
    return out[: min(n, cap)].copy()


def ref_plan_search_tree(rmap, start, goal, update_goal=True, cap=1 << 22):
    """Plan() + getSearchTree() (hybrid_a_star.cpp:1114-1164): [n, 4] {next_x, next_y, curr_x, curr_y}."""
    s = np.array(start, dtype=np.float64)
    g = np.array(goal, dtype=np.float64)
    out = np.zeros((cap, 4), dtype=np.float64)
    L = rmap.lib.L
    L.ref_plan_search_tree.restype = C.c_long
    n = L.ref_plan_search_tree(C.c_void_p(rmap.h), C.c_void_p(s.ctypes.data), C.c_void_p(g.ctypes.data), C.c_int(int(update_goal)),
                               C.c_void_p(out.ctypes.data), C.c_long(cap))
    assert n <= cap
    return out[:n].copy()


def ref_rs_solve(lib, cfg7, start5, goal3, traj_cap=256):
    cfg = np.array(cfg7, dtype=np.float64)
    s = np.array(start5, dtype=np.float64)
    g = np.array(goal3, dtype=np.float64)
    seg = np.zeros((5, 3), dtype=np.float64)
    nseg = C.c_int(0)
    cost = C.c_double(0)
    traj = np.zeros((traj_cap, 5), dtype=np.float64)
    n = lib.L.ref_rs_solve(cfg.ctypes.data, s.ctypes.data, g.ctypes.data, seg.ctypes.data, C.byref(nseg), C.byref(cost), traj.ctypes.data, traj_cap)
    return seg[: nseg.value].copy(), cost.value, traj[: min(n, traj_cap)].copy()


def ref_rs_distance(lib, rho, a3, b3):
    a = np.array(a3, dtype=np.float64)
    b = np.array(b3, dtype=np.float64)
    return lib.L.ref_rs_distance(rho, a.ctypes.data, b.ctypes.data)


def ref_rs_word(lib, word, xyphi):
    """RS::PathFunction::path<word> (1..12, rs_curve.cpp:353-599) for normalised goals [n, 3]: (segs [n, 5, 3], nseg [n])."""
    p = np.ascontiguousarray(xyphi, dtype=np.float64).reshape(-1, 3)
    segs = np.zeros((len(p), 5, 3), dtype=np.float64)
    nseg = np.zeros(len(p), dtype=np.int32)
    lib.L.ref_rs_word(C.c_int(word), C.c_void_p(p.ctypes.data), C.c_long(len(p)), C.c_void_p(segs.ctypes.data), C.c_void_p(nseg.ctypes.data))
    return segs, nseg
