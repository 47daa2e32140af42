"""TEST INFRASTRUCTURE ONLY — ctypes access to the CPU restatement (oracle/libmpros_oracle.so, kind "port").
Mirrors refapi's classes so tests can run the same checks against either oracle."""
import ctypes as C
import os

import numpy as np

from . import refapi

LAYERS = refapi.LAYERS
PLANNER_ORDER = [
    ("/mpros/vehicle/max_curvature", 0.4), ("/mpros/vehicle/width", 2.0), ("/mpros/vehicle/wheelbase", 3.0),
    ("/mpros/vehicle/length", 4.2), ("/mpros/vehicle/center_to_rotate_center_longitudinal", 0.0),
    ("/mpros/vehicle/max_steering_angle", 0.35), ("/mpros/planner/gear_penalty", 10.0),
    ("/mpros/planner/reverse_penalty", 0.0), ("/mpros/planner/steering_penalty", 0.1),
    ("/mpros/planner/steering_change_penalty", 10.0), ("/mpros/planner/analytic_solution_range", 10.0),
    ("/mpros/planner/yaw_discretization", 36), ("/mpros/planner/step_size", 1.1),
    ("/mpros/planner/path_point_distance", 0.2), ("/mpros/planner/steering_discretization", 3),
    ("/mpros/planner/max_iteration", 10000), ("/mpros/planner/optimal_path", 0),
]
MAP_ORDER = [
    ("/mpros/map/occupied_thresh", 0.68), ("/mpros/map/free_thresh", 0.2), ("/mpros/map/voro_fallout_rate", 10.0),
    ("/mpros/map/voro_max_region", 50.0), ("/mpros/map/downsampling_factor", 1), ("/mpros/map/obstacle_inflation_radius", 0.0),
]


class PortLib:
    def __init__(self, path=refapi.PORT_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle port`)")
        L = C.CDLL(path)
        vp, d, i = C.c_void_p, C.c_double, C.c_int
        L.port_map_create.restype = vp
        L.port_map_create.argtypes = [vp]
        L.port_map_destroy.argtypes = [vp]
        L.port_map_set_occupancy.argtypes = [vp, vp, i, i, d, d, d, d]
        L.port_map_update_goal.argtypes = [vp, d, d]
        L.port_map_info.argtypes = [vp, vp, vp]
        L.port_map_layer.restype = C.c_long
        L.port_map_layer.argtypes = [vp, i, vp]
        L.port_map_set_voronoi.argtypes = [vp, vp]
        L.port_planner_create.restype = vp
        L.port_planner_create.argtypes = [vp, vp]
        L.port_planner_destroy.argtypes = [vp]
        L.port_planner_collision4.argtypes = [vp, vp, C.c_long, vp]
        L.port_rs_solve.argtypes = [vp, vp, vp, vp, vp, vp, vp, i]
        L.port_rs_distance.restype = d
        L.port_rs_distance.argtypes = [d, vp, vp]
        L.port_expand.argtypes = [vp, vp, vp, vp, vp, vp]
        L.port_plan_query.argtypes = [vp, vp, vp, vp, vp, i]
        self.L = L


class PortMap:
    def __init__(self, lib, params):
        self.lib = lib
        self.params = dict(params)
        mp = np.array([params.get(k, dv) for k, dv in MAP_ORDER], dtype=np.float64)
        self.h = lib.L.port_map_create(mp.ctypes.data)

    def close(self):
        if self.h:
            self.lib.L.port_map_destroy(self.h)
            self.h = None

    def set_occupancy(self, occ, res, origin, exact_geometry=False):
        """res/origin as the message carries them: resolution is rounded to float32 like nav_msgs (pass
        exact_geometry=True when `res`/`origin` already are NavMap's stored doubles)."""
        occ = np.ascontiguousarray(occ, dtype=np.int8)
        H, W = occ.shape
        r = float(res) if exact_geometry else float(np.float32(res))
        self.lib.L.port_map_set_occupancy(self.h, occ.ctypes.data, W, H, r, origin[0], origin[1], origin[2])

    def update_goal(self, gx, gy):
        self.lib.L.port_map_update_goal(self.h, gx, gy)

    def info(self):
        oi = np.zeros(7, dtype=np.int32)
        od = np.zeros(7, dtype=np.float64)
        self.lib.L.port_map_info(self.h, oi.ctypes.data, od.ctypes.data)
        d = {k: int(v) for k, v in zip(["H", "W", "H_orig", "W_orig", "get_map", "get_goal", "num_regions"], oi)}
        d.update({k: float(v) for k, v in zip(["res", "res_orig", "ox", "oy", "oyaw", "sin", "cos"], od)})
        return d

    def layer(self, name):
        lid, dt, orig = LAYERS[name]
        n = self.lib.L.port_map_layer(self.h, lid, None)
        out = np.zeros(n, dtype=dt)
        if n:
            self.lib.L.port_map_layer(self.h, lid, out.ctypes.data)
        inf = self.info()
        shape = (inf["H_orig"], inf["W_orig"]) if orig else (inf["H"], inf["W"])
        return out.reshape(shape) if n == shape[0] * shape[1] else out

    def set_voronoi(self, v):
        v = np.ascontiguousarray(v, dtype=np.float32)
        self.lib.L.port_map_set_voronoi(self.h, v.ctypes.data)


class PortPlanner:
    def __init__(self, pmap, params=None):
        self.map = pmap
        self.lib = pmap.lib
        params = pmap.params if params is None else params
        pp = np.array([params.get(k, dv) for k, dv in PLANNER_ORDER], dtype=np.float64)
        self.h = self.lib.L.port_planner_create(pmap.h, pp.ctypes.data)

    def close(self):
        if self.h:
            self.lib.L.port_planner_destroy(self.h)
            self.h = None

    def collision4(self, poses):
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(poses), dtype=np.uint8)
        self.lib.L.port_planner_collision4(self.h, poses.ctypes.data, len(poses), out.ctypes.data)
        return out

    def expand(self, start3, goal3, parent6, n_prim):
        s = np.array(start3, dtype=np.float64)
        g = np.array(goal3, dtype=np.float64)
        p = np.array(parent6, dtype=np.float64)
        prim = np.zeros((n_prim, 7), dtype=np.float64)
        child = np.zeros((n_prim, 10), dtype=np.float64)
        nc = self.lib.L.port_expand(self.h, s.ctypes.data, g.ctypes.data, p.ctypes.data, prim.ctypes.data, child.ctypes.data)
        return prim, child[: max(nc, 0)], nc

    def plan(self, start3, goal3, path_cap=4096):
        s = np.array(start3, dtype=np.float64)
        g = np.array(goal3, dtype=np.float64)
        stats = np.zeros(8, dtype=np.float64)
        path = np.zeros((path_cap, 5), dtype=np.float64)
        self.lib.L.port_plan_query(self.h, s.ctypes.data, g.ctypes.data, stats.ctypes.data, path.ctypes.data, path_cap)
        keys = ["status", "iterations", "goal_g", "n_primitives", "num_nodes", "path_len", "n_rs_shots", "n_collision"]
        d = dict(zip(keys, stats.tolist()))
        d["path"] = path[: int(min(d["path_len"], path_cap))].copy()
        return d


def rs_solve(lib, cfg7, start5, goal3, traj_cap=256):
    cfg = np.array(cfg7, dtype=np.float64)
    s = np.array(start5, dtype=np.float64)
    g = np.array(goal3, dtype=np.float64)
    seg = np.zeros((5, 3), dtype=np.float64)
    nseg = C.c_int(0)
    cost = C.c_double(0)
    traj = np.zeros((traj_cap, 5), dtype=np.float64)
    n = lib.L.port_rs_solve(cfg.ctypes.data, s.ctypes.data, g.ctypes.data, seg.ctypes.data, C.byref(nseg), C.byref(cost), traj.ctypes.data, traj_cap)
    return seg[: nseg.value].copy(), cost.value, traj[: min(n, traj_cap)].copy()


def rs_distance(lib, rho, a3, b3):
    a = np.array(a3, dtype=np.float64)
    b = np.array(b3, dtype=np.float64)
    return lib.L.port_rs_distance(rho, a.ctypes.data, b.ctypes.data)


def collision4_on_map(mp, params, poses):
    """Convenience for smoke(): restated footPrintInCollision4 verdicts on a fixture map."""
    lib = PortLib()
    pm = PortMap(lib, params)
    pm.set_occupancy(mp["occupancy"], mp["resolution"], mp["origin"])
    pl = PortPlanner(pm)
    return pl.collision4(poses)
