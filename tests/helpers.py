"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from oracle import refapi


def make_pair(reflib, ctx, map_name, ds=1, num_steer=None, goal=None, occupancy=None, resolution=None, origin=None):
    """Configure the reference oracle and the GPU context identically on one map. Returns (RefMap, info)."""
    if occupancy is None:
        m = refapi.load_map(map_name)
    else:
        m = {"occupancy": occupancy, "resolution": resolution, "origin": origin, "free_thresh": 0.2, "occupied_thresh": 0.68}
    params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"], ds=ds, num_steer=num_steer)
    rm = refapi.RefMap(reflib, params)
    rm.set_occupancy(m["occupancy"], m["resolution"], m["origin"])
    inf = rm.info()
    ctx.map_config(params)
    # the GPU side receives exactly the geometry NavMap stored (float32 resolution, tf2 yaw)
    ctx.L.mpros_map_set_occupancy(ctx.h, m["occupancy"].ctypes.data, m["occupancy"].shape[1], m["occupancy"].shape[0],
                                  inf["res_orig"], inf["ox"], inf["oy"], inf["oyaw"])
    ctx.planner_config(params)
    if goal is not None:
        rm.update_goal(goal[0], goal[1])
        ctx.update_goal(goal[0], goal[1])
    return rm, params, m


def random_poses(rng, n, info, margin=3.0):
    """Uniform poses over the map's bounding box (plus a margin so some fall outside), yaw in (-pi, pi]."""
    H, W, res = info["H_orig"], info["W_orig"], info["res_orig"]
    # corners of the (possibly rotated) map in world frame
    c, s = info["cos"], info["sin"]
    u = rng.uniform(-margin, W * res + margin, n)
    v = rng.uniform(-margin, H * res + margin, n)
    x = info["ox"] + u * c - v * s
    y = info["oy"] + u * s + v * c
    yaw = rng.uniform(-np.pi, np.pi, n)
    return np.stack([x, y, yaw], axis=1)
