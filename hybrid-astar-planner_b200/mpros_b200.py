"""ctypes binding of libmpros_b200.so (include/mpros_gpu.h) — the host-side handle used by tests/ and bench.py.

There is no CPU fallback: if the CUDA library is missing or no CUDA device is present every call raises.
The C++ façade with the reference's class names (NavMap / HybridAstar / RS::RSCurve) lives in host/.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPROS_LIB", os.path.join(HERE, "libmpros_b200.so"))  # MPROS_LIB: developer builds (phase profile)

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


class MapParams(C.Structure):
    _fields_ = [
        ("occupied_thresh", C.c_double),
        ("free_thresh", C.c_double),
        ("voro_fallout_rate", C.c_float),
        ("voro_max_region", C.c_float),
        ("downsampling_factor", C.c_int),
        ("obstacle_inflation_radius", C.c_double),
    ]


class PlannerParams(C.Structure):
    _fields_ = [
        ("max_curvature", C.c_double),
        ("width", C.c_double),
        ("wheelbase", C.c_double),
        ("length", C.c_double),
        ("steer_offset", C.c_double),
        ("max_steering_angle", C.c_double),
        ("gear_penalty", C.c_double),
        ("reverse_penalty", C.c_double),
        ("steering_penalty", C.c_double),
        ("steering_change_penalty", C.c_double),
        ("analytic_solution_range", C.c_double),
        ("yaw_discretization", C.c_int),
        ("step_size", C.c_double),
        ("path_point_distance", C.c_double),
        ("steering_discretization", C.c_int),
        ("max_iteration", C.c_int),
        ("optimal_path", C.c_int),
        ("interpolation_enable", C.c_int),
    ]


class MapInfo(C.Structure):
    _fields_ = [
        ("get_map", C.c_int),
        ("get_goal", C.c_int),
        ("map_height", C.c_int),
        ("map_width", C.c_int),
        ("map_height_orig", C.c_int),
        ("map_width_orig", C.c_int),
        ("map_resolution", C.c_double),
        ("map_resolution_orig", C.c_double),
        ("map_origin_x", C.c_double),
        ("map_origin_y", C.c_double),
        ("map_origin_yaw", C.c_double),
        ("map_origin_sin", C.c_double),
        ("map_origin_cos", C.c_double),
        ("x_max", C.c_double),
        ("x_min", C.c_double),
        ("y_max", C.c_double),
        ("y_min", C.c_double),
        ("num_regions", C.c_int),
    ]


class RsParams(C.Structure):
    _fields_ = [
        ("radius", C.c_double),
        ("steering_angle", C.c_double),
        ("step_size", C.c_double),
        ("pen_gear", C.c_double),
        ("pen_backward", C.c_double),
        ("pen_steer_change", C.c_double),
        ("pen_steer_angle", C.c_double),
    ]


PLAN_FOUND, PLAN_LIMIT, PLAN_OPEN_EMPTY, PLAN_NOT_INIT, PLAN_OVERFLOW = 1, 0, -1, -2, -3


class MprosError(RuntimeError):
    pass


_lib = None


def load_library(path=LIB_PATH):
    """Load libmpros_b200.so; raises (loudly) when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise MprosError("CUDA extension missing: %s (run __graft_entry__.build()); there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, i, d, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    L.mpros_ctx_create.argtypes = [i, C.POINTER(vp)]
    L.mpros_ctx_destroy.argtypes = [vp]
    L.mpros_last_error_string.restype = C.c_char_p
    L.mpros_ctx_stream.restype = vp
    L.mpros_ctx_stream.argtypes = [vp]
    L.mpros_ctx_synchronize.argtypes = [vp]
    L.mpros_ctx_launch_count.restype = C.c_uint64
    L.mpros_ctx_launch_count.argtypes = [vp]
    L.mpros_map_params_default.argtypes = [C.POINTER(MapParams)]
    L.mpros_map_params_default.restype = None
    L.mpros_planner_params_default.argtypes = [C.POINTER(PlannerParams)]
    L.mpros_planner_params_default.restype = None
    L.mpros_map_config.argtypes = [vp, C.POINTER(MapParams)]
    L.mpros_planner_config.argtypes = [vp, C.POINTER(PlannerParams)]
    L.mpros_map_set_occupancy.argtypes = [vp, vp, i, i, d, d, d, d]
    L.mpros_map_set_occupancy_dev.argtypes = [vp, vp, i, i, d, d, d, d]
    L.mpros_map_update_goal.argtypes = [vp, d, d]
    L.mpros_map_band_begin.argtypes = [vp, i, i, d, d, d, d, C.POINTER(C.c_int)]
    L.mpros_map_band_threshold.argtypes = [vp, vp, i, i]
    L.mpros_map_band_threshold_dev.argtypes = [vp, vp, i, i]
    L.mpros_map_band_layer_ptr.argtypes = [vp, i, C.POINTER(vp), C.POINTER(sz)]
    L.mpros_map_band_inflate.argtypes = [vp, i, i]
    L.mpros_map_band_finish.argtypes = [vp]
    L.mpros_map_get_info.argtypes = [vp, C.POINTER(MapInfo)]
    L.mpros_map_download.argtypes = [vp, i, vp, sz]
    L.mpros_map_upload.argtypes = [vp, i, vp, sz]
    L.mpros_map_layer_msg.argtypes = [vp, i, vp, sz, C.POINTER(sz)]
    L.mpros_map_point_in_collision.argtypes = [vp, i, vp, sz, vp]
    L.mpros_collision_check_poses.argtypes = [vp, vp, sz, vp]
    L.mpros_collision_check_poses_dev.argtypes = [vp, vp, sz, vp]
    L.mpros_collision_check_paths.argtypes = [vp, vp, vp, sz, vp]
    L.mpros_collision_check_poses_method.argtypes = [vp, i, vp, sz, vp]
    L.mpros_rs_default_config.argtypes = [vp, C.POINTER(RsParams)]
    L.mpros_rs_solve_batch.argtypes = [vp, C.POINTER(RsParams), vp, vp, sz, vp, vp, vp, vp, i, vp]
    L.mpros_rs_shot_batch.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp, i, vp, vp]
    L.mpros_rs_shot_batch_dev.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp, vp]
    L.mpros_rs_distance_batch.argtypes = [vp, d, vp, vp, sz, vp]
    L.mpros_expand_batch.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp]
    L.mpros_expand_batch_dev.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp]
    L.mpros_expand_batch_compact.argtypes = [vp, vp, vp, vp, sz, vp, sz, C.POINTER(sz), vp]
    L.mpros_expand_batch_compact_dev.argtypes = [vp, vp, vp, vp, sz, vp, sz, vp, vp]
    L.mpros_plan_batch.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp, i, vp]
    L.mpros_path_postprocess_batch.argtypes = [vp, vp, vp, sz, vp, sz, vp]
    L.mpros_plan_search_tree.argtypes = [vp, vp, vp, vp, vp, vp, sz, C.POINTER(sz)]
    L.mpros_plan_batch_dev.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp, i, vp]
    L.mpros_plan_batch_submit.argtypes = [vp, i, vp, vp, sz, vp, vp, vp, vp, i, vp]
    L.mpros_plan_batch_submit_dev.argtypes = [vp, i, vp, vp, sz, vp, vp, vp, vp, i, vp]
    L.mpros_plan_batch_wait.argtypes = [vp, i]
    L.mpros_plan_lane_stream.restype = vp
    L.mpros_plan_lane_stream.argtypes = [vp, i]
    L.mpros_ctx_set_option.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.mpros_plan_release_workspaces.argtypes = [vp]
    L.mpros_comm_unique_id.argtypes = [vp]
    L.mpros_comm_create_nccl.argtypes = [vp, i, i, vp, C.POINTER(vp)]
    L.mpros_comm_adopt_nccl.argtypes = [i, i, vp, C.POINTER(vp)]
    L.mpros_comm_create_local.argtypes = [i, C.POINTER(vp)]
    L.mpros_comm_destroy.argtypes = [vp]
    L.mpros_comm_info.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.mpros_map_band_rows.argtypes = [vp, vp, i, C.POINTER(i), C.POINTER(i)]
    L.mpros_map_set_occupancy_banded.argtypes = [vp, vp, vp, i, i, d, d, d, d]
    L.mpros_map_set_occupancy_banded_dev.argtypes = [vp, vp, vp, i, i, d, d, d, d]
    L.mpros_plan_shard_range.argtypes = [vp, sz, C.POINTER(sz), C.POINTER(sz)]
    L.mpros_plan_batch_sharded.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp, vp, i, vp, i]
    L.mpros_plan_share_create.argtypes = [vp, vp, sz, i, C.POINTER(vp)]
    L.mpros_plan_share_destroy.argtypes = [vp, vp]
    L.mpros_plan_batch_shared.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp, vp, i, vp]
    _lib = L
    return L


PLAN_LANES = 8  # MPROS_PLAN_LANES

# mpros_child_record (include/mpros_gpu.h): 80 bytes
CHILD_RECORD = np.dtype([("parent", "<u4"), ("prim", "<u2"), ("dir", "<i2"), ("i", "<i4"), ("j", "<i4"), ("x", "<f8"), ("y", "<f8"),
                         ("yaw", "<f8"), ("steer", "<f8"), ("g", "<f8"), ("h", "<f8"), ("f", "<f8"), ("order", "<u4"), ("reserved", "<u4")])
assert CHILD_RECORD.itemsize == 80


def exported_symbols():
    """Names declared in include/mpros_gpu.h (parsed from the header)."""
    import re

    hdr = os.path.join(os.path.dirname(HERE), "include", "mpros_gpu.h")
    return sorted(set(re.findall(r"MPROS_API[^;(]*?\b(mpros_\w+)\s*\(", open(hdr).read())))


# parameter-server keys of the reference -> struct fields
_MAP_KEYS = {
    "/mpros/map/occupied_thresh": "occupied_thresh",
    "/mpros/map/free_thresh": "free_thresh",
    "/mpros/map/voro_fallout_rate": "voro_fallout_rate",
    "/mpros/map/voro_max_region": "voro_max_region",
    "/mpros/map/downsampling_factor": "downsampling_factor",
    "/mpros/map/obstacle_inflation_radius": "obstacle_inflation_radius",
}
_PLANNER_KEYS = {
    "/mpros/vehicle/max_curvature": "max_curvature",
    "/mpros/vehicle/width": "width",
    "/mpros/vehicle/wheelbase": "wheelbase",
    "/mpros/vehicle/length": "length",
    "/mpros/vehicle/center_to_rotate_center_longitudinal": "steer_offset",
    "/mpros/vehicle/max_steering_angle": "max_steering_angle",
    "/mpros/planner/gear_penalty": "gear_penalty",
    "/mpros/planner/reverse_penalty": "reverse_penalty",
    "/mpros/planner/steering_penalty": "steering_penalty",
    "/mpros/planner/steering_change_penalty": "steering_change_penalty",
    "/mpros/planner/analytic_solution_range": "analytic_solution_range",
    "/mpros/planner/yaw_discretization": "yaw_discretization",
    "/mpros/planner/step_size": "step_size",
    "/mpros/planner/path_point_distance": "path_point_distance",
    "/mpros/planner/steering_discretization": "steering_discretization",
    "/mpros/planner/max_iteration": "max_iteration",
    "/mpros/planner/optimal_path": "optimal_path",
    "/mpros/planner/interpolation_enable": "interpolation_enable",
}


def _fill(struct, keys, params):
    for k, v in params.items():
        f = keys.get(k)
        if f is None:
            continue
        ftype = dict(struct._fields_)[f]
        setattr(struct, f, int(v) if ftype is C.c_int else float(v))


class Comm:
    """mpros_comm: one rank of a multi-GPU job. Comm.nccl(ctx, world, rank, id) on an 8 x B200 box (id from
    Comm.unique_id() on rank 0, handed round by the launcher); Comm.local(world) -> the ranks of one process (one host
    thread and one Context each)."""

    def __init__(self, handle):
        self.L = load_library()
        self.h = handle

    @staticmethod
    def unique_id():
        L = load_library()
        buf = (C.c_char * 128)()
        rc = L.mpros_comm_unique_id(buf)
        if rc != 0:
            raise MprosError("mpros error %d: %s" % (rc, L.mpros_last_error_string().decode()))
        return bytes(buf)

    @staticmethod
    def nccl(ctx, world, rank, unique_id):
        L = load_library()
        h = C.c_void_p()
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        rc = L.mpros_comm_create_nccl(ctx.h, world, rank, buf, C.byref(h))
        if rc != 0:
            raise MprosError("mpros error %d: %s" % (rc, L.mpros_last_error_string().decode()))
        return Comm(h)

    @staticmethod
    def from_torch_distributed(ctx, dist, device):
        """One mpros communicator over the ranks of an initialised torch.distributed job (the id travels by broadcast)."""
        import torch

        world, rank = dist.get_world_size(), dist.get_rank()
        t = torch.zeros(128, dtype=torch.uint8, device=device)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return Comm.nccl(ctx, world, rank, bytes(t.cpu().numpy().tobytes()))

    @staticmethod
    def local(world):
        L = load_library()
        arr = (C.c_void_p * world)()
        rc = L.mpros_comm_create_local(world, arr)
        if rc != 0:
            raise MprosError("mpros error %d: %s" % (rc, L.mpros_last_error_string().decode()))
        return [Comm(C.c_void_p(arr[r])) for r in range(world)]

    def info(self):
        w, r, b, e = C.c_int(), C.c_int(), C.c_uint64(), C.c_uint64()
        self.L.mpros_comm_info(self.h, C.byref(w), C.byref(r), C.byref(b), C.byref(e))
        return {"world": w.value, "rank": r.value, "bytes_sent": b.value, "exchanges": e.value}

    def close(self):
        if getattr(self, "h", None):
            self.L.mpros_comm_destroy(self.h)
            self.h = None


def run_local_ranks(world, fn):
    """Run fn(rank) on `world` host threads (the local communicator's ranks); re-raises the first failure."""
    import threading

    out, err = [None] * world, [None] * world

    def work(r):
        try:
            out[r] = fn(r)
        except BaseException as e:  # noqa: BLE001 - reported to the caller
            err[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out


class Context:
    """One GPU context = one NavMap + one planner configuration + one CUDA stream."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        self._check(self.L.mpros_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def _check(self, rc):
        if rc != 0:
            raise MprosError("mpros error %d: %s" % (rc, self.L.mpros_last_error_string().decode()))

    def set_option(self, key, value=None):
        """mpros_ctx_set_option: "c1_mode", "goal_mode", "plan_mode", "solo_nodes", "solo_min_queries", "plan_grow", "plan_policy", "plan_age_limit", "debug" (value None = default)."""
        self._check(self.L.mpros_ctx_set_option(self.h, key.encode(), None if value is None else str(value).encode()))

    def release_workspaces(self):
        self._check(self.L.mpros_plan_release_workspaces(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.mpros_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration (ROS parameter-server style dict, same keys as the reference) ----
    def map_config(self, params):
        p = MapParams()
        self.L.mpros_map_params_default(C.byref(p))
        _fill(p, _MAP_KEYS, params)
        self._check(self.L.mpros_map_config(self.h, C.byref(p)))

    def planner_config(self, params):
        p = PlannerParams()
        self.L.mpros_planner_params_default(C.byref(p))
        _fill(p, _PLANNER_KEYS, params)
        self._check(self.L.mpros_planner_config(self.h, C.byref(p)))
        self.planner_params = p

    # ---- NavMap ----
    def set_occupancy(self, occ, resolution, origin):
        """occ: int8 HxW (-1/0/100), row 0 = bottom. resolution: the float32 message value widened to double
        (what NavMap stores); origin: [x, y, yaw]."""
        occ = np.ascontiguousarray(occ, dtype=np.int8)
        H, W = occ.shape
        res = float(np.float32(resolution))
        self._check(self.L.mpros_map_set_occupancy(self.h, occ.ctypes.data, W, H, res, origin[0], origin[1], origin[2]))

    def set_occupancy_dev(self, dptr, W, H, resolution, origin):
        res = float(np.float32(resolution))
        self._check(self.L.mpros_map_set_occupancy_dev(self.h, dptr, W, H, res, origin[0], origin[1], origin[2]))

    # ---- NavMap::setOccupancyMap tiled by row bands (one context per GPU; sharding.set_occupancy_banded drives it) ----
    def band_begin(self, W, H, resolution, origin):
        """Geometry + allocation; returns the halo (rows of the raw layer a band needs from each neighbour for N3)."""
        halo = C.c_int(0)
        res = float(np.float32(resolution))
        self._check(self.L.mpros_map_band_begin(self.h, W, H, res, origin[0], origin[1], origin[2], C.byref(halo)))
        return halo.value

    def band_threshold(self, rows, row0):
        rows = np.ascontiguousarray(rows, dtype=np.int8)
        self._check(self.L.mpros_map_band_threshold(self.h, rows.ctypes.data, row0, rows.shape[0]))

    def band_rows(self, layer, row0, nrows):
        """Rows [row0, row0 + nrows) of the bit-packed static_orig / inflate_orig layer as a torch int32 CUDA tensor
        [nrows, pitch_words] that ALIASES the layer's device memory (for the halo exchange / all-gather)."""
        import torch

        ptr, pitch = C.c_void_p(0), C.c_size_t(0)
        self._check(self.L.mpros_map_band_layer_ptr(self.h, LAYERS[layer][0], C.byref(ptr), C.byref(pitch)))
        words = pitch.value // 4

        class _Alias:  # CUDA array interface: lets torch wrap memory owned by the context
            __cuda_array_interface__ = {"shape": (nrows, words), "typestr": "<i4", "data": (ptr.value + row0 * pitch.value, False),
                                        "version": 3, "strides": None}

        self.synchronize()  # the tensor is used on torch's stream
        return torch.as_tensor(_Alias(), device="cuda:%d" % self.device)

    def band_inflate(self, row0, nrows):
        self._check(self.L.mpros_map_band_inflate(self.h, row0, nrows))

    def band_finish(self):
        self._check(self.L.mpros_map_band_finish(self.h))

    # ---- the same with every stage banded and the exchanges inside the library (mpros_map_set_occupancy_banded) ----
    def band_of(self, comm, H):
        """Rows [row0, row0 + nrows) of the occupancy grid this rank provides."""
        r0, n = C.c_int(), C.c_int()
        self._check(self.L.mpros_map_band_rows(self.h, comm.h if comm else None, H, C.byref(r0), C.byref(n)))
        return r0.value, n.value

    def set_occupancy_banded(self, comm, rows, W, H, resolution, origin):
        rows = np.ascontiguousarray(rows, dtype=np.int8)
        keep = rows if rows.size else np.zeros(1, dtype=np.int8)
        res = float(np.float32(resolution))
        self._check(self.L.mpros_map_set_occupancy_banded(self.h, comm.h if comm else None, keep.ctypes.data, W, H, res, origin[0],
                                                          origin[1], origin[2]))

    def update_goal(self, gx, gy):
        self._check(self.L.mpros_map_update_goal(self.h, gx, gy))

    def info(self):
        o = MapInfo()
        self._check(self.L.mpros_map_get_info(self.h, C.byref(o)))
        return {f: getattr(o, f) for f, _ in MapInfo._fields_}

    def layer(self, name):
        lid, dt, orig = LAYERS[name]
        inf = self.info()
        shape = (inf["map_height_orig"], inf["map_width_orig"]) if orig else (inf["map_height"], inf["map_width"])
        out = np.zeros(shape, dtype=dt)
        self._check(self.L.mpros_map_download(self.h, lid, out.ctypes.data, out.nbytes))
        return out

    def layer_msg(self, name):
        """NavMap::get*MapMsg (nav_map.cpp:654-716): the int8 data of the published OccupancyGrid for one layer (flat)."""
        n = C.c_size_t(0)
        self._check(self.L.mpros_map_layer_msg(self.h, LAYERS[name][0], None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int8)
        if n.value:
            self._check(self.L.mpros_map_layer_msg(self.h, LAYERS[name][0], out.ctypes.data, out.nbytes, C.byref(n)))
        return out

    def upload_voronoi(self, v):
        v = np.ascontiguousarray(v, dtype=np.float32)
        self._check(self.L.mpros_map_upload(self.h, LAYERS["voronoi"][0], v.ctypes.data, v.nbytes))

    def point_in_collision(self, layer, ij):
        ij = np.ascontiguousarray(ij, dtype=np.int32).reshape(-1, 2)
        out = np.zeros(len(ij), dtype=np.uint8)
        self._check(self.L.mpros_map_point_in_collision(self.h, LAYERS[layer][0], ij.ctypes.data, len(ij), out.ctypes.data))
        return out

    # ---- collision ----
    def collision_check(self, poses):
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(poses), dtype=np.uint8)
        self._check(self.L.mpros_collision_check_poses(self.h, poses.ctypes.data, len(poses), out.ctypes.data))
        return out

    def collision_check_dev(self, d_poses, n, d_out):
        self._check(self.L.mpros_collision_check_poses_dev(self.h, d_poses, n, d_out))

    def collision_check_paths(self, poses, offsets):
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        out = np.zeros(len(offsets) - 1, dtype=np.uint8)
        self._check(self.L.mpros_collision_check_paths(self.h, poses.ctypes.data, offsets.ctypes.data, len(out), out.ctypes.data))
        return out

    # ---- Reeds-Shepp ----
    def rs_default_config(self):
        p = RsParams()
        self._check(self.L.mpros_rs_default_config(self.h, C.byref(p)))
        return p

    def rs_solve_batch(self, starts5, goals3, traj_cap=64, rs=None, shot=False):
        """RSCurve FindOptimalPath + GetTraj for n shots. Returns dict(segs[n,5,3], nseg, cost, traj[n,cap,5], npts
        [, verdicts])."""
        s = np.ascontiguousarray(starts5, dtype=np.float64).reshape(-1, 5)
        g = np.ascontiguousarray(goals3, dtype=np.float64).reshape(-1, 3)
        n = len(s)
        segs = np.zeros((n, 5, 3))
        nseg = np.zeros(n, dtype=np.int32)
        cost = np.zeros(n)
        traj = np.zeros((n, traj_cap, 5))
        npts = np.zeros(n, dtype=np.int32)
        out = {"segs": segs, "nseg": nseg, "cost": cost, "traj": traj, "npts": npts}
        tp = traj.ctypes.data if traj_cap else None
        if shot:
            ver = np.zeros(n, dtype=np.uint8)
            self._check(self.L.mpros_rs_shot_batch(self.h, s.ctypes.data, g.ctypes.data, n, segs.ctypes.data, nseg.ctypes.data,
                                                   cost.ctypes.data, tp, traj_cap, npts.ctypes.data, ver.ctypes.data))
            out["verdicts"] = ver
        else:
            rp = C.byref(rs) if rs is not None else None
            self._check(self.L.mpros_rs_solve_batch(self.h, rp, s.ctypes.data, g.ctypes.data, n, segs.ctypes.data, nseg.ctypes.data,
                                                    cost.ctypes.data, tp, traj_cap, npts.ctypes.data))
        return out

    def rs_distance_batch(self, rho, a3, b3):
        a = np.ascontiguousarray(a3, dtype=np.float64).reshape(-1, 3)
        b = np.ascontiguousarray(b3, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(a))
        self._check(self.L.mpros_rs_distance_batch(self.h, rho, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data))
        return out

    # ---- expansion ----
    def n_primitives(self):
        return 2 * (2 * (self.planner_params.steering_discretization // 2) + 1)

    def expand_batch(self, start3, goal3, parents6):
        p = np.ascontiguousarray(parents6, dtype=np.float64).reshape(-1, 6)
        n, P = len(p), self.n_primitives()
        g = np.array(goal3, dtype=np.float64)
        s = np.array(start3, dtype=np.float64) if start3 is not None else None
        prims = np.zeros((n, P, 7))
        childs = np.zeros((n, P, 10))
        nchild = np.zeros(n, dtype=np.int32)
        self._check(self.L.mpros_expand_batch(self.h, s.ctypes.data if s is not None else None, g.ctypes.data, p.ctypes.data, n,
                                              prims.ctypes.data, childs.ctypes.data, nchild.ctypes.data))
        return prims, childs, nchild

    def expand_batch_compact(self, start3, goal3, parents6, cap=None):
        """mpros_expand_batch_compact: the inserted children only, as CHILD_RECORD rows sorted by (parent, order)."""
        par = np.ascontiguousarray(parents6, dtype=np.float64).reshape(-1, 6)
        n = len(par)
        cap = n * self.n_primitives() if cap is None else cap
        rec = np.zeros(max(cap, 1), dtype=CHILD_RECORD)
        nchild = np.zeros(n, dtype=np.int32)
        cnt = C.c_size_t(0)
        s = None if start3 is None else np.ascontiguousarray(start3, dtype=np.float64)
        g = np.ascontiguousarray(goal3, dtype=np.float64)
        self._check(self.L.mpros_expand_batch_compact(self.h, None if s is None else s.ctypes.data, g.ctypes.data, par.ctypes.data, n,
                                                      rec.ctypes.data, cap, C.byref(cnt), nchild.ctypes.data))
        rec = rec[: cnt.value]
        return rec[np.lexsort((rec["order"], rec["parent"]))], nchild

    # ---- planner ----
    def plan_batch(self, starts3, goals3, path_cap=256, truncate_ok=False):
        """mpros_plan_batch. A path longer than path_cap is an error (the C entry stores only the first path_cap points
        and reports the full length in path_len) unless truncate_ok is set or no paths are asked for (path_cap == 0)."""
        s = np.ascontiguousarray(starts3, dtype=np.float64).reshape(-1, 3)
        g = np.ascontiguousarray(goals3, dtype=np.float64).reshape(-1, 3)
        nq = len(s)
        status = np.zeros(nq, dtype=np.int32)
        cost = np.zeros(nq)
        plen = np.zeros(nq, dtype=np.int32)
        paths = np.zeros((nq, path_cap, 5))
        stats = np.zeros((nq, 6), dtype=np.int64)
        self._check(self.L.mpros_plan_batch(self.h, s.ctypes.data, g.ctypes.data, nq, status.ctypes.data, cost.ctypes.data,
                                            plen.ctypes.data, paths.ctypes.data if path_cap else None, path_cap, stats.ctypes.data))
        if path_cap and not truncate_ok and nq and int(plen.max()) > path_cap:
            raise MprosError("plan_batch: a path has %d points, path_cap is %d (MPROS_ERR_CAPACITY); re-run with a larger path_cap"
                             % (int(plen.max()), path_cap))
        return {"status": status, "cost": cost, "path_len": plen, "paths": paths, "iterations": stats[:, 0],
                "primitives": stats[:, 1], "rs_shots": stats[:, 2], "collision_checks": stats[:, 3],
                "goal_map_cells": stats[:, 4], "nodes_inserted": stats[:, 5]}

    def plan_batch_sharded(self, comm, starts3, goals3, path_cap=256, gather_paths=False):
        """mpros_plan_batch_sharded: the global query set, this rank plans its block, summaries gathered on every rank."""
        s = np.ascontiguousarray(starts3, dtype=np.float64).reshape(-1, 3)
        g = np.ascontiguousarray(goals3, dtype=np.float64).reshape(-1, 3)
        nq = len(s)
        status = np.zeros(nq, dtype=np.int32)
        cost = np.zeros(nq)
        plen = np.zeros(nq, dtype=np.int32)
        paths = np.zeros((nq, path_cap, 5))
        stats = np.zeros((nq, 6), dtype=np.int64)
        self._check(self.L.mpros_plan_batch_sharded(self.h, comm.h if comm else None, s.ctypes.data, g.ctypes.data, nq, status.ctypes.data,
                                                    cost.ctypes.data, plen.ctypes.data, paths.ctypes.data if path_cap else None, path_cap,
                                                    stats.ctypes.data, 1 if gather_paths else 0))
        lo, hi = C.c_size_t(), C.c_size_t()
        self.L.mpros_plan_shard_range(comm.h if comm else None, nq, C.byref(lo), C.byref(hi))
        return {"status": status, "cost": cost, "path_len": plen, "paths": paths, "iterations": stats[:, 0], "primitives": stats[:, 1],
                "owned": (lo.value, hi.value)}

    def plan_share_create(self, comm, max_queries, path_cap):
        """mpros_plan_share_create (collective): rank 0's ticket counter and result arrays, mapped by every rank."""
        h = C.c_void_p()
        self._check(self.L.mpros_plan_share_create(self.h, comm.h, max_queries, path_cap, C.byref(h)))
        return h

    def plan_share_destroy(self, share):
        self._check(self.L.mpros_plan_share_destroy(self.h, share))

    def plan_batch_shared(self, share, starts3, goals3, path_cap, want_results=True):
        """mpros_plan_batch_shared (collective): all ranks draw tickets from rank 0's counter and store into rank 0's arrays."""
        s = np.ascontiguousarray(starts3, dtype=np.float64).reshape(-1, 3)
        g = np.ascontiguousarray(goals3, dtype=np.float64).reshape(-1, 3)
        nq = len(s)
        if not want_results:
            self._check(self.L.mpros_plan_batch_shared(self.h, share, s.ctypes.data, g.ctypes.data, nq, None, None, None, None, path_cap, None))
            return None
        status = np.zeros(nq, dtype=np.int32)
        cost = np.zeros(nq)
        plen = np.zeros(nq, dtype=np.int32)
        paths = np.zeros((nq, path_cap, 5))
        stats = np.zeros((nq, 6), dtype=np.int64)
        self._check(self.L.mpros_plan_batch_shared(self.h, share, s.ctypes.data, g.ctypes.data, nq, status.ctypes.data, cost.ctypes.data,
                                                   plen.ctypes.data, paths.ctypes.data if path_cap else None, path_cap, stats.ctypes.data))
        return {"status": status, "cost": cost, "path_len": plen, "paths": paths, "iterations": stats[:, 0], "primitives": stats[:, 1]}

    def plan_batch_dev(self, d_starts, d_goals, nq, d_status, d_cost, d_len, d_paths, path_cap, d_stats):
        self._check(self.L.mpros_plan_batch_dev(self.h, d_starts, d_goals, nq, d_status, d_cost, d_len, d_paths, path_cap, d_stats))

    # batches in flight: lane = 0 .. PLAN_LANES-1, raw pointers (device for *_dev, page-locked host otherwise)
    def plan_batch_submit_dev(self, lane, d_starts, d_goals, nq, d_status, d_cost, d_len, d_paths, path_cap, d_stats):
        self._check(self.L.mpros_plan_batch_submit_dev(self.h, lane, d_starts, d_goals, nq, d_status, d_cost, d_len, d_paths, path_cap,
                                                       d_stats))

    def plan_batch_submit(self, lane, h_starts, h_goals, nq, h_status, h_cost, h_len, h_paths, path_cap, h_stats):
        self._check(self.L.mpros_plan_batch_submit(self.h, lane, h_starts, h_goals, nq, h_status, h_cost, h_len, h_paths, path_cap,
                                                   h_stats))

    def plan_batch_wait(self, lane=-1):
        self._check(self.L.mpros_plan_batch_wait(self.h, lane))

    def plan_lane_stream(self, lane):
        return self.L.mpros_plan_lane_stream(self.h, lane)

    def plan_search_tree(self, start, goal):
        """HybridAstar::getSearchTree for one query: (status, cost, [n, 4] {next_x, next_y, curr_x, curr_y})."""
        s = np.ascontiguousarray(start, dtype=np.float64)
        g = np.ascontiguousarray(goal, dtype=np.float64)
        status, cost, n = C.c_int32(0), C.c_double(0), C.c_size_t(0)
        self._check(self.L.mpros_plan_search_tree(self.h, s.ctypes.data, g.ctypes.data, C.byref(status), C.byref(cost), None, 0, C.byref(n)))
        tree = np.zeros((n.value, 4), dtype=np.float64)
        if n.value:
            self._check(self.L.mpros_plan_search_tree(self.h, s.ctypes.data, g.ctypes.data, C.byref(status), C.byref(cost), tree.ctypes.data,
                                                      n.value, C.byref(n)))
        return status.value, cost.value, tree

    def collision_check_method(self, method, poses):
        """footPrintInCollision (1), footPrintInCollision2 (2), footPrintInCollision3 (3) or footPrintInCollision4 (4)."""
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 3)
        out = np.zeros(len(poses), dtype=np.uint8)
        self._check(self.L.mpros_collision_check_poses_method(self.h, method, poses.ctypes.data, len(poses), out.ctypes.data))
        return out

    def path_postprocess(self, paths):
        """getNewPath() for a list of raw paths (each [n_k, 5] {x, y, yaw, steer, direction}): list of [m_k, 8] arrays
        {x, y, yaw, curv, direc, steering, is_cusp, dist} (setPath's post-processing incl. Bezier up-sampling)."""
        paths = [np.ascontiguousarray(p, dtype=np.float64).reshape(-1, 5) for p in paths]
        offs = np.zeros(len(paths) + 1, dtype=np.int64)
        offs[1:] = np.cumsum([len(p) for p in paths])
        flat = np.concatenate(paths) if paths else np.zeros((0, 5))
        out_offs = np.zeros(len(paths) + 1, dtype=np.int64)
        cap = 64 * max(1, int(offs[-1]))
        while True:
            out = np.zeros((cap, 8), dtype=np.float64)
            rc = self.L.mpros_path_postprocess_batch(self.h, flat.ctypes.data, offs.ctypes.data, len(paths), out.ctypes.data, cap, out_offs.ctypes.data)
            if rc == 3 and out_offs[-1] > cap:  # MPROS_ERR_CAPACITY: sizes are known now
                cap = int(out_offs[-1])
                continue
            self._check(rc)
            break
        return [out[out_offs[k]:out_offs[k + 1]].copy() for k in range(len(paths))]

    # ---- plumbing ----
    def stream(self):
        return self.L.mpros_ctx_stream(self.h)

    def synchronize(self):
        self._check(self.L.mpros_ctx_synchronize(self.h))

    def launch_count(self):
        return int(self.L.mpros_ctx_launch_count(self.h))
