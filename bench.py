#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native mpros hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload collision]

Workload "collision" (SURVEY.md §8(d), C1 standalone): synthetic 8192x8192 grid @0.1 m (seed 20261019, 1-m border
wall + 1100 rectangles), inflation 1.0 m, down-sampling 8; one STEP = one pass of batched footPrintInCollision4
over 2^24 seeded uniform poses per GPU (403 MB of poses, larger than the 126 MB L2). Multi-GPU: poses are
sharded across ranks, map replicated, no data-path collective (scaling "weak").

One JSON line on stdout (rank 0). The oracle (oracle/) is used only as the CPU baseline / reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hybrid-astar-planner_b200"))

SYN_SIZE = 8192
SYN_RES = 0.1
SYN_DS = 8
POSES_PER_STEP = 1 << 24
POSE_SEED = 1
ALGO_BYTES_PER_CHECK = 24 + 1 + (22 + 1 + 4)  # pose in + verdict out + (n+1+4) one-byte map cells, n=22 @0.1 m


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(s) > 2 + k and s[2 + k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def synthetic_setup():
    from oracle import refapi  # map generator only (shared fixture recipe)

    occ = refapi.synthetic_occupancy(SYN_SIZE, 1100, 20261019, SYN_RES)
    params = refapi.shipped_params(0.2, 0.68, ds=SYN_DS)
    return occ, params


def gen_poses(n, seed, size_m):
    rng = np.random.default_rng(seed)
    p = np.empty((n, 3), dtype=np.float64)
    p[:, 0] = rng.uniform(0.0, size_m, n)
    p[:, 1] = rng.uniform(0.0, size_m, n)
    p[:, 2] = rng.uniform(-np.pi, np.pi, n)
    return p


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref) — one process per core, one NavMap copy each, because the reference
# is single-threaded and NavMap is mutable shared state.
_W = {}


def _cpu_worker_init(so_path, params, shm_name, shape, res):
    from multiprocessing import shared_memory

    from oracle import refapi

    shm = shared_memory.SharedMemory(name=shm_name)
    occ = np.ndarray(shape, dtype=np.int8, buffer=shm.buf)
    lib = refapi.RefLib(so_path)
    rm = refapi.RefMap(lib, params)
    rm.set_occupancy(occ, res, [0.0, 0.0, 0.0])
    _W["rp"] = refapi.RefPlanner(rm)
    _W["rm"] = rm
    shm.close()


def _cpu_worker_collision(poses):
    t0 = time.perf_counter()
    v = _W["rp"].collision4(poses)
    return time.perf_counter() - t0, int(v.sum()), len(poses)


class CpuArm:
    def __init__(self, occ, params, cores):
        from multiprocessing import get_context, shared_memory

        from oracle import refapi

        if not os.path.exists(refapi.REF_SO):
            raise RuntimeError("oracle/_ref/libmpros_ref.so missing")
        self.cores = cores
        self.shm = shared_memory.SharedMemory(create=True, size=occ.nbytes)
        np.ndarray(occ.shape, dtype=np.int8, buffer=self.shm.buf)[:] = occ
        self.pool = get_context("spawn").Pool(cores, initializer=_cpu_worker_init,
                                              initargs=(refapi.REF_SO, params, self.shm.name, occ.shape, SYN_RES))

    def collision_rate(self, n_poses, seed):
        poses = gen_poses(n_poses, seed, SYN_SIZE * SYN_RES)
        res = self.pool.map(_cpu_worker_collision, np.array_split(poses, self.cores))
        # workers run concurrently: throughput = total checks / slowest worker's footPrintInCollision4 loop
        return sum(r[2] for r in res) / max(r[0] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()
        self.shm.close()
        self.shm.unlink()


def cpu_sample_text(n, cores):
    return ("%d seeded uniform poses on the synthetic 8192^2 map, %d process(es) x 1 thread, unmodified reference "
            "footPrintInCollision4 loop timed (map preprocessing excluded)" % (n, cores))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    occ, params = synthetic_setup()
    cores = os.cpu_count() or 1
    n = (1 << 20) * cores
    arm = CpuArm(occ, params, cores)
    vals = []
    t0 = time.perf_counter()
    try:
        for s in range(args.warmup + args.steps):
            v = arm.collision_rate(n, POSE_SEED + s)
            if s >= args.warmup:
                vals.append(v)
            if time.perf_counter() - t0 > 200 and vals:
                break
    finally:
        arm.close()
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "collision_checks_per_sec", "value": value, "unit": "checks/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * n / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "collision: synthetic 8192x8192 @0.1 m (seed 20261019), inflation 1.0 m, seeded uniform poses",
                   "poses_per_step": n},
        "cpu_baseline": {"value": value, "unit": "checks/s", "cores": cores, "kind": "reference", "sample": cpu_sample_text(n, cores)},
        "e2e": {"value": value, "unit": "checks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_b200(args):
    import torch
    import torch.distributed as dist

    import mpros_b200

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    occ, params = synthetic_setup()
    ctx = mpros_b200.Context(local_rank)
    ctx.map_config(params)
    t0 = time.perf_counter()
    ctx.set_occupancy(occ, SYN_RES, [0.0, 0.0, 0.0])
    t_pre = time.perf_counter() - t0
    ctx.planner_config(params)

    n = POSES_PER_STEP
    poses_h = torch.from_numpy(gen_poses(n, POSE_SEED + 1000 * rank, SYN_SIZE * SYN_RES)).pin_memory()
    poses_d = poses_h.cuda(non_blocking=False)
    verd_d = torch.zeros(n, dtype=torch.uint8, device="cuda")
    verd_h = torch.zeros(n, dtype=torch.uint8).pin_memory()
    stream = torch.cuda.ExternalStream(ctx.stream())

    def step_dev():
        ctx.collision_check_dev(poses_d.data_ptr(), n, verd_d.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_dev()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    # device-resident timing: one event pair per step (per-kernel durations) + one around the whole region
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_all0.record(stream)
    for k in range(args.steps):
        ev[k][0].record(stream)
        step_dev()
        ev[k][1].record(stream)
    e_all1.record(stream)
    barrier()
    launches = ctx.launch_count() - launches0
    total_ms = e_all0.elapsed_time(e_all1)
    kern_ms = [a.elapsed_time(b) for a, b in ev]

    # end-to-end through the public host-pointer C-ABI call: pinned host poses -> H2D -> kernel -> D2H verdicts
    L = ctx.L
    for _ in range(2):
        ctx._check(L.mpros_collision_check_poses(ctx.h, poses_h.data_ptr(), n, verd_h.data_ptr()))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx._check(L.mpros_collision_check_poses(ctx.h, poses_h.data_ptr(), n, verd_h.data_ptr()))
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    hit_rate = float(verd_h.float().mean())
    assert torch.equal(verd_h, verd_d.cpu()), "host-API and device-API verdicts differ"

    t_max = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    total_ms_max, e2e_ms_max = t_max.tolist()

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        value = world * n * args.steps / (total_ms_max * 1e-3)
        e2e_value = world * n * args.steps / (e2e_ms_max * 1e-3)
        k_ms = float(np.mean(kern_ms))
        achieved = n * ALGO_BYTES_PER_CHECK / (k_ms * 1e-3) / 1e9
        cpu = None
        if args.gpus == 1 and not args.no_cpu:
            try:
                arm = CpuArm(occ, params, 1)
                try:
                    arm.collision_rate(1 << 18, POSE_SEED + 5)
                    v = arm.collision_rate(1 << 22, POSE_SEED + 6)
                finally:
                    arm.close()
                cpu = {"value": v, "unit": "checks/s", "cores": 1, "kind": "reference", "sample": cpu_sample_text(1 << 22, 1)}
            except Exception as e:  # oracle absent: report it, do not fail the GPU bench
                cpu = {"value": None, "unit": "checks/s", "cores": 0, "kind": "unavailable", "sample": str(e)}
        line = {
            "metric": "collision_checks_per_sec", "value": value, "unit": "checks/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "collision: synthetic 8192x8192 @0.1 m (seed 20261019), inflation 1.0 m, 2^24 seeded uniform poses per GPU per step",
                       "poses_per_step_per_gpu": n, "l2_policy": "inputs larger than L2 (403 MB poses per step)",
                       "hit_rate": hit_rate, "preprocess_s_first_call": t_pre},
            "e2e": {"value": e2e_value, "unit": "checks/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "collision_poses_kernel", "kernel_ms": k_ms, "peak_source": peak_kind,
                         "algorithmic_bytes_per_check": ALGO_BYTES_PER_CHECK},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="collision")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
