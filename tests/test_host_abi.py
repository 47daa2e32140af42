"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/mpros_gpu.h declares, and
fails LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

import mpros_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = mpros_b200.load_library()
    names = mpros_b200.exported_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_header_cites_reference_interfaces():
    hdr = open(os.path.join(ROOT, "include", "mpros_gpu.h")).read()
    # every entry point names the reference interface it replaces (file:line)
    for cite in ["nav_map.cpp:42-162", "nav_map.cpp:279-349", "hybrid_a_star.cpp:617-680", "hybrid_a_star.cpp:304-402",
                 "rs_curve.cpp:111-222", "rs_statespace.h:37-54", "hybrid_a_star.cpp:52-118"]:
        assert cite in hdr, cite
    assert "torch" not in hdr.lower()


def test_defaults_mirror_reference_config():
    lib = mpros_b200.load_library()
    mp = mpros_b200.MapParams()
    lib.mpros_map_params_default(C.byref(mp))
    assert (mp.free_thresh, mp.occupied_thresh, mp.downsampling_factor, mp.obstacle_inflation_radius) == (0.2, 0.68, 1, 0.0)
    assert (mp.voro_fallout_rate, mp.voro_max_region) == (10.0, 50.0)
    pp = mpros_b200.PlannerParams()
    lib.mpros_planner_params_default(C.byref(pp))
    assert (pp.max_curvature, pp.width, pp.wheelbase, pp.length, pp.max_steering_angle) == (0.4, 2.0, 3.0, 4.2, 0.35)
    assert (pp.yaw_discretization, pp.step_size, pp.steering_discretization, pp.max_iteration) == (36, 1.1, 3, 10000)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(mpros_b200.MprosError) as e:
        mpros_b200.Context(0)
    assert "CUDA" in str(e.value)


def test_product_never_touches_the_oracle():
    """Nothing under the package (the product path) may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "hybrid-astar-planner_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(from|import)\s+oracle|#include\s+\"[^\"]*oracle|libmpros_oracle|libmpros_ref", txt), os.path.join(dirpath, f)


def test_plan_lanes_constant_matches_header():
    import re

    hdr = open(os.path.join(ROOT, "include", "mpros_gpu.h")).read()
    m = re.search(r"#define\s+MPROS_PLAN_LANES\s+(\d+)", hdr)
    assert m and int(m.group(1)) == mpros_b200.PLAN_LANES


def test_bench_workloads_are_documented():
    """Every --workload of bench.py appears in its usage line and in DESIGN.md's measurement section."""
    import re

    src = open(os.path.join(ROOT, "bench.py")).read()
    choices = re.search(r'"--workload", default="plan", choices=\[([^\]]*)\]', src).group(1)
    names = re.findall(r'"(\w+)"', choices)
    assert set(names) >= {"plan", "collision", "expand", "preprocess", "rs", "arena"}
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    for w in names:
        if w != "plan":
            assert "--workload %s" % w in design, w
