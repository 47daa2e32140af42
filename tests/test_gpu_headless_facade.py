"""GPU test of the host side: the C++ mirror classes (NavMap / HybridAstar) driven by the headless replacement of
main_planner.cpp, from a PGM + YAML on disk (map_server semantics), against the golden plan recorded from the reference."""
import json
import os
import subprocess

import numpy as np
import pytest

import golden_inputs as gi
from oracle import refapi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "hybrid-astar-planner_b200", "mpros_headless")
GOLD = os.path.join(ROOT, "tests", "golden")


def write_map(tmp_path, name):
    m = refapi.load_map(name)
    occ = m["occupancy"]
    img = np.full(occ.shape, 153, dtype=np.uint8)  # p = 0.4: neither free nor occupied -> -1
    img[occ == 100] = 0
    img[occ == 0] = 254
    img = img[::-1]  # image row 0 is the top
    pgm = tmp_path / (name + ".pgm")
    with open(pgm, "wb") as f:
        f.write(b"P5\n# test fixture\n%d %d\n255\n" % (occ.shape[1], occ.shape[0]))
        f.write(img.tobytes())
    yml = tmp_path / (name + ".yaml")
    yml.write_text("image: ./%s.pgm\nresolution: %r # comment\norigin: [%r, %r, %r]\nnegate: 0\noccupied_thresh: %r\nfree_thresh: %r\n" % (
        name, m["resolution"], m["origin"][0], m["origin"][1], m["origin"][2], m["occupied_thresh"], m["free_thresh"]))
    return str(yml)


@pytest.mark.parametrize("ci", [0, 1, 3])
def test_headless_driver_matches_golden_plan(tmp_path, ci):
    if not os.path.exists(BIN):
        pytest.fail("mpros_headless not built (run __graft_entry__.build())")
    name, ds, num_steer, start, goal = gi.PLAN_CASES[ci]
    yml = write_map(tmp_path, name)
    args = [BIN, yml] + [repr(v) for v in start + goal]
    if num_steer:
        args.append("/mpros/planner/steering_discretization=%d" % num_steer)
    out = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    z = np.load(os.path.join(GOLD, "plans.npz"))
    ok, g_ref, prims_ref, nodes_ref, plen = z["stats_%d" % ci]
    assert r["plan_ok"] == int(ok) and r["path_in_collision"] == 0
    assert abs(r["goal_g"] - g_ref) <= 1e-4 * g_ref  # canonical Voronoi tie-breaks: within the stated tolerance
    path = np.array(r["path"])
    assert np.allclose(path[0, :3], start) and np.allclose(path[-1, :3], goal, atol=0.6)
    if ci in (0, 1, 3):  # these cases do not touch a tie-broken Voronoi cell: identical search
        assert r["n_primitives"] == int(prims_ref)
        assert np.allclose(path, z["path_%d" % ci], rtol=0, atol=1e-8)
        # HybridAstar::getNewPath(): the up-sampled path of setPath (hybrid_a_star.cpp:814-838)
        newp = np.array(r["new_path"])
        want = z["newpath_%d" % ci]
        assert newp.shape == want.shape
        assert np.allclose(newp, want, rtol=0, atol=1e-7)

    # the five map messages main_planner.cpp:216-220 publishes (NavMap::get*MapMsg -> vecToMsg, nav_map.cpp:654-716), built by
    # the facade from mpros_map_layer_msg: geometry as the reference writes it, data byte-exact (FNV-1a-64) against the
    # unmodified reference (the Voronoi message follows the canonical tie-break contract, tests/test_gpu_layer_msg.py)
    m = refapi.load_map(name)
    H, W = m["occupancy"].shape
    for layer in ["static", "voronoi", "static_orig", "inflate_orig", "goal"]:
        g = r["msgs"][layer]
        assert g["frame_id"] == "map" and (g["height"], g["width"]) == (H, W) and g["n"] == H * W
        assert abs(g["resolution"] - m["resolution"]) < 1e-6 and g["origin"] == [m["origin"][0], m["origin"][1]]
    if os.path.exists(refapi.REF_SO):
        params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"], ds=ds, num_steer=num_steer)
        rm = refapi.RefMap(refapi.RefLib(), params)
        rm.set_occupancy(m["occupancy"], m["resolution"], m["origin"])
        rm.update_goal(goal[0], goal[1])
        for layer in ["static", "static_orig", "inflate_orig", "goal"]:
            data, _ = refapi.ref_layer_msg(rm, layer)
            h = 1469598103934665603
            for b in np.asarray(data, dtype=np.int8).view(np.uint8).tolist():
                h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
            assert r["msgs"][layer]["fnv"] == "%016x" % h, layer
        rm.close()
    # HybridAstar::getSearchTree() -> vis.PublishSearchedTree (main_planner.cpp:259): leaf_num segments per evaluated primitive
    assert r["search_tree_segments"] == r["n_primitives"] * int(np.ceil(2.0 / 0.1))
    assert r["new_path_in_collision"] == 0 and r["rs_shot_from_start"] in (0, 1)
