"""the map-layer messages of nav_map.cpp:654-716 (vecToMsg) through mpros_map_layer_msg, byte for byte
against the unmodified reference, including its int8 wrap-around of goal-map values above 127."""
import numpy as np
import pytest

from helpers import make_pair
from oracle import refapi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,ds,goal", [("reverse_parking", 1, (11.5, 2.1)), ("ARENA_part", 4, (5.0, 5.0)), ("parking4", 1, (10.0, -2.0))])
def test_layer_messages_match_reference(reflib, gpu_ctx_factory, name, ds, goal):
    ctx = gpu_ctx_factory()
    rm, params, m = make_pair(reflib, ctx, name, ds=ds, goal=goal)
    ctx.upload_voronoi(rm.layer("voronoi"))  # the float field itself is compared elsewhere (canonical tie-breaks)
    inf = ctx.info()
    for layer in ["static", "goal", "voronoi", "static_orig", "inflate_orig"]:
        want, (h, w, res) = refapi.ref_layer_msg(rm, layer)
        got = ctx.layer_msg(layer)
        orig = layer.endswith("_orig")
        assert (h, w) == ((inf["map_height_orig"], inf["map_width_orig"]) if orig else (inf["map_height"], inf["map_width"]))
        assert np.float32(res) == np.float32(inf["map_resolution_orig"] if orig else inf["map_resolution"])
        assert want.shape == got.shape
        assert np.array_equal(want, got), "%s: %d bytes differ" % (layer, int((want != got).sum()))
    assert (refapi.ref_layer_msg(rm, "goal")[0] < 0).any()  # the wrap-around really is exercised
    rm.close()


def test_layer_message_is_empty_before_the_layer_exists(gpu_ctx_factory, reflib):
    import mpros_b200

    ctx = mpros_b200.Context(0)
    assert ctx.layer_msg("static").size == 0  # no map yet: vecToMsg returns an empty message
    m = refapi.load_map("reverse_parking")
    params = refapi.shipped_params(m["free_thresh"], m["occupied_thresh"])
    ctx.map_config(params)
    ctx.set_occupancy(m["occupancy"], m["resolution"], m["origin"])
    assert ctx.layer_msg("static").size == 150 * 200 and ctx.layer_msg("goal").size == 0  # no goal map yet
    ctx.close()
