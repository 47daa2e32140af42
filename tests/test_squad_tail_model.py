"""tools/squad_tail_model.py, the CPU model the hand-over rules of the squad pass were chosen with, on one
eighth of the headline batch (every eighth query, 18 CTAs of 24 warps): the qualitative claims the two policies rest on.
CPU only; the kernels themselves are held to identical results in tests/test_gpu_synthetic_fullsize.py."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_model():
    spec = importlib.util.spec_from_file_location("squad_tail_model", os.path.join(ROOT, "tools", "squad_tail_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_policies_of_the_squad_pass():
    m = load_model()
    its = m.ITS[::8]
    C = 18
    total = int(its.sum())
    rate = 28e6 * C / 148  # the team pass's share of the device
    lat = m.model(adopt=False, C=C, its=its, team_rate=rate)
    thr = m.model(adopt=True, age_limit=30000, C=C, its=its, team_rate=rate)
    # latency policy: a CTA hands over at most 8 queries, once; roughly a third of the iterations are left for the team pass
    assert lat["hand_overs"] <= 8 * C and lat["team_queries"] == lat["hand_overs"]
    assert 0.2 * total < lat["team_iterations"] < 0.5 * total
    # throughput policy: the same queries migrate, most are adopted; the team pass keeps the old ones (age limit) and a few
    # leftovers: well under half of the latency policy's iterations, at the price of a longer squad pass and lone batch
    assert thr["team_queries"] < lat["team_queries"] // 3
    assert thr["team_iterations"] < 0.5 * lat["team_iterations"]
    assert thr["squad_end_s"] > lat["squad_end_s"] and thr["lone_batch_s"] > lat["lone_batch_s"]
    assert thr["plans_per_s"] > 1.1 * lat["plans_per_s"]
