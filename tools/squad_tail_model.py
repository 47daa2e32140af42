"""CPU model of the end of a squad pass (hybrid-astar-planner_b200/csrc/plan_squad.cuh).

Input: the iteration counts of the 16 384 queries of the headline batch (tests/golden/synthetic_plans.npz). Model: 148 CTAs
of 24 warps, every active warp advances its query by one iteration per round, a round takes 20 us + 1.3 us per active warp,
tickets are drawn as warps fall idle; after the tickets have run out a CTA with at most `evict_at` active warps hands its
queries over (to the team pass, or - `adopt` - onto a migration list from which idle warps of CTAs with at least `adopt_min`
active warps take them), and queries older than `age_limit` iterations go to the team pass for good. The team pass is
modelled as 28 M iterations/s and 7.5 us per iteration for its longest query.

With the latency rules it reproduces what the kernels do (squad pass 0.48 s, 1 183 queries and 11.2 M iterations handed over;
measured 0.45-0.53 s, 1 184, 11.1 M), so it was used to choose the rules of the throughput policy before GPU time was spent
on them. `plans_per_s` is the streaming rate with perfect packing of the SMs (an upper bound), `lone_batch` the time of one
batch alone.

    python tools/squad_tail_model.py
"""
import heapq
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ITS = np.load(os.path.join(ROOT, "tests", "golden", "synthetic_plans.npz"))["port_iterations"].astype(np.int64)


def model(evict_at=8, adopt_min=9, W=24, C=148, adopt=True, age_limit=None, cap=10 ** 9, team_rate=28e6, its=None):
    ITS = globals()["ITS"] if its is None else np.asarray(its, dtype=np.int64)
    order = list(range(len(ITS)))
    nxt = 0
    rem = ITS.copy() + 30  # ~30 rounds of goal wavefront before the first expansion
    done_it = np.zeros(len(ITS), dtype=np.int64)
    ctas = [[] for _ in range(C)]
    queue, team, moves = [], [], 0
    for c in range(C):
        for _ in range(W):
            if nxt < len(order):
                ctas[c].append(order[nxt])
                nxt += 1
    pq = [(0.0, c) for c in range(C)]
    heapq.heapify(pq)
    sm_seconds, tend = 0.0, 0.0
    while pq:
        tc, c = heapq.heappop(pq)
        exhausted = nxt >= len(order)
        if age_limit is not None or cap < 10 ** 9:
            keep = []
            for q in ctas[c]:
                if (age_limit is not None and exhausted and done_it[q] >= age_limit) or done_it[q] >= cap:
                    team.append(q)
                else:
                    keep.append(q)
            ctas[c] = keep
        A = len(ctas[c])
        if exhausted and 0 < A <= evict_at:
            keep = []
            for q in ctas[c]:
                if done_it[q] >= 32:
                    (queue if adopt else team).append(q)
                    moves += 1
                else:
                    keep.append(q)
            ctas[c] = keep
            A = len(keep)
        if exhausted and adopt and A >= adopt_min:
            while len(ctas[c]) < W and queue:
                ctas[c].append(queue.pop(0))
            A = len(ctas[c])
        if A == 0 and exhausted:
            sm_seconds += tc
            tend = max(tend, tc)
            continue
        dur = 20e-6 + 1.3e-6 * A
        newl = []
        for q in ctas[c]:
            rem[q] -= 1
            done_it[q] += 1
            if rem[q] > 0:
                newl.append(q)
        while len(newl) < W and nxt < len(order):
            newl.append(order[nxt])
            nxt += 1
        ctas[c] = newl
        heapq.heappush(pq, (tc + dur, c))
    team += queue
    left = int(sum(rem[q] for q in team))
    longest = max([rem[q] for q in team], default=0)
    step = (sm_seconds + left * C / team_rate) / C
    return {"squad_end_s": round(tend, 3), "hand_overs": moves, "team_queries": len(team), "team_iterations": left,
            "team_tail_s": round(float(longest) * 7.5e-6, 3), "plans_per_s": round(len(ITS) / step),
            "lone_batch_s": round(tend + max(left / team_rate, float(longest) * 7.5e-6), 3)}


if __name__ == "__main__":
    for name, kw in [("latency policy: evict at <= 8 -> team pass", dict(adopt=False)),
                     ("evict at <= 4 -> team pass", dict(adopt=False, evict_at=4)),
                     ("evict at <= 12 -> team pass", dict(adopt=False, evict_at=12)),
                     ("throughput policy: adopt, age limit 20 000", dict(age_limit=20000)),
                     ("throughput policy: adopt, age limit 30 000", dict(age_limit=30000)),
                     ("throughput policy: adopt, age limit 40 000", dict(age_limit=40000))]:
        print("%-48s %s" % (name, model(**kw)))
