"""Round-2 design study (CPU, oracle): how well does a CA step's visit sequence predict the next step's?
For the heaviest queries of a fixture, replay each step's sequence of visited node pairs against the previous
step's as a STREAM: a cursor walks the previous sequence; a visit found at the cursor is a sequential hit, one found
elsewhere in the previous sequence a jump (the memoised record exists but the stream has to be re-positioned), one
not found a miss (the pair has to be evaluated on the spot).  Usage: python tests/analysis/replay_study.py [fixture] [n]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
from c2a_b200 import api, meshes
fx = sys.argv[1] if len(sys.argv) > 1 else "ref_knot_512x32"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 6
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests/golden", fx + ".npz"))
nu, nv = (int(x) for x in fx.split("_")[2].split("x"))
bvh = api.build_bvh(meshes.torus_knot(nu, nv)[0]); P = oracle.port()
tot = dict(visits=0, seq=0, jump=0, miss=0, steps=0, runs=0)
for i in np.argsort(-g["num_bv_tests"])[:nq]:
    res, steps = P.solve_visits(bvh, bvh, g["poses"][i])
    assert res["num_bv_tests"] == g["num_bv_tests"][i]
    q = dict(visits=0, seq=0, jump=0, miss=0, runs=0)
    for k in range(6, len(steps)):  # exact-mode steps only (numCA > 5)
        prev, cur = steps[k - 1], steps[k]
        pos = {int(v): j for j, v in enumerate(prev)}
        c = 0
        for v in cur:
            j = pos.get(int(v), -1)
            if j < 0: q["miss"] += 1
            elif j == c: q["seq"] += 1; c += 1
            else: q["jump"] += 1; q["runs"] += 1; c = j + 1
        q["visits"] += len(cur)
    tot["steps"] += max(0, len(steps) - 6)
    for k in q: tot[k] += q[k]
    n = max(1, q["visits"])
    print(f"q={i} numCA={res['numCA']} visits/step {n / max(1, len(steps) - 6):.0f}: sequential hit {q['seq'] / n:.4f} jump {q['jump'] / n:.4f} miss {q['miss'] / n:.4f}"
          f" mean run {q['seq'] / max(1, q['runs'] + q['miss']):.1f}")
n = max(1, tot["visits"])
print(f"ALL {tot['steps']} steps, {n} visits: sequential {tot['seq'] / n:.4f} jump {tot['jump'] / n:.4f} miss {tot['miss'] / n:.4f}")
for c_hit, c_jump in ((100, 800), (200, 1500)):
    c_miss = 13700 / 2.85  # today's cost of one committed expansion for a query alone on its warp
    cyc = tot["seq"] * c_hit + tot["jump"] * c_jump + tot["miss"] * c_miss
    print(f"  replay at {c_hit} cycles per streamed record, {c_jump} per re-positioning, {c_miss:.0f} per miss: "
          f"{cyc / n:.0f} cycles per visit vs {c_miss:.0f} today -> {c_miss * n / cyc:.1f}x")
