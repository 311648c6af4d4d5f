"""Round-2 design study (CPU, oracle): exact-mode CA steps as a depth-first traversal that pops W node pairs per round
(oracle wide_step, the CPU statement of c2a_wide.cuh).  Checks bit-exactness against the sequential port and prints how
much the windows over-evaluate, how many rounds a step takes, how often the ancestor anomaly forces a sequential redo.
Usage: python tests/analysis/wide_study.py [fixture] [n_heaviest] [W ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle
from c2a_b200 import api, meshes
fx = sys.argv[1] if len(sys.argv) > 1 else "ref_knot_512x32"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 6
Ws = [int(x) for x in sys.argv[3:]] or [16, 64, 128]
LB = int(os.environ.get('LEAF_BATCH', '0'))
g = np.load(os.path.join(ROOT, "tests/golden", fx + ".npz"))
nu, nv = (int(x) for x in fx.split("_")[2].split("x"))
bvh = api.build_bvh(meshes.torus_knot(nu, nv)[0]); P = oracle.port()
idx = np.argsort(-g["num_bv_tests"])[:nq] if nq > 0 else np.arange(min(-nq, len(g["poses"])))
poses = g["poses"][idx]
seq = P.solve_batch(bvh, bvh, poses)
print("queries", len(idx), "numCA", seq["numCA"].tolist()[:8], "nbv", seq["num_bv_tests"].tolist()[:8])
for W in Ws:
    out, s = P.solve_wide(bvh, bvh, poses, window=W, leaf_batch=LB)
    for k in seq.dtype.names:
        assert np.array_equal(seq[k], out[k], equal_nan=True) if seq[k].dtype.kind == "f" else np.array_equal(seq[k], out[k]), k
    ws = max(1, s["steps"])
    print(f"W={W}: bit-exact; {ws} exact-mode steps: tests/step {s['wide_tests'] / ws:.0f} (visited {s['wide_tests_visited'] / ws:.0f}"
          f" -> x{s['wide_tests'] / max(1, s['wide_tests_visited']):.2f}), leaves/step {s['wide_leaves'] / ws:.0f} (visited "
          f"{s['wide_leaves_visited'] / ws:.0f}), events/step {s['events'] / ws:.1f}, rounds/step {s['rounds'] / ws:.1f} + {s['leaf_passes'] / ws:.1f} leaf passes "
          f"(sequential: {s['wide_tests_visited'] / 2 / ws + s['wide_leaves_visited'] / ws:.0f}), max stack {s['max_stack']}, "
          f"max unresolved {s['max_unresolved']}, redo {s['redo']} (anomalies {s['anomalies']}, closure {s['closure_fail']})")
