"""Candidate / trip histograms of rss_rect_dist over a knot batch (variant built with -DC2A_RD_STATS; development aid).
C2A_B200_LIB=variants/rdstats.so python scripts/rd_stats.py [n]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import api, meshes, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
model = api.Model(api.build_bvh(meshes.torus_knot(512, 32)[0]), 0)
poses = workloads.approach_batch(n, 20260002, radius=workloads.KNOT_RADIUS)
out = api.solve_batch(model, model, poses, fields=("status", "num_bv_tests"))
st = (C.c_uint64 * 64)()
assert api.lib().c2a_b200_rd_stats(st) == 0
s = np.array(list(st), dtype=np.float64)
tot = s[0:17].sum()
print("rect-dist calls", int(tot), "BV tests", int(out["num_bv_tests"].sum()))
print("candidates per call (entry predicates passed):", " ".join(f"{i}:{s[i] / tot:.3f}" for i in range(17) if s[i]))
print("trips per call:", " ".join(f"{i}:{s[17 + i] / tot:.3f}" for i in range(9) if s[17 + i]))
print(f"accepted: first of its trip {s[26] / tot:.3f}, second {s[27] / tot:.3f}, none (face separation) {s[28] / tot:.3f}")
print("accepted pair k:", " ".join(f"{k}:{s[32 + k] / tot:.3f}" for k in range(16)))
print(f"lanes active: at loop entry {s[61] / max(1, s[60]):.1f}, per trip {s[63] / max(1, s[62]):.1f}; warp-level trips per warp-level call {s[62] / max(1, s[60]):.2f}")
