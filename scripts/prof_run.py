"""Short run for ncu captures: c2a_solve_kernel over a batch of knot queries whose CA length is bounded
(so that one launch is short enough to be replayed ~40 times by `ncu --set full`).
Launch 1 (unprofiled, skipped with --launch-skip 1) finds the per-query CA counts; launch 2 is the capture."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import api, meshes, workloads
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--max-ca", type=int, default=14)
a = ap.parse_args()
bvh = api.build_bvh(meshes.torus_knot(512, 32)[0])
model = api.Model(bvh, 0)
poses = workloads.approach_batch(3 * a.batch, 20260002, radius=workloads.KNOT_RADIUS)
f = ("status", "num_ca", "num_bv_tests", "num_tri_tests")
first = api.solve_batch(model, model, poses, fields=f)
keep = np.where(first["num_ca"] <= a.max_ca)[0][:a.batch]
out = api.solve_batch(model, model, poses[keep], fields=f)
print("profiled launch:", len(keep), "queries, mean nbv", out["num_bv_tests"].mean(), "mean ntri", out["num_tri_tests"].mean(),
      "mean numCA", out["num_ca"].mean())
