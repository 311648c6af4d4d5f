"""Exactly the bench workload, one c2a_solve_kernel launch over --batch queries (after a tiny warm-up launch).
For single-metric ncu passes over the real launch (e.g. DRAM traffic):  ncu --launch-skip 1 -c 1 ..."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import api, meshes, workloads
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=1000000)
ap.add_argument("--bunny", action="store_true", help="config 2's models and poses instead of config 3's")
a = ap.parse_args()
if a.bunny:
    m = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/bunny_mesh.npz"))
    bvh = api.build_bvh(m["verts"][m["vidx"]].reshape(-1, 9).copy()); model = api.Model(bvh, 0)
    poses = workloads.approach_batch(a.batch, 20260001)
else:
    bvh = api.build_bvh(meshes.torus_knot(512, 32)[0]); model = api.Model(bvh, 0)
    poses = workloads.approach_batch(a.batch, 20260002, radius=workloads.KNOT_RADIUS)
f = ("status", "num_bv_tests", "num_tri_tests", "num_ca")
api.solve_batch(model, model, poses[:256], fields=f)
out = api.solve_batch(model, model, poses, fields=f)
nbv, ntri = int(out["num_bv_tests"].sum()), int(out["num_tri_tests"].sum())
print("launch:", a.batch, "queries nbv", nbv, "ntri", ntri, "algorithmic bytes", 208 * nbv + 144 * ntri + 448 * a.batch)
