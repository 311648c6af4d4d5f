import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import api, meshes, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
bvh = api.build_bvh(meshes.torus_knot(512, 32)[0]); model = api.Model(bvh, 0)
poses = workloads.approach_batch(n, 20260002, radius=workloads.KNOT_RADIUS)
f = ("status", "num_ca", "num_bv_tests", "num_tri_tests", "collisionfree")
out = api.solve_batch(model, model, poses, fields=f)
nbv = out["num_bv_tests"].astype(np.int64); nca = out["num_ca"]
print("nbv percentiles 50/90/99/99.9/99.99/max:", [int(x) for x in np.percentile(nbv, [50, 90, 99, 99.9, 99.99, 100])])
print("numCA percentiles:", [int(x) for x in np.percentile(nca, [50, 90, 99, 99.9, 99.99, 100])], "count numCA>=100:", int((nca >= 100).sum()), "==151:", int((nca >= 151).sum()))
top = np.argsort(-nbv)[:10]
print("top10 nbv:", nbv[top], "numCA:", nca[top], "ntri:", out["num_tri_tests"][top])
order = api.schedule_order(model, model, api.motions_from_poses(poses))
rank = np.empty(n, dtype=np.int64); rank[order] = np.arange(n)
print("claim rank of top10 (0 = first):", rank[top])
print("mean claim rank of top 1000 by nbv / n:", rank[np.argsort(-nbv)[:1000]].mean() / n)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/tail_stats_1M.npz", nbv=nbv.astype(np.int32), nca=nca.astype(np.int16), cf=out["collisionfree"].astype(np.int8))
