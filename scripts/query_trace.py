"""Per-query timeline of one batch (development aid): when each query was claimed and finished, against its cost.
Usage: python scripts/query_trace.py [n]   (env C2A_B200_PRIO=heavy,lead to vary the in-warp priority)
Needs a library built with the solve kernel's counters: python scripts/build_variant.py stats -DC2A_SOLVE_STATS=1 -DC2A_WIDE_STATS=1, then C2A_B200_LIB=$PWD/variants/stats.so (the product build has them compiled out: they cost 6 %)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import api, meshes, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
bvh = api.build_bvh(meshes.torus_knot(512, 32)[0]); model = api.Model(bvh, 0)
poses = workloads.approach_batch(n, 20260002, radius=workloads.KNOT_RADIUS)
f = ("status", "num_ca", "num_bv_tests", "num_tri_tests")
api.solve_batch(model, model, poses[:4096], fields=f)
L = api.lib()
api._check(L.c2a_b200_query_trace(C.c_int64(n), None))
out = api.solve_batch(model, model, poses, fields=f)
tr = np.zeros((n, 2), dtype=np.uint64)
api._check(L.c2a_b200_query_trace(C.c_int64(0), tr.ctypes.data_as(C.c_void_p)))
L.c2a_b200_query_trace(C.c_int64(-1), None)
t0 = tr[:, 0].min()
claim = (tr[:, 0] - t0) / 1e9; done = (tr[:, 1] - t0) / 1e9
nbv = out["num_bv_tests"].astype(np.int64); nca = out["num_ca"]
print(f"n={n} last claim {claim.max():.3f}s last finish {done.max():.3f}s")
last = np.argsort(-done)[:15]
print("the 15 queries that finish last:")
for i in last:
    print(f"  q={i} claim {claim[i]:.3f} done {done[i]:.3f} dur {done[i]-claim[i]:.3f} nbv {nbv[i]} numCA {nca[i]} ntri {out['num_tri_tests'][i]}")
for thr in (0.25, 0.5, 1.0, 1.5, 2.0):
    late = done > claim.max() + thr
    print(f"  finishing more than {thr}s after the last claim: {int(late.sum())} queries, median nbv {int(np.median(nbv[late])) if late.any() else 0}, median claim {np.median(claim[late]) if late.any() else 0:.2f}")
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/query_trace.npz", claim=claim.astype(np.float32), done=done.astype(np.float32), nbv=nbv.astype(np.int32), nca=nca.astype(np.int16))
