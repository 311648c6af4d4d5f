"""Kernel times (CUDA events inside the library, no in-kernel counters) of knot-512x32 approach batches of several sizes:
python scripts/kernel_ms.py [n ...]   (development aid for A/B runs under the C2A_B200_* environment knobs)"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import api, meshes, workloads
sizes = [int(x) for x in sys.argv[1:]] or [65536, 131072, 262144, 1048576]
tris = meshes.torus_knot(512, 32)[0]
model = api.Model(api.build_bvh(tris), 0)
f = ("status", "collisionfree", "num_ca", "num_bv_tests", "toc", "distance")
kt = (C.c_double * 3)()
poses_all = workloads.approach_batch(max(sizes), 20260002, radius=workloads.KNOT_RADIUS)
api.solve_batch(model, model, poses_all[:4096], fields=f)
for n in sizes:
    best = None
    for rep in range(int(os.environ.get("REPS", "2"))):
        out = api.solve_batch(model, model, poses_all[:n], fields=f)
        api.lib().c2a_b200_kernel_times(kt)
        if best is None or kt[0] + kt[1] < best[0] + best[1]: best = list(kt)
    print(f"n={n}: solve {best[0]:.1f} ms + wide {best[1]:.1f} ms = {best[0] + best[1]:.1f} ms  (checksum {int(out['num_bv_tests'].sum())})", flush=True)
