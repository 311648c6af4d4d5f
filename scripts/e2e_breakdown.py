"""Where the host-buffer call spends its wall clock (development aid): python scripts/e2e_breakdown.py [batch]"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from c2a_b200 import api, meshes, workloads

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
model = api.Model(api.build_bvh(meshes.torus_knot(512, 32)[0]), 0)
poses = workloads.approach_batch(B, 20260002, radius=workloads.KNOT_RADIUS)
fields = ("status", "collisionfree", "num_ca", "num_bv_tests", "num_tri_tests", "toc", "distance", "pose_toc")
names = ("alloc", "motions", "order", "enqueue", "kernel_wait", "d2h", "release", "total")
for rep in range(3):
    t0 = time.perf_counter()
    out = api.solve_batch(model, model, poses, fields=fields)
    wall = time.perf_counter() - t0
    t = (C.c_double * 8)()
    api.lib().c2a_b200_host_timing(t)
    print(f"rep {rep}: python wall {wall:.3f} s | " + " ".join(f"{k}={v:.3f}" for k, v in zip(names, t)), flush=True)
