"""Phase fill statistics of the solve kernel on a bounded knot batch (development aid).
Needs a library built with the solve kernel's counters: python scripts/build_variant.py stats -DC2A_SOLVE_STATS=1 -DC2A_WIDE_STATS=1, then C2A_B200_LIB=$PWD/variants/stats.so (the product build has them compiled out: they cost 6 %)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import api, meshes, workloads
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
bvh = api.build_bvh(meshes.torus_knot(512, 32)[0]); model = api.Model(bvh, 0)
poses = workloads.approach_batch(n, 20260002, radius=workloads.KNOT_RADIUS)
f = ("status", "num_ca", "num_bv_tests", "num_tri_tests")
api.solve_batch(model, model, poses[:4096], fields=f)
L = api.lib(); st = (C.c_uint64 * 20)()
for rep in range(2):
    L.c2a_b200_phase_stats(1, None)
    t = time.time(); out = api.solve_batch(model, model, poses, fields=f); dt = time.time() - t
    L.c2a_b200_phase_stats(1, st)
    s = list(st)
    print(f"n={n} {dt:.3f}s {n/dt:.0f} q/s nbv={out['num_bv_tests'].sum()} ntri={out['num_tri_tests'].sum()}")
    print(f"  kernel timeline: drained at {(s[7]-s[6])/1e9:.3f}s, last slot retired at {(s[8]-s[6])/1e9:.3f}s")
    for i, name in enumerate(("expand", "leaf", "advance")):
        print(f"  {name}: passes {s[2*i]} lanes/pass {s[2*i+1]/max(1,s[2*i]):.2f} cycles/pass {s[11+i]/max(1,s[2*i]):.0f} share of warp cycles {s[11+i]/max(1,sum(s[11:14])):.3f}")
    print(f"  alone on a warp: look-ahead passes {s[9]} levels/pass {s[10]/max(1,s[9]):.2f}; " + " ".join(f"{nm} {s[14+2*i]} passes x {s[15+2*i]/max(1,s[14+2*i]):.0f} cycles" for i, nm in enumerate(("expand", "leaf", "advance"))))
L.c2a_b200_phase_stats(0, None)
