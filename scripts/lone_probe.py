import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from c2a_b200 import api, meshes, workloads
n = 262144
bvh = api.build_bvh(meshes.torus_knot(512, 32)[0]); model = api.Model(bvh, 0)
poses = workloads.approach_batch(n, 20260002, radius=workloads.KNOT_RADIUS)
f = ("status", "num_ca", "num_bv_tests", "num_tri_tests")
api.solve_batch(model, model, poses[:4096], fields=f)
L = api.lib(); st = (C.c_uint64 * 32)()
L.c2a_b200_phase_stats(1, None)
api.solve_batch(model, model, poses, fields=f)
L.c2a_b200_phase_stats(1, st); s = list(st)
k = max(1, s[24])
print("lone level-3 lane probes:", s[24], "cycles: prologue+pop+meta %.0f | walk %.0f | rect dist %.0f | to-world+bounds %.0f | replay+commit %.0f" % tuple(x / k for x in s[25:30]))
print("lone expand passes", s[14], "cycles/pass %.0f" % (s[15] / max(1, s[14])))
