"""Per-pass latency of a lonely query: the longest query of the knot fixture, alone and 16 copies."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import api, meshes
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/ref_knot_512x32.npz"))
bvh = api.build_bvh(meshes.torus_knot(512, 32)[0]); model = api.Model(bvh, 0)
i = int(np.argmax(g["num_bv_tests"])); f = ("status", "num_bv_tests", "num_tri_tests", "num_ca")
api.solve_batch(model, model, g["poses"][:64], fields=f)
for copies in (1, 16, 32, 128, 4096):
    p = np.repeat(g["poses"][i:i + 1], copies, 0)
    t = time.time(); out = api.solve_batch(model, model, p, fields=f); dt = time.time() - t
    nbv, ntri, nca = int(out["num_bv_tests"][0]), int(out["num_tri_tests"][0]), int(out["num_ca"][0])
    passes = nbv // 2 + ntri + nca
    print(f"copies={copies} {dt:.3f}s nbv={nbv} ntri={ntri} numCA={nca} -> {dt/passes*1e6:.2f} us/pass = {dt/passes*1.9e9:.0f} cycles/pass")
