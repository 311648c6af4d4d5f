import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import api, meshes
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/ref_knot_512x32.npz"))
bvh = api.build_bvh(meshes.torus_knot(512, 32)[0]); model = api.Model(bvh, 0)
i = int(np.argmax(g["num_bv_tests"])); f = ("status", "num_bv_tests", "num_tri_tests", "num_ca")
api.solve_batch(model, model, g["poses"][:64], fields=f)
L = api.lib(); st = (C.c_uint64 * 20)()
L.c2a_b200_phase_stats(1, None)
t = time.time(); out = api.solve_batch(model, model, g["poses"][i:i + 1], fields=f); dt = time.time() - t
L.c2a_b200_phase_stats(1, st); s = list(st)
print(f"lone query {dt:.3f}s nbv={out['num_bv_tests'][0]} ntri={out['num_tri_tests'][0]} numCA={out['num_ca'][0]}")
print(f"expand passes {s[0]}, leaf passes {s[2]}, advance passes {s[4]}; 32-lane look-ahead passes {s[9]} committing {s[10]} levels ({s[10]/max(1,s[9]):.2f}/pass)")
print(f"time per pass overall {dt/(s[0]+s[2]+s[4])*1e6:.2f} us")
L.c2a_b200_phase_stats(0, None)
