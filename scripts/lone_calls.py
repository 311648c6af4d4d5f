"""Where a single C2A_Solve spends its time on the GPU: the CCDDemo's 303 frames, one call each, seeds carried
(development aid): kernel times of the main and the wide kernel per call next to the work the call did."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from c2a_b200 import api
mesh = np.load(os.path.join(ROOT, "tests/golden/bunny_mesh.npz"))
tris = mesh["verts"][mesh["vidx"]].reshape(-1, 9).copy()
m = api.Model(api.build_bvh(tris), 0)
demo = np.load(os.path.join(ROOT, "tests/golden/demo_poses.npy"))
F = ("status", "collisionfree", "num_ca", "num_bv_tests", "num_tri_tests", "toc", "distance", "last_tri")
kt = (C.c_double * 3)()
api.solve_batch(m, m, demo[:1], fields=F)
rows = []
sa = sb = 0
for i in range(len(demo)):
    t = time.perf_counter()
    o = api.solve_batch(m, m, demo[i:i + 1], [sa], [sb], fields=F)
    dt = time.perf_counter() - t
    api.lib().c2a_b200_kernel_times(kt)
    rows.append((dt * 1e3, kt[0], kt[1], kt[2], o["num_bv_tests"][0], o["num_ca"][0], o["num_tri_tests"][0]))
    if o["last_tri"][0, 0] >= 0: sa = int(o["last_tri"][0, 0])
    if o["last_tri"][0, 1] >= 0: sb = int(o["last_tri"][0, 1])
r = np.array(rows)
print(f"{len(r)} calls: mean call {r[:, 0].mean():.3f} ms = main kernel {r[:, 1].mean():.3f} + wide kernel {r[:, 2].mean():.3f} + translation {r[:, 3].mean():.3f} + host/launch {(r[:, 0] - r[:, 1] - r[:, 2] - r[:, 3]).mean():.3f}")
print(f"mean BV tests {r[:, 4].mean():.0f}, mean numCA {r[:, 5].mean():.1f}, mean tri tests {r[:, 6].mean():.0f}")
for lo, hi in ((0, 2), (2, 6), (6, 12), (12, 1000)):
    s = (r[:, 5] >= lo) & (r[:, 5] < hi)
    if s.any(): print(f"  numCA in [{lo},{hi}): {s.sum()} calls, call {r[s, 0].mean():.3f} ms, main {r[s, 1].mean():.3f}, wide {r[s, 2].mean():.3f}, BV tests {r[s, 4].mean():.0f}")
if os.environ.get("WIDE_STATS"):
    # with a library built with -DC2A_WIDE_STATS=1 (scripts/build_variant.py): cycle shares of the wide kernel's phases
    L = api.lib(); ws = (C.c_uint64 * 32)()
    L.c2a_b200_wide_stats.argtypes = [C.c_int32, C.c_void_p]
    L.c2a_b200_wide_stats(1, None)
    sa = sb = 0
    for i in range(len(demo)):
        o = api.solve_batch(m, m, demo[i:i + 1], [sa], [sb], fields=F)
        if o["last_tri"][0, 0] >= 0: sa = int(o["last_tri"][0, 0])
        if o["last_tri"][0, 1] >= 0: sb = int(o["last_tri"][0, 1])
    L.c2a_b200_wide_stats(1, ws); w = list(ws)
    cyc = sum(w[7:12])
    print(f"wide kernel over the 303 calls: {w[12]} queries, {w[0]} steps ({w[13]} one-pair, {w[1]} redone); per step: {w[2] / w[0]:.1f} rounds x {w[7] / max(1, w[2]):.0f} cyc, "
          f"{w[3] / w[0]:.1f} leaf passes x {w[8] / max(1, w[3]):.0f} cyc, {w[4] / w[0]:.0f} tests, {w[5] / w[0]:.0f} tri tests, {w[6] / w[0]:.1f} events; "
          f"cycles per step {cyc / w[0]:.0f}: expand {w[7] / cyc:.2f} leaf {w[8] / cyc:.2f} resolve {w[9] / cyc:.2f} fold {w[10] / cyc:.2f} setup {w[11] / cyc:.2f}")
    r_ = max(1, w[2])
    print(f"inside an EXPAND pass (lane 0, cycles per round): window select {w[16] / r_:.0f}, pop + node meta {w[17] / r_:.0f}, child fetch + transform {w[18] / r_:.0f}, "
          f"rectangle distance {w[19] / r_:.0f}, motion bounds {w[20] / r_:.0f}, wait for the slowest lane {w[21] / r_:.0f}, records {w[22] / r_:.0f}, "
          f"leaf list + pushes {w[23] / r_:.0f}, rest (counters) {(w[7] - sum(w[16:24])) / r_:.0f}")
