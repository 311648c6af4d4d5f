"""Timeline and counters of c2a_solve_kernel + c2a_wide_kernel on knot batches of several sizes (development aid).
Usage: python scripts/wide_stats.py [n ...]   (env: C2A_B200_NO_WIDE, C2A_B200_SPILL_LIVE, C2A_B200_WIDE_WINDOW, CHECK=k)
Needs a library built with the solve kernel's counters: python scripts/build_variant.py stats -DC2A_SOLVE_STATS=1 -DC2A_WIDE_STATS=1, then C2A_B200_LIB=$PWD/variants/stats.so (the product build has them compiled out: they cost 6 %)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import api, meshes, workloads
sizes = [int(x) for x in sys.argv[1:]] or [1, 4096, 65536, 262144]
knot = tuple(int(x) for x in os.environ.get("KNOT", "512x32").split("x"))
tris = meshes.torus_knot(*knot)[0]
bvh = api.build_bvh(tris); model = api.Model(bvh, 0)
f = ("status", "collisionfree", "num_ca", "num_bv_tests", "num_tri_tests", "toc", "distance", "mint", "p1p2", "pose_toc", "last_tri")
L = api.lib(); st = (C.c_uint64 * 20)(); ws = (C.c_uint64 * 32)(); kt = (C.c_double * 3)()
L.c2a_b200_wide_stats.argtypes = [C.c_int32, C.c_void_p]
poses_all = workloads.approach_batch(max(sizes), 20260002, radius=workloads.KNOT_RADIUS)
api.solve_batch(model, model, poses_all[:4096], fields=f)
check = int(os.environ.get("CHECK", "0"))
for n in sizes:
    poses = poses_all[:n]
    L.c2a_b200_phase_stats(1, None); L.c2a_b200_wide_stats(1, None)
    t = time.time(); out = api.solve_batch(model, model, poses, fields=f); dt = time.time() - t
    L.c2a_b200_phase_stats(1, st); L.c2a_b200_wide_stats(1, ws)
    kms = list(kt) if hasattr(L, 'c2a_b200_kernel_times') and L.c2a_b200_kernel_times(kt) == 0 else [0, 0, 0]
    s, w = list(st), list(ws)
    nbv = int(out["num_bv_tests"].sum())
    print(f"n={n}: {dt:.4f}s  {n / dt:.0f} q/s  {nbv / dt / 1e9:.3f} G BV tests/s  (nbv {nbv}, max per query {int(out['num_bv_tests'].max())});"
          f" kernels: solve {kms[0]:.2f} ms, wide {kms[1]:.2f} ms, translation {kms[2]:.3f} ms")
    print(f"  solve kernel: drained at {(s[7] - s[6]) / 1e9:.3f}s, last slot at {(s[8] - s[6]) / 1e9:.3f}s; "
          f"EXPAND {s[11] / max(1, s[0]):.0f} cyc/pass, LEAF {s[12] / max(1, s[2]):.0f} cyc/pass at {s[3] / max(1, s[2]):.1f} lanes")
    if w[12]:
        cyc = sum(w[7:12])
        print(f"  wide kernel: {w[12]} queries, {w[0]} steps ({w[13]} one-pair, {w[1]} redone), {(w[15] - w[14]) / 1e6:.2f} ms from first block to last; "
              f"per step: {w[2] / w[0]:.1f} rounds x {w[7] / max(1, w[2]):.0f} cyc, {w[3] / w[0]:.1f} leaf passes x {w[8] / max(1, w[3]):.0f} cyc, "
              f"{w[4] / w[0]:.0f} tests, {w[5] / w[0]:.0f} tri tests, {w[6] / w[0]:.1f} events; "
              f"cycle share expand {w[7] / cyc:.2f} leaf {w[8] / cyc:.2f} resolve {w[9] / cyc:.2f} fold {w[10] / cyc:.2f} setup {w[11] / cyc:.2f}")
    if check:
        import oracle
        k = min(check, n)
        idx = np.unique(np.concatenate([np.argsort(-out["num_bv_tests"])[:k // 2], np.random.default_rng(1).choice(n, k - k // 2, replace=False)])) if n > k else np.arange(n)
        ref = oracle.port().solve_batch(bvh, bvh, poses[idx], threads=os.cpu_count())  # (the port also reports p1/p2 and last_tri)
        bad = {}
        for name, rn in (("collisionfree", "collisionfree"), ("num_ca", "numCA"), ("num_bv_tests", "num_bv_tests"), ("num_tri_tests", "num_tri_tests"),
                         ("toc", "toc"), ("distance", "distance"), ("mint", "mint")):
            a, b = out[name][idx], ref[rn]
            ne = ~((a == b) | (np.isnan(a.astype(float)) & np.isnan(b.astype(float))))
            if ne.any(): bad[name] = int(ne.sum())
        pp = np.concatenate([ref["p1"], ref["p2"]], axis=1)
        if not np.array_equal(out["p1p2"][idx], pp): bad["p1p2"] = int((out["p1p2"][idx] != pp).any(axis=1).sum())
        lt = np.stack([ref["last_tri_a"], ref["last_tri_b"]], axis=1)
        if not np.array_equal(out["last_tri"][idx], lt): bad["last_tri"] = int((out["last_tri"][idx] != lt).any(axis=1).sum())
        hit = ref["collisionfree"] == 0
        if not np.array_equal(out["pose_toc"][idx][hit], ref["pose_toc"][hit]): bad["pose_toc"] = 1
        print(f"  check vs port on {len(idx)} queries (heaviest {k // 2} + random): {'BIT-EXACT' if not bad else 'MISMATCH ' + str(bad)}")
L.c2a_b200_phase_stats(0, None); L.c2a_b200_wide_stats(0, None)
