"""Ad-hoc GPU parity + timing check against the golden fixtures (development helper)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from c2a_b200 import api, meshes

G = os.path.join(ROOT, "tests", "golden")


def compare(name, got, ref, tol_t):
    n = len(ref["toc"])
    ok_status = (got["status"] == 0).all()
    verdict = (got["collisionfree"] == ref["collisionfree"])
    toc_ok = np.abs(got["toc"] - ref["toc"]) <= tol_t
    dist_ok = np.abs(got["distance"] - ref["distance"]) <= 1e-9 * np.maximum(1.0, np.abs(ref["distance"]))
    exact = {k: float(np.mean(got[g] == ref[k])) for g, k in (("toc", "toc"), ("distance", "distance"), ("num_ca", "numCA"),
             ("num_bv_tests", "num_bv_tests"), ("num_tri_tests", "num_tri_tests"), ("mint", "mint"))}
    pose_exact = float(np.mean((got["pose_toc"] == ref["pose_toc"]).all(1)))
    print(f"[{name}] n={n} status_ok={ok_status} verdict={verdict.mean():.6f} toc_ok={toc_ok.mean():.6f} "
          f"dist_ok={dist_ok.mean():.6f} bit-exact: {exact} pose={pose_exact:.4f}")
    bad = np.where(~(verdict & toc_ok & dist_ok))[0]
    for i in bad[:5]:
        print("   bad", i, "cf", got["collisionfree"][i], ref["collisionfree"][i], "toc", got["toc"][i], ref["toc"][i],
              "dist", got["distance"][i], ref["distance"][i], "nca", got["num_ca"][i], ref["numCA"][i])
    return len(bad) == 0


def sincos_check():
    import ctypes as C, math
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-3.2, 3.2, 400000), rng.uniform(-0.2, 0.2, 100000), rng.uniform(-50, 50, 100000)])
    s = np.zeros_like(x); c = np.zeros_like(x)
    api._check(api.lib().c2a_b200_test_sincos(x.ctypes.data_as(C.c_void_p), C.c_int64(len(x)), s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p)))
    print("device sin/cos vs libm mismatches:", int((s != np.sin(x)).sum()), int((c != np.cos(x)).sum()), "of", len(x))


def main():
    sincos_check()
    R = oracle.ref()
    mesh = np.load(os.path.join(G, "bunny_mesh.npz"))
    btris = mesh["verts"][mesh["vidx"]].reshape(-1, 9).copy()
    t = time.time(); bunny = R.model(btris, mesh["vidx"]).export(); print("ref build bunny", time.time() - t)
    mb = api.Model(bunny, 0)
    print(mb.info())
    allok = True
    for case in ("ref_demo_bunny", "ref_bunny_approach"):
        ref = np.load(os.path.join(G, case + ".npz"))
        got = api.solve_batch(mb, mb, ref["poses"], tol_d=float(ref["tol_d"]), tol_t=float(ref["tol_t"]))
        t = time.time(); got = api.solve_batch(mb, mb, ref["poses"], tol_d=float(ref["tol_d"]), tol_t=float(ref["tol_t"])); dt = time.time() - t
        print(case, "e2e time", dt, "q/s", len(ref["toc"]) / dt)
        allok &= compare(case, got, ref, float(ref["tol_t"]))
    for nu, nv in ((128, 16), (512, 32), (1024, 32)):
        tris, vi = meshes.torus_knot(nu, nv)
        knot = R.model(tris, vi).export()
        mk = api.Model(knot, 0)
        ref = np.load(os.path.join(G, f"ref_knot_{nu}x{nv}.npz"))
        t = time.time(); got = api.solve_batch(mk, mk, ref["poses"]); dt = time.time() - t
        print("knot", nu, nv, "e2e time", dt, "q/s", len(ref["toc"]) / dt)
        allok &= compare(f"knot_{nu}x{nv}", got, ref, 1e-4)
        if nu == 512:
            ref = np.load(os.path.join(G, "ref_bunny_vs_knot_seeded.npz"))
            got = api.solve_batch(mb, mk, ref["poses"], ref["seed_a"], ref["seed_b"], float(ref["tol_d"]), float(ref["tol_t"]))
            allok &= compare("bunny_vs_knot_seeded", got, ref, float(ref["tol_t"]))
            # bigger batch for timing
            from c2a_b200 import workloads
            poses = workloads.approach_batch(50000, 99, radius=workloads.KNOT_RADIUS)
            for rep in range(2):
                t = time.time(); got = api.solve_batch(mk, mk, poses, fields=("status", "collisionfree", "toc", "distance", "num_ca", "num_bv_tests", "num_tri_tests")); dt = time.time() - t
                print("knot 512x32 50k e2e", dt, "q/s", len(poses) / dt, "nbv/q", got["num_bv_tests"].mean(), "hits", (got["collisionfree"] == 0).mean())
    print("ALL OK" if allok else "MISMATCHES")


if __name__ == "__main__":
    main()
