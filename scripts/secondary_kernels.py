"""Time the one-thread-per-query kernels on batches where they are all there is: translation-only CCD queries,
the contact pass, the discrete distance query (development aid; kernel times from CUDA events inside the library)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from c2a_b200 import api, meshes, workloads
import oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
tris = meshes.torus_knot(512, 32)[0]
bvh = api.build_bvh(tris); m = api.Model(bvh, 0)
kt = (C.c_double * 3)()
f = ("status", "collisionfree", "num_ca", "num_bv_tests", "num_tri_tests", "toc", "distance")
tp = workloads.translation_batch(n, 31, radius=workloads.KNOT_RADIUS, move_b=True)
api.solve_batch(m, m, tp[:256], fields=f)
t = time.perf_counter(); out = api.solve_batch(m, m, tp, fields=f); dt = time.perf_counter() - t
api.lib().c2a_b200_kernel_times(kt)
cores = os.cpu_count() or 1
ns = min(n, 256 * cores)
t = time.perf_counter(); ref = oracle.port().solve_batch(bvh, bvh, tp[:ns], threads=1); dc = time.perf_counter() - t  # (the reference runs this branch single-threaded: global flag)
ok = np.array_equal(out["toc"][:ns], ref["toc"]) and np.array_equal(out["num_bv_tests"][:ns], ref["num_bv_tests"])
print(f"translation-only: {n} queries, call {dt * 1e3:.1f} ms, c2a_translation_kernel {kt[2]:.1f} ms = {n / (kt[2] * 1e-3):.0f} queries/s, "
      f"mean BV tests {out['num_bv_tests'].mean():.0f}; port on one core {ns / dc:.0f} queries/s; bit-exact {ok}")
sp = workloads.static_pose_batch(n, 32, radius=workloads.KNOT_RADIUS)
api.distance_batch(m, m, sp[:256])
t = time.perf_counter(); d = api.distance_batch(m, m, sp); dt = time.perf_counter() - t
t = time.perf_counter(); rd = oracle.port().distance(bvh, bvh, sp[:512]); dc = time.perf_counter() - t
print(f"C2A_Distance: {n} queries, call {dt * 1e3:.1f} ms = {n / dt:.0f} queries/s, mean BV tests {d['num_bv_tests'].mean():.0f}; port on one core {512 / dc:.0f} queries/s; "
      f"bit-exact {np.array_equal(d['distance'][:512], rd['distance'])}")
thr = np.full(n, 2.0)
api.contacts_batch(m, m, sp[:256], thr[:256], max_contacts=16)
t = time.perf_counter(); num, recs = api.contacts_batch(m, m, sp, thr, max_contacts=16); dt = time.perf_counter() - t
print(f"contact pass: {n} queries at threshold 2.0, call {dt * 1e3:.1f} ms = {n / dt:.0f} queries/s, mean contacts {num.mean():.2f}")
t = time.perf_counter(); c = api.collide_batch(m, m, sp, max_pairs=64); dt = time.perf_counter() - t
t = time.perf_counter(); rn, rp, rbv, rtr = oracle.port().collide(bvh, bvh, sp[:512], max_pairs=64); dc = time.perf_counter() - t
print(f"C2A_Collide (all contacts): {n} queries, call {dt * 1e3:.1f} ms = {n / dt:.0f} queries/s, colliding {(c['num_pairs'] > 0).mean():.2f}, mean BV tests {c['num_bv_tests'].mean():.0f}, "
      f"mean pairs {c['num_pairs'].mean():.0f}; port on one core {512 / dc:.0f} queries/s; bit-exact {np.array_equal(c['num_pairs'][:512], rn) and np.array_equal(c['num_bv_tests'][:512], rbv)}")
t = time.perf_counter(); c1 = api.collide_batch(m, m, sp, flag=api.FIRST_CONTACT, max_pairs=1); dt = time.perf_counter() - t
print(f"C2A_Collide (first contact): call {dt * 1e3:.1f} ms = {n / dt:.0f} queries/s, mean BV tests {c1['num_bv_tests'].mean():.0f}")
t = time.perf_counter(); cd = api.collide_distance_batch(m, m, sp); dt = time.perf_counter() - t
print(f"C2A_Collide (distance overload): call {dt * 1e3:.1f} ms = {n / dt:.0f} queries/s, mean BV tests {cd['num_bv_tests'].mean():.0f}")
