"""Why is c2a_b200_broadphase slow after a launch that used the early wide kernel?  (development aid)"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from c2a_b200 import api, meshes, workloads
rt = C.CDLL("libcudart.so.12")
def dsync():
    t = time.perf_counter(); rt.cudaDeviceSynchronize(); return (time.perf_counter() - t) * 1e3
mesh = np.load(os.path.join(ROOT, "tests/golden/bunny_mesh.npz"))
bunny = api.build_bvh(mesh["verts"][mesh["vidx"]].reshape(-1, 9).copy()); knot = api.build_bvh(meshes.torus_knot(512, 32)[0])
mb, mk = api.Model(bunny, 0), api.Model(knot, 0)
radii = np.array([np.linalg.norm(b["tris"].reshape(-1, 3), axis=1).max() for b in (bunny, knot)])
sc = workloads.scene(4096, 101, radii)
def bp():
    t = time.perf_counter(); p = api.broadphase(sc["begin"][:, 9:], sc["end"][:, 9:], radii[sc["model"]]); return (time.perf_counter() - t) * 1e3, p
F = ("status", "collisionfree", "num_ca", "num_bv_tests", "toc")
print("broadphase cold %.1f ms" % bp()[0], "then %.1f ms" % bp()[0])
poses = workloads.approach_batch(10000, 20260001)
for rep in range(2):
    t = time.perf_counter(); api.solve_batch(mb, mb, poses, fields=F); print("solve_batch 10000: %.1f ms" % ((time.perf_counter() - t) * 1e3), "device sync after it: %.2f ms" % dsync())
    print("  broadphase %.1f ms" % bp()[0], "again %.1f ms" % bp()[0])
ms, pairs = bp()
q, ma, mbb = workloads.scene_queries(sc, pairs)
for rep in range(3):
    t = time.perf_counter(); api.solve_pairs([mb, mk], ma, mbb, q, fields=F); print("solve_pairs %d: %.1f ms" % (len(q), (time.perf_counter() - t) * 1e3), "device sync after it: %.2f ms" % dsync())
    print("  broadphase %.1f ms" % bp()[0], "again %.1f ms" % bp()[0])
