"""BASELINE config 4: a scene of moving instances (50/50 bunny / torus knot 512x32) -> swept-sphere broadphase ->
heterogeneous CCD batch through c2a_b200_solve_pairs.  Prints one JSON line: candidate pairs/s and the frame time,
with a sample checked against the oracle port.   python scripts/config4_scene.py [instances] [frames]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from c2a_b200 import api, meshes, workloads

n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mesh = np.load(os.path.join(ROOT, "tests/golden/bunny_mesh.npz"))
bunny = mesh["verts"][mesh["vidx"]].reshape(-1, 9)
bvhs = [api.build_bvh(bunny), api.build_bvh(meshes.torus_knot(512, 32)[0])]
models = [api.Model(b, 0) for b in bvhs]
radii = np.array([np.linalg.norm(b["tris"].reshape(-1, 3), axis=1).max() for b in bvhs])
fields = ("status", "collisionfree", "num_ca", "num_bv_tests", "num_tri_tests", "toc", "distance")
rows = []
for f in range(frames + 1):  # frame 0 warms up
    sc = workloads.scene(n_inst, 100 + f, radii)
    t0 = time.perf_counter()
    pairs = api.broadphase(sc["begin"][:, 9:], sc["end"][:, 9:], radii[sc["model"]])
    t1 = time.perf_counter()
    poses, ma, mb = workloads.scene_queries(sc, pairs)
    t2 = time.perf_counter()
    out = api.solve_pairs(models, ma, mb, poses, fields=fields)
    t3 = time.perf_counter()
    assert (out["status"] == 0).all()
    if f:
        rows.append((len(pairs), t1 - t0, t2 - t1, t3 - t2, int((out["collisionfree"] == 0).sum()), float(out["num_bv_tests"].mean())))
# parity sample: the last frame's first 256 candidates per group against the oracle port
import oracle
checked = 0
for a in range(2):
    for b in range(2):
        g = np.nonzero((ma == a) & (mb == b))[0][:256]
        if len(g):
            ref = oracle.port().solve_batch(bvhs[a], bvhs[b], poses[g], threads=os.cpu_count() or 1)
            assert np.array_equal(out["toc"][g], ref["toc"]) and np.array_equal(out["distance"][g], ref["distance"])
            assert np.array_equal(out["collisionfree"][g], ref["collisionfree"]) and np.array_equal(out["num_bv_tests"][g], ref["num_bv_tests"])
            checked += len(g)
r = np.array(rows)
frame = r[:, 1:4].sum(1)
print(json.dumps({"workload": f"config 4: {n_inst} instances (bunny 69664 tris / torus knot 32768 tris, alternating), swept-sphere broadphase, "
                  "heterogeneous batch via c2a_b200_solve_pairs", "frames": frames, "candidate_pairs_per_frame": float(r[:, 0].mean()),
                  "neighbours_per_instance": float(2 * r[:, 0].mean() / n_inst), "colliding_pairs_per_frame": float(r[:, 4].mean()),
                  "mean_bv_tests_per_pair": float(r[:, 5].mean()),
                  "frame_time_s": float(frame.mean()), "broadphase_s": float(r[:, 1].mean()), "assembly_s": float(r[:, 2].mean()),
                  "ccd_s": float(r[:, 3].mean()), "pairs_per_sec": float(r[:, 0].sum() / frame.sum()),
                  "verified_against_oracle_port": {"n": checked, "bit_exact": True}}))
