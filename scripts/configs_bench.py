"""BASELINE.json configs 1, 2, 4, 5 (config 3 is bench.py): this repo's CUDA path through the host-buffer C ABI next to
the reference's own CPU code (oracle/_ref when it was built, else the port) on the same inputs, timed in the same run.
One JSON line per config on stdout; `python scripts/configs_bench.py [1 2 4 5]`.

  config 1  the CCDDemo's 303 frames (bunny vs bunny), ONE C2A_Solve per call in demo mode (seeds carried through
            last_tri): ms per call, and BV tests/s of a lone query on the GPU beside one CPU core
  config 2  bunny vs bunny, 10 000 random approach queries (one batch)
  config 4  4 096 instances (bunny / torus knot alternating): swept-sphere broadphase -> heterogeneous batch
  config 5  2 000 grazing re-poses of colliding bunny pairs, tolerance_t 1e-3 .. 1e-6
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from c2a_b200 import api, meshes, workloads
from conftest import load_golden

which = [int(x) for x in sys.argv[1:]] or [1, 2, 4, 5]
cores = os.cpu_count() or 1
mesh = np.load(os.path.join(ROOT, "tests/golden/bunny_mesh.npz"))
bunny_tris = mesh["verts"][mesh["vidx"]].reshape(-1, 9).copy()
knot_tris = meshes.torus_knot(512, 32)[0]
bvh = {"bunny": api.build_bvh(bunny_tris), "knot": api.build_bvh(knot_tris)}
model = {k: api.Model(v, 0) for k, v in bvh.items()}
have_ref = oracle.have_ref()
if have_ref:
    R = oracle.ref()
    rmodel = {"bunny": R.model(bunny_tris), "knot": R.model(knot_tris), "bunny2": R.model(bunny_tris)}  # (demo mode: two objects, like the demo)
else:
    oracle.build_oracle()
P = oracle.port()
KIND = "reference" if have_ref else "port"
F = ("status", "collisionfree", "num_ca", "num_bv_tests", "num_tri_tests", "toc", "distance", "last_tri")


def cpu(a, b, poses, seed_a=None, seed_b=None, tol_d=1e-4, tol_t=1e-4, threads=cores):
    if have_ref:
        return R.solve_batch(rmodel[a], rmodel[b], poses, seed_a, seed_b, tol_d=tol_d, tol_t=tol_t, threads=threads)
    return P.solve_batch(bvh[a], bvh[b], poses, seed_a, seed_b, tol_d, tol_t, threads=threads)


def same(got, ref):
    return bool(np.array_equal(got["collisionfree"], ref["collisionfree"]) and np.array_equal(got["toc"], ref["toc"]) and
                np.array_equal(got["distance"], ref["distance"]) and np.array_equal(got["num_ca"], ref["numCA"]) and
                np.array_equal(got["num_bv_tests"], ref["num_bv_tests"]))


api.solve_batch(model["bunny"], model["bunny"], workloads.approach_batch(256, 1), fields=F)  # warm: context, pools

if 1 in which:
    demo = np.load(os.path.join(ROOT, "tests/golden/demo_poses.npy"))
    n = len(demo)
    # GPU: one call per frame, seeds carried like the demo does (CCDDemo/mainTorusknot.cpp:314-315)
    sa = sb = 0
    rows = []
    t0 = time.perf_counter()
    for i in range(n):
        o = api.solve_batch(model["bunny"], model["bunny"], demo[i:i + 1], [sa], [sb], fields=F)
        rows.append((o["collisionfree"][0], o["toc"][0], o["distance"][0], o["num_ca"][0], o["num_bv_tests"][0]))
        if o["last_tri"][0, 0] >= 0: sa = int(o["last_tri"][0, 0])
        if o["last_tri"][0, 1] >= 0: sb = int(o["last_tri"][0, 1])
    gpu_s = time.perf_counter() - t0
    # CPU: the same sequence, one thread (the reference's demo is single-threaded)
    sa = sb = 0
    refrows = []
    t0 = time.perf_counter()
    for i in range(n):
        r = cpu("bunny", "bunny2", demo[i:i + 1], [sa], [sb], threads=1) if have_ref else P.solve_batch(bvh["bunny"], bvh["bunny"], demo[i:i + 1], [sa], [sb], threads=1)
        refrows.append((r["collisionfree"][0], r["toc"][0], r["distance"][0], r["numCA"][0], r["num_bv_tests"][0]))
        if r["last_tri_a"][0] >= 0: sa = int(r["last_tri_a"][0])
        if r["last_tri_b"][0] >= 0: sb = int(r["last_tri_b"][0])
    cpu_s = time.perf_counter() - t0
    nbv = float(sum(r[4] for r in rows))
    # the same frames as ONE batch (what a caller with many pairs per frame would do)
    t0 = time.perf_counter()
    ob = api.solve_batch(model["bunny"], model["bunny"], demo, fields=F)
    gpu_batch_s = time.perf_counter() - t0
    print(json.dumps({"config": 1, "workload": f"CCDDemo: {n} frames bunny vs bunny, one C2A_Solve per call, seeds carried (demo mode)",
                      "gpu_ms_per_call": 1e3 * gpu_s / n, "cpu_ms_per_call": 1e3 * cpu_s / n, "cpu_kind": KIND, "cpu_threads": 1,
                      "gpu_over_cpu": cpu_s / gpu_s, "mean_bv_tests_per_call": nbv / n,
                      "lone_query_bv_tests_per_sec": {"gpu": nbv / gpu_s, "one_cpu_core": nbv / cpu_s},
                      "gpu_ms_all_frames_as_one_batch": 1e3 * gpu_batch_s, "bit_exact": rows == refrows}), flush=True)

if 2 in which:
    poses = workloads.approach_batch(10000, 20260001)
    api.solve_batch(model["bunny"], model["bunny"], poses[:2000], fields=F)
    t0 = time.perf_counter(); got = api.solve_batch(model["bunny"], model["bunny"], poses, fields=F); gpu_s = time.perf_counter() - t0
    t0 = time.perf_counter(); ref = cpu("bunny", "bunny", poses); cpu_s = time.perf_counter() - t0
    t0 = time.perf_counter(); cpu("bunny", "bunny", poses[:400], threads=1); cpu1_s = (time.perf_counter() - t0) * len(poses) / 400
    print(json.dumps({"config": 2, "workload": "bunny vs bunny, 10000 random approach queries, one host-buffer call",
                      "gpu_s": gpu_s, "gpu_queries_per_sec": len(poses) / gpu_s, "cpu_s": cpu_s, "cpu_queries_per_sec": len(poses) / cpu_s,
                      "cpu_kind": KIND, "cpu_threads": cores, "cpu_single_thread_queries_per_sec": len(poses) / cpu1_s,
                      "gpu_over_cpu": cpu_s / gpu_s, "mean_bv_tests": float(got["num_bv_tests"].mean()),
                      "bvtt_pairs_per_sec_gpu": float(got["num_bv_tests"].sum()) / gpu_s, "bit_exact": same(got, ref)}), flush=True)

if 4 in which:
    names = ["bunny", "knot"]
    models = [model["bunny"], model["knot"]]
    radii = np.array([np.linalg.norm(bvh[k]["tris"].reshape(-1, 3), axis=1).max() for k in names])
    rows = []
    for f in range(3):  # frame 0 warms up
        sc = workloads.scene(4096, 100 + f, radii)
        t0 = time.perf_counter()
        pairs = api.broadphase(sc["begin"][:, 9:], sc["end"][:, 9:], radii[sc["model"]])
        t1 = time.perf_counter()
        poses, ma, mb = workloads.scene_queries(sc, pairs)
        t2 = time.perf_counter()
        out = api.solve_pairs(models, ma, mb, poses, fields=F)
        t3 = time.perf_counter()
        if f:
            rows.append((len(pairs), t1 - t0, t2 - t1, t3 - t2, float(out["num_bv_tests"].mean())))
    # the reference on the last frame's pair list, group by group, all cores
    t0 = time.perf_counter()
    ok = True
    for a in range(2):
        for b in range(2):
            g = np.nonzero((ma == a) & (mb == b))[0]
            if len(g):
                ref = cpu(names[a], names[b], poses[g])
                ok = ok and same({k: out[k][g] for k in out}, ref)
    cpu_s = time.perf_counter() - t0
    r = np.array(rows)
    frame = r[:, 1:4].sum(1).mean()
    print(json.dumps({"config": 4, "workload": "4096 instances (bunny / torus knot 512x32 alternating), swept-sphere broadphase + heterogeneous CCD batch",
                      "candidate_pairs_per_frame": float(r[:, 0].mean()), "frame_s": float(frame), "broadphase_s": float(r[:, 1].mean()),
                      "assembly_s": float(r[:, 2].mean()), "ccd_s": float(r[:, 3].mean()), "gpu_pairs_per_sec": float(r[:, 0].mean() / frame),
                      "cpu_ccd_s": cpu_s, "cpu_pairs_per_sec": float(len(pairs) / cpu_s), "cpu_kind": KIND, "cpu_threads": cores,
                      "gpu_ccd_over_cpu_ccd": cpu_s / float(r[-1, 3]), "mean_bv_tests_per_pair": float(r[:, 4].mean()), "bit_exact": bool(ok)}), flush=True)

if 5 in which:
    res = []
    for tol in ("0.001", "0.0001", "1e-05", "1e-06"):
        g = load_golden(f"ref_bunny_grazing_2k_tol{tol}")
        poses, tol_t = g["poses"], float(g["tol_t"])
        api.solve_batch(model["bunny"], model["bunny"], poses[:256], tol_d=1e-4, tol_t=tol_t, fields=F)
        t0 = time.perf_counter(); got = api.solve_batch(model["bunny"], model["bunny"], poses, tol_d=1e-4, tol_t=tol_t, fields=F); gpu_s = time.perf_counter() - t0
        t0 = time.perf_counter(); ref = cpu("bunny", "bunny", poses, tol_d=1e-4, tol_t=tol_t); cpu_s = time.perf_counter() - t0
        fix = bool(np.array_equal(got["toc"], g["toc"]) and np.array_equal(got["collisionfree"], g["collisionfree"]) and np.array_equal(got["num_bv_tests"], g["num_bv_tests"]))
        res.append({"tolerance_t": tol_t, "gpu_s": gpu_s, "cpu_s": cpu_s, "gpu_over_cpu": cpu_s / gpu_s, "hits": int((got["collisionfree"] == 0).sum()),
                    "mean_num_ca": float(got["num_ca"].mean()), "mean_bv_tests": float(got["num_bv_tests"].mean()), "bit_exact": same(got, ref),
                    "matches_committed_fixture": fix})
    print(json.dumps({"config": 5, "workload": "2000 grazing re-poses of colliding bunny pairs, tolerance_d 1e-4, tolerance_t swept",
                      "cpu_kind": KIND, "cpu_threads": cores, "sweep": res}), flush=True)
