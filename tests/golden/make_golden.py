"""Generates the committed fixtures under tests/golden/ by running the REFERENCE's own object code
(oracle/_ref/libc2a_ref.so = /root/reference/C2A/src/*.cpp compiled verbatim against oracle/pqp_shim).

Run in the container that has /root/reference:   python tests/golden/make_golden.py

Outputs (all consumed by tests/ and bench.py; nothing under /root/reference is read at test time):
  bunny_mesh.npz        verts float64 [34834,3], vidx int32 [69664,3] parsed from tri_models/bunny_noholes.tri
  demo_poses.npy        the 303 queries of the CCDDemo (config 1), from models/torusknot{1,2}.ani
  ref_translation_*.npz the same for pure translations (the reference's translation-only branch)
  ref_<case>.npz        per-query results of the reference (collisionfree, toc, distance, numCA, counters,
                        p1p2, pose_toc) + the poses / tolerances that produced them
  bvh_digest.json       sha256 of every flattened BVH array built by the reference's builder
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from c2a_b200 import meshes, workloads  # noqa: E402

REF = "/root/reference"
THREADS = os.cpu_count() or 1


def save_results(name, res, poses, tol_d, tol_t, extra=None):
    d = {k: res[k] for k in res.dtype.names if k not in ("last_tri_a", "last_tri_b")}
    d["p1p2"] = np.concatenate([res["p1"], res["p2"]], 1)
    del d["p1"], d["p2"]
    d["poses"] = poses
    d["tol_d"] = np.float64(tol_d)
    d["tol_t"] = np.float64(tol_t)
    if extra:
        d.update(extra)
    np.savez_compressed(os.path.join(HERE, name), **d)
    hits = int((res["collisionfree"] == 0).sum())
    print(f"{name}: n={len(res)} hits={hits} mean numCA={res['numCA'].mean():.2f} "
          f"mean nbv={res['num_bv_tests'].mean():.0f} mean ntri={res['num_tri_tests'].mean():.0f}")


def digest(bvh):
    return {k: hashlib.sha256(np.ascontiguousarray(bvh[k]).tobytes()).hexdigest() for k in sorted(bvh)}


def carry_fixture(R, name, mA, mB, poses):
    """Sequential single-thread calls; each call's seeds are the models' last_tri as the previous call left them."""
    sa = sb = 0
    rows, seeds = [], []
    for i in range(len(poses)):
        r = R.solve_batch(mA, mB, poses[i:i + 1], seedA=[sa], seedB=[sb], threads=1)
        rows.append(r[0]); seeds.append((sa, sb))
        if r["last_tri_a"][0] >= 0:
            sa = int(r["last_tri_a"][0])
        if r["last_tri_b"][0] >= 0:
            sb = int(r["last_tri_b"][0])
    rows = np.array(rows, dtype=oracle.RESULT_DTYPE)
    save_results(name, rows, poses, 1e-4, 1e-4,
                 {"seed_a": np.array([s[0] for s in seeds], np.int32), "seed_b": np.array([s[1] for s in seeds], np.int32),
                  "last_tri": np.stack([rows["last_tri_a"], rows["last_tri_b"]], 1)})


def translation_fixtures(R, bunny):
    """Translation-only branch (both angular speeds < 1e-8, C2A.cpp:2391-2395, :1362-1521): pure translations,
    object 2 static or drifting, run one at a time (the reference keeps the branch flag in a global)."""
    tris, vi = meshes.torus_knot(128, 16)
    knot = R.model(tris, vi)
    poses = np.concatenate([workloads.translation_batch(300, 20260011, radius=workloads.KNOT_RADIUS),
                            workloads.translation_batch(300, 20260012, radius=workloads.KNOT_RADIUS, move_b=True)])
    res = R.solve_batch(knot, knot, poses, threads=1)
    assert (res["numCA"] == 0).all()
    # the full, unmodified C2A_Solve agrees with the TOC-only wrapper and gives the contact counts
    full, ncont = R.solve_batch(knot, knot, poses[:120], mode=1)
    for k in ("collisionfree", "numCA", "num_bv_tests", "num_tri_tests", "toc", "distance", "pose_toc"):
        assert np.array_equal(full[k], res[k][:120]), k
    save_results("ref_translation_knot_128x16.npz", res, poses, 1e-4, 1e-4,
                 {"last_tri": np.stack([res["last_tri_a"], res["last_tri_b"]], 1), "num_contact": ncont})
    # heterogeneous pair, explicit seeds, non-default tolerances (m_toc_delta = tolerance_d)
    tris, vi = meshes.torus_knot(512, 32)
    knot512 = R.model(tris, vi)
    rng = np.random.default_rng(8)
    n = 160
    sa = rng.integers(0, bunny.n_tris, n).astype(np.int32)
    sb = rng.integers(0, knot512.n_tris, n).astype(np.int32)
    poses = workloads.translation_batch(n, 20260013, radius=workloads.KNOT_RADIUS, move_b=True)
    res = R.solve_batch(bunny, knot512, poses, seedA=sa, seedB=sb, tol_d=1e-3, tol_t=1e-5, threads=1)
    save_results("ref_translation_bunny_vs_knot_seeded.npz", res, poses, 1e-3, 1e-5,
                 {"seed_a": sa, "seed_b": sb, "last_tri": np.stack([res["last_tri_a"], res["last_tri_b"]], 1)})


def bunny_grazing_fixture(R, bunny):
    """Config 5 on the model BASELINE.json names for it: colliding bunny pairs re-posed so that the motion only just
    reaches contact at its end (end pose = TOC pose pushed 0..1e-3 past contact), tolerance_t 1e-3 and 1e-6."""
    g = np.load(os.path.join(HERE, "ref_bunny_approach.npz"))
    hit = np.where((g["collisionfree"] == 0) & (g["num_tri_tests"] > 0))[0][:120]
    rng = np.random.default_rng(56)
    gp = workloads.grazing_batch(g["poses"][hit], g["toc"][hit], g["pose_toc"][hit], g["p1p2"][hit], rng.uniform(0, 1e-3, len(hit)))
    for tol_t in (1e-3, 1e-6):
        res = R.solve_batch(bunny, bunny, gp, tol_d=1e-4, tol_t=tol_t, threads=THREADS)
        save_results(f"ref_bunny_grazing_tol{tol_t:g}.npz", res, gp, 1e-4, tol_t)


def distance_fixture(R):
    """The reference's C2A_Distance (C2A_PQP.cpp:970-1056, depth-first routine) on static pose pairs, random seeds."""
    tris, vi = meshes.torus_knot(128, 16)
    a, b = R.model(tris, vi), R.model(tris, vi)  # two objects: the query reads and writes each model's last_tri
    n = 400
    poses24 = workloads.static_pose_batch(n, 20260021, radius=workloads.KNOT_RADIUS)
    rng = np.random.default_rng(9)
    sa = rng.integers(0, a.n_tris, n).astype(np.int32); sb = rng.integers(0, b.n_tris, n).astype(np.int32)
    out = {"poses24": poses24, "seed_a": sa, "seed_b": sb}
    for tag, rel, ab in (("exact", 0.0, 0.0), ("approx", 0.25, 2.0)):
        r = R.distance(a, b, poses24, sa, sb, rel, ab)
        for k in r.dtype.names:
            out[f"{tag}_{k}"] = r[k]
        out[f"{tag}_err"] = np.array([rel, ab])
        print(f"ref_distance_knot_128x16 {tag}: touching {int((r['distance'] == 0).sum())}/{n} mean nbv {r['num_bv_tests'].mean():.0f}")
    np.savez_compressed(os.path.join(HERE, "ref_distance_knot_128x16.npz"), **out)


def distance_queue_fixture(R):
    """The reference's C2A_Distance with qsize > 2 (C2ADistanceQueueRecurse, C2A_PQP.cpp:624-787; the queue is the PQP
    stand-in of oracle/pqp_shim) on the poses and seeds of ref_distance_knot_128x16."""
    g = np.load(os.path.join(HERE, "ref_distance_knot_128x16.npz"))
    tris, vi = meshes.torus_knot(128, 16)
    a, b = R.model(tris, vi), R.model(tris, vi)
    out = {}
    for qs in (3, 10):
        for tag, rel, ab in (("exact", 0.0, 0.0), ("approx", 0.25, 2.0)):
            r = R.distance(a, b, g["poses24"], g["seed_a"], g["seed_b"], rel, ab, qsize=qs)
            for k in r.dtype.names:
                out[f"q{qs}_{tag}_{k}"] = r[k]
            print(f"ref_distance_queue_knot_128x16 qsize {qs} {tag}: mean nbv {r['num_bv_tests'].mean():.0f} (depth-first {g[tag + '_num_bv_tests'].mean():.0f}), "
                  f"other pair than depth-first {int((r['tri_a'] != g[tag + '_tri_a']).sum())}/{len(r)}")
    np.savez_compressed(os.path.join(HERE, "ref_distance_queue_knot_128x16.npz"), **out)


def collide_fixture(R):
    """The reference's C2A_Collide, both overloads (C2A_PQP.cpp:798-968, 1060-1280), on static pose pairs: a knot against
    itself and the bunny against a larger knot.  Pairs are the reference's Tri::id values in its reporting order."""
    m = np.load(os.path.join(HERE, "bunny_mesh.npz"))
    bunny_tris = m["verts"][m["vidx"]].reshape(-1, 9).copy()
    cases = (("knot_128x16", meshes.torus_knot(128, 16)[0], meshes.torus_knot(128, 16)[0], 400, workloads.KNOT_RADIUS),
             ("bunny_vs_knot_512x32", bunny_tris, meshes.torus_knot(512, 32)[0], 150, workloads.BUNNY_RADIUS))
    for tag, ta, tb, n, radius in cases:
        a, b = R.model(ta), R.model(tb)  # two objects: the distance overload reads and writes each model's last_tri
        poses24 = workloads.static_pose_batch(n, 20260031, radius=radius)
        out = {"poses24": poses24}
        for name, flag in (("all", 1), ("first", 2)):
            num, pairs, nbv, ntri = R.collide(a, b, poses24, flag=flag, max_pairs=1 << 16)
            assert num.max() < (1 << 16)
            out[f"{name}_num_pairs"] = num; out[f"{name}_num_bv_tests"] = nbv; out[f"{name}_num_tri_tests"] = ntri
            out[f"{name}_pairs"] = np.concatenate(pairs).astype(np.int32).reshape(-1, 2)
            print(f"ref_collide_{tag} {name}: colliding {int((num > 0).sum())}/{n}, pairs {int(num.sum())} (max {int(num.max())}), mean nbv {nbv.mean():.0f}")
        rng = np.random.default_rng(11)
        sa = rng.integers(0, a.n_tris, n).astype(np.int32); sb = rng.integers(0, b.n_tris, n).astype(np.int32)
        out["seed_a"] = sa; out["seed_b"] = sb
        for name, rel, ab in (("exact", 0.0, 0.0), ("approx", 0.25, 2.0)):
            r = R.collide_distance(a, b, poses24, sa, sb, rel, ab)
            for k in r.dtype.names:
                out[f"dist_{name}_{k}"] = r[k]
            out[f"dist_{name}_err"] = np.array([rel, ab])
            print(f"ref_collide_{tag} distance overload {name}: reached a leaf {int((r['num_tri_tests'] > 0).sum())}/{n}, touching {int((r['distance'] == 0).sum())}")
        np.savez_compressed(os.path.join(HERE, f"ref_collide_{tag}.npz"), **out)


def digest_only(R):
    """bvh_digest.json alone (sha256 of every array the reference's builder made, OBB fields included)."""
    m = np.load(os.path.join(HERE, "bunny_mesh.npz"))
    digests = {"bunny": digest(R.model(m["verts"][m["vidx"]].reshape(-1, 9).copy(), m["vidx"]).export())}
    for nu, nvv in ((128, 16), (512, 32), (1024, 32)):
        tris, vi = meshes.torus_knot(nu, nvv)
        digests[f"knot_{nu}x{nvv}"] = digest(R.model(tris, vi).export())
    with open(os.path.join(HERE, "bvh_digest.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)


def save_compact(name, res, gen, poses, tol_d, tol_t, extra=None):
    """Large fixtures keep the reference's outputs and a recipe for the poses (generator + seed, or a source fixture
    plus the doubles that differ) instead of the poses themselves; tests/conftest.py rebuilds them and checks the
    sha256.  pose_toc is dropped (a pure function of toc, covered by the small fixtures)."""
    d = {"collisionfree": res["collisionfree"].astype(np.int8), "numCA": res["numCA"].astype(np.int16),
         "num_bv_tests": res["num_bv_tests"], "num_tri_tests": res["num_tri_tests"], "toc": res["toc"],
         "distance": res["distance"], "mint": res["mint"], "p1p2": np.concatenate([res["p1"], res["p2"]], 1),
         "gen": np.array(json.dumps(gen)), "pose_sha256": np.array(hashlib.sha256(np.ascontiguousarray(poses).tobytes()).hexdigest()),
         "tol_d": np.float64(tol_d), "tol_t": np.float64(tol_t)}
    if extra:
        d.update(extra)
    np.savez_compressed(os.path.join(HERE, name), **d)
    print(f"{name}: n={len(res)} hits={int((res['collisionfree'] == 0).sum())} mean numCA={res['numCA'].mean():.2f} max numCA={res['numCA'].max()} "
          f"mean nbv={res['num_bv_tests'].mean():.0f} max nbv={res['num_bv_tests'].max()}")


def big_fixtures(R, bunny):
    """SURVEY section 8d at full size: config 2 (10 000 bunny approach queries), config 5 (2 000 grazing re-poses of
    its colliding queries x tolerance_t 1e-3..1e-6), and the heaviest queries of the 1M-query bench batch (config 3):
    every query that runs into the reference's 150-iteration cap (C2A.cpp:2079) plus the 200 with the most BV tests.
    The heavy indices were picked from a GPU run's per-query counters (tests/golden/heavy_indices.npy); what is pinned
    here is the reference's own output for those poses."""
    n = 10000
    poses = workloads.approach_batch(n, 20260001)
    res = R.solve_batch(bunny, bunny, poses, threads=THREADS)
    save_compact("ref_bunny_approach_10k.npz", res, {"fn": "approach_batch", "n": n, "seed": 20260001, "radius": "bunny"}, poses, 1e-4, 1e-4)
    hit = np.where((res["collisionfree"] == 0) & (res["num_tri_tests"] > 0))[0][:2000]
    rng = np.random.default_rng(57)
    p1p2 = np.concatenate([res["p1"], res["p2"]], 1)
    gp = workloads.grazing_batch(poses[hit], res["toc"][hit], res["pose_toc"][hit], p1p2[hit], rng.uniform(0, 1e-3, len(hit)))
    np.savez_compressed(os.path.join(HERE, "bunny_grazing_2k_poses.npz"), src=hit.astype(np.int32), end1=gp[:, 12:24].copy())
    for tol_t in (1e-3, 1e-4, 1e-5, 1e-6):
        r = R.solve_batch(bunny, bunny, gp, tol_d=1e-4, tol_t=tol_t, threads=THREADS)
        save_compact(f"ref_bunny_grazing_2k_tol{tol_t:g}.npz", r, {"fn": "grazing_of", "src": "ref_bunny_approach_10k", "poses": "bunny_grazing_2k_poses"},
                     gp, 1e-4, tol_t)
    idx = np.load(os.path.join(HERE, "heavy_indices.npy"))
    tris, vi = meshes.torus_knot(512, 32)
    knot = R.model(tris, vi)
    batch = workloads.approach_batch(1000000, 20260002, radius=workloads.KNOT_RADIUS)
    hp = np.ascontiguousarray(batch[idx])
    r = R.solve_batch(knot, knot, hp, threads=THREADS)
    save_results("ref_knot_512x32_heavy.npz", r, hp, 1e-4, 1e-4, {"batch_index": idx.astype(np.int32)})


def main():
    oracle.build_oracle()
    R = oracle.ref()
    if "--only-big" in sys.argv:
        m = np.load(os.path.join(HERE, "bunny_mesh.npz"))
        big_fixtures(R, R.model(m["verts"][m["vidx"]].reshape(-1, 9).copy(), m["vidx"]))
        return
    if "--only-bunny-grazing" in sys.argv:
        m = np.load(os.path.join(HERE, "bunny_mesh.npz"))
        bunny_grazing_fixture(R, R.model(m["verts"][m["vidx"]].reshape(-1, 9).copy(), m["vidx"]))
        return
    if "--only-distance" in sys.argv:
        distance_fixture(R)
        return
    if "--only-distance-queue" in sys.argv:
        distance_queue_fixture(R)
        return
    if "--only-collide" in sys.argv:
        collide_fixture(R)
        return
    if "--only-digest" in sys.argv:
        digest_only(R)
        return
    if "--only-translation" in sys.argv:
        m = np.load(os.path.join(HERE, "bunny_mesh.npz"))
        translation_fixtures(R, R.model(m["verts"][m["vidx"]].reshape(-1, 9).copy(), m["vidx"]))
        return

    # --- meshes
    with open(os.path.join(REF, "tri_models/bunny_noholes.tri")) as f:
        tok = f.read().split()
    nv, nt = int(tok[1]), int(tok[2])
    verts = np.array(tok[3:3 + 3 * nv], dtype=np.float64).reshape(nv, 3)
    vidx = np.array(tok[3 + 3 * nv:3 + 3 * nv + 3 * nt], dtype=np.int32).reshape(nt, 3)
    np.savez_compressed(os.path.join(HERE, "bunny_mesh.npz"), verts=verts, vidx=vidx)
    bunny_tris = verts[vidx].reshape(nt, 9).copy()

    digests = {}
    bunny = R.model(bunny_tris, vidx)
    digests["bunny"] = digest(bunny.export())

    # --- config 1: the demo's 303 queries
    R1, T1 = meshes.load_ani(os.path.join(REF, "models/torusknot1.ani"))
    R2, T2 = meshes.load_ani(os.path.join(REF, "models/torusknot2.ani"))
    demo = workloads.demo_batch(R1, T1, R2, T2)
    np.save(os.path.join(HERE, "demo_poses.npy"), demo)
    res = R.solve_batch(bunny, bunny, demo, threads=THREADS)
    save_results("ref_demo_bunny.npz", res, demo, 1e-4, 1e-4)
    # the full, unmodified C2A_Solve (incl. contact pass) must agree with the TOC-only wrapper
    full, ncont = R.solve_batch(bunny, bunny, demo[:40], mode=1)
    for k in ("collisionfree", "numCA", "num_bv_tests", "num_tri_tests", "toc", "distance", "pose_toc"):
        assert np.array_equal(full[k], res[k][:40]), k
    np.save(os.path.join(HERE, "ref_demo_num_contact.npy"), ncont)

    # --- config 2 (sample): bunny vs bunny random approach
    poses = workloads.approach_batch(2000, 20260001)
    res = R.solve_batch(bunny, bunny, poses, threads=THREADS)
    save_results("ref_bunny_approach.npz", res, poses, 1e-4, 1e-4)

    # --- config 3 (samples): torus knots at three resolutions
    for nu, nvv, n in ((128, 16, 1500), (512, 32, 2000), (1024, 32, 500)):
        tris, vi = meshes.torus_knot(nu, nvv)
        knot = R.model(tris, vi)
        digests[f"knot_{nu}x{nvv}"] = digest(knot.export())
        poses = workloads.approach_batch(n, 20260002, radius=workloads.KNOT_RADIUS)
        res = R.solve_batch(knot, knot, poses, threads=THREADS)
        save_results(f"ref_knot_{nu}x{nvv}.npz", res, poses, 1e-4, 1e-4)
        if nu == 512:
            # heterogeneous pair + explicit seeds + non-default tolerances
            rng = np.random.default_rng(7)
            sa = rng.integers(0, bunny.n_tris, 400).astype(np.int32)
            sb = rng.integers(0, knot.n_tris, 400).astype(np.int32)
            poses = workloads.approach_batch(400, 20260003, radius=workloads.KNOT_RADIUS)
            res = R.solve_batch(bunny, knot, poses, seedA=sa, seedB=sb, tol_d=1e-3, tol_t=1e-5, threads=THREADS)
            save_results("ref_bunny_vs_knot_seeded.npz", res, poses, 1e-3, 1e-5, {"seed_a": sa, "seed_b": sb})

    # --- config 5 (sample): grazing re-poses of colliding knot queries, tolerance_t swept
    tris, vi = meshes.torus_knot(128, 16)
    knot = R.model(tris, vi)
    g = np.load(os.path.join(HERE, "ref_knot_128x16.npz"))
    hit = np.where((g["collisionfree"] == 0) & (g["num_tri_tests"] > 0))[0][:240]
    rng = np.random.default_rng(55)
    gp = workloads.grazing_batch(g["poses"][hit], g["toc"][hit], g["pose_toc"][hit], g["p1p2"][hit], rng.uniform(0, 1e-3, len(hit)))
    for tol_t in (1e-3, 1e-6):
        res = R.solve_batch(knot, knot, gp, tol_d=1e-4, tol_t=tol_t, threads=THREADS)
        save_results(f"ref_knot_128x16_grazing_tol{tol_t:g}.npz", res, gp, 1e-4, tol_t)

    # --- "demo mode": seeds carried from call to call through o->last_tri (CCDDemo/mainTorusknot.cpp:314-315),
    # two separate model objects like the demo's object1_tested / object2_tested
    tris, vi = meshes.torus_knot(128, 16)
    g = np.load(os.path.join(HERE, "ref_knot_128x16.npz"))
    carry_fixture(R, "ref_knot_128x16_carry.npz", R.model(tris, vi), R.model(tris, vi), g["poses"][:64])
    # config 1 in demo mode: the 303 frames played twice (the second pass starts from the first pass's seeds)
    carry_fixture(R, "ref_demo_bunny_carry.npz", R.model(bunny_tris, vidx), R.model(bunny_tris, vidx),
                  np.concatenate([demo, demo]))

    # --- contact pass: the full, unmodified C2A_Solve with its ContactF list exported (list order)
    tris, vi = meshes.torus_knot(128, 16)
    knot = R.model(tris, vi)
    g = np.load(os.path.join(HERE, "ref_knot_128x16.npz"))
    nq = 200
    counts, recs = [], []
    for i in range(nq):
        res, n, rr = R.solve_contacts(knot, knot, g["poses"][i])
        assert res["toc"] == g["toc"][i] and res["collisionfree"] == g["collisionfree"][i]
        counts.append(n)
        recs.append(rr.copy())
    allr = np.concatenate(recs) if recs else np.zeros(0, dtype=oracle.CONTACT_DTYPE)
    np.savez_compressed(os.path.join(HERE, "ref_contacts_knot_128x16.npz"), num_contact=np.array(counts, np.int32),
                        **{k: allr[k] for k in allr.dtype.names})
    print(f"ref_contacts_knot_128x16.npz: {nq} queries, {int(np.sum(counts))} contacts, max {int(np.max(counts))}")

    translation_fixtures(R, bunny)
    distance_fixture(R)
    distance_queue_fixture(R)
    collide_fixture(R)
    bunny_grazing_fixture(R, bunny)

    with open(os.path.join(HERE, "bvh_digest.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
