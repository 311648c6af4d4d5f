"""GPU parity tests (pytest -m gpu, on the B200 box).  Everything goes through the C ABI of
include/c2a_b200.h; the oracle port and the committed reference fixtures are the checkers.

Bar: the path is FP64 with data-dependent branching, and every device expression keeps the reference's
rounding sequence, so the comparison is BIT-EXACT (stronger than the contract of BASELINE.json:
identical verdict, |toc - toc_ref| <= tolerance_t, distance within 1e-9 relative)."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle
from c2a_b200 import api, meshes, workloads
from conftest import GOLDEN_CASES
from test_oracle import rect_cases, tri_cases, TRANSLATION_CASES, COLLIDE_CASES, split_pairs

pytestmark = pytest.mark.gpu

FIELDS = (("collisionfree", "collisionfree"), ("num_ca", "numCA"), ("num_bv_tests", "num_bv_tests"),
          ("num_tri_tests", "num_tri_tests"), ("toc", "toc"), ("distance", "distance"), ("mint", "mint"),
          ("pose_toc", "pose_toc"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def assert_contract(got, ref, tol_t):
    """BASELINE.json's contract, checked explicitly besides the bit-exact comparison."""
    assert (got["status"] == 0).all()
    assert np.array_equal(got["collisionfree"], ref["collisionfree"])
    assert (np.abs(got["toc"] - ref["toc"]) <= tol_t).all()
    assert (np.abs(got["distance"] - ref["distance"]) <= 1e-9 * np.maximum(1.0, np.abs(ref["distance"]))).all()


# ---------------------------------------------------------------- device functions ----------
def test_device_sincos_match_libm():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-3.3, 3.3, 400000), rng.uniform(-0.13, 0.13, 100000), rng.uniform(-1e4, 1e4, 100000),
                        np.array([0.0, -0.0, 0.126, 1 / 128, 0.85546875, 2.426265, math.pi, 2 ** -26, 2 ** -27])])
    s = np.zeros_like(x); c = np.zeros_like(x)
    api._check(api.lib().c2a_b200_test_sincos(_p(x), C.c_int64(len(x)), _p(s), _p(c)))
    ms = np.array([math.sin(v) for v in x]); mc = np.array([math.cos(v) for v in x])
    assert np.array_equal(s, ms) and np.array_equal(c, mc)


def test_device_rect_dist_matches_oracle():
    rng = np.random.default_rng(21)
    cases = rect_cases(rng, 20000)
    n = len(cases)
    R = np.array([c[0] for c in cases]); T = np.array([c[1] for c in cases])
    ab = np.array([np.concatenate([c[2], c[3]]) for c in cases])
    dist = np.zeros(n); S = np.full((n, 3), np.nan)
    api._check(api.lib().c2a_b200_test_rect_dist(_p(R), _p(T), _p(ab), C.c_int64(n), _p(dist), _p(S)))
    P = oracle.port()
    for i in range(n):
        d, _, _, s = P.rect_dist(R[i], T[i], ab[i, :2], ab[i, 2:])
        assert d == dist[i], i
        assert np.array_equal(s, S[i], equal_nan=True), i
    assert (dist == 0).sum() > 100 and (dist > 0).sum() > 1000


def test_device_tri_distance_matches_oracle():
    rng = np.random.default_rng(22)
    cases = tri_cases(rng, 20000)
    n = len(cases)
    R = np.array([c[0] for c in cases]); T = np.array([c[1] for c in cases])
    t1 = np.array([c[2] for c in cases]); t2 = np.array([c[3] for c in cases])
    dist = np.zeros(n); pq = np.zeros((n, 6))
    api._check(api.lib().c2a_b200_test_tri_distance(_p(R), _p(T), _p(t1), _p(t2), C.c_int64(n), _p(dist), _p(pq)))
    P = oracle.port()
    for i in range(n):
        d, p, q = P.tri_distance(R[i], T[i], t1[i], t2[i])
        assert d == dist[i] or (math.isnan(d) and math.isnan(dist[i])), i
        assert np.array_equal(np.concatenate([p, q]), pq[i], equal_nan=True), i
    assert (dist == 0).sum() > 100


def test_device_motion_matches_oracle():
    """integrate() and both motion bounds against the oracle port."""
    rng = np.random.default_rng(23)
    n = 5000
    poses = workloads.approach_batch(n, 99)
    rec = np.ascontiguousarray(api.motions_from_poses(poses)[:, :24])
    t = rng.uniform(0, 1.2, n); ar = rng.uniform(0, 200, n); d = rng.normal(size=(n, 3))
    d[::50] = 0.0  # zero direction -> NaN normalisation, like a zero triangle distance (quirk Q7)
    out = np.zeros((n, 14))
    api._check(api.lib().c2a_b200_test_motion(_p(rec), _p(t), _p(ar), _p(d), C.c_int64(n), _p(out)))

    class M(C.Structure):
        _fields_ = [("Rs", C.c_double * 9), ("Ts", C.c_double * 3), ("Re", C.c_double * 9), ("Te", C.c_double * 3),
                    ("cv", C.c_double * 3), ("axis", C.c_double * 3), ("w", C.c_double), ("Rc", C.c_double * 9), ("Tc", C.c_double * 3)]
    L = oracle.port().lib
    for i in range(n):
        p = np.ascontiguousarray(poses[i, :24]); m = M()
        L.orc_motion_init(C.byref(m), _p(p[0:]), _p(p[9:]), _p(p[12:]), _p(p[21:]))
        L.orc_motion_integrate(C.byref(m), C.c_double(t[i]), None)
        assert list(m.Rc) == list(out[i, :9]) and list(m.Tc) == list(out[i, 9:12]), i
        n1 = d[i].copy(); n2 = d[i].copy()
        b1 = L.orc_motion_bound_bv(C.byref(m), C.c_double(ar[i]), _p(n1))
        b2 = L.orc_motion_bound_leaf(C.byref(m), C.c_double(ar[i]), _p(n2))
        assert np.array_equal(np.array([b1, b2]), out[i, 12:14], equal_nan=True), i


# ---------------------------------------------------------------- whole queries -------------
@pytest.mark.parametrize("case,ma,mb", GOLDEN_CASES)
def test_golden_bit_exact(case, ma, mb, golden, models):
    """Every query of every committed reference fixture, through c2a_b200_solve_batch (host buffers)."""
    g = golden(case)
    tol_d, tol_t = float(g["tol_d"]), float(g["tol_t"])
    got = api.solve_batch(models(ma), models(mb), g["poses"], g.get("seed_a"), g.get("seed_b"), tol_d, tol_t)
    assert_contract(got, g, tol_t)
    for a, b in FIELDS:
        if b in g:  # (the compact fixtures carry no pose_toc)
            assert np.array_equal(got[a], g[b]), (case, a, int((got[a] != g[b]).sum()))
    upd = g["num_tri_tests"] > 0
    same = (got["p1p2"] == g["p1p2"]).all(1)
    assert same[upd & (got["p1p2"] != 0).any(1)].all()
    if "last_tri" in g:
        assert np.array_equal(got["last_tri"], g["last_tri"])


@pytest.mark.parametrize("case,ma,mb", TRANSLATION_CASES)
def test_translation_golden_bit_exact(case, ma, mb, golden, models):
    """The translation-only branch (C2A.cpp:2391-2395, :1362-1521; c2a_translation_kernel) against the reference's
    object code on pure translations: every field bit-exact, incl. the res->last_triA/B the traversal leaves."""
    g = golden(case)
    tol_d, tol_t = float(g["tol_d"]), float(g["tol_t"])
    got = api.solve_batch(models(ma), models(mb), g["poses"], g.get("seed_a"), g.get("seed_b"), tol_d, tol_t,
                          max_contacts=8 if "num_contact" in g else 0)
    assert_contract(got, g, tol_t)
    for a, b in FIELDS:
        assert np.array_equal(got[a], g[b]), (case, a, int((got[a] != g[b]).sum()))
    assert np.array_equal(got["last_tri"], g["last_tri"])
    assert (got["p1p2"] == 0).all()
    if "num_contact" in g:  # the full, unmodified C2A_Solve's number_of_contact on the first queries
        k = len(g["num_contact"])
        assert np.array_equal(got["num_contact"][:k], g["num_contact"])


def test_translation_mixed_batch_and_device_entry(models, bvhs):
    """Rotational and translation-only queries interleaved in one batch, through the host entry and the
    device-resident entry (claim order given), vs the oracle port."""
    import torch
    poses = np.concatenate([workloads.translation_batch(150, 41, radius=workloads.KNOT_RADIUS, move_b=True),
                            workloads.approach_batch(150, 42, radius=workloads.KNOT_RADIUS)])
    poses = np.ascontiguousarray(poses[np.random.default_rng(1).permutation(len(poses))])
    n = len(poses)
    ref = oracle.port().solve_batch(bvhs("knot_128x16"), bvhs("knot_128x16"), poses, threads=8)
    m = models("knot_128x16")
    got = api.solve_batch(m, m, poses)
    assert_contract(got, ref, 1e-4)
    for a, b in FIELDS:
        assert np.array_equal(got[a], ref[b]), a
    assert (ref["numCA"] == 0).sum() == 150
    mot = torch.from_numpy(api.motions_from_poses(poses)).cuda()
    out = {"status": torch.full((n,), -7, dtype=torch.int32, device="cuda"),
           "collisionfree": torch.zeros(n, dtype=torch.int32, device="cuda"),
           "num_ca": torch.full((n,), -1, dtype=torch.int32, device="cuda"),
           "toc": torch.zeros(n, dtype=torch.float64, device="cuda"),
           "distance": torch.zeros(n, dtype=torch.float64, device="cuda")}
    api.solve_batch_device(m, m, mot.data_ptr(), n, {k: v.data_ptr() for k, v in out.items()},
                           stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for k, v in out.items():
        assert np.array_equal(v.cpu().numpy(), got[k]), k


@pytest.mark.parametrize("tag", ["exact", "approx"])
def test_distance_query_bit_exact(tag, golden, models, bvhs):
    """Batched C2A_Distance (c2a_distance_kernel) against the reference's object code, and on fresh inputs against the port."""
    g = golden("ref_distance_knot_128x16")
    rel, ab = g[f"{tag}_err"]
    m = models("knot_128x16")
    got = api.distance_batch(m, m, g["poses24"], g["seed_a"], g["seed_b"], rel, ab)
    assert np.array_equal(got["distance"], g[f"{tag}_distance"])
    assert np.array_equal(got["p1p2"], np.concatenate([g[f"{tag}_p1"], g[f"{tag}_p2"]], 1))
    assert np.array_equal(got["tri_pair"], np.stack([g[f"{tag}_tri_a"], g[f"{tag}_tri_b"]], 1))
    assert np.array_equal(got["num_bv_tests"], g[f"{tag}_num_bv_tests"]) and np.array_equal(got["num_tri_tests"], g[f"{tag}_num_tri_tests"])
    poses = workloads.static_pose_batch(257, 99, radius=workloads.KNOT_RADIUS)
    ref = oracle.port().distance(bvhs("knot_128x16"), bvhs("knot_128x16"), poses, None, None, rel, ab)
    got = api.distance_batch(m, m, poses, None, None, rel, ab)
    assert np.array_equal(got["distance"], ref["distance"]) and np.array_equal(got["num_bv_tests"], ref["num_bv_tests"])
    assert np.array_equal(got["tri_pair"], np.stack([ref["tri_a"], ref["tri_b"]], 1))
    assert api.distance_batch(m, m, np.zeros((0, 24)))["distance"].shape == (0,)
    # coincident models (distance 0 at once) and models very far apart
    sp = workloads.static_pose_batch(20, 77, radius=workloads.KNOT_RADIUS)
    same = sp.copy(); same[:, 12:24] = same[:, 0:12]
    far = sp.copy(); far[:, 9:12] *= 50.0
    poses = np.ascontiguousarray(np.concatenate([same, far]))
    ref = oracle.port().distance(bvhs("knot_128x16"), bvhs("knot_128x16"), poses, None, None, rel, ab)
    got = api.distance_batch(m, m, poses, None, None, rel, ab)
    assert np.array_equal(got["distance"], ref["distance"]) and np.array_equal(got["num_bv_tests"], ref["num_bv_tests"])
    assert np.array_equal(got["p1p2"], np.concatenate([ref["p1"], ref["p2"]], 1)) and (got["distance"][:20] == 0).all()


@pytest.mark.parametrize("qsize", [3, 10])
@pytest.mark.parametrize("tag", ["exact", "approx"])
def test_distance_queue_routine_bit_exact(qsize, tag, golden, models, bvhs):
    """C2A_Distance with qsize > 2 (c2a_distance_queue_kernel: best-first over a bounded queue, a frame per nested call)
    against the reference's object code, and on a heterogeneous pair against the port."""
    g, gq = golden("ref_distance_knot_128x16"), golden("ref_distance_queue_knot_128x16")
    rel, ab = g[f"{tag}_err"]
    m = models("knot_128x16")
    got = api.distance_batch(m, m, g["poses24"], g["seed_a"], g["seed_b"], rel, ab, qsize=qsize)
    pre = f"q{qsize}_{tag}_"
    assert np.array_equal(got["distance"], gq[pre + "distance"])
    assert np.array_equal(got["p1p2"], np.concatenate([gq[pre + "p1"], gq[pre + "p2"]], 1))
    assert np.array_equal(got["tri_pair"], np.stack([gq[pre + "tri_a"], gq[pre + "tri_b"]], 1))
    assert np.array_equal(got["num_bv_tests"], gq[pre + "num_bv_tests"]) and np.array_equal(got["num_tri_tests"], gq[pre + "num_tri_tests"])
    poses = workloads.static_pose_batch(200, 123, radius=workloads.BUNNY_RADIUS)
    ref = oracle.port().distance(bvhs("bunny"), bvhs("knot_512x32"), poses, None, None, rel, ab, qsize=qsize)
    got = api.distance_batch(models("bunny"), models("knot_512x32"), poses, None, None, rel, ab, qsize=qsize)
    assert np.array_equal(got["distance"], ref["distance"]) and np.array_equal(got["num_bv_tests"], ref["num_bv_tests"])
    assert np.array_equal(got["tri_pair"], np.stack([ref["tri_a"], ref["tri_b"]], 1)) and np.array_equal(got["num_tri_tests"], ref["num_tri_tests"])
    # qsize <= 2 is the depth-first routine, whichever entry is called
    a = api.distance_batch(m, m, g["poses24"][:50], g["seed_a"][:50], g["seed_b"][:50], rel, ab)
    b = api.distance_batch(m, m, g["poses24"][:50], g["seed_a"][:50], g["seed_b"][:50], rel, ab, _entry="c2a_b200_distance_queue_batch", qsize=2)
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_degenerate_motions_against_oracle_port(models, bvhs):
    """No motion at all, rotation in place, contact at the start pose, vanishing motions (both branches)."""
    poses = workloads.degenerate_batch(radius=workloads.KNOT_RADIUS)
    ref = oracle.port().solve_batch(bvhs("knot_128x16"), bvhs("knot_128x16"), poses, threads=8)
    got = api.solve_batch(models("knot_128x16"), models("knot_128x16"), poses)
    assert (got["status"] == 0).all()
    for a, b in FIELDS:
        assert np.array_equal(got[a], ref[b], equal_nan=True), (a, np.nonzero([not np.array_equal(x, y, equal_nan=True) for x, y in zip(got[a], ref[b])])[0][:8])


def test_fresh_batch_against_oracle_port(models, bvhs):
    """Inputs that are in no fixture: GPU vs the oracle port run here, bit-exact."""
    poses = workloads.approach_batch(300, 777, radius=workloads.KNOT_RADIUS, max_turn=3.1)
    ref = oracle.port().solve_batch(bvhs("knot_128x16"), bvhs("knot_128x16"), poses, threads=8)
    got = api.solve_batch(models("knot_128x16"), models("knot_128x16"), poses)
    assert_contract(got, ref, 1e-4)
    for a, b in FIELDS:
        assert np.array_equal(got[a], ref[b]), a
    # the traversal's side effect o->last_tri (the seeds of the demo's next call)
    assert np.array_equal(got["last_tri"][:, 0], ref["last_tri_a"]) and np.array_equal(got["last_tri"][:, 1], ref["last_tri_b"])


@pytest.mark.parametrize("case,model,n", [("ref_knot_128x16_carry", "knot_128x16", 24), ("ref_demo_bunny_carry", "bunny", 340)])
def test_carried_seeds_sequence(models, golden, case, model, n):
    """Demo mode: each call's seeds are the previous call's last_tri (cross-query state)."""
    g = golden(case)
    m = models(model)
    sa = sb = 0
    for i in range(n):
        assert (sa, sb) == (int(g["seed_a"][i]), int(g["seed_b"][i]))
        got = api.solve_batch(m, m, g["poses"][i:i + 1], [sa], [sb])
        for a, b in FIELDS:
            assert np.array_equal(got[a][0], g[b][i]), (i, a)
        assert np.array_equal(got["last_tri"][0], g["last_tri"][i])
        if got["last_tri"][0, 0] >= 0:
            sa = int(got["last_tri"][0, 0])
        if got["last_tri"][0, 1] >= 0:
            sb = int(got["last_tri"][0, 1])


def test_tolerance_sweep_against_oracle_port(models, bvhs):
    """config 5 flavour: tolerance_t from 1e-3 to 1e-6 through the exposed tolerances."""
    poses = workloads.approach_batch(120, 31337, radius=workloads.KNOT_RADIUS)
    for tol_t in (1e-3, 1e-5, 1e-6):
        ref = oracle.port().solve_batch(bvhs("knot_128x16"), bvhs("knot_128x16"), poses, tol_d=1e-4, tol_t=tol_t, threads=8)
        got = api.solve_batch(models("knot_128x16"), models("knot_128x16"), poses, tol_d=1e-4, tol_t=tol_t)
        assert_contract(got, ref, tol_t)
        for a, b in FIELDS:
            assert np.array_equal(got[a], ref[b]), (tol_t, a)


def test_device_resident_entry_matches_host_entry(models, golden):
    """c2a_b200_solve_batch_device with torch tensors == the host-buffer entry."""
    import torch
    g = golden("ref_knot_512x32")
    n = 700
    m = models("knot_512x32")
    host = api.solve_batch(m, m, g["poses"][:n])
    mot = torch.from_numpy(api.motions_from_poses(g["poses"][:n])).cuda()
    out = {"status": torch.full((n,), -7, dtype=torch.int32, device="cuda"),
           "collisionfree": torch.zeros(n, dtype=torch.int32, device="cuda"),
           "num_ca": torch.zeros(n, dtype=torch.int32, device="cuda"),
           "toc": torch.zeros(n, dtype=torch.float64, device="cuda"),
           "distance": torch.zeros(n, dtype=torch.float64, device="cuda"),
           "pose_toc": torch.zeros(n, 24, dtype=torch.float64, device="cuda")}
    stream = torch.cuda.current_stream()
    api.solve_batch_device(m, m, mot.data_ptr(), n, {k: v.data_ptr() for k, v in out.items()}, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    for k, v in out.items():
        assert np.array_equal(v.cpu().numpy(), host[k]), k


def test_edge_cases(models):
    m = models("knot_128x16")
    # empty batch
    got = api.solve_batch(m, m, np.zeros((0, 48)))
    assert got["toc"].shape == (0,)
    # a single query
    one = workloads.approach_batch(1, 1, radius=workloads.KNOT_RADIUS)
    assert api.solve_batch(m, m, one)["status"][0] == 0
    # ragged sizes around the warp / block size
    for n in (31, 32, 33, 127, 129):
        p = workloads.approach_batch(n, n, radius=workloads.KNOT_RADIUS)
        a = api.solve_batch(m, m, p)
        b = api.solve_batch(m, m, p[::-1].copy())
        assert np.array_equal(a["toc"], b["toc"][::-1]) and np.array_equal(a["num_bv_tests"], b["num_bv_tests"][::-1])
    # pure translation inside a rotational batch: the reference switches branch (C2A.cpp:2391-2395); solved (num_ca == 0)
    p = workloads.approach_batch(4, 2, radius=workloads.KNOT_RADIUS)
    p[1, 12:21] = p[1, 0:9]
    got = api.solve_batch(m, m, p)
    assert list(got["status"]) == [0, 0, 0, 0] and list(got["num_ca"] == 0) == [False, True, False, False]
    # bad seeds / mismatched devices are argument errors, not crashes
    with pytest.raises(api.C2AError):
        api._check(api.lib().c2a_b200_solve_batch(m.h, None, None, None, None, C.c_int64(1), C.c_double(1e-4), C.c_double(1e-4), None))


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_small_meshes_every_entry_against_port(seed):
    """Random triangle soups of 1 .. 40 triangles (hierarchies of depth 0 .. ~8, root-leaf pairs, slivers, one model much
    larger than the other): the CCD query (both branches), the contact pass, both distance routines and both C2A_Collide
    overloads against the oracle port, bit for bit."""
    rng = np.random.default_rng(seed)
    P = oracle.port()
    for trial in range(6):
        na, nb = int(rng.integers(1, 41)), int(rng.integers(1, 41))
        scale_b = float(rng.choice([0.2, 1.0, 5.0]))
        ta = rng.normal(size=(na, 3, 3)) + rng.normal(scale=2.0, size=(na, 1, 3))
        tb = (rng.normal(size=(nb, 3, 3)) + rng.normal(scale=2.0, size=(nb, 1, 3))) * scale_b
        if trial % 3 == 0:
            ta[0, 2] = ta[0, 1] + 1e-9 * (ta[0, 1] - ta[0, 0])   # a sliver
        ba, bb = api.build_bvh(ta.reshape(na, 9)), api.build_bvh(tb.reshape(nb, 9))
        ma, mb = api.Model(ba, 0), api.Model(bb, 0)
        rad = max(float(np.linalg.norm(ba["tris"].reshape(-1, 3), axis=1).max()), float(np.linalg.norm(bb["tris"].reshape(-1, 3), axis=1).max()))
        poses = workloads.approach_batch(64, seed * 100 + trial, radius=rad)
        tp = workloads.translation_batch(24, seed * 100 + trial, radius=rad, move_b=True)
        sa = rng.integers(0, na, 64).astype(np.int32); sb = rng.integers(0, nb, 64).astype(np.int32)
        ref = P.solve_batch(ba, bb, poses, sa, sb, threads=4)
        got = api.solve_batch(ma, mb, poses, sa, sb)
        for a, b in FIELDS:
            assert np.array_equal(got[a], ref[b]), (seed, trial, a)
        ref = P.solve_batch(ba, bb, tp, threads=1)
        got = api.solve_batch(ma, mb, tp)
        for a, b in FIELDS:
            assert np.array_equal(got[a], ref[b]), (seed, trial, "translation", a)
        sp = workloads.static_pose_batch(48, seed * 100 + trial, radius=rad)
        for qs in (2, 4):
            rd = P.distance(ba, bb, sp, sa[:48], sb[:48], qsize=qs); gd = api.distance_batch(ma, mb, sp, sa[:48], sb[:48], qsize=qs)
            assert np.array_equal(gd["distance"], rd["distance"]) and np.array_equal(gd["num_bv_tests"], rd["num_bv_tests"]), (seed, trial, qs)
            assert np.array_equal(gd["tri_pair"], np.stack([rd["tri_a"], rd["tri_b"]], 1)), (seed, trial, qs)
        rn, rp, rbv, rtr = P.collide(ba, bb, sp, max_pairs=2048); gc = api.collide_batch(ma, mb, sp, max_pairs=2048)
        assert np.array_equal(gc["num_pairs"], rn) and np.array_equal(gc["num_bv_tests"], rbv) and np.array_equal(gc["num_tri_tests"], rtr), (seed, trial)
        for i in range(len(sp)):
            assert np.array_equal(gc["pairs"][i, :rn[i]], rp[i]), (seed, trial, i)
        rd = P.collide_distance(ba, bb, sp, sa[:48], sb[:48]); gd = api.collide_distance_batch(ma, mb, sp, sa[:48], sb[:48])
        assert np.array_equal(gd["distance"], rd["distance"]) and np.array_equal(gd["num_tri_tests"], rd["num_tri_tests"]), (seed, trial)
        thr = np.full(len(sp), 0.3 * rad)
        num, recs = api.contacts_batch(ma, mb, sp, thr, max_contacts=32)
        for i in range(0, len(sp), 6):
            n_ref, r_ref = P.contacts(ba, bb, sp[i, :12], sp[i, 12:], thr[i], max_out=32)
            assert num[i] == n_ref, (seed, trial, i)
            k = min(n_ref, 32)
            assert np.array_equal(recs[i][:k]["dist"], r_ref["dist"][:k]) and np.array_equal(recs[i][:k]["tri_a"], r_ref["tri_a"][:k]), (seed, trial, i)
        ma.free(); mb.free()


def test_large_batch_properties(models):
    """At a size the CPU oracle cannot check query by query: size-independent properties.
    (i) the dynamic scheduler is order independent: a permuted batch gives the permuted results;
    (ii) repeat runs are identical; (iii) verdict/toc invariants; (iv) a CPU-checked random sample."""
    n = 120000
    m = models("knot_512x32")
    poses = workloads.approach_batch(n, 20260002, radius=workloads.KNOT_RADIUS)
    f = ("status", "collisionfree", "toc", "distance", "num_ca", "num_bv_tests", "num_tri_tests")
    a = api.solve_batch(m, m, poses, fields=f)
    perm = np.random.default_rng(1).permutation(n)
    b = api.solve_batch(m, m, poses[perm], fields=f)
    for k in f:
        assert np.array_equal(a[k][perm], b[k]), k
    assert (a["status"] == 0).all()
    free = a["collisionfree"] == 1
    assert (a["toc"][free] == 0).all() and ((a["toc"] >= 0) & (a["toc"] < 1)).all()
    assert (a["num_ca"] >= 1).all() and (a["num_ca"] <= 152).all()
    assert (a["distance"][~free] >= 0).all()
    assert 0.85 < (~free).mean() < 0.999


def test_large_batch_sample_against_oracle(models, bvhs):
    n = 120000
    poses = workloads.approach_batch(n, 20260002, radius=workloads.KNOT_RADIUS)
    idx = np.random.default_rng(2).choice(n, 400, replace=False)
    m = models("knot_512x32")
    got = api.solve_batch(m, m, poses[idx])
    ref = oracle.port().solve_batch(bvhs("knot_512x32"), bvhs("knot_512x32"), poses[idx], threads=16)
    assert_contract(got, ref, 1e-4)
    for a, b in FIELDS:
        assert np.array_equal(got[a], ref[b]), a


def test_contact_pass_matches_reference(models, golden):
    """The contact pass of C2A_Solve (number_of_contact + ContactF list) fused behind c2a_b200_solve_batch,
    and the standalone c2a_b200_contacts_batch, against the lists the unmodified reference produced."""
    from test_oracle import contacts_equal, golden_contacts
    counts, lists = golden_contacts(golden)
    g = golden("ref_knot_128x16")
    n = len(counts)
    tris, vidx = meshes.torus_knot(128, 16)
    m = api.Model(api.build_bvh(tris, vidx), 0)
    got = api.solve_batch(m, m, g["poses"][:n], max_contacts=32)
    assert np.array_equal(got["num_contact"], counts)
    assert np.array_equal(got["toc"], g["toc"][:n])
    for i in range(n):
        recs = got["contacts"][i][:counts[i]][::-1]  # device order = visiting order; the reference's list is reversed
        for a, c in zip(lists[i], recs):
            assert contacts_equal(a, c), i
    # truncation keeps the count
    small = api.solve_batch(m, m, g["poses"][:n], max_contacts=2, fields=("status",))
    assert np.array_equal(small["num_contact"], counts)
    # standalone entry at the TOC poses
    hits = np.where(g["collisionfree"][:n] == 0)[0]
    num, recs = api.contacts_batch(m, m, g["pose_toc"][hits], 2 * g["distance"][hits] + 0.001, max_contacts=32)
    assert np.array_equal(num, counts[hits])
    for j, i in enumerate(hits[:40]):
        for a, c in zip(lists[i], recs[j][:num[j]][::-1]):
            assert contacts_equal(a, c), i


def skinny_mesh(n=50, ratio=64.0):
    """Triangles whose positions and sizes grow geometrically with a ratio larger than their number: the builder's mean
    split peels off one triangle per level, so the hierarchy is a chain of depth n - 1."""
    tris = []
    for i in range(n):
        x = ratio ** i * 1e-40
        s = 0.3 * x
        tris.append([x, 0, 0, x + s, 0.1 * s, 0, x, s, 0.05 * s])
    return np.array(tris)


def test_deep_hierarchies_run_on_global_memory_stacks():
    """depth(A) + depth(B) + 2 = 100 > 96: deeper than the local-memory stacks of the translation-only, contact and
    distance kernels, and than the 56 key levels of the wide kernel.  Round 1 returned a status / ERR_DEPTH there; now
    the same kernels run with their stacks in global memory (and the wide kernel one pair per round) -- checked
    bit for bit against the oracle port, which recurses without a depth limit."""
    tris = skinny_mesh()
    bvh = api.build_bvh(tris)
    assert 2 * bvh["depth"] + 2 > 96
    m = api.Model(bvh, 0)
    rad = float(np.linalg.norm(bvh["tris"].reshape(-1, 3), axis=1).max())
    P = oracle.port()
    # rotational queries (main kernel; the ones past their fifth CA step finish in the wide kernel)
    poses = workloads.approach_batch(300, 7, radius=rad)
    ref = P.solve_batch(bvh, bvh, poses, threads=8)
    got = api.solve_batch(m, m, poses)
    assert (got["status"] == 0).all()
    for a, b in FIELDS:
        assert np.array_equal(got[a], ref[b]), a
    assert (ref["numCA"] > 6).sum() >= 3
    # translation-only queries
    tp = workloads.translation_batch(120, 8, radius=rad, move_b=True)
    ref = P.solve_batch(bvh, bvh, tp, threads=1)
    got = api.solve_batch(m, m, tp)
    assert (got["status"] == 0).all() and (ref["numCA"] == 0).all()
    for a, b in FIELDS:
        assert np.array_equal(got[a], ref[b]), a
    assert np.array_equal(got["last_tri"], np.stack([ref["last_tri_a"], ref["last_tri_b"]], 1))
    # discrete distance and the contact pass
    sp = workloads.static_pose_batch(100, 9, radius=rad)
    rd = P.distance(bvh, bvh, sp)
    gd = api.distance_batch(m, m, sp)
    assert np.array_equal(gd["distance"], rd["distance"]) and np.array_equal(gd["num_bv_tests"], rd["num_bv_tests"])
    assert np.array_equal(gd["tri_pair"], np.stack([rd["tri_a"], rd["tri_b"]], 1))
    thr = np.full(len(sp), 0.05 * rad)
    num, recs = api.contacts_batch(m, m, sp, thr, max_contacts=64)
    for i in range(len(sp)):
        n_ref, r_ref = P.contacts(bvh, bvh, sp[i, :12], sp[i, 12:], thr[i], max_out=64)
        assert num[i] == n_ref, i
        k = min(n_ref, 64)
        assert np.array_equal(recs[i][:k]["tri_a"], r_ref["tri_a"][:k]) and np.array_equal(recs[i][:k]["dist"], r_ref["dist"][:k]), i


    rq = P.distance(bvh, bvh, sp, qsize=5)
    gq = api.distance_batch(m, m, sp, qsize=5)     # 100 frames of nested calls available, a few used
    assert np.array_equal(gq["distance"], rq["distance"]) and np.array_equal(gq["num_bv_tests"], rq["num_bv_tests"])
    assert np.array_equal(gq["tri_pair"], np.stack([rq["tri_a"], rq["tri_b"]], 1))
    # C2A_Collide, both overloads, on the same chain-shaped hierarchy
    rn, rp, rbv, rtr = P.collide(bvh, bvh, sp, max_pairs=64)
    gc = api.collide_batch(m, m, sp, max_pairs=64)
    assert np.array_equal(gc["num_pairs"], rn) and np.array_equal(gc["num_bv_tests"], rbv) and np.array_equal(gc["num_tri_tests"], rtr)
    for i in range(len(sp)):
        assert np.array_equal(gc["pairs"][i, :min(rn[i], 64)], rp[i]), i
    rd = P.collide_distance(bvh, bvh, sp)
    gd = api.collide_distance_batch(m, m, sp)
    assert np.array_equal(gd["distance"], rd["distance"]) and np.array_equal(gd["num_bv_tests"], rd["num_bv_tests"])


@pytest.mark.parametrize("case,ma,mb", COLLIDE_CASES)
def test_collide_bit_exact(case, ma, mb, golden, models, bvhs):
    """Batched C2A_Collide, PQP_CollideResult overload (c2a_collide_kernel): pair lists in the reference's reporting order,
    counters, both flags, against the reference's object code; truncated and count-only output."""
    g = golden(f"ref_collide_{case}")
    a, b = models(ma), models(mb)
    ia, ib = bvhs(ma)["tri_ids"], bvhs(mb)["tri_ids"]
    cap = int(g["all_num_pairs"].max())
    for name, flag in (("all", api.ALL_CONTACTS), ("first", api.FIRST_CONTACT)):
        got = api.collide_batch(a, b, g["poses24"], flag=flag, max_pairs=cap)
        assert np.array_equal(got["num_pairs"], g[f"{name}_num_pairs"])
        assert np.array_equal(got["num_bv_tests"], g[f"{name}_num_bv_tests"]) and np.array_equal(got["num_tri_tests"], g[f"{name}_num_tri_tests"])
        for i, want in enumerate(split_pairs(g[f"{name}_num_pairs"], g[f"{name}_pairs"])):
            k = len(want)
            assert np.array_equal(np.stack([ia[got["pairs"][i, :k, 0]], ib[got["pairs"][i, :k, 1]]], 1), want), (name, i)
            assert (got["pairs"][i, k:] == -1).all()
    few = api.collide_batch(a, b, g["poses24"], max_pairs=3)      # more pairs than room: all counted, the first three kept
    full = api.collide_batch(a, b, g["poses24"], max_pairs=cap)
    assert np.array_equal(few["num_pairs"], g["all_num_pairs"]) and np.array_equal(few["pairs"], full["pairs"][:, :3])
    none = api.collide_batch(a, b, g["poses24"], max_pairs=0)
    assert np.array_equal(none["num_pairs"], g["all_num_pairs"])
    assert api.collide_batch(a, b, np.zeros((0, 24)))["num_pairs"].shape == (0,)
    with pytest.raises(api.C2AError):
        api.collide_batch(a, b, g["poses24"], flag=3)


@pytest.mark.parametrize("case,ma,mb", COLLIDE_CASES)
@pytest.mark.parametrize("tag", ["exact", "approx"])
def test_collide_distance_overload_bit_exact(case, ma, mb, tag, golden, models):
    """C2A_Collide, C2A_DistanceResult overload: c2a_distance_kernel behind the box-overlap gate."""
    g = golden(f"ref_collide_{case}")
    rel, ab = g[f"dist_{tag}_err"]
    got = api.collide_distance_batch(models(ma), models(mb), g["poses24"], g["seed_a"], g["seed_b"], rel, ab)
    assert np.array_equal(got["distance"], g[f"dist_{tag}_distance"])
    assert np.array_equal(got["p1p2"], np.concatenate([g[f"dist_{tag}_p1"], g[f"dist_{tag}_p2"]], 1))
    assert np.array_equal(got["tri_pair"], np.stack([g[f"dist_{tag}_tri_a"], g[f"dist_{tag}_tri_b"]], 1))
    assert np.array_equal(got["num_bv_tests"], g[f"dist_{tag}_num_bv_tests"]) and np.array_equal(got["num_tri_tests"], g[f"dist_{tag}_num_tri_tests"])


def test_collide_needs_box_data(bvhs):
    """A model uploaded without obb_d / obb_To serves every other query; the C2A_Collide entries refuse it."""
    b = {k: v for k, v in bvhs("knot_128x16").items() if not k.startswith("obb_")}
    m = api.Model(b, 0)
    sp = workloads.static_pose_batch(8, 5, radius=workloads.KNOT_RADIUS)
    assert api.distance_batch(m, m, sp)["distance"].shape == (8,)
    with pytest.raises(api.C2AError):
        api.collide_batch(m, m, sp)
    with pytest.raises(api.C2AError):
        api.collide_distance_batch(m, m, sp)


def test_multi_device_entry_matches_single_device(models, bvhs, golden):
    """c2a_b200_solve_batch_multi: the same batch sharded over every GPU of the box (interleaved cost-sorted shards, one
    host thread + stream per device, results scattered back) gives exactly the single-device results.  With one GPU the
    entry degenerates to c2a_b200_solve_batch; the argument checks are exercised either way."""
    g = golden("ref_knot_128x16")
    n_dev = api.device_count()
    bvh = bvhs("knot_128x16")
    reps = [models("knot_128x16")] + [api.Model(bvh, d) for d in range(1, min(n_dev, 4))]
    rng = np.random.default_rng(4)
    sa = rng.integers(0, len(bvh["tris"]), len(g["poses"])).astype(np.int32)
    one = api.solve_batch(reps[0], reps[0], g["poses"], sa, None)
    many = api.solve_batch_multi(reps, reps, g["poses"], sa, None)
    for k in one:
        assert np.array_equal(one[k], many[k], equal_nan=True) if one[k].dtype.kind == "f" else np.array_equal(one[k], many[k]), k
    # a batch large enough for the claim order to be cost-sorted (>= 4096), on fresh poses
    poses = workloads.approach_batch(6000, 77, radius=workloads.KNOT_RADIUS)
    one = api.solve_batch(reps[0], reps[0], poses, fields=("status", "toc", "distance", "num_ca", "num_bv_tests", "pose_toc"))
    many = api.solve_batch_multi(reps, reps, poses, fields=("status", "toc", "distance", "num_ca", "num_bv_tests", "pose_toc"))
    for k in one:
        assert np.array_equal(one[k], many[k]), k
    with pytest.raises(api.C2AError):   # two shards on one device
        api.solve_batch_multi([reps[0], reps[0]], [reps[0], reps[0]], poses[:8])
    with pytest.raises(api.C2AError):   # a seed that is not a triangle of the model
        api.solve_batch_multi(reps, reps, poses[:8], np.full(8, 10 ** 6, np.int32), None)
    with pytest.raises(api.C2AError):
        api.solve_batch(reps[0], reps[0], poses[:8], np.full(8, -1, np.int32), None)
