"""CPU tests of the oracle: the port (oracle/c2a_oracle.cpp) against the reference's own object code
(oracle/_ref, when it was built here), against the committed golden fixtures that object code
produced, and against the closed-form known answers of SURVEY.md section 4."""
import math

import numpy as np
import pytest

import oracle
from c2a_b200 import workloads
from conftest import GOLDEN_CASES

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def rand_rot(rng):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    x, y, z, w = q
    return np.array([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)])


def rect_cases(rng, n):
    """Random rectangle pairs at mixed scales: far, near, overlapping, axis-aligned, degenerate."""
    out = []
    for i in range(n):
        R = rand_rot(rng) if i % 5 else np.eye(3).ravel()
        scale = 10.0 ** rng.uniform(-1, 1.5)
        T = rng.normal(size=3) * scale * rng.choice([0.05, 0.5, 2.0])
        a = rng.uniform(0, 4, 2) * (i % 7 != 0)
        b = rng.uniform(0, 4, 2) * (i % 11 != 0)
        out.append((R, T, a, b))
    return out


def test_rect_dist_known_answers():
    """a=b=(1,1), Rab=I."""
    P = oracle.port()
    I = np.eye(3).ravel(); ab = np.array([1.0, 1.0])
    d, p, q, s = P.rect_dist(I, [0, 0, 2], ab, ab)
    assert d == 2 and list(s) == [0, 0, 2] and list(p) == [0, 0, 0] and list(q) == [0, 0, 2]
    d, p, q, s = P.rect_dist(I, [3, 0, 0], ab, ab)
    assert d == 2 and list(s) == [2, 0, 0] and list(p) == [1, 0, 0] and list(q) == [3, 0, 0]
    d, p, q, s = P.rect_dist(I, [3, 4, 12], ab, ab)
    assert d == math.sqrt(157.0) and list(s) == [2, 3, 12]
    d, p, q, s = P.rect_dist(I, [0.25, 0.25, 0], ab, ab)
    assert d == 0 and list(s) == [0, 0, 0]


@needs_ref
def test_rect_dist_port_vs_reference_header():
    """bit-exact against C2ARectDist compiled from /root/reference/C2A/C2A_RectDist.h"""
    P, R = oracle.port(), oracle.ref()
    rng = np.random.default_rng(11)
    for Rab, T, a, b in rect_cases(rng, 4000):
        d0, p0, q0, s0 = R.rect_dist(Rab, T, a, b)
        d1, p1, q1, s1 = P.rect_dist(Rab, T, a, b)
        assert d0 == d1
        assert np.array_equal(s0, s1, equal_nan=True)
        if d0 > 0:
            assert np.array_equal(p0, p1) and np.array_equal(q0, q1)


def tri_cases(rng, n):
    out = []
    for i in range(n):
        t1 = rng.normal(size=9)
        t2 = rng.normal(size=9)
        if i % 6 == 0:
            t2[:3] = t1[:3]                      # shared vertex
        if i % 10 == 0:
            t2 = t1 + np.tile(rng.normal(size=3) * 0.1, 3)   # parallel copies
        if i % 13 == 0:
            t1[3:6] = t1[0:3]                    # degenerate triangle
        R = rand_rot(rng)
        T = rng.normal(size=3) * rng.choice([0.0, 0.3, 3.0])
        out.append((R, T, t1, t2))
    return out


@needs_ref
def test_tri_distance_port_vs_reference_intree():
    """The PQP TriDistance restatement against the reference's in-tree copy (C2A.cpp:165-424)."""
    P, R = oracle.port(), oracle.ref()
    rng = np.random.default_rng(5)
    n_overlap = 0
    for Rm, T, t1, t2 in tri_cases(rng, 4000):
        d0, p0, q0, col = R.tri_distance_intree(Rm, T, t1, t2)
        d1, p1, q1 = P.tri_distance(Rm, T, t1, t2)
        assert d0 == d1 or (math.isnan(d0) and math.isnan(d1))
        assert np.array_equal(p0, p1, equal_nan=True) and np.array_equal(q0, q1, equal_nan=True)
        n_overlap += col
    assert n_overlap > 50  # the overlap branch was exercised


@needs_ref
def test_port_vs_reference_fresh_batch(bunny_tris):
    """Port and reference on a batch that is NOT in the fixtures, bit-exact, including the full
    unmodified C2A_Solve (printf + contact pass) on a few queries."""
    from c2a_b200 import workloads
    R = oracle.ref()
    m = R.model(bunny_tris)
    b = m.export()
    poses = workloads.approach_batch(64, 424242)
    r = R.solve_batch(m, m, poses, threads=4)
    p = oracle.port().solve_batch(b, b, poses, threads=4)
    for k in ("collisionfree", "numCA", "num_bv_tests", "num_tri_tests", "toc", "distance", "mint", "pose_toc"):
        assert np.array_equal(r[k], p[k]), k
    full, ncont = R.solve_batch(m, m, poses[:6], mode=1)
    for k in ("collisionfree", "numCA", "toc", "distance", "pose_toc"):
        assert np.array_equal(full[k], r[k][:6]), k


@pytest.mark.parametrize("case,ma,mb", GOLDEN_CASES)
def test_port_matches_golden(case, ma, mb, golden, bvhs):
    """The port reproduces the reference's committed outputs bit for bit (a slice, to stay fast)."""
    g = golden(case)
    n = min(160 if "heavy" not in case else 24, len(g["toc"]))  # (a heavy query costs the port ~0.15 s)
    sl = slice(0, n)
    out = oracle.port().solve_batch(bvhs(ma), bvhs(mb), g["poses"][sl],
                                    g["seed_a"][sl] if "seed_a" in g else None,
                                    g["seed_b"][sl] if "seed_b" in g else None,
                                    float(g["tol_d"]), float(g["tol_t"]), threads=8)
    for k in ("collisionfree", "numCA", "num_bv_tests", "num_tri_tests", "toc", "distance", "mint", "pose_toc"):
        if k in g:  # (the compact fixtures carry no pose_toc)
            assert np.array_equal(out[k], g[k][sl]), (case, k)
    upd = (out["p1"] != 0).any(1)  # p1/p2 are only defined once a leaf updated them
    assert np.array_equal(np.concatenate([out["p1"], out["p2"]], 1)[upd], g["p1p2"][sl][upd])
    if "last_tri" in g:  # demo-mode fixtures: the traversal's o->last_tri side effect and the seed chain it feeds
        lt = np.stack([out["last_tri_a"], out["last_tri_b"]], 1)
        assert np.array_equal(lt, g["last_tri"][sl])
        nxt_a = np.where(lt[:-1, 0] >= 0, lt[:-1, 0], g["seed_a"][sl][:-1])
        nxt_b = np.where(lt[:-1, 1] >= 0, lt[:-1, 1], g["seed_b"][sl][:-1])
        assert np.array_equal(nxt_a, g["seed_a"][sl][1:]) and np.array_equal(nxt_b, g["seed_b"][sl][1:])


TRANSLATION_CASES = [("ref_translation_knot_128x16", "knot_128x16", "knot_128x16"),
                     ("ref_translation_bunny_vs_knot_seeded", "bunny", "knot_512x32")]


@pytest.mark.parametrize("case,ma,mb", TRANSLATION_CASES)
def test_port_translation_matches_golden(case, ma, mb, golden, bvhs):
    """Translation-only branch (C2A.cpp:2391-2395, :1362-1521): the port reproduces the reference's object code
    bit for bit on pure translations -- including where that code reads an uninitialised direction (see
    oracle/c2a_oracle.cpp: any finite non-zero garbage gives the same outcome)."""
    g = golden(case)
    out = oracle.port().solve_batch(bvhs(ma), bvhs(mb), g["poses"], g["seed_a"] if "seed_a" in g else None,
                                    g["seed_b"] if "seed_b" in g else None, float(g["tol_d"]), float(g["tol_t"]), threads=8)
    for k in ("collisionfree", "numCA", "num_bv_tests", "num_tri_tests", "toc", "distance", "mint", "pose_toc"):
        assert np.array_equal(out[k], g[k]), (case, k)
    assert np.array_equal(np.stack([out["last_tri_a"], out["last_tri_b"]], 1), g["last_tri"])
    # the branch's semantics: one traversal, no outer CA loop; the verdict is toc < 1; free queries keep toc = mint >= 1
    assert (g["numCA"] == 0).all() and np.array_equal(g["collisionfree"] == 1, g["toc"] >= 1.0)
    assert (g["collisionfree"] == 1).sum() > 20 and (g["collisionfree"] == 0).sum() > 20


@needs_ref
def test_port_translation_matches_ref_fresh_and_mixed(bvhs):
    """Inputs in no fixture, and a batch mixing rotational and translation-only queries (the reference's threaded
    wrapper defers the latter to a serial phase because the branch flag is a global, C2A.cpp:33)."""
    from c2a_b200 import meshes
    tris, vi = meshes.torus_knot(128, 16)
    R = oracle.ref()
    m = R.model(tris, vi)
    poses = np.concatenate([workloads.translation_batch(60, 5, radius=workloads.KNOT_RADIUS, move_b=True),
                            workloads.approach_batch(60, 6, radius=workloads.KNOT_RADIUS)])
    poses = np.ascontiguousarray(poses[np.random.default_rng(0).permutation(len(poses))])
    ref = R.solve_batch(m, m, poses, threads=4)
    out = oracle.port().solve_batch(bvhs("knot_128x16"), bvhs("knot_128x16"), poses, threads=4)
    for k in ("collisionfree", "numCA", "num_bv_tests", "num_tri_tests", "toc", "distance", "mint", "pose_toc"):
        assert np.array_equal(out[k], ref[k]), k
    assert (ref["numCA"] == 0).sum() == 60


@pytest.mark.parametrize("tag", ["exact", "approx"])
def test_port_distance_matches_golden(tag, golden, bvhs):
    """C2A_Distance (C2A_PQP.cpp:970-1056, depth-first routine): the port against the reference's object code."""
    g = golden("ref_distance_knot_128x16")
    rel, ab = g[f"{tag}_err"]
    out = oracle.port().distance(bvhs("knot_128x16"), bvhs("knot_128x16"), g["poses24"], g["seed_a"], g["seed_b"], rel, ab)
    for k in out.dtype.names:
        assert np.array_equal(out[k], g[f"{tag}_{k}"]), (tag, k)
    assert (g[f"{tag}_distance"] == 0).sum() > 20 and (g[f"{tag}_distance"] > 1).sum() > 20
    if tag == "approx":  # both bounds non-zero: fewer BV tests, distance within the bounds of the exact one
        assert g["approx_num_bv_tests"].sum() < g["exact_num_bv_tests"].sum()
        assert (g["approx_distance"] >= g["exact_distance"]).all()


@pytest.mark.parametrize("qsize", [3, 10])
@pytest.mark.parametrize("tag", ["exact", "approx"])
def test_port_distance_queue_matches_golden(qsize, tag, golden, bvhs):
    """C2A_Distance with qsize > 2 (C2ADistanceQueueRecurse, C2A_PQP.cpp:624-787): the port against the reference's object
    code (linked with the queue stand-in of oracle/pqp_shim).  With exact bounds the distance is the depth-first one."""
    g, gq = golden("ref_distance_knot_128x16"), golden("ref_distance_queue_knot_128x16")
    rel, ab = g[f"{tag}_err"]
    b = bvhs("knot_128x16")
    out = oracle.port().distance(b, b, g["poses24"], g["seed_a"], g["seed_b"], rel, ab, qsize=qsize)
    for k in out.dtype.names:
        assert np.array_equal(out[k], gq[f"q{qsize}_{tag}_{k}"]), (qsize, tag, k)
    if tag == "exact":
        assert np.array_equal(out["distance"], g["exact_distance"])
    assert (out["num_bv_tests"] != g[f"{tag}_num_bv_tests"]).sum() > 100      # another visiting order than the depth-first walk


COLLIDE_CASES = [("knot_128x16", "knot_128x16", "knot_128x16"), ("bunny_vs_knot_512x32", "bunny", "knot_512x32")]


def split_pairs(num, flat):
    """Per-query pair lists from the fixture's concatenated [sum(num), 2] array."""
    ends = np.cumsum(num)
    return [flat[e - k:e] for k, e in zip(num, ends)]


@pytest.mark.parametrize("case,ma,mb", COLLIDE_CASES)
def test_port_collide_matches_golden(case, ma, mb, golden, bvhs):
    """C2A_Collide, PQP_CollideResult overload (C2A_PQP.cpp:798-968), both flags: the port against the reference's object
    code (pairs as Tri::id in the reference's reporting order)."""
    g = golden(f"ref_collide_{case}")
    a, b = bvhs(ma), bvhs(mb)
    for name, flag in (("all", 1), ("first", 2)):
        num, pairs, nbv, ntri = oracle.port().collide(a, b, g["poses24"], flag=flag, max_pairs=4096)
        assert np.array_equal(num, g[f"{name}_num_pairs"]) and np.array_equal(nbv, g[f"{name}_num_bv_tests"])
        assert np.array_equal(ntri, g[f"{name}_num_tri_tests"])
        for got, want in zip(pairs, split_pairs(g[f"{name}_num_pairs"], g[f"{name}_pairs"])):
            assert np.array_equal(np.stack([a["tri_ids"][got[:, 0]], b["tri_ids"][got[:, 1]]], 1), want)
    assert (g["all_num_pairs"] > 0).sum() > 50 and (g["all_num_pairs"] == 0).sum() > 50 and g["first_num_pairs"].max() == 1
    assert g["first_num_bv_tests"].sum() < g["all_num_bv_tests"].sum()


@pytest.mark.parametrize("case,ma,mb", COLLIDE_CASES)
@pytest.mark.parametrize("tag", ["exact", "approx"])
def test_port_collide_distance_matches_golden(case, ma, mb, tag, golden, bvhs):
    """C2A_Collide, C2A_DistanceResult overload (C2A_PQP.cpp:1060-1280): the distance walk behind the box-overlap gate."""
    g = golden(f"ref_collide_{case}")
    rel, ab = g[f"dist_{tag}_err"]
    out = oracle.port().collide_distance(bvhs(ma), bvhs(mb), g["poses24"], g["seed_a"], g["seed_b"], rel, ab)
    for k in out.dtype.names:
        assert np.array_equal(out[k], g[f"dist_{tag}_{k}"]), (case, tag, k)
    assert 20 < (g[f"dist_{tag}_num_tri_tests"] > 0).sum() < len(out)   # gate open for some queries, shut at the root for others


def test_port_box_and_triangle_overlap_vs_shim():
    """The port's obb_disjoint / TriContact against the stand-ins the compiled reference links (oracle/pqp_shim), through
    the reference's own C2A_Collide on two-triangle models: every random placement gives the same verdict and counters."""
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref (built where /root/reference exists)")
    from c2a_b200 import api
    rng = np.random.default_rng(3)
    R, P = oracle.ref(), oracle.port()
    for trial in range(40):
        ta = rng.normal(size=(2, 9)); tb = rng.normal(size=(2, 9))
        ba, bb = api.build_bvh(ta), api.build_bvh(tb)
        ra, rb = R.model(ta), R.model(tb)
        q = rng.normal(size=(8, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
        poses = np.zeros((8, 24)); poses[:, 0:9] = np.eye(3).reshape(9)
        w, x, y, z = q.T
        poses[:, 12:21] = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
                                    2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)
        poses[:, 21:24] = rng.normal(scale=0.7, size=(8, 3))
        n0, p0, v0, t0 = R.collide(ra, rb, poses); n1, p1, v1, t1 = P.collide(ba, bb, poses)
        assert np.array_equal(n0, n1) and np.array_equal(v0, v1) and np.array_equal(t0, t1)
        for u, w_ in zip(p0, p1):
            assert np.array_equal(u, np.stack([ba["tri_ids"][w_[:, 0]], bb["tri_ids"][w_[:, 1]]], 1))


def test_speculative_step_split_is_exact(golden, bvhs):
    """Round-2 design study (oracle/c2a_oracle.cpp, orc_solve_spec): CA steps split into subtrees run under a guessed entry
    distance + validity interval and stitched in the reference's order reproduce the sequential result bit for bit."""
    g = golden("ref_knot_128x16")
    idx = np.argsort(-g["num_bv_tests"])[:12]
    for K in (3, 7):
        res, st = oracle.port().solve_spec(bvhs("knot_128x16"), bvhs("knot_128x16"), g["poses"][idx], K)
        for k in ("collisionfree", "numCA", "num_bv_tests", "num_tri_tests", "toc", "distance", "mint", "pose_toc"):
            assert np.array_equal(res[k], g[k][idx]), (K, k)
        assert st["steps"] > 50 and 0 < st["valid"] < st["reached"]  # both the accepted and the re-run path are exercised


@needs_ref
def test_port_matches_ref_on_degenerate_motions(bvhs):
    """No motion, rotation in place, contact at the start pose, vanishing motions: the port against the reference's
    object code (both branches; NaN-tolerant comparison, none occurs)."""
    from c2a_b200 import meshes
    tris, vi = meshes.torus_knot(128, 16)
    R = oracle.ref()
    m = R.model(tris, vi)
    poses = workloads.degenerate_batch(radius=workloads.KNOT_RADIUS)
    ref = R.solve_batch(m, m, poses, threads=1)
    out = oracle.port().solve_batch(bvhs("knot_128x16"), bvhs("knot_128x16"), poses, threads=1)
    for k in ("collisionfree", "numCA", "num_bv_tests", "num_tri_tests", "toc", "distance", "mint", "pose_toc"):
        assert np.array_equal(out[k], ref[k], equal_nan=True), k
    assert (ref["numCA"] == 0).sum() == 50 and (ref["collisionfree"] == 0).sum() > 30


def test_replay_over_previous_visit_list_is_exact(golden, bvhs):
    """Round-2 design study (orc_solve_replay): CA steps walked over the previous step's visit list, with the records
    re-evaluated for the current poses beforehand and misses evaluated on the spot, give the sequential result."""
    for case, model in (("ref_knot_128x16", "knot_128x16"), ("ref_knot_128x16_grazing_tol1e-06", "knot_128x16")):
        g = golden(case)
        idx = np.argsort(-g["num_bv_tests"])[:25]
        res, st = oracle.port().solve_replay(bvhs(model), bvhs(model), g["poses"][idx], tol_d=float(g["tol_d"]), tol_t=float(g["tol_t"]))
        for k in ("collisionfree", "numCA", "num_bv_tests", "num_tri_tests", "toc", "distance", "mint", "pose_toc"):
            assert np.array_equal(res[k], g[k][idx]), (case, k)
        upd = (res["p1"] != 0).any(1)
        assert np.array_equal(np.concatenate([res["p1"], res["p2"]], 1)[upd], g["p1p2"][idx][upd])
        assert st["visits"] == st["hits"] + st["misses"] and st["hits"] > st["misses"] > 0


def test_wide_traversal_is_exact(golden, bvhs):
    """The round-2 device algorithm stated on the CPU (orc_solve_wide, the checker of c2a_wide.cuh): exact-mode CA steps
    as a depth-first traversal that pops W node pairs per round, with keys, the M rule, in-order event resolution and
    the fold, give the reference's results bit for bit -- for any window, with and without batched leaf passes, on
    the bunny as well (seeded, non-default tolerances)."""
    P = oracle.port()
    cases = (("ref_knot_128x16", "knot_128x16", "knot_128x16", 40), ("ref_knot_128x16_grazing_tol1e-06", "knot_128x16", "knot_128x16", 30),
             ("ref_bunny_vs_knot_seeded", "bunny", "knot_512x32", 25), ("ref_bunny_grazing_tol1e-06", "bunny", "bunny", 20))
    for case, ma, mb, k in cases:
        g = golden(case)
        idx = np.concatenate([np.argsort(-g["num_bv_tests"])[:k], np.arange(k)])
        for W, lb in ((1, 0), (16, 32), (5, 7), (64, 0)):
            res, st = P.solve_wide(bvhs(ma), bvhs(mb), g["poses"][idx], None if "seed_a" not in g else g["seed_a"][idx],
                                   None if "seed_b" not in g else g["seed_b"][idx], tol_d=float(g["tol_d"]), tol_t=float(g["tol_t"]),
                                   window=W, leaf_batch=lb)
            for f in ("collisionfree", "numCA", "num_bv_tests", "num_tri_tests", "toc", "distance", "mint", "pose_toc"):
                assert np.array_equal(res[f], g[f][idx]), (case, W, f)
            upd = (res["p1"] != 0).any(1)
            assert np.array_equal(np.concatenate([res["p1"], res["p2"]], 1)[upd], g["p1p2"][idx][upd]), (case, W)
            if "last_tri" in g:
                assert np.array_equal(np.stack([res["last_tri_a"], res["last_tri_b"]], 1), g["last_tri"][idx]), (case, W)
            assert st["steps"] > 0 and st["closure_fail"] == 0
            assert st["wide_tests_visited"] <= st["wide_tests"]
            if W == 1:
                assert st["wide_tests_visited"] == st["wide_tests"] and st["redo"] == 0  # the sequential walk itself


def test_visit_sequences_account_for_the_counters(golden, bvhs):
    """Round-2 design study hook (orc_solve_visits): every visited node pair is an expansion (2 BV tests) or a leaf
    pair (1 triangle test), so the logged sequences must add up to the query's counters."""
    g = golden("ref_knot_128x16")
    i = int(np.argmax(g["num_bv_tests"]))
    res, steps = oracle.port().solve_visits(bvhs("knot_128x16"), bvhs("knot_128x16"), g["poses"][i])
    assert res["toc"] == g["toc"][i] and len(steps) == g["numCA"][i]
    assert sum(len(s) for s in steps) == g["num_bv_tests"][i] // 2 + g["num_tri_tests"][i]
    assert all(s[0] == 0 for s in steps)  # every step starts at the root pair


def test_golden_fixtures_are_sane(golden):
    """Verdict semantics: toc == 0 for free queries, hits end within tolerance."""
    for case, _, _ in GOLDEN_CASES:
        g = golden(case)
        free = g["collisionfree"] == 1
        assert (g["toc"][free] == 0).all()
        assert ((g["toc"] >= 0) & (g["toc"] < 1)).all()
        assert (g["numCA"] >= 1).all() and (g["numCA"] <= 152).all()
        assert free.sum() > 0 and (~free).sum() > 0


def contacts_equal(a, c):
    """Two contact records; FeatureID entries beyond the feature's arity are undefined in the reference."""
    na, nb = int(a["type_a"]), int(a["type_b"])
    return (a["type_a"] == c["type_a"] and a["type_b"] == c["type_b"] and a["tri_a"] == c["tri_a"] and a["tri_b"] == c["tri_b"]
            and a["dist"] == c["dist"] and np.array_equal(a["pa"], c["pa"]) and np.array_equal(a["pb"], c["pb"])
            and np.array_equal(a["fid_a"][:na], c["fid_a"][:na]) and np.array_equal(a["fid_b"][:nb], c["fid_b"][:nb]))


def golden_contacts(golden):
    g = golden("ref_contacts_knot_128x16")
    names = [k for k in g if k != "num_contact"]
    recs = np.zeros(len(g["dist"]), dtype=oracle.CONTACT_DTYPE)
    for k in names:
        recs[k] = g[k]
    offs = np.concatenate([[0], np.cumsum(g["num_contact"])])
    return g["num_contact"], [recs[offs[i]:offs[i + 1]] for i in range(len(g["num_contact"]))]


def test_port_contact_pass_matches_golden(golden, bvhs):
    """Contact pass of C2A_Solve: port vs the reference's exported ContactF lists (list order = reversed visiting order)."""
    from c2a_b200 import api, meshes
    counts, lists = golden_contacts(golden)
    g = golden("ref_knot_128x16")
    tris, vidx = meshes.torus_knot(128, 16)
    b = api.build_bvh(tris, vidx)
    P = oracle.port()
    assert counts.sum() > 100
    for i in range(len(counts)):
        if g["collisionfree"][i]:
            assert counts[i] == 0
            continue
        thr = 2 * g["distance"][i] + 0.001
        n, recs = P.contacts(b, b, g["pose_toc"][i][:12], g["pose_toc"][i][12:], thr, b["tri_vidx"], b["tri_vidx"])
        assert n == counts[i], i
        for a, c in zip(lists[i], recs[::-1]):
            assert contacts_equal(a, c), i
