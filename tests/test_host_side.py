"""CPU tests of the product's host side: the BVH builder, the host half of the motion model, the libm
mirror, the C-ABI surface and the sharding logic.  No compute call needs a GPU here."""
import ctypes as C
import hashlib
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle
from c2a_b200 import api, meshes, sharding, workloads
from conftest import ROOT


def test_library_exports_every_declared_symbol():
    lib = api.lib()
    names = set()
    for h in ("c2a_b200.h", "c2a_b200_testing.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(c2a_b200_\w+)\s*\(", src))
    assert len(names) >= 15
    for n in sorted(names):
        assert hasattr(lib, n), n


def test_bvh_matches_reference_digest(bvhs, bvh_digest):
    """The product's builder reproduces the tree the reference's builder made (sha256 of every array,
    fixture written by tests/golden/make_golden.py from the reference's own object code)."""
    for name, dg in bvh_digest.items():
        b = bvhs(name)
        for k, want in dg.items():
            assert hashlib.sha256(np.ascontiguousarray(b[k]).tobytes()).hexdigest() == want, (name, k)


def test_bvh_structure(bvhs):
    b = bvhs("knot_128x16")
    n = len(b["tris"])
    assert len(b["first_child"]) == 2 * n - 1
    leaves = b["first_child"] < 0
    assert leaves.sum() == n
    assert sorted(-b["first_child"][leaves] - 1) == list(range(n))       # every triangle in exactly one leaf
    assert sorted(b["tri_ids"]) == list(range(n))                        # a permutation of the input
    orig = meshes.torus_knot(128, 16)[0]
    assert np.array_equal(b["tris"], orig[b["tri_ids"]])
    assert (b["l"] >= 0).all() and (b["r"] >= 0).all()
    assert b["ang_radius"][0] == np.sqrt((orig.reshape(-1, 3) ** 2).sum(1)).max()


def test_bvh_tiny_models():
    one = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], dtype=np.float64)
    b = api.build_bvh(one)
    assert list(b["first_child"]) == [-1] and b["depth"] == 0
    two = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0], [5, 0, 0, 6, 0, 0, 5, 1, 0]], dtype=np.float64)
    b = api.build_bvh(two)
    assert list(b["first_child"]) == [1, -1, -2] and b["depth"] == 1


def test_host_libm_mirror_matches_libm():
    """c2a_libm.cuh (host twin of the device code) == the C library's sin/cos, bit for bit."""
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-3.3, 3.3, 200000), rng.uniform(-0.13, 0.13, 50000), rng.uniform(-1e5, 1e5, 50000),
                        np.array([0.0, -0.0, 0.126, 1 / 128, 0.85546875, 2.426265, math.pi, 2 ** -26, 2 ** -27, 1e-300])])
    s = np.zeros_like(x); c = np.zeros_like(x)
    api._check(api.lib().c2a_b200_host_sincos(x.ctypes.data_as(C.c_void_p), C.c_int64(len(x)),
                                              s.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p)))
    ms = np.array([math.sin(v) for v in x]); mc = np.array([math.cos(v) for v in x])
    assert np.array_equal(s, ms) and np.array_equal(c, mc)


def test_motion_records_match_oracle():
    """Host half of the motion model against the oracle port's orc_motion_init (and the reference's
    CInterpMotion_Linear when oracle/_ref is present)."""
    poses = np.concatenate([workloads.approach_batch(3000, 17), np.load(os.path.join(ROOT, "tests/golden/demo_poses.npy"))])
    rec = api.motions_from_poses(poses, threads=3)

    class M(C.Structure):
        _fields_ = [("Rs", C.c_double * 9), ("Ts", C.c_double * 3), ("Re", C.c_double * 9), ("Te", C.c_double * 3),
                    ("cv", C.c_double * 3), ("axis", C.c_double * 3), ("w", C.c_double), ("Rc", C.c_double * 9), ("Tc", C.c_double * 3)]
    P = oracle.port().lib
    for i in range(0, len(poses), 5):
        for o in (0, 1):
            p = np.ascontiguousarray(poses[i, 24 * o:24 * o + 24])
            m = M()
            P.orc_motion_init(C.byref(m), p[0:].ctypes.data_as(C.c_void_p), p[9:].ctypes.data_as(C.c_void_p),
                              p[12:].ctypes.data_as(C.c_void_p), p[21:].ctypes.data_as(C.c_void_p))
            r = rec[i, 24 * o:24 * o + 24]
            assert np.array_equal(r[0:12], p[0:12])
            assert list(r[12:15]) == list(m.cv) and list(r[15:18]) == list(m.axis) and r[18] == m.w
    if oracle.have_ref():
        R = oracle.ref()
        for i in range(0, len(poses), 37):
            pr = R.motion_probe(poses[i, :24], 0.5, 1.0, np.array([1.0, 0, 0]))
            assert np.array_equal(rec[i, 12:19], pr[0:7])


def test_abi_argument_errors():
    L = api.lib()
    assert L.c2a_b200_model_upload(None, 0, None) == -1
    assert L.c2a_b200_bvh_build(None, 3, None) == -1
    assert L.c2a_b200_motions_from_poses(None, C.c_int64(5), None, 1) == -1
    assert L.c2a_b200_solve_batch(None, None, None, None, None, C.c_int64(1), C.c_double(1e-4), C.c_double(1e-4), None) == -1
    assert b"NULL" in L.c2a_b200_last_error()


def test_no_cpu_fallback_without_gpu(bvhs):
    """Without a CUDA device the product fails loudly instead of computing on the CPU."""
    try:
        n = api.device_count()
    except api.C2AError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.C2AError):
        api.Model(bvhs("knot_128x16"), 0)


def test_shard_indices_partition():
    rng = np.random.default_rng(3)
    for n in (1, 7, 8, 1003):
        order = rng.permutation(n)
        for w in (1, 2, 3, 8):
            parts = [sharding.shard_indices(order, r, w) for r in range(w)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            assert all(np.array_equal(p, order[r::w]) for r, p in enumerate(parts))


def test_shard_bounds_cover():
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch.distributed as dist
import oracle
from c2a_b200 import api, meshes, sharding, workloads
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
bvh = api.build_bvh(meshes.torus_knot(64, 8)[0])
poses = workloads.approach_batch(41, 5, radius=workloads.KNOT_RADIUS)
# the split bench.py makes: every world-th entry of the cost-sorted claim order (host half of the product, no GPU needed)
order = np.random.default_rng(0).permutation(len(poses))  # (the real order comes from api.schedule_order, which needs device models)
idx = sharding.shard_indices(order, rank, world)
# stand-in solver for the CPU test: the oracle port plays the GPU's role (the checker, not the product)
r = oracle.port().solve_batch(bvh, bvh, poses[idx])
local = {"toc": r["toc"].copy(), "collisionfree": r["collisionfree"].copy(), "num_ca": r["numCA"].copy(), "pose_toc": r["pose_toc"].copy()}
full = sharding.gather_results(local, idx, len(poses), rank, world, dist)
if rank == 0:
    ref = oracle.port().solve_batch(bvh, bvh, poses)
    assert np.array_equal(full["toc"], ref["toc"]) and np.array_equal(full["collisionfree"], ref["collisionfree"])
    assert np.array_equal(full["num_ca"], ref["numCA"]) and np.array_equal(full["pose_toc"], ref["pose_toc"])
    print("GATHER_OK")
else:
    assert full is None
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_shard_and_gather(tmp_path):
    """world_size-2 gloo run of the N>1 host logic: interleaved shards of the claim order, per-rank solve, gather +
    scatter to batch order on rank 0."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


def test_dropin_header_compiles_standalone(tmp_path):
    """include/C2A/C2A.h and the alias headers a reference-style program includes are self-contained C++."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "PQP.h"\n#include "C2A/C2A.h"\n#include "C2A/LinearMath.h"\n#include "C2A/InterpMotion.h"\n'
                   '#include "C2A/C2A_Internal.h"\n#include "c2a_b200.h"\n#include "c2a_b200_testing.h"\nint main() { C2A_TimeOfContactResult r; (void)r; return 0; }\n')
    out = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "include"), str(src)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_scene_workload_and_broadphase_statement():
    """Config 4 plumbing on the CPU: scene generator, query assembly and the numpy broadphase statement."""
    from c2a_b200 import workloads
    # two unit spheres passing each other: closest approach 1.0 at t = 0.5
    c0 = np.array([[-5.0, 0.5, 0.0], [5.0, -0.5, 0.0]]); c1 = np.array([[5.0, 0.5, 0.0], [-5.0, -0.5, 0.0]])
    p, gap = workloads.broadphase_reference(c0, c1, np.array([0.4, 0.4]))
    assert len(p) == 0
    p, gap = workloads.broadphase_reference(c0, c1, np.array([0.5, 0.6]))
    assert p.tolist() == [[0, 1]] and abs(gap[0] - 0.1) < 1e-12
    p, _ = workloads.broadphase_reference(c0, c0, np.array([0.5, 0.6]))   # static and 10.05 apart
    assert len(p) == 0
    sc = workloads.scene(300, 4, [131.4, 198.0])
    r = np.array([131.4, 198.0])[sc["model"]]
    pairs, _ = workloads.broadphase_reference(sc["begin"][:, 9:], sc["end"][:, 9:], r)
    assert 3.0 < 2.0 * len(pairs) / 300 < 14.0            # "about 8 neighbours"
    poses, ma, mb = workloads.scene_queries(sc, pairs)
    assert poses.shape == (len(pairs), 48) and np.array_equal(ma, sc["model"][pairs[:, 0]])
    assert np.array_equal(poses[:, 24:36], sc["begin"][pairs[:, 1]])
    R = poses[:, :9].reshape(-1, 3, 3)
    assert np.allclose(np.einsum("nij,nkj->nik", R, R), np.eye(3), atol=1e-12)
