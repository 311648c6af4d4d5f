import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build the product library (nvcc cross-compiles without a GPU) and the CPU oracle once."""
    from c2a_b200 import build as c2a_build
    import oracle
    c2a_build.build()
    oracle.build_oracle()


def load_golden(name):
    """A committed fixture as a dict.  The large ones are stored compact (tests/golden/make_golden.py save_compact):
    the reference's outputs plus a recipe for the poses, rebuilt here and checked against the stored sha256."""
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    if "gen" in d:
        import hashlib
        from c2a_b200 import workloads
        gen = json.loads(str(d["gen"]))
        if gen["fn"] == "approach_batch":
            poses = workloads.approach_batch(gen["n"], gen["seed"])
        elif gen["fn"] == "grazing_of":
            src = load_golden(gen["src"])
            g = np.load(os.path.join(GOLDEN, gen["poses"] + ".npz"))
            poses = src["poses"][g["src"]].copy()
            poses[:, 12:24] = g["end1"]
        else:
            raise ValueError(gen["fn"])
        poses = np.ascontiguousarray(poses)
        assert hashlib.sha256(poses.tobytes()).hexdigest() == str(d["pose_sha256"]), f"{name}: the pose generator no longer reproduces the fixture's inputs"
        d["poses"] = poses
        for k in ("collisionfree", "numCA"):
            d[k] = d[k].astype(np.int32)
    return d


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def bunny_tris():
    m = np.load(os.path.join(GOLDEN, "bunny_mesh.npz"))
    return m["verts"][m["vidx"]].reshape(-1, 9).copy()


@pytest.fixture(scope="session")
def bvhs(bunny_tris):
    """Hierarchies built by the PRODUCT's host builder (bit-identical to the reference's, see
    test_host_side.py::test_bvh_matches_reference_digest), cached per session."""
    from c2a_b200 import api, meshes
    cache = {}

    def get(name):
        if name not in cache:
            if name == "bunny":
                cache[name] = api.build_bvh(bunny_tris)
            else:
                nu, nv = (int(x) for x in name.split("_")[1].split("x"))
                cache[name] = api.build_bvh(meshes.torus_knot(nu, nv)[0])
        return cache[name]
    return get


@pytest.fixture(scope="session")
def models(bvhs):
    """Models resident on cuda:0 (GPU tests only), cached per session."""
    from c2a_b200 import api
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = api.Model(bvhs(name), 0)
        return cache[name]
    return get


@pytest.fixture(scope="session")
def bvh_digest():
    with open(os.path.join(GOLDEN, "bvh_digest.json")) as f:
        return json.load(f)


GOLDEN_CASES = [  # (fixture, model A, model B)
    ("ref_demo_bunny", "bunny", "bunny"),
    ("ref_bunny_approach", "bunny", "bunny"),
    ("ref_knot_128x16", "knot_128x16", "knot_128x16"),
    ("ref_knot_512x32", "knot_512x32", "knot_512x32"),
    ("ref_knot_1024x32", "knot_1024x32", "knot_1024x32"),
    ("ref_bunny_vs_knot_seeded", "bunny", "knot_512x32"),
    ("ref_knot_128x16_grazing_tol0.001", "knot_128x16", "knot_128x16"),   # config 5 flavour: grazing end poses,
    ("ref_knot_128x16_grazing_tol1e-06", "knot_128x16", "knot_128x16"),   # tolerance_t swept (verdicts flip with it)
    ("ref_bunny_grazing_tol0.001", "bunny", "bunny"),                     # config 5 on the bunny, as BASELINE names it
    ("ref_bunny_grazing_tol1e-06", "bunny", "bunny"),
    ("ref_knot_128x16_carry", "knot_128x16", "knot_128x16"),              # "demo mode": seeds carried through
    ("ref_demo_bunny_carry", "bunny", "bunny"),                           # o->last_tri (quirk Q4), 303 frames twice
    ("ref_knot_512x32_heavy", "knot_512x32", "knot_512x32"),              # the 1M bench batch's heaviest: all 81 queries at the
                                                                          # reference's 150-iteration cap + the 200 with most BV tests
    ("ref_bunny_approach_10k", "bunny", "bunny"),                         # config 2 at its full 10 000 (compact fixture)
    ("ref_bunny_grazing_2k_tol0.001", "bunny", "bunny"),                  # config 5 at its full 2 000 x four tolerances
    ("ref_bunny_grazing_2k_tol0.0001", "bunny", "bunny"),
    ("ref_bunny_grazing_2k_tol1e-05", "bunny", "bunny"),
    ("ref_bunny_grazing_2k_tol1e-06", "bunny", "bunny"),
]
