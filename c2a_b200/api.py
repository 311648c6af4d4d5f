"""ctypes bindings of the C ABI in ``include/c2a_b200.h`` (``c2a_b200/csrc/libc2a_b200.so``).

Plumbing only: the compute is the CUDA library.  There is no CPU fallback -- if the library is
missing or no CUDA device is present, calls raise."""
import ctypes as C
import os

import numpy as np

# C2A_B200_LIB: development override (A/B timing of two builds of the same library in one GPU session)
_LIB_PATH = os.environ.get("C2A_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libc2a_b200.so")
_lib = None

RESULT_FIELDS = (("status", np.int32, ()), ("collisionfree", np.int32, ()), ("num_ca", np.int32, ()),
                 ("num_bv_tests", np.int32, ()), ("num_tri_tests", np.int32, ()), ("toc", np.float64, ()),
                 ("distance", np.float64, ()), ("mint", np.float64, ()), ("p1p2", np.float64, (6,)),
                 ("pose_toc", np.float64, (24,)))


class C2AError(RuntimeError):
    pass


class Bvh(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_tris", C.c_int32),
                ("R", C.c_void_p), ("Tr", C.c_void_p), ("l", C.c_void_p), ("r", C.c_void_p),
                ("R_loc", C.c_void_p), ("ang_radius", C.c_void_p), ("first_child", C.c_void_p),
                ("tris", C.c_void_p), ("tri_vidx", C.c_void_p), ("obb_d", C.c_void_p), ("obb_To", C.c_void_p)]


# mirrors struct c2a_b200_contact
CONTACT_DTYPE = np.dtype([("type_a", np.int32), ("type_b", np.int32), ("fid_a", np.int32, 3), ("fid_b", np.int32, 3),
                          ("tri_a", np.int32), ("tri_b", np.int32), ("pa", np.float64, 3), ("pb", np.float64, 3),
                          ("dist", np.float64)], align=True)


class Results(C.Structure):
    _fields_ = [(name, C.c_void_p) for name, _, _ in RESULT_FIELDS] + [("num_contact", C.c_void_p), ("contacts", C.c_void_p),
                                                                       ("max_contacts", C.c_int32), ("last_tri", C.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise C2AError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        L = C.CDLL(_LIB_PATH)
        L.c2a_b200_last_error.restype = C.c_char_p
        L.c2a_b200_launch_count.restype = C.c_int64
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise C2AError(f"c2a_b200 error {rc}: {lib().c2a_b200_last_error().decode()}")


def device_count():
    n = C.c_int32(0)
    _check(lib().c2a_b200_device_count(C.byref(n)))
    return n.value


def launch_count():
    return int(lib().c2a_b200_launch_count())


def build_bvh(tris9, vidx=None):
    """Host-side RSS BVH build (the product's own builder, c2a_host_model.cpp).  tris9: [n,9] float64;
    vidx: optional [n,3] int32 vertex indices per triangle (labels of contact features).
    Returns a dict of numpy arrays with the keys of ``struct c2a_b200_bvh`` plus ``tri_ids``/``depth``."""
    tris9 = np.ascontiguousarray(tris9, dtype=np.float64).reshape(-1, 9)
    n = tris9.shape[0]
    h = C.c_void_p()
    vi = None if vidx is None else np.ascontiguousarray(vidx, dtype=np.int32).reshape(n, 3)
    _check(lib().c2a_b200_bvh_build_indexed(tris9.ctypes.data_as(C.c_void_p),
                                            vi.ctypes.data_as(C.c_void_p) if vi is not None else None, C.c_int32(n), C.byref(h)))
    try:
        v = Bvh()
        ids = C.c_void_p()
        depth = C.c_int32()
        _check(lib().c2a_b200_bvh_view(h, C.byref(v), C.byref(ids), C.byref(depth)))
        nn = v.n_nodes

        def arr(ptr, count, ctype, dtype):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).astype(dtype, copy=True)
        out = {"R": arr(v.R, 9 * nn, C.c_double, np.float64).reshape(nn, 9),
               "Tr": arr(v.Tr, 3 * nn, C.c_double, np.float64).reshape(nn, 3),
               "l": arr(v.l, 2 * nn, C.c_double, np.float64).reshape(nn, 2),
               "r": arr(v.r, nn, C.c_double, np.float64),
               "R_loc": arr(v.R_loc, 9 * nn, C.c_double, np.float64).reshape(nn, 9),
               "ang_radius": arr(v.ang_radius, nn, C.c_double, np.float64),
               "first_child": arr(v.first_child, nn, C.c_int32, np.int32),
               "tris": arr(v.tris, 9 * n, C.c_double, np.float64).reshape(n, 9),
               "tri_ids": arr(ids, n, C.c_int32, np.int32),
               "obb_d": arr(v.obb_d, 3 * nn, C.c_double, np.float64).reshape(nn, 3),
               "obb_To": arr(v.obb_To, 3 * nn, C.c_double, np.float64).reshape(nn, 3),
               "depth": depth.value}
        if v.tri_vidx:
            out["tri_vidx"] = arr(v.tri_vidx, 3 * n, C.c_int32, np.int32).reshape(n, 3)
    finally:
        lib().c2a_b200_bvh_free(h)
    return out


class Model:
    """A C2A model resident on one GPU.  ``bvh`` is a dict of contiguous numpy arrays with the keys of
    ``struct c2a_b200_bvh`` (R, Tr, l, r, R_loc, ang_radius float64; first_child int32; tris float64; optionally
    tri_vidx int32 and obb_d / obb_To float64)."""

    def __init__(self, bvh, device=0):
        s = Bvh()
        s.n_nodes = int(bvh["first_child"].shape[0])
        s.n_tris = int(bvh["tris"].shape[0])
        keep = []
        for k in ("R", "Tr", "l", "r", "R_loc", "ang_radius", "tris"):
            a = np.ascontiguousarray(bvh[k], dtype=np.float64)
            keep.append(a)
            setattr(s, k, a.ctypes.data)
        fc = np.ascontiguousarray(bvh["first_child"], dtype=np.int32)
        keep.append(fc)
        s.first_child = fc.ctypes.data
        if bvh.get("tri_vidx") is not None:
            tv = np.ascontiguousarray(bvh["tri_vidx"], dtype=np.int32)
            keep.append(tv)
            s.tri_vidx = tv.ctypes.data
        if bvh.get("obb_d") is not None and bvh.get("obb_To") is not None:   # (C2A_Collide only)
            for k in ("obb_d", "obb_To"):
                a = np.ascontiguousarray(bvh[k], dtype=np.float64)
                keep.append(a)
                setattr(s, k, a.ctypes.data)
        h = C.c_void_p()
        _check(lib().c2a_b200_model_upload(C.byref(s), C.c_int32(device), C.byref(h)))
        self.h = h
        self.device = device
        self.n_nodes, self.n_tris = s.n_nodes, s.n_tris

    def info(self):
        d, nn, nt, dep = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _check(lib().c2a_b200_model_info(self.h, C.byref(d), C.byref(nn), C.byref(nt), C.byref(dep)))
        return {"device": d.value, "n_nodes": nn.value, "n_tris": nt.value, "depth": dep.value}

    def free(self):
        if self.h:
            lib().c2a_b200_model_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def solve_batch(model_a, model_b, poses, seed_a=None, seed_b=None, tol_d=1e-4, tol_t=1e-4, fields=None, max_contacts=0):
    """Host-buffer entry (H2D + kernel + D2H inside the call).  Returns a dict of numpy arrays.
    max_contacts > 0 also runs the contact pass of C2A_Solve: ``num_contact`` [n] and ``contacts``
    [n, max_contacts] (CONTACT_DTYPE, visiting order)."""
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 48)
    n = poses.shape[0]
    res = Results()
    out = {}
    for name, dt, shape in RESULT_FIELDS:
        if fields is not None and name not in fields:
            continue
        a = np.zeros((n,) + shape, dtype=dt)
        out[name] = a
        setattr(res, name, a.ctypes.data)
    if fields is None or "last_tri" in fields:
        out["last_tri"] = np.full((n, 2), -1, dtype=np.int32)
        res.last_tri = out["last_tri"].ctypes.data
    if max_contacts > 0:
        out["num_contact"] = np.zeros(n, dtype=np.int32)
        out["contacts"] = np.zeros((n, max_contacts), dtype=CONTACT_DTYPE)
        res.num_contact = out["num_contact"].ctypes.data
        res.contacts = out["contacts"].ctypes.data
        res.max_contacts = max_contacts
    sa = None if seed_a is None else np.ascontiguousarray(seed_a, dtype=np.int32)
    sb = None if seed_b is None else np.ascontiguousarray(seed_b, dtype=np.int32)
    _check(lib().c2a_b200_solve_batch(model_a.h, model_b.h, poses.ctypes.data_as(C.c_void_p),
                                      sa.ctypes.data_as(C.c_void_p) if sa is not None else None,
                                      sb.ctypes.data_as(C.c_void_p) if sb is not None else None,
                                      C.c_int64(n), C.c_double(tol_d), C.c_double(tol_t), C.byref(res)))
    return out


def solve_batch_multi(models_a, models_b, poses, seed_a=None, seed_b=None, tol_d=1e-4, tol_t=1e-4, fields=None):
    """One batch sharded over several GPUs of the box: ``models_a[d]`` / ``models_b[d]`` are replicas of the two
    models on distinct devices.  Same results as ``solve_batch`` for any number of devices."""
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 48)
    n = poses.shape[0]
    assert len(models_a) == len(models_b) and len(models_a) >= 1
    res = Results()
    out = {}
    for name, dt, shape in RESULT_FIELDS:
        if fields is not None and name not in fields:
            continue
        a = np.zeros((n,) + shape, dtype=dt)
        out[name] = a
        setattr(res, name, a.ctypes.data)
    if fields is None or "last_tri" in fields:
        out["last_tri"] = np.full((n, 2), -1, dtype=np.int32)
        res.last_tri = out["last_tri"].ctypes.data
    sa = None if seed_a is None else np.ascontiguousarray(seed_a, dtype=np.int32)
    sb = None if seed_b is None else np.ascontiguousarray(seed_b, dtype=np.int32)
    ha = (C.c_void_p * len(models_a))(*[m.h for m in models_a])
    hb = (C.c_void_p * len(models_b))(*[m.h for m in models_b])
    _check(lib().c2a_b200_solve_batch_multi(ha, hb, C.c_int32(len(models_a)), poses.ctypes.data_as(C.c_void_p),
                                            sa.ctypes.data_as(C.c_void_p) if sa is not None else None,
                                            sb.ctypes.data_as(C.c_void_p) if sb is not None else None,
                                            C.c_int64(n), C.c_double(tol_d), C.c_double(tol_t), C.byref(res)))
    return out


def solve_pairs(models, model_a, model_b, poses, seed_a=None, seed_b=None, tol_d=1e-4, tol_t=1e-4, fields=None):
    """Heterogeneous batch (per-query model handles): ``models`` is a list of Model on one device,
    ``model_a`` / ``model_b`` [n] index into it.  Returns the same dict as ``solve_batch``."""
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 48)
    n = poses.shape[0]
    ma = np.ascontiguousarray(model_a, dtype=np.int32)
    mb = np.ascontiguousarray(model_b, dtype=np.int32)
    assert ma.shape == (n,) and mb.shape == (n,)
    res = Results()
    out = {}
    for name, dt, shape in RESULT_FIELDS:
        if fields is not None and name not in fields:
            continue
        a = np.zeros((n,) + shape, dtype=dt)
        out[name] = a
        setattr(res, name, a.ctypes.data)
    if fields is None or "last_tri" in fields:
        out["last_tri"] = np.full((n, 2), -1, dtype=np.int32)
        res.last_tri = out["last_tri"].ctypes.data
    sa = None if seed_a is None else np.ascontiguousarray(seed_a, dtype=np.int32)
    sb = None if seed_b is None else np.ascontiguousarray(seed_b, dtype=np.int32)
    handles = (C.c_void_p * len(models))(*[m.h for m in models])
    _check(lib().c2a_b200_solve_pairs(handles, C.c_int32(len(models)), ma.ctypes.data_as(C.c_void_p), mb.ctypes.data_as(C.c_void_p),
                                      poses.ctypes.data_as(C.c_void_p),
                                      sa.ctypes.data_as(C.c_void_p) if sa is not None else None,
                                      sb.ctypes.data_as(C.c_void_p) if sb is not None else None,
                                      C.c_int64(n), C.c_double(tol_d), C.c_double(tol_t), C.byref(res)))
    return out


def broadphase(c0, c1, radius, margin=0.0, device=0, max_pairs=None):
    """Swept-sphere candidate pairs of a scene: c0, c1 [n,3] (begin / end position of each instance's model
    origin), radius [n] (max |vertex| of its model).  Returns int32 [m,2] with i < j, sorted."""
    c0 = np.ascontiguousarray(c0, dtype=np.float64).reshape(-1, 3)
    c1 = np.ascontiguousarray(c1, dtype=np.float64).reshape(-1, 3)
    r = np.ascontiguousarray(radius, dtype=np.float64).reshape(-1)
    n = c0.shape[0]
    assert c1.shape == (n, 3) and r.shape == (n,)
    cap = int(max_pairs) if max_pairs is not None else max(1024, 32 * n)
    while True:
        pairs = np.empty((cap, 2), dtype=np.int32)
        found = C.c_int64(0)
        _check(lib().c2a_b200_broadphase(c0.ctypes.data_as(C.c_void_p), c1.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p),
                                         C.c_int32(n), C.c_double(margin), C.c_int32(device), pairs.ctypes.data_as(C.c_void_p),
                                         C.c_int64(cap), C.byref(found)))
        if found.value <= cap or max_pairs is not None:
            pairs = pairs[:min(found.value, cap)]
            return pairs[np.lexsort((pairs[:, 1], pairs[:, 0]))]
        cap = int(found.value)


def contacts_batch(model_a, model_b, poses24, threshold, max_contacts=64):
    """Batched C2A_QueryContact: poses24 [n,24] (pose of A, pose of B), threshold [n]."""
    poses24 = np.ascontiguousarray(poses24, dtype=np.float64).reshape(-1, 24)
    n = poses24.shape[0]
    thr = np.ascontiguousarray(np.broadcast_to(np.asarray(threshold, dtype=np.float64), (n,)))
    num = np.zeros(n, dtype=np.int32)
    recs = np.zeros((n, max_contacts), dtype=CONTACT_DTYPE)
    _check(lib().c2a_b200_contacts_batch(model_a.h, model_b.h, poses24.ctypes.data_as(C.c_void_p), thr.ctypes.data_as(C.c_void_p),
                                         C.c_int64(n), C.c_int32(max_contacts), num.ctypes.data_as(C.c_void_p),
                                         recs.ctypes.data_as(C.c_void_p)))
    return num, recs


def distance_batch(model_a, model_b, poses24, seed_a=None, seed_b=None, rel_err=0.0, abs_err=0.0, _entry="c2a_b200_distance_batch", qsize=2):
    """Batched C2A_Distance (depth-first routine; qsize > 2: the priority-queue routine): poses24 [n,24] = pose of A, pose
    of B.  Returns a dict with distance [n], p1p2 [n,6], tri_pair [n,2], num_bv_tests [n], num_tri_tests [n]."""
    if qsize > 2:
        _entry = "c2a_b200_distance_queue_batch"
    poses24 = np.ascontiguousarray(poses24, dtype=np.float64).reshape(-1, 24)
    n = poses24.shape[0]
    sa = None if seed_a is None else np.ascontiguousarray(seed_a, dtype=np.int32)
    sb = None if seed_b is None else np.ascontiguousarray(seed_b, dtype=np.int32)
    out = {"distance": np.zeros(n), "p1p2": np.zeros((n, 6)), "tri_pair": np.zeros((n, 2), dtype=np.int32),
           "num_bv_tests": np.zeros(n, dtype=np.int32), "num_tri_tests": np.zeros(n, dtype=np.int32)}
    _check(getattr(lib(), _entry)(model_a.h, model_b.h, poses24.ctypes.data_as(C.c_void_p),
                                         sa.ctypes.data_as(C.c_void_p) if sa is not None else None,
                                         sb.ctypes.data_as(C.c_void_p) if sb is not None else None,
                                         C.c_int64(n), C.c_double(rel_err), C.c_double(abs_err),
                                         *([C.c_int32(qsize)] if _entry == "c2a_b200_distance_queue_batch" else []),
                                         out["distance"].ctypes.data_as(C.c_void_p), out["p1p2"].ctypes.data_as(C.c_void_p),
                                         out["tri_pair"].ctypes.data_as(C.c_void_p), out["num_bv_tests"].ctypes.data_as(C.c_void_p),
                                         out["num_tri_tests"].ctypes.data_as(C.c_void_p)))
    return out


def collide_distance_batch(model_a, model_b, poses24, seed_a=None, seed_b=None, rel_err=0.0, abs_err=0.0):
    """Batched C2A_Collide, C2A_DistanceResult overload: ``distance_batch`` restricted to node pairs whose boxes overlap."""
    return distance_batch(model_a, model_b, poses24, seed_a, seed_b, rel_err, abs_err, _entry="c2a_b200_collide_distance_batch")


ALL_CONTACTS, FIRST_CONTACT = 1, 2   # C2A_ALL_CONTACTS / C2A_FIRST_CONTACT (C2A/C2A.h:246-247)


def collide_batch(model_a, model_b, poses24, flag=ALL_CONTACTS, max_pairs=256):
    """Batched C2A_Collide, PQP_CollideResult overload: poses24 [n,24] = pose of A, pose of B.  Returns a dict with
    num_pairs [n], pairs [n,max_pairs,2] (builder-order triangle indices in the reference's reporting order, -1 beyond
    min(num_pairs, max_pairs)), num_bv_tests [n], num_tri_tests [n]."""
    poses24 = np.ascontiguousarray(poses24, dtype=np.float64).reshape(-1, 24)
    n = poses24.shape[0]
    out = {"num_pairs": np.zeros(n, dtype=np.int32), "pairs": np.full((n, max_pairs, 2), -1, dtype=np.int32),
           "num_bv_tests": np.zeros(n, dtype=np.int32), "num_tri_tests": np.zeros(n, dtype=np.int32)}
    _check(lib().c2a_b200_collide_batch(model_a.h, model_b.h, poses24.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int32(flag),
                                        C.c_int32(max_pairs), out["num_pairs"].ctypes.data_as(C.c_void_p),
                                        out["pairs"].ctypes.data_as(C.c_void_p) if max_pairs else None,
                                        out["num_bv_tests"].ctypes.data_as(C.c_void_p), out["num_tri_tests"].ctypes.data_as(C.c_void_p)))
    return out


def motions_from_poses(poses, threads=0, out=None):
    """Host half of the motion model (acos via the host libm): poses [n,48] -> motion records [n,48]."""
    poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 48)
    n = poses.shape[0]
    if out is None:
        out = np.empty((n, 48), dtype=np.float64)
    _check(lib().c2a_b200_motions_from_poses(poses.ctypes.data_as(C.c_void_p), C.c_int64(n),
                                             out.ctypes.data_as(C.c_void_p), C.c_int32(threads)))
    return out


def schedule_order(model_a, model_b, motions):
    """Claim order (longest-expected queries first) for ``solve_batch_device``: int32 [n]."""
    motions = np.ascontiguousarray(motions, dtype=np.float64).reshape(-1, 48)
    order = np.empty(motions.shape[0], dtype=np.int32)
    _check(lib().c2a_b200_schedule_order(model_a.h, model_b.h, motions.ctypes.data_as(C.c_void_p),
                                         C.c_int64(motions.shape[0]), order.ctypes.data_as(C.c_void_p)))
    return order


def solve_batch_device(model_a, model_b, poses_ptr, n, out_ptrs, seed_a_ptr=None, seed_b_ptr=None,
                       tol_d=1e-4, tol_t=1e-4, stream=None, order_ptr=None):
    """Device-pointer entry: ``poses_ptr`` (MOTION RECORDS from ``motions_from_poses``, resident on the
    device) and the values of ``out_ptrs`` (dict field -> int address) are CUDA device addresses (e.g. ``torch.Tensor.data_ptr()``); enqueues on ``stream`` (a
    cudaStream_t address or None) and returns without synchronising."""
    res = Results()
    for name, _, _ in RESULT_FIELDS:
        if name in out_ptrs and out_ptrs[name]:
            setattr(res, name, out_ptrs[name])
    _check(lib().c2a_b200_solve_batch_device(model_a.h, model_b.h, C.c_void_p(poses_ptr),
                                             C.c_void_p(seed_a_ptr) if seed_a_ptr else None,
                                             C.c_void_p(seed_b_ptr) if seed_b_ptr else None,
                                             C.c_void_p(order_ptr) if order_ptr else None,
                                             C.c_int64(n), C.c_double(tol_d), C.c_double(tol_t), C.byref(res),
                                             C.c_void_p(stream) if stream else None))
