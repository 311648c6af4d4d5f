"""ctypes bindings for the CPU oracle: the port (libc2a_oracle.so) and, when it was
built in the container that has /root/reference, the compiled reference
(oracle/_ref/libc2a_ref.so).  TEST INFRASTRUCTURE -- not imported by c2a_b200."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "libc2a_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libc2a_ref.so")

# mirrors struct orc_result (oracle/c2a_oracle.h)
RESULT_DTYPE = np.dtype([
    ("collisionfree", np.int32), ("numCA", np.int32), ("num_bv_tests", np.int32), ("num_tri_tests", np.int32),
    ("toc", np.float64), ("distance", np.float64), ("mint", np.float64),
    ("p1", np.float64, 3), ("p2", np.float64, 3), ("pose_toc", np.float64, 24),
    ("last_tri_a", np.int32), ("last_tri_b", np.int32),
], align=True)
OrcResult = RESULT_DTYPE


# mirrors struct orc_contact
CONTACT_DTYPE = np.dtype([("type_a", np.int32), ("type_b", np.int32), ("fid_a", np.int32, 3), ("fid_b", np.int32, 3),
                          ("tri_a", np.int32), ("tri_b", np.int32), ("pa", np.float64, 3), ("pb", np.float64, 3),
                          ("dist", np.float64)], align=True)


DISTANCE_DTYPE = np.dtype([("distance", np.float64), ("p1", np.float64, 3), ("p2", np.float64, 3), ("tri_a", np.int32),
                           ("tri_b", np.int32), ("num_bv_tests", np.int32), ("num_tri_tests", np.int32)], align=True)


SPEC_STATS_DTYPE = np.dtype([("steps", np.int64), ("tasks", np.int64), ("reached", np.int64), ("valid", np.int64),
                             ("work_seq", np.int64), ("work_fallback", np.int64), ("work_wasted", np.int64),
                             ("work_max_task", np.int64), ("work_top", np.int64), ("par_time", np.float64, 3),
                             ("depth", np.int32)], align=True)


REPLAY_STATS_DTYPE = np.dtype([("steps", np.int64), ("visits", np.int64), ("hits", np.int64), ("seq_hits", np.int64),
                               ("misses", np.int64), ("preeval", np.int64), ("wasted", np.int64)], align=True)


WIDE_STATS_DTYPE = np.dtype([(k, np.int64) for k in ("window", "leaf_batch", "leaf_passes", "steps", "redo", "anomalies", "closure_fail", "events", "rounds",
                                                    "max_width", "max_stack", "max_unresolved", "wide_tests",
                                                    "wide_tests_visited", "wide_leaves", "wide_leaves_visited")], align=True)


class _Bvh(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_tris", C.c_int32),
                ("R", C.c_void_p), ("Tr", C.c_void_p), ("l", C.c_void_p), ("r", C.c_void_p),
                ("R_loc", C.c_void_p), ("ang_radius", C.c_void_p), ("first_child", C.c_void_p),
                ("tris", C.c_void_p)]


def build_oracle(verbose=False):
    """make -C oracle: always builds the port; builds _ref only where /root/reference exists."""
    out = subprocess.run(["make", "-C", _HERE, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def have_ref():
    return os.path.exists(_REF_SO)


_port = None
_ref = None


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def bvh_struct(bvh):
    """bvh: dict of contiguous numpy arrays (R, Tr, l, r, R_loc, ang_radius, first_child, tris)."""
    s = _Bvh()
    s.n_nodes = int(bvh["first_child"].shape[0])
    s.n_tris = int(bvh["tris"].shape[0])
    for k in ("R", "Tr", "l", "r", "R_loc", "ang_radius", "tris"):
        a = bvh[k]
        assert a.dtype == np.float64 and a.flags.c_contiguous, k
        setattr(s, k, a.ctypes.data)
    a = bvh["first_child"]
    assert a.dtype == np.int32 and a.flags.c_contiguous
    s.first_child = a.ctypes.data
    s._keep = bvh
    return s


class _Port:
    def __init__(self):
        if not os.path.exists(_PORT_SO):
            build_oracle()
        self.lib = C.CDLL(_PORT_SO)
        L = self.lib
        L.orc_rect_dist.restype = C.c_double
        L.orc_tri_distance.restype = C.c_double
        L.orc_tri_dist.restype = C.c_double
        L.orc_motion_bound_bv.restype = C.c_double
        L.orc_motion_bound_leaf.restype = C.c_double

    def solve_batch(self, bvhA, bvhB, poses, seedA=None, seedB=None, tol_d=1e-4, tol_t=1e-4, threads=1):
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 48)
        n = poses.shape[0]
        out = np.zeros(n, dtype=RESULT_DTYPE)
        sA, sB = bvh_struct(bvhA), bvh_struct(bvhB)
        seedA = None if seedA is None else np.ascontiguousarray(seedA, dtype=np.int32)
        seedB = None if seedB is None else np.ascontiguousarray(seedB, dtype=np.int32)
        self.lib.orc_solve_batch(C.byref(sA), C.byref(sB), _ptr(poses), C.c_int64(n),
                                 _ptr(seedA) if seedA is not None else None,
                                 _ptr(seedB) if seedB is not None else None,
                                 C.c_double(tol_d), C.c_double(tol_t), _ptr(out), C.c_int32(threads))
        return out

    def solve_spec(self, bvhA, bvhB, poses, K, tol_d=1e-4, tol_t=1e-4):
        """Round-2 design study: every exact-mode CA step through the speculative subtree split at frontier depth K.
        Returns (results like solve_batch, dict of accumulated statistics)."""
        sA, sB = bvh_struct(bvhA), bvh_struct(bvhB)
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 48)
        out = np.zeros(len(poses), dtype=RESULT_DTYPE)
        st = np.zeros(1, dtype=SPEC_STATS_DTYPE)
        for i in range(len(poses)):
            self.lib.orc_solve_spec(C.byref(sA), C.byref(sB), _ptr(poses[i]), C.c_int32(0), C.c_int32(0), C.c_double(tol_d),
                                    C.c_double(tol_t), C.c_int32(K), C.c_void_p(out[i:i + 1].ctypes.data), _ptr(st))
        return out, {k: (st[k][0].tolist() if st[k].ndim > 1 else st[k][0].item()) for k in st.dtype.names}

    def solve_replay(self, bvhA, bvhB, poses, seedA=None, seedB=None, tol_d=1e-4, tol_t=1e-4):
        """Round-2 design study: every CA step replayed over the previous step's visit list.
        Returns (results like solve_batch, dict of accumulated statistics)."""
        sA, sB = bvh_struct(bvhA), bvh_struct(bvhB)
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 48)
        out = np.zeros(len(poses), dtype=RESULT_DTYPE)
        st = np.zeros(1, dtype=REPLAY_STATS_DTYPE)
        for i in range(len(poses)):
            self.lib.orc_solve_replay(C.byref(sA), C.byref(sB), _ptr(poses[i]), C.c_int32(0 if seedA is None else int(seedA[i])),
                                      C.c_int32(0 if seedB is None else int(seedB[i])), C.c_double(tol_d), C.c_double(tol_t),
                                      C.c_void_p(out[i:i + 1].ctypes.data), _ptr(st))
        return out, {k: st[k][0].item() for k in st.dtype.names}

    def solve_wide(self, bvhA, bvhB, poses, seedA=None, seedB=None, tol_d=1e-4, tol_t=1e-4, window=16, leaf_batch=0):
        """Round-2 algorithm on the CPU: exact-mode CA steps as a depth-first traversal popping `window` pairs per round.
        Returns (results like solve_batch, dict of accumulated statistics)."""
        sA, sB = bvh_struct(bvhA), bvh_struct(bvhB)
        poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 48)
        out = np.zeros(len(poses), dtype=RESULT_DTYPE)
        st = np.zeros(1, dtype=WIDE_STATS_DTYPE)
        st["window"] = window
        st["leaf_batch"] = leaf_batch
        for i in range(len(poses)):
            self.lib.orc_solve_wide(C.byref(sA), C.byref(sB), _ptr(poses[i]), C.c_int32(0 if seedA is None else int(seedA[i])),
                                    C.c_int32(0 if seedB is None else int(seedB[i])), C.c_double(tol_d), C.c_double(tol_t),
                                    C.c_void_p(out[i:i + 1].ctypes.data), _ptr(st))
        return out, {k: st[k][0].item() for k in st.dtype.names}

    def solve_visits(self, bvhA, bvhB, pose48, tol_d=1e-4, tol_t=1e-4, cap=1 << 23):
        """Round-2 design study: one query's result plus the per-CA-step sequences of visited node pairs
        (list of uint64 arrays, b1 << 32 | b2)."""
        sA, sB = bvh_struct(bvhA), bvh_struct(bvhB)
        p = np.ascontiguousarray(pose48, np.float64).reshape(48)
        out = np.zeros(1, dtype=RESULT_DTYPE)
        buf = np.zeros(cap, dtype=np.uint64)
        self.lib.orc_solve_visits.restype = C.c_int64
        n = self.lib.orc_solve_visits(C.byref(sA), C.byref(sB), _ptr(p), C.c_int32(0), C.c_int32(0), C.c_double(tol_d),
                                      C.c_double(tol_t), _ptr(out), _ptr(buf), C.c_int64(cap))
        v = buf[:min(n, cap)]
        cuts = np.nonzero(v == np.uint64(0xFFFFFFFFFFFFFFFF))[0]
        steps = [v[a + 1:b] for a, b in zip(cuts, list(cuts[1:]) + [len(v)])]
        return out[0], steps

    def contacts(self, bvhA, bvhB, pose1, pose2, threshold, vidx_a=None, vidx_b=None, max_out=4096):
        """Contact pass at the given poses; records in visiting order. Returns (count, records)."""
        sA, sB = bvh_struct(bvhA), bvh_struct(bvhB)
        p1 = np.ascontiguousarray(pose1, np.float64); p2 = np.ascontiguousarray(pose2, np.float64)
        va = None if vidx_a is None else np.ascontiguousarray(vidx_a, np.int32)
        vb = None if vidx_b is None else np.ascontiguousarray(vidx_b, np.int32)
        out = np.zeros(max_out, dtype=CONTACT_DTYPE)
        self.lib.orc_contacts.restype = C.c_int64
        n = self.lib.orc_contacts(C.byref(sA), C.byref(sB), _ptr(va) if va is not None else None,
                                  _ptr(vb) if vb is not None else None, _ptr(p1), _ptr(p2), C.c_double(threshold),
                                  C.c_int64(max_out), _ptr(out))
        return int(n), out[:min(int(n), max_out)]

    def distance(self, bvhA, bvhB, poses24, seedA=None, seedB=None, rel_err=0.0, abs_err=0.0, qsize=2):
        """C2A_Distance per query (depth-first routine; the priority-queue one for qsize > 2): poses24 [n,24] = pose of A, pose of B."""
        sA, sB = bvh_struct(bvhA), bvh_struct(bvhB)
        poses24 = np.ascontiguousarray(poses24, np.float64).reshape(-1, 24)
        out = np.zeros(len(poses24), dtype=DISTANCE_DTYPE)
        for i in range(len(poses24)):
            self.lib.orc_distance_queue(C.byref(sA), C.byref(sB), _ptr(poses24[i]), C.c_int32(0 if seedA is None else int(seedA[i])),
                                        C.c_int32(0 if seedB is None else int(seedB[i])), C.c_double(rel_err), C.c_double(abs_err),
                                        C.c_int32(qsize), C.c_void_p(out[i:i + 1].ctypes.data))
        return out

    def collide(self, bvhA, bvhB, poses24, flag=1, max_pairs=4096):
        """C2A_Collide (PQP_CollideResult overload) per query.  Returns (num_pairs [n], list of [k,2] builder-order
        triangle index pairs in traversal order, num_bv_tests [n], num_tri_tests [n])."""
        sA, sB = bvh_struct(bvhA), bvh_struct(bvhB)
        obb = [np.ascontiguousarray(x, np.float64) for x in (bvhA["obb_d"], bvhA["obb_To"], bvhB["obb_d"], bvhB["obb_To"])]
        poses24 = np.ascontiguousarray(poses24, np.float64).reshape(-1, 24)
        n = len(poses24)
        num = np.zeros(n, np.int32); nbv = np.zeros(n, np.int32); ntri = np.zeros(n, np.int32); pairs = []
        buf = np.zeros((max_pairs, 2), np.int32)
        self.lib.orc_collide.restype = C.c_int32
        for i in range(n):
            a, b = C.c_int32(), C.c_int32()
            num[i] = self.lib.orc_collide(C.byref(sA), C.byref(sB), _ptr(obb[0]), _ptr(obb[1]), _ptr(obb[2]), _ptr(obb[3]),
                                          _ptr(poses24[i]), C.c_int32(flag), C.c_int32(max_pairs), _ptr(buf), C.byref(a), C.byref(b))
            nbv[i], ntri[i] = a.value, b.value
            pairs.append(buf[:min(int(num[i]), max_pairs)].copy())
        return num, pairs, nbv, ntri

    def collide_distance(self, bvhA, bvhB, poses24, seedA=None, seedB=None, rel_err=0.0, abs_err=0.0):
        """C2A_Collide (C2A_DistanceResult overload) per query; records like distance()."""
        sA, sB = bvh_struct(bvhA), bvh_struct(bvhB)
        dA = np.ascontiguousarray(bvhA["obb_d"], np.float64); dB = np.ascontiguousarray(bvhB["obb_d"], np.float64)
        poses24 = np.ascontiguousarray(poses24, np.float64).reshape(-1, 24)
        out = np.zeros(len(poses24), dtype=DISTANCE_DTYPE)
        for i in range(len(poses24)):
            self.lib.orc_collide_distance(C.byref(sA), C.byref(sB), _ptr(dA), _ptr(dB), _ptr(poses24[i]),
                                          C.c_int32(0 if seedA is None else int(seedA[i])), C.c_int32(0 if seedB is None else int(seedB[i])),
                                          C.c_double(rel_err), C.c_double(abs_err), C.c_void_p(out[i:i + 1].ctypes.data))
        return out

    def rect_dist(self, Rab, Tab, a, b):
        Rab = np.ascontiguousarray(Rab, np.float64); Tab = np.ascontiguousarray(Tab, np.float64)
        a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
        P = np.zeros(3); Q = np.zeros(3); S = np.full(3, np.nan)
        d = self.lib.orc_rect_dist(_ptr(Rab), _ptr(Tab), _ptr(a), _ptr(b), _ptr(P), _ptr(Q), _ptr(S))
        return d, P, Q, S

    def tri_distance(self, R, T, t1, t2):
        R = np.ascontiguousarray(R, np.float64); T = np.ascontiguousarray(T, np.float64)
        t1 = np.ascontiguousarray(t1, np.float64); t2 = np.ascontiguousarray(t2, np.float64)
        p = np.zeros(3); q = np.zeros(3)
        d = self.lib.orc_tri_distance(_ptr(R), _ptr(T), _ptr(t1), _ptr(t2), _ptr(p), _ptr(q))
        return d, p, q


class RefModel:
    """A C2A_Model built by the reference's own BeginModel/AddTri/EndModel."""

    def __init__(self, lib, tris9, vidx=None):
        tris9 = np.ascontiguousarray(tris9, dtype=np.float64).reshape(-1, 9)
        self.lib = lib
        vi = None if vidx is None else np.ascontiguousarray(vidx, dtype=np.int32)
        lib.ref_model_build.restype = C.c_void_p
        self.h = C.c_void_p(lib.ref_model_build(_ptr(tris9), _ptr(vi) if vi is not None else None,
                                                C.c_int32(tris9.shape[0])))
        nn, nt = C.c_int32(), C.c_int32()
        lib.ref_model_counts(self.h, C.byref(nn), C.byref(nt))
        self.n_nodes, self.n_tris = nn.value, nt.value

    def export(self):
        nn, nt = self.n_nodes, self.n_tris
        b = {"R": np.zeros((nn, 9)), "Tr": np.zeros((nn, 3)), "l": np.zeros((nn, 2)), "r": np.zeros(nn),
             "R_loc": np.zeros((nn, 9)), "ang_radius": np.zeros(nn), "first_child": np.zeros(nn, np.int32),
             "tris": np.zeros((nt, 9)), "tri_ids": np.zeros(nt, np.int32)}
        self.lib.ref_model_export(self.h, _ptr(b["R"]), _ptr(b["Tr"]), _ptr(b["l"]), _ptr(b["r"]), _ptr(b["R_loc"]),
                                  _ptr(b["ang_radius"]), _ptr(b["first_child"]), _ptr(b["tris"]), _ptr(b["tri_ids"]))
        b["obb_d"] = np.zeros((nn, 3)); b["obb_To"] = np.zeros((nn, 3))
        self.lib.ref_model_export_obb(self.h, _ptr(b["obb_d"]), _ptr(b["obb_To"]))
        return b


class _Ref:
    def __init__(self):
        if not os.path.exists(_REF_SO):
            raise RuntimeError("oracle/_ref/libc2a_ref.so not built (needs /root/reference; run make -C oracle)")
        self.lib = C.CDLL(_REF_SO)
        self.lib.ref_rect_dist.restype = C.c_double
        self.lib.ref_tri_distance_intree.restype = C.c_double

    def model(self, tris9, vidx=None):
        return RefModel(self.lib, tris9, vidx)

    def solve_batch(self, mA, mB, poses, seedA=None, seedB=None, mode=0, tol_d=1e-4, tol_t=1e-4, threads=1):
        poses = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, 48)
        n = poses.shape[0]
        out = np.zeros(n, dtype=RESULT_DTYPE)
        ncont = np.zeros(n, dtype=np.int32)
        seedA = None if seedA is None else np.ascontiguousarray(seedA, dtype=np.int32)
        seedB = None if seedB is None else np.ascontiguousarray(seedB, dtype=np.int32)
        self.lib.ref_solve_batch(mA.h, mB.h, _ptr(poses), C.c_int64(n),
                                 _ptr(seedA) if seedA is not None else None,
                                 _ptr(seedB) if seedB is not None else None,
                                 C.c_int32(mode), C.c_double(tol_d), C.c_double(tol_t), _ptr(out), _ptr(ncont),
                                 C.c_int32(threads))
        return (out, ncont) if mode == 1 else out

    def distance(self, mA, mB, poses24, seedA=None, seedB=None, rel_err=0.0, abs_err=0.0, qsize=2):
        """The reference's own C2A_Distance, one query at a time (it reads and writes the models' last_tri)."""
        poses24 = np.ascontiguousarray(poses24, np.float64).reshape(-1, 24)
        out = np.zeros(len(poses24), dtype=DISTANCE_DTYPE)
        for i in range(len(poses24)):
            self.lib.ref_distance(mA.h, mB.h, _ptr(poses24[i]), C.c_int32(0 if seedA is None else int(seedA[i])),
                                  C.c_int32(0 if seedB is None else int(seedB[i])), C.c_double(rel_err), C.c_double(abs_err),
                                  C.c_int32(qsize), C.c_void_p(out[i:i + 1].ctypes.data))
        return out

    def collide(self, mA, mB, poses24, flag=1, max_pairs=4096):
        """The reference's own C2A_Collide (PQP_CollideResult overload), one query at a time.  Pairs are AddTri ids."""
        poses24 = np.ascontiguousarray(poses24, np.float64).reshape(-1, 24)
        n = len(poses24)
        num = np.zeros(n, np.int32); nbv = np.zeros(n, np.int32); ntri = np.zeros(n, np.int32); pairs = []
        buf = np.zeros((max_pairs, 2), np.int32)
        self.lib.ref_collide.restype = C.c_int32
        for i in range(n):
            a, b = C.c_int32(), C.c_int32()
            num[i] = self.lib.ref_collide(mA.h, mB.h, _ptr(poses24[i]), C.c_int32(flag), C.c_int32(max_pairs), _ptr(buf), C.byref(a), C.byref(b))
            nbv[i], ntri[i] = a.value, b.value
            pairs.append(buf[:min(int(num[i]), max_pairs)].copy())
        return num, pairs, nbv, ntri

    def collide_distance(self, mA, mB, poses24, seedA=None, seedB=None, rel_err=0.0, abs_err=0.0):
        """The reference's own C2A_Collide (C2A_DistanceResult overload), one query at a time."""
        poses24 = np.ascontiguousarray(poses24, np.float64).reshape(-1, 24)
        out = np.zeros(len(poses24), dtype=DISTANCE_DTYPE)
        for i in range(len(poses24)):
            self.lib.ref_collide_distance(mA.h, mB.h, _ptr(poses24[i]), C.c_int32(0 if seedA is None else int(seedA[i])),
                                          C.c_int32(0 if seedB is None else int(seedB[i])), C.c_double(rel_err), C.c_double(abs_err),
                                          C.c_void_p(out[i:i + 1].ctypes.data))
        return out

    def rect_dist(self, Rab, Tab, a, b):
        Rab = np.ascontiguousarray(Rab, np.float64); Tab = np.ascontiguousarray(Tab, np.float64)
        a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
        P = np.zeros(3); Q = np.zeros(3); S = np.full(3, np.nan)
        d = self.lib.ref_rect_dist(_ptr(Rab), _ptr(Tab), _ptr(a), _ptr(b), _ptr(P), _ptr(Q), _ptr(S))
        return d, P, Q, S

    def solve_contacts(self, mA, mB, poses48, seedA=0, seedB=0, max_out=4096):
        """The full unmodified C2A_Solve for one query: (result record, num_contact, contact records in LIST order)."""
        p = np.ascontiguousarray(poses48, np.float64)
        res = np.zeros(1, dtype=RESULT_DTYPE)
        out = np.zeros(max_out, dtype=CONTACT_DTYPE)
        self.lib.ref_solve_contacts.restype = C.c_int64
        n = self.lib.ref_solve_contacts(mA.h, mB.h, _ptr(p), C.c_int32(seedA), C.c_int32(seedB), _ptr(res), C.c_int64(max_out), _ptr(out))
        return res[0], int(n), out[:min(int(n), max_out)]

    def tri_distance_intree(self, R, T, t1, t2):
        R = np.ascontiguousarray(R, np.float64); T = np.ascontiguousarray(T, np.float64)
        t1 = np.ascontiguousarray(t1, np.float64); t2 = np.ascontiguousarray(t2, np.float64)
        p = np.zeros(3); q = np.zeros(3); col = C.c_int32(0)
        d = self.lib.ref_tri_distance_intree(_ptr(R), _ptr(T), _ptr(t1), _ptr(t2), _ptr(p), _ptr(q), C.byref(col))
        return d, p, q, col.value

    def motion_probe(self, poses24, t, ang_radius, direction):
        poses24 = np.ascontiguousarray(poses24, np.float64); direction = np.ascontiguousarray(direction, np.float64)
        out = np.zeros(21)
        self.lib.ref_motion_probe(_ptr(poses24), C.c_double(t), C.c_double(ang_radius), _ptr(direction), _ptr(out))
        return out


def port():
    global _port
    if _port is None:
        _port = _Port()
    return _port


def ref():
    global _ref
    if _ref is None:
        _ref = _Ref()
    return _ref
