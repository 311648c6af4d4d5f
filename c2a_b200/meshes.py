"""Mesh inputs for the CCD workloads.

* ``load_tri``   -- the ``.tri`` text format the reference demo reads
                    (/root/reference/CCDDemo/mainTorusknot.cpp:448-485): ``TRI``, vertex count,
                    triangle count, ``x y z`` lines, then 0-based ``i j k`` lines.
* ``torus_knot`` -- the synthetic (2,3) torus-knot tube of SURVEY.md section 8d config 3 (the
                    reference ships no torus-knot mesh; its ``torusknot*.ani`` files are motions).
* ``load_ani``   -- the ``.ani`` rigid-motion format (CCDDemo/mainTorusknot.cpp:565-598): R is
                    stored column-major in the file.

All return ``tris9``: float64 ``[n_tris, 9]`` = p1,p2,p3 per triangle, plus ``vidx`` int32 ``[n_tris, 3]``.
"""
import numpy as np


def load_tri(path):
    with open(path, "r") as f:
        tok = f.read().split()
    if not tok or tok[0] != "TRI":
        raise ValueError(f"{path}: not a .tri file")
    nv, nt = int(tok[1]), int(tok[2])
    verts = np.array(tok[3:3 + 3 * nv], dtype=np.float64).reshape(nv, 3)
    vidx = np.array(tok[3 + 3 * nv:3 + 3 * nv + 3 * nt], dtype=np.int32).reshape(nt, 3)
    return verts[vidx].reshape(nt, 9).copy(), vidx


def save_tri(path, verts, vidx):
    with open(path, "w") as f:
        f.write("TRI\n\n%d\n\n%d\n" % (len(verts), len(vidx)))
        for v in verts:
            f.write("%.17g %.17g %.17g\n" % tuple(v))
        for t in vidx:
            f.write("%d %d %d\n" % tuple(t))


def torus_knot_verts(nu=512, nv=32, scale=60.0, tube=18.0):
    """(2,3) torus knot: c(u) = scale*((2+cos3u)cos2u, (2+cos3u)sin2u, sin3u); tube of radius ``tube``
    swept with a parallel-transport-free Frenet-like frame (normal from the second derivative)."""
    u = np.arange(nu, dtype=np.float64) * (2.0 * np.pi / nu)
    c = scale * np.stack([(2 + np.cos(3 * u)) * np.cos(2 * u), (2 + np.cos(3 * u)) * np.sin(2 * u), np.sin(3 * u)], 1)
    # analytic tangent
    d = scale * np.stack([
        -3 * np.sin(3 * u) * np.cos(2 * u) - 2 * (2 + np.cos(3 * u)) * np.sin(2 * u),
        -3 * np.sin(3 * u) * np.sin(2 * u) + 2 * (2 + np.cos(3 * u)) * np.cos(2 * u),
        3 * np.cos(3 * u)], 1)
    t = d / np.linalg.norm(d, axis=1, keepdims=True)
    # a stable frame: project the radial direction (from the z axis) off the tangent
    radial = np.stack([c[:, 0], c[:, 1], np.zeros(nu)], 1)
    n = radial - (radial * t).sum(1, keepdims=True) * t
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    b = np.cross(t, n)
    v = np.arange(nv, dtype=np.float64) * (2.0 * np.pi / nv)
    ring = np.cos(v)[None, :, None] * n[:, None, :] + np.sin(v)[None, :, None] * b[:, None, :]
    verts = c[:, None, :] + tube * ring
    return verts.reshape(nu * nv, 3)


def torus_knot(nu=512, nv=32, scale=60.0, tube=18.0):
    """nu*nv quads -> 2*nu*nv triangles (512x32 -> 32768, the config-3 mesh)."""
    verts = torus_knot_verts(nu, nv, scale, tube)
    i = np.arange(nu)[:, None]
    j = np.arange(nv)[None, :]
    v00 = (i * nv + j).ravel()
    v10 = (((i + 1) % nu) * nv + j).ravel()
    v01 = (i * nv + (j + 1) % nv).ravel()
    v11 = (((i + 1) % nu) * nv + (j + 1) % nv).ravel()
    vidx = np.concatenate([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], 0).astype(np.int32)
    # interleave the two triangles of each quad so neighbouring triangles stay adjacent in the list
    nq = nu * nv
    order = np.arange(2 * nq).reshape(2, nq).T.ravel()
    vidx = vidx[order]
    return verts[vidx].reshape(-1, 9).copy(), vidx


def load_ani(path):
    """Returns R [n,9] row-major and T [n,3]."""
    with open(path, "r") as f:
        tok = f.read().split()
    n = int(tok[0].rstrip("f"))
    R = np.zeros((n, 9)); T = np.zeros((n, 3))
    k = 1
    for i in range(n):
        k += 1  # "<i>f"
        col_major = np.array(tok[k:k + 9], dtype=np.float64); k += 9
        R[i] = col_major.reshape(3, 3).T.ravel()
        T[i] = np.array(tok[k:k + 3], dtype=np.float64); k += 3
    return R, T
