"""Synthetic pose batches for the BASELINE.json configs.

A query is 48 float64: trans00, trans01 (object 1 begin/end), trans10, trans11 (object 2 begin/end),
each R (9, row-major) + T (3) -- the four ``Transform*`` arguments of ``C2A_Solve``
(/root/reference/C2A/C2A.h:23-35).
"""
import numpy as np

BUNNY_RADIUS = 131.4   # max |v| of tri_models/bunny_noholes.tri
KNOT_RADIUS = 198.0    # 60*3 + 18 for meshes.torus_knot defaults


def quat_to_matrix(q):
    """q: [n,4] (x,y,z,w), need not be normalised. Returns [n,9] row-major rotation matrices."""
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                  2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                  2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1)
    return R


def axis_angle_to_matrix(axis, theta):
    axis = axis / np.linalg.norm(axis, axis=1, keepdims=True)
    h = 0.5 * theta
    q = np.concatenate([axis * np.sin(h)[:, None], np.cos(h)[:, None]], 1)
    return quat_to_matrix(q)


def _matmul9(A, B):
    return np.einsum("nij,njk->nik", A.reshape(-1, 3, 3), B.reshape(-1, 3, 3)).reshape(-1, 9)


def approach_batch(n, seed, radius=BUNNY_RADIUS, max_turn=2.0):
    """Configs 2/3: object 2 static (random rotation, T=0); object 1 starts 300 units out (in bunny
    units; scaled by radius/BUNNY_RADIUS) along a random direction u and ends at s*u, s~U(-300,200),
    while turning by theta~U(0,max_turn) about a random body axis."""
    rng = np.random.default_rng(seed)
    k = radius / BUNNY_RADIUS
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    T0 = 300.0 * k * u
    s = rng.uniform(-300.0, 200.0, size=n) * k
    T1 = s[:, None] * u
    R0 = quat_to_matrix(rng.normal(size=(n, 4)))
    turn = axis_angle_to_matrix(rng.normal(size=(n, 3)), rng.uniform(0.0, max_turn, size=n))
    R1 = _matmul9(R0, turn)
    R2 = quat_to_matrix(rng.normal(size=(n, 4)))
    Z = np.zeros((n, 3))
    return np.ascontiguousarray(np.concatenate([R0, T0, R1, T1, R2, Z, R2, Z], 1))


def translation_batch(n, seed, radius=BUNNY_RADIUS, move_b=False):
    """Pure translations (both angular speeds exactly 0): the reference's translation-only branch
    (/root/reference/C2A/src/C2A.cpp:2391-2395, :1362-1521).  Object 1 keeps a random rotation and moves
    from 300 units out (scaled) along u to s*u + a lateral offset, so hits, grazes and misses all occur;
    object 2 keeps a random rotation and is static, or (``move_b``) drifts by up to 40 units."""
    rng = np.random.default_rng(seed)
    k = radius / BUNNY_RADIUS
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    T0 = 300.0 * k * u
    lat = rng.normal(size=(n, 3)); lat -= (lat * u).sum(1, keepdims=True) * u
    lat /= np.linalg.norm(lat, axis=1, keepdims=True)
    s = rng.uniform(-300.0, 200.0, size=n) * k
    T1 = s[:, None] * u + lat * rng.uniform(0.0, 220.0, size=(n, 1)) * k
    R1 = quat_to_matrix(rng.normal(size=(n, 4)))
    R2 = quat_to_matrix(rng.normal(size=(n, 4)))
    Z = np.zeros((n, 3))
    Z1 = rng.normal(size=(n, 3)) * (40.0 * k / 3.0) if move_b else Z
    return np.ascontiguousarray(np.concatenate([R1, T0, R1, T1, R2, Z, R2, Z1], 1))


def static_pose_batch(n, seed, radius=BUNNY_RADIUS):
    """Static pose pairs for the discrete distance query: the poses an approach_batch motion passes through at
    a random time (far apart, close, and interpenetrating cases all occur).  Returns [n,24] = pose of A, pose of B."""
    ap = approach_batch(n, seed, radius=radius, max_turn=0.0)
    lam = np.random.default_rng(seed + 1).uniform(0.0, 1.0, size=(n, 1))
    pa = ap[:, 0:12].copy()
    pa[:, 9:12] = ap[:, 9:12] * (1.0 - lam) + ap[:, 21:24] * lam
    return np.ascontiguousarray(np.concatenate([pa, ap[:, 24:36]], 1))


def degenerate_batch(seed=5, radius=BUNNY_RADIUS):
    """Edge-case motions (70 queries): no motion at all at three separations (far apart, close, interpenetrating);
    pure rotation in place; start pose already deep in contact (with and without rotation); motions of 1e-9 of the
    approach.  Zero or vanishing velocities exercise the floors and NaN paths of the motion bounds."""
    base = approach_batch(60, seed, radius=radius)
    out = []
    for lam in (0.0, 0.45, 0.9):
        p = base[:10].copy()
        T = p[:, 9:12] * (1 - lam) + p[:, 21:24] * lam
        p[:, 9:12] = T; p[:, 12:21] = p[:, 0:9]; p[:, 21:24] = T
        out.append(p)
    p = base[10:20].copy(); T = p[:, 9:12] * 0.4 + p[:, 21:24] * 0.6; p[:, 9:12] = T; p[:, 21:24] = T
    out.append(p)
    p = base[20:30].copy(); p[:, 9:12] = 0.02 * p[:, 9:12]
    out.append(p)
    q = p.copy(); q[:, 12:21] = q[:, 0:9]
    out.append(q)
    p = base[30:40].copy(); p[:, 21:24] = p[:, 9:12] + 1e-9 * (p[:, 21:24] - p[:, 9:12]); p[:, 12:21] = p[:, 0:9]
    out.append(p)
    return np.ascontiguousarray(np.concatenate(out))


def demo_batch(R1f, T1f, R2f, T2f):
    """Config 1: the 303 queries ``cb_display`` builds from torusknot1.ani / torusknot2.ani
    (/root/reference/CCDDemo/mainTorusknot.cpp:216-266): object 1 moves from frame step1 to step2 of
    file 1; object 2 keeps the rotation of frame ``iframe`` of file 2 with the translations of
    frames step1 -> step2 of file 2."""
    n = R1f.shape[0]
    out = np.zeros((n, 48))
    for i in range(n):
        s1, s2 = (0, 1) if i < 101 else ((101, 102) if i < 202 else (202, 203))
        out[i, 0:9] = R1f[s1]; out[i, 9:12] = T1f[s1]
        out[i, 12:21] = R1f[s2]; out[i, 21:24] = T1f[s2]
        out[i, 24:33] = R2f[i]; out[i, 33:36] = T2f[s1]
        out[i, 36:45] = R2f[i]; out[i, 45:48] = T2f[s2]
    return out


def grazing_batch(poses, toc, pose_toc, p1p2, push, seed=0):
    """Config 5 (adversarial near-contact sweep): from colliding queries (their poses, time of contact,
    object poses at contact and closest points p1/p2 in model-1 frame) build new queries whose END pose of
    object 1 is its TOC pose pushed ``push`` units (array, e.g. U(0, 1e-3)) past contact along the contact
    normal; the start pose and object 2 are unchanged.  The motion then only just reaches contact at t ~ 1:
    deep BVTT fronts and many conservative-advancement iterations."""
    poses = np.asarray(poses, dtype=np.float64).reshape(-1, 48)
    n = poses.shape[0]
    out = poses.copy()
    R1 = pose_toc[:, 0:9].reshape(n, 3, 3)
    d = np.einsum("nij,nj->ni", R1, p1p2[:, 3:6] - p1p2[:, 0:3])  # contact direction in the world frame
    nrm = np.linalg.norm(d, axis=1, keepdims=True)
    rng = np.random.default_rng(seed)
    fallback = rng.normal(size=(n, 3)); fallback /= np.linalg.norm(fallback, axis=1, keepdims=True)
    d = np.where(nrm > 1e-12, d / np.maximum(nrm, 1e-300), fallback)
    out[:, 12:21] = pose_toc[:, 0:9]
    out[:, 21:24] = pose_toc[:, 9:12] + np.asarray(push, dtype=np.float64).reshape(n, 1) * d
    return np.ascontiguousarray(out)


def scene(n_instances, seed, radii, side=None, speed=0.35, max_turn=1.0):
    """Config 4: ``n_instances`` objects, model k = i % len(radii), at random poses in a cube
    of edge ``side`` moving with random velocities (|displacement| ~ U(0, speed*side)) and turning by up to
    ``max_turn`` rad.  ``radii``: max |vertex| per model.  side=None sizes the cube for about 8 swept-sphere
    neighbours per instance.  Returns dict(model [n], begin [n,12], end [n,12]) with R(9)+T(3) poses."""
    rng = np.random.default_rng(seed)
    radii = np.asarray(radii, dtype=np.float64)
    model = (np.arange(n_instances) % len(radii)).astype(np.int32)
    if side is None:
        # expected neighbours ~ n * (4/3) pi (2 r_eff)^3 / side^3 with r_eff inflated by the sweep; solved for 8
        r_eff = float(np.mean(radii)) * 1.08
        side = (n_instances * (4.0 / 3.0) * np.pi * (2 * r_eff) ** 3 / 8.0) ** (1.0 / 3.0)
    T0 = rng.uniform(0.0, side, size=(n_instances, 3))
    d = rng.normal(size=(n_instances, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    T1 = T0 + d * rng.uniform(0.0, speed * side / n_instances ** (1.0 / 3.0) * 2.0, size=(n_instances, 1))
    R0 = quat_to_matrix(rng.normal(size=(n_instances, 4)))
    R1 = _matmul9(R0, axis_angle_to_matrix(rng.normal(size=(n_instances, 3)), rng.uniform(0.05, max_turn, size=n_instances)))
    return {"model": model, "begin": np.ascontiguousarray(np.concatenate([R0, T0], 1)),
            "end": np.ascontiguousarray(np.concatenate([R1, T1], 1)), "side": side}


def scene_queries(sc, pairs):
    """Assemble the heterogeneous CCD batch of a scene's candidate pairs: poses [m,48] = trans00, trans01 of
    instance i and trans10, trans11 of instance j, plus the two model-index arrays."""
    i, j = pairs[:, 0], pairs[:, 1]
    poses = np.concatenate([sc["begin"][i], sc["end"][i], sc["begin"][j], sc["end"][j]], 1)
    return np.ascontiguousarray(poses), sc["model"][i].copy(), sc["model"][j].copy()


def broadphase_reference(c0, c1, radius, margin=0.0):
    """numpy statement of c2a_b200_broadphase (same formula, FP64); returns (pairs sorted, |gap| per pair) where gap
    is how far inside the reach the pair is -- tests skip pairs whose gap is within rounding of zero."""
    n = len(radius)
    out, gaps = [], []
    v = c1 - c0
    for i in range(n - 1):
        p = c0[i] - c0[i + 1:]
        w = v[i] - v[i + 1:]
        vv = (w * w).sum(1)
        t = np.where(vv > 0, -(p * w).sum(1) / np.where(vv > 0, vv, 1.0), 0.0)
        t = np.clip(t, 0.0, 1.0)
        qv = p + t[:, None] * w
        reach = radius[i] + radius[i + 1:] + margin
        gap = reach - np.sqrt((qv * qv).sum(1))
        for k in np.nonzero(gap >= -1e-9)[0]:
            out.append((i, i + 1 + int(k))); gaps.append(float(gap[k]))
    return np.array(out, dtype=np.int32).reshape(-1, 2), np.array(gaps)
