"""Seeded synthetic inputs for the hot path.  numpy only, deterministic.

cloud "surface-box": N points on the surface of an axis-aligned box with side ratios
(1, 0.6, 0.3) scaled to unit diagonal (the NOCS convention, data_transforms.py:25), random
rotation, sigma = 0.002 noise, 10 % outliers uniform in the radius-0.6 ball
(config_track.yml:27), shuffled.  cloud "uniform": U(-0.5, 0.5)^3.  "tiled": a few unique
points repeated until N (the data loaders do exactly this for small crops,
nocs_data_process.py:105-106) -- the FPS tie-rule stress.
"""
import numpy as np


def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def surface_box(n, rng, num_parts=1, outlier_frac=0.10, noise=0.002, rotate=True):
    """Returns (points [n,3] float32, labels [n] int64); outliers get label num_parts."""
    sides = np.array([1.0, 0.6, 0.3])
    sides = sides / np.linalg.norm(sides)
    n_out = int(round(n * outlier_frac))
    n_in = n - n_out
    areas = np.array([sides[1] * sides[2], sides[0] * sides[2], sides[0] * sides[1]])
    face_axis = rng.choice(3, size=n_in, p=areas / areas.sum())
    pts = (rng.random((n_in, 3)) - 0.5) * sides
    sign = rng.choice([-0.5, 0.5], size=n_in)
    pts[np.arange(n_in), face_axis] = sign * sides[face_axis]
    labels = np.zeros(n_in, dtype=np.int64)
    if num_parts > 1:  # split along the longest side into equal slabs
        slab = np.floor((pts[:, 0] / sides[0] + 0.5) * num_parts).astype(np.int64)
        labels = np.clip(slab, 0, num_parts - 1)
    pts += rng.normal(scale=noise, size=pts.shape)
    if rotate:
        pts = pts @ random_rotation(rng).T
    d = rng.normal(size=(n_out, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    out = d * (0.6 * rng.random((n_out, 1)) ** (1.0 / 3.0))
    pts = np.concatenate([pts, out], 0)
    labels = np.concatenate([labels, np.full(n_out, num_parts, dtype=np.int64)])
    perm = rng.permutation(n)
    return pts[perm].astype(np.float32), labels[perm]


def batch_surface_box(b, n, seed=0, num_parts=1):
    rng = np.random.default_rng(seed)
    pl = [surface_box(n, rng, num_parts) for _ in range(b)]
    return np.stack([p for p, _ in pl]), np.stack([l for _, l in pl])


def batch_uniform(b, n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(-0.5, 0.5, size=(b, n, 3)).astype(np.float32)


def batch_tiled(b, n, unique, seed=0):
    """`unique` distinct points tiled (whole copies, in order) until n -- exact duplicates."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(-0.5, 0.5, size=(b, unique, 3)).astype(np.float32)
    reps = -(-n // unique)
    return np.ascontiguousarray(np.tile(base, (1, reps, 1))[:, :n])


def pose_fit_case(b, p, n, seed=0, nocs_noise=0.01, sym=False):
    """Synthetic NOCS predictions for the pose fit:
    nocs = R^T (cam - t) / s + noise for a known (R, s, t) per part; returns a dict of float32
    arrays: cam [b,n,3], labels [b,n], nocs [b,p,n,3], R [b,p,3,3], s [b,p], t [b,p,3]."""
    rng = np.random.default_rng(seed)
    cam = np.zeros((b, n, 3), np.float32)
    labels = np.zeros((b, n), np.int64)
    nocs = np.zeros((b, p, n, 3), np.float32)
    R = np.zeros((b, p, 3, 3), np.float32)
    s = np.zeros((b, p), np.float32)
    t = np.zeros((b, p, 3), np.float32)
    for bi in range(b):
        canon, lab = surface_box(n, rng, num_parts=p, rotate=False)
        labels[bi] = lab
        base_t = rng.normal(size=3)
        base_t = base_t / np.linalg.norm(base_t)
        cam_pts = np.zeros((n, 3))
        for pi in range(p):
            Rp = random_rotation(rng)
            if sym:  # keep y as the symmetry axis but leave a residual rotation about it
                th = rng.uniform(-np.pi, np.pi)
                Ry = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
                Rp = Rp @ Ry
            sp = rng.uniform(0.2, 0.5)
            tp = base_t + rng.normal(scale=0.05, size=3)
            R[bi, pi], s[bi, pi], t[bi, pi] = Rp, sp, tp
            sel = lab == pi
            cam_pts[sel] = (sp * (canon[sel] @ Rp.T) + tp)
        out = lab == p
        cam_pts[out] = base_t + canon[out]
        cam[bi] = cam_pts
        for pi in range(p):
            nocs[bi, pi] = ((cam_pts - t[bi, pi]) @ R[bi, pi]) / s[bi, pi] + rng.normal(scale=nocs_noise, size=(n, 3))
    return dict(cam=cam, labels=labels, nocs=nocs, R=R, s=s, t=t)


def depth_scene(seed=0, obj_radius=0.12, obj_depth=0.8, height=480, width=640, holes=0.02,
                intrinsics=((591.0125, 0, 322.525), (0, 590.16775, 244.11084), (0, 0, 1))):
    """A synthetic depth frame for the data-side crop (nocs_data_process.py:148-164): a sphere of `obj_radius` metres
    whose centre is `obj_depth` metres in front of the camera (slightly off-axis), in front of a noisy back wall at
    ~2 m; integer millimetres as a depth sensor gives them, `holes` of the pixels without measurement (0).
    Returns depth [H,W] float32 (mm), mask [H,W] int32 (1 = object), the sphere centre in the reference's camera
    frame (z negative, nocs_utils.py:29) and the intrinsics."""
    rng = np.random.default_rng(seed)
    K = np.array(intrinsics, dtype=np.float64)
    Kinv = np.linalg.inv(K)
    rows, cols = np.mgrid[0:height, 0:width]
    uv = np.stack([cols.ravel(), height - rows.ravel(), np.ones(height * width)])
    ray = (Kinv @ uv).T                                   # (x, y, 1): the point at depth d is (x d, y d, -d)
    ray[:, 2] = -1.0
    c = np.array([0.05 * rng.normal(), 0.04 * rng.normal(), -obj_depth])
    a = (ray * ray).sum(1)
    bq = -2.0 * ray @ c
    cq = c @ c - obj_radius ** 2
    disc = bq * bq - 4 * a * cq
    hit = disc > 0
    d = np.full(height * width, 2.0) + 0.01 * rng.normal(size=height * width)
    d[hit] = (-bq[hit] - np.sqrt(disc[hit])) / (2 * a[hit])
    depth = np.round(d * 1000.0).astype(np.float32)
    depth[rng.random(height * width) < holes] = 0.0
    mask = hit.astype(np.int32)
    return depth.reshape(height, width), mask.reshape(height, width), c, K
