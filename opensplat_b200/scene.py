"""Seeded synthetic scenes in the reference's simple_trainer camera convention.

simple_trainer.cpp:84-136: viewmat = identity with t_z = 8, projmat = viewmat (so w == 1 and
the NDC depth equals view-space z), fx = fy = 0.5*W/tan(fovX/2) with fovX = 90 deg, principal point
at the image centre.  Under that convention the reference's CPU back end and its CUDA tile
semantics agree on depth order and pixel centres, which is what the
parity tests need.  Host-side plumbing only (numpy); no kernels here.
"""
import math
import numpy as np

SH_C0 = 0.28209479177387814  # spherical_harmonics.cpp:18


def _distinct_depths(n, zlo, zhi, rng):
    """n strictly distinct fp32 values in [zlo, zhi] (bit-pattern grid, randomly permuted).
    Distinct view depths make the (tile | depth) sort order unique."""
    lo = np.array([zlo], np.float32).view(np.int32)[0]
    hi = np.array([zhi], np.float32).view(np.int32)[0]
    if hi - lo + 1 < n:
        raise ValueError(f"only {hi - lo + 1} distinct fp32 depths in [{zlo},{zhi}] for n={n}")
    bits = lo + (np.arange(n, dtype=np.int64) * (int(hi) - int(lo))) // max(n - 1, 1)
    z = bits.astype(np.int32).view(np.float32)
    return z[rng.permutation(n)]


def make_camera(W, H, t_z=8.0):
    view = np.eye(4, dtype=np.float32)
    view[2, 3] = t_z
    focal = 0.5 * float(W) / math.tan(0.5 * math.pi / 2.0)
    return dict(viewmat=view, projmat=view.copy(), fx=float(focal), fy=float(focal),
                cx=float(W // 2), cy=float(H // 2), W=int(W), H=int(H))


def make_scene(n, W, H, scale=0.02, sh_degree=3, opacity=(0.05, 0.35), seed=0, t_z=8.0,
               zrange=None, xy_extent=1.0):
    """Returns dict of float32 numpy arrays: means [n,3], scales [n,3] (already exp'ed),
    quats [n,4] (unit, w first), coeffs [n,K,3], opacities [n,1] (already sigmoid'ed),
    viewdirs [n,3] (unit), plus the camera."""
    rng = np.random.default_rng(seed)
    cam = make_camera(W, H, t_z)
    if zrange is None:
        zrange = (7.0, 9.0) if n <= 3_000_000 else (6.0, 10.0)
    tz = _distinct_depths(n, zrange[0], zrange[1], rng)
    means = np.empty((n, 3), np.float32)
    means[:, 0] = rng.uniform(-xy_extent, xy_extent, n)
    means[:, 1] = rng.uniform(-xy_extent, xy_extent, n)
    means[:, 2] = tz - np.float32(t_z)  # exact (Sterbenz); mean_z + t_z reproduces tz bit-for-bit
    scales = (scale * rng.uniform(0.25, 1.0, (n, 3))).astype(np.float32)
    # random unit quaternions, simple_trainer.cpp:110-126
    u, v, w = rng.uniform(size=(3, n))
    quats = np.stack([np.sqrt(1 - u) * np.sin(2 * np.pi * v), np.sqrt(1 - u) * np.cos(2 * np.pi * v),
                      np.sqrt(u) * np.sin(2 * np.pi * w), np.sqrt(u) * np.cos(2 * np.pi * w)], -1)
    quats = (quats / np.linalg.norm(quats, axis=-1, keepdims=True)).astype(np.float32)
    K = (sh_degree + 1) ** 2
    coeffs = (0.3 * rng.standard_normal((n, K, 3))).astype(np.float32)
    coeffs[:, 0, :] = ((rng.uniform(size=(n, 3)) - 0.5) / SH_C0).astype(np.float32)  # rgb2sh
    opac = rng.uniform(opacity[0], opacity[1], (n, 1)).astype(np.float32)
    cam_pos = np.array([0.0, 0.0, -t_z], np.float32)  # camera centre for viewmat above
    vd = means - cam_pos
    vd = (vd / np.linalg.norm(vd, axis=-1, keepdims=True)).astype(np.float32)
    out = dict(means=means, scales=scales, quats=quats, coeffs=coeffs, opacities=opac, viewdirs=vd)
    out.update(cam)
    return out


def rotated_camera(W, H, k, n_views=8, t_z=8.0):
    """View k of n_views: camera orbiting the scene's y axis by k*360/n_views degrees, keeping the
    w == 1 convention (projmat = viewmat).  Used for the data-parallel multi-view config (C4)."""
    cam = make_camera(W, H, t_z)
    a = 2.0 * math.pi * k / n_views
    R = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]], np.float32)
    view = np.eye(4, dtype=np.float32)
    view[:3, :3] = R
    view[2, 3] = t_z
    cam["viewmat"] = view
    cam["projmat"] = view.copy()
    cam["cam_pos"] = (-R.T @ np.array([0, 0, t_z], np.float32)).astype(np.float32)
    return cam


# The 8 data-parallel views of config C4.  The synthetic scene fills the cube [-1,1]^3 uniformly, so the rotations of
# the cube's symmetry group map it onto itself: every view below sees the same footprint on screen (the whole
# image), the same depth range and, statistically, the same number of intersections -- ranks of a weak-scaling run
# then carry comparable work, and N-GPU throughput can be compared with N x the 1-GPU number.  (Generic orbit
# views, rotated_camera(), crop the cube differently per view and skew the per-rank work by up to ~25 %.)
_CUBE_VIEWS = [
    np.eye(3),                                      # 0: front (the single-GPU view)
    [[0, 0, 1], [0, 1, 0], [-1, 0, 0]],             # 1: 90 deg about y
    [[-1, 0, 0], [0, 1, 0], [0, 0, -1]],            # 2: 180 deg about y (back)
    [[0, 0, -1], [0, 1, 0], [1, 0, 0]],             # 3: 270 deg about y
    [[1, 0, 0], [0, 0, -1], [0, 1, 0]],             # 4: 90 deg about x (top)
    [[1, 0, 0], [0, 0, 1], [0, -1, 0]],             # 5: 270 deg about x (bottom)
    [[0, -1, 0], [1, 0, 0], [0, 0, 1]],             # 6: front, rolled 90 deg
    [[0, 1, 0], [1, 0, 0], [0, 0, -1]],             # 7: back, rolled (180 deg about the x=y diagonal)
]


def cube_view_camera(W, H, k, t_z=8.0):
    """View k (mod 8) of the C4 view set: a rotation of the cube's symmetry group applied to the scene, camera
    convention as make_camera (projmat = viewmat, w == 1)."""
    cam = make_camera(W, H, t_z)
    R = np.asarray(_CUBE_VIEWS[k % 8], np.float32)
    assert abs(float(np.linalg.det(R)) - 1.0) < 1e-6
    view = np.eye(4, dtype=np.float32)
    view[:3, :3] = R
    view[2, 3] = t_z
    cam["viewmat"] = view
    cam["projmat"] = view.copy()
    cam["cam_pos"] = (-R.T @ np.array([0, 0, t_z], np.float32)).astype(np.float32)
    return cam
