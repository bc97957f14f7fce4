"""Initial splats (host side): brush-train/src/splat_init.rs.

  estimate_scene_scale   <- splat_init.rs:20-46   (3x the mean nearest-neighbour camera spacing, floor 1.0)
  create_random_splats   <- splat_init.rs:48-128  (points in random camera frusta, log-uniform depth)
  compute_knn_scales     <- splat_init.rs:180-222 (log of half the mean distance to the 2 nearest neighbours,
                                                   clamped to [1e-3, 0.1 * median bounds size])
  to_init_splats         <- splat_init.rs:224-252 (defaults for the fields a point cloud does not carry)
  with_sh_degree         <- gaussian_splats.rs:145-163
The reference draws from an unseeded `rand::rng()`; here every draw comes from the caller's numpy Generator."""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

from .camera import Camera, _mat3_from_quat_xyzw
from .ply import SplatData


def inverse_sigmoid(x: float) -> float:
    return math.log(x / (1.0 - x))


def estimate_scene_scale(cameras: Sequence[Camera]) -> float:
    if len(cameras) < 2:
        return 1.0
    pos = np.array([c.position for c in cameras], np.float32)
    d = np.linalg.norm(pos[:, None, :] - pos[None, :, :], axis=2).astype(np.float32)
    np.fill_diagonal(d, np.inf)
    avg_nn = float(d.min(axis=1).astype(np.float32).sum() / np.float32(len(cameras)))
    return max(avg_nn * 3.0, 1.0)


def create_random_splats(init_count: int, cameras: Sequence[Camera], rng: np.random.Generator,
                         scene_scale: Optional[float] = None):
    """-> (transforms [N,10], sh [N,1,3], raw_opac [N])."""
    n = int(init_count)
    scale = float(scene_scale) if scene_scale is not None else estimate_scene_scale(cameras)
    near, far = scale * 0.05, scale
    cam_idx = rng.integers(0, len(cameras), n)
    pos = np.empty((n, 3), np.float32)
    for ci, cam in enumerate(cameras):
        sel = np.nonzero(cam_idx == ci)[0]
        if sel.size == 0:
            continue
        hx, hy = np.float32(cam.fov_x * 0.5), np.float32(cam.fov_y * 0.5)
        dx = np.tan(rng.uniform(-hx, hx, sel.size).astype(np.float32))
        dy = np.tan(rng.uniform(-hy, hy, sel.size).astype(np.float32))
        depth = np.exp(rng.uniform(math.log(near), math.log(far), sel.size)).astype(np.float32)
        local = np.stack([dx * depth, dy * depth, -depth], 1)          # as written in the reference (:88-90)
        R = _mat3_from_quat_xyzw(cam.rotation).T                        # columns of the local -> world rotation
        pos[sel] = local @ R.T + np.array(cam.position, np.float32)
    sh = rng.uniform(0.0, 1.0, (n, 1, 3)).astype(np.float32)
    q = rng.uniform(-1.0, 1.0, (n, 4)).astype(np.float32)
    q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-6)
    # drawn as (x, y, z, w) then stored in that order by the reference; the packed row wants (w, x, y, z)
    op = rng.uniform(inverse_sigmoid(0.1), inverse_sigmoid(0.25), n).astype(np.float32)
    log_scale = np.float32(math.log(scale / float(n) ** (1.0 / 3.0)))
    t = np.concatenate([pos, q, np.full((n, 3), log_scale, np.float32)], 1).astype(np.float32)
    return np.ascontiguousarray(t), sh, op


def compute_knn_scales(means: np.ndarray) -> np.ndarray:
    """-> log scales [N,3] (the same value on the three axes)."""
    from scipy.spatial import cKDTree
    from .train import bounds_from_pos
    n = means.shape[0]
    if n < 3:
        return np.zeros((n, 3), np.float32)
    m = np.ascontiguousarray(means, np.float32)
    median_size = max(bounds_from_pos(0.75, m).median_size(), 0.01)
    d, _ = cKDTree(m.astype(np.float64)).query(m.astype(np.float64), k=3)      # self + the two nearest neighbours
    dist = ((d[:, 1].astype(np.float32) + d[:, 2].astype(np.float32)) / np.float32(4.0))
    ls = np.log(np.clip(dist, np.float32(1e-3), np.float32(median_size * 0.1))).astype(np.float32)
    return np.repeat(ls[:, None], 3, axis=1)


def to_init_splats(data: SplatData):
    """-> (transforms [N,10], sh [N,K,3], raw_opac [N]); like SplatData.into_arrays but with KNN scales."""
    n = data.num_splats()
    ls = data.log_scales if data.log_scales is not None else compute_knn_scales(data.means)
    rot = data.rotations if data.rotations is not None else np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1))
    op = data.raw_opacities if data.raw_opacities is not None else np.full(n, inverse_sigmoid(0.5), np.float32)
    sh = data.sh_coeffs if data.sh_coeffs is not None else np.full((n, 1, 3), 0.5, np.float32)
    t = np.concatenate([data.means, rot, ls], 1).astype(np.float32)
    return np.ascontiguousarray(t), np.ascontiguousarray(sh, np.float32), np.ascontiguousarray(op, np.float32)


def with_sh_degree(sh: np.ndarray, degree: int) -> np.ndarray:
    """Pad with zeros or truncate to (degree+1)^2 coefficients."""
    k = (degree + 1) ** 2
    n, cur = sh.shape[0], sh.shape[1]
    if cur < k:
        return np.concatenate([sh, np.zeros((n, k - cur, 3), sh.dtype)], 1)
    return np.ascontiguousarray(sh[:, :k])
