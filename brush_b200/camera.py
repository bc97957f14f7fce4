"""Host mirror of the reference camera (no device code).

Follows /root/reference/crates/brush-render/src/camera.rs:
  Camera                         camera.rs:11-82
  fov_to_focal / focal_to_fov    camera.rs:85-119  (f64, radians)
  calculate_jacobian_clamp_limits camera.rs:200-254 (pinhole branch)
and the uniform construction of render.rs:70-99 (ProjectUniforms).

`world_to_local` restates glam 0.30's `Affine3A::from_rotation_translation(..).inverse()`
(a dependency that is not vendored under /root/reference): `Mat3A::from_quat`, then the
cross-product 3x3 inverse, all in f32.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

F32 = np.float32

PINHOLE = 0


def fov_to_focal(fov: float, pixels: int) -> float:
    """camera.rs:85-101, pinhole branch.  f64 in, f64 out."""
    return (float(pixels) / 2.0) / math.tan(fov / 2.0)


def focal_to_fov(focal: float, pixels: int) -> float:
    """camera.rs:104-119, pinhole branch."""
    return 2.0 * math.atan((float(pixels) / 2.0) / focal)


def _mat3_from_quat_xyzw(q) -> np.ndarray:
    """glam Mat3A::from_quat; q = (x, y, z, w).  Returns columns as rows of a [3,3] f32 array (m[i] = column i)."""
    x, y, z, w = (F32(v) for v in q)
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz = x * x2, x * y2, x * z2
    yy, yz, zz = y * y2, y * z2, z * z2
    wx, wy, wz = w * x2, w * y2, w * z2
    one = F32(1.0)
    return np.array(
        [
            [one - (yy + zz), xy + wz, xz - wy],
            [xy - wz, one - (xx + zz), yz + wx],
            [xz + wy, yz - wx, one - (xx + yy)],
        ],
        dtype=F32,
    )


def _cross(a, b):
    return np.array(
        [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], dtype=F32
    )


@dataclass
class Camera:
    """camera.rs:11-19.  `rotation` is a glam quaternion (x, y, z, w), local -> world."""

    position: tuple = (0.0, 0.0, 0.0)
    rotation: tuple = (0.0, 0.0, 0.0, 1.0)
    fov_x: float = 0.0
    fov_y: float = 0.0
    center_uv: tuple = (0.5, 0.5)
    camera_model: int = PINHOLE

    def is_valid(self) -> bool:
        vals = [self.fov_x, self.fov_y, *self.center_uv, *self.position, *self.rotation]
        return all(math.isfinite(float(v)) for v in vals)

    def focal(self, img_w: int, img_h: int):
        return (F32(fov_to_focal(self.fov_x, img_w)), F32(fov_to_focal(self.fov_y, img_h)))

    def center(self, img_w: int, img_h: int):
        return (F32(self.center_uv[0]) * F32(img_w), F32(self.center_uv[1]) * F32(img_h))

    def world_to_local(self) -> np.ndarray:
        """Returns the 3x4 view matrix as 12 f32, column major (c0, c1, c2, translation)."""
        cols = _mat3_from_quat_xyzw(self.rotation)
        xa, ya, za = cols[0], cols[1], cols[2]
        t0, t1, t2 = _cross(ya, za), _cross(za, xa), _cross(xa, ya)
        det = F32(za[0] * t2[0] + za[1] * t2[1] + za[2] * t2[2])
        inv_det = F32(1.0) / det
        # from_cols(t0*inv, t1*inv, t2*inv).transpose()
        m = np.stack([t0 * inv_det, t1 * inv_det, t2 * inv_det]).astype(F32)  # m[i] = column i (pre-transpose)
        inv_cols = m.T.copy()  # after transpose: column i = (m[0][i], m[1][i], m[2][i])
        p = np.array(self.position, dtype=F32)
        # matrix3 * translation = col0*p.x + col1*p.y + col2*p.z
        mt = inv_cols[0] * p[0] + inv_cols[1] * p[1] + inv_cols[2] * p[2]
        trans = (-mt).astype(F32)
        return np.concatenate([inv_cols[0], inv_cols[1], inv_cols[2], trans]).astype(F32)


@dataclass
class ProjectUniforms:
    """Host mirror of shaders.rs:17-66 / kernels/types.rs:51-80 (pinhole)."""

    viewmat: np.ndarray  # [12] f32 column-major 3x4
    fx: float
    fy: float
    cx: float
    cy: float
    cam_pos: tuple
    lim_pos_x: float
    lim_pos_y: float
    lim_neg_x: float
    lim_neg_y: float
    half_max_render_fov: float
    camera_model: int = PINHOLE
    img_w: int = 0
    img_h: int = 0


def build_uniforms(camera: Camera, img_w: int, img_h: int) -> ProjectUniforms:
    """render.rs:70-99 + camera.rs:200-254."""
    assert img_w > 0 and img_h > 0, "Can't render images with 0 size."
    fx, fy = camera.focal(img_w, img_h)
    cx, cy = camera.center(img_w, img_h)
    wf, hf = F32(img_w), F32(img_h)
    lim_pos_x = (F32(1.15) * wf - cx) / fx
    lim_pos_y = (F32(1.15) * hf - cy) / fy
    lim_neg_x = (F32(-0.15) * wf - cx) / fx
    lim_neg_y = (F32(-0.15) * hf - cy) / fy
    hyp = F32(math.hypot(float(F32(camera.fov_x)), float(F32(camera.fov_y))))
    half = F32(min(float(hyp * F32(1.05)), float(F32(2.0) * F32(math.pi) - F32(1e-6)))) * F32(0.5)
    return ProjectUniforms(
        viewmat=camera.world_to_local(),
        fx=float(fx), fy=float(fy), cx=float(cx), cy=float(cy),
        cam_pos=tuple(float(F32(v)) for v in camera.position),
        lim_pos_x=float(lim_pos_x), lim_pos_y=float(lim_pos_y),
        lim_neg_x=float(lim_neg_x), lim_neg_y=float(lim_neg_y),
        half_max_render_fov=float(half),
        camera_model=camera.camera_model,
        img_w=img_w, img_h=img_h,
    )
