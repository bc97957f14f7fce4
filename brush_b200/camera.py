"""Host mirror of the reference camera (no device code).

Follows /root/reference/crates/brush-render/src/camera.rs:
  Camera                         camera.rs:11-82
  fov_to_focal / focal_to_fov    camera.rs:85-198  (f64, radians; all four camera models)
  calculate_jacobian_clamp_limits camera.rs:200-254
  camera models                  kernels/camera_model/mod.rs:32-39: `camera_model` is the model id and
                                 `model_params` its distortion coefficients (KB4: k1..k4; RT8: k1..k6, p1, p2;
                                 thin-prism fisheye: k1..k4, p1, p2, sx1, sy1)
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

PINHOLE, KANNALA_BRANDT_4, RADIAL_TANGENTIAL_8, THIN_PRISM_FISHEYE = 0, 1, 2, 3
MODEL_PARAM_COUNT = {PINHOLE: 0, KANNALA_BRANDT_4: 4, RADIAL_TANGENTIAL_8: 8, THIN_PRISM_FISHEYE: 8}


def _kb4_d(theta: float, k) -> float:
    """camera.rs:121-130 (coefficients are f32 in the reference, widened to f64)."""
    t2 = theta * theta
    t3 = t2 * theta
    t5 = t3 * t2
    t7 = t5 * t2
    t9 = t7 * t2
    return theta + k[0] * t3 + k[1] * t5 + k[2] * t7 + k[3] * t9


def _kb4_dd(theta: float, k) -> float:
    t2 = theta * theta
    t4 = t2 * t2
    t6 = t4 * t2
    t8 = t6 * t2
    return 1.0 + 3.0 * k[0] * t2 + 5.0 * k[1] * t4 + 7.0 * k[2] * t6 + 9.0 * k[3] * t8


def _kb4_invert_d(target: float, k) -> float:
    """camera.rs:146-169: Newton on d(theta) = target, theta in [0, pi]."""
    if target <= 0.0:
        return 0.0
    theta = min(target, math.pi - 1e-6)
    for _ in range(50):
        f = _kb4_d(theta, k) - target
        fp = _kb4_dd(theta, k)
        if abs(fp) < 1e-12:
            break
        nxt = min(max(theta - f / fp, 0.0), math.pi)
        if abs(nxt - theta) < 1e-12:
            theta = nxt
            break
        theta = nxt
    return theta


def _rt8_radial(r: float, p) -> float:
    """camera.rs:172-180."""
    r2 = r * r
    r4 = r2 * r2
    r6 = r4 * r2
    return (1.0 + p[0] * r2 + p[1] * r4 + p[2] * r6) / (1.0 + p[3] * r2 + p[4] * r4 + p[5] * r6)


def rt8_undistort_radius(r_d: float, p) -> float:
    """camera.rs:184-198: fixed-point iteration r = r_d / radial(r)."""
    r = r_d
    for _ in range(30):
        factor = _rt8_radial(r, p)
        if abs(factor) < 1e-12:
            break
        r_new = r_d / factor
        if abs(r_new - r) < 1e-12:
            r = r_new
            break
        r = r_new
    return r


def _params64(model_params, camera_model=None):
    if camera_model is not None and len(model_params) != MODEL_PARAM_COUNT.get(camera_model, -1):
        raise ValueError("model_params does not match the camera model")
    return [float(F32(v)) for v in model_params]


def fov_to_focal(fov: float, pixels: int, camera_model: int = PINHOLE, model_params=()) -> float:
    """camera.rs:85-101.  f64 in, f64 out."""
    half = fov / 2.0
    p = _params64(model_params, camera_model)
    if camera_model == PINHOLE:
        projected = math.tan(half)
    elif camera_model in (KANNALA_BRANDT_4, THIN_PRISM_FISHEYE):
        projected = _kb4_d(half, p)
    elif camera_model == RADIAL_TANGENTIAL_8:
        r = math.tan(half)
        projected = r * _rt8_radial(r, p)
    else:
        raise ValueError("unknown camera model")
    return (float(pixels) / 2.0) / projected


def focal_to_fov(focal: float, pixels: int, camera_model: int = PINHOLE, model_params=()) -> float:
    """camera.rs:104-119."""
    r_norm = (float(pixels) / 2.0) / focal
    p = _params64(model_params, camera_model)
    if camera_model == PINHOLE:
        half = math.atan(r_norm)
    elif camera_model in (KANNALA_BRANDT_4, THIN_PRISM_FISHEYE):
        half = _kb4_invert_d(r_norm, p)
    elif camera_model == RADIAL_TANGENTIAL_8:
        half = math.atan(rt8_undistort_radius(r_norm, p))
    else:
        raise ValueError("unknown camera model")
    return 2.0 * half


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
    model_params: tuple = ()

    def is_valid(self) -> bool:
        vals = [self.fov_x, self.fov_y, *self.center_uv, *self.position, *self.rotation]
        return all(math.isfinite(float(v)) for v in vals)

    def focal(self, img_w: int, img_h: int):
        return (F32(fov_to_focal(self.fov_x, img_w, self.camera_model, self.model_params)),
                F32(fov_to_focal(self.fov_y, img_h, self.camera_model, self.model_params)))

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
    model_params: tuple = ()
    img_w: int = 0
    img_h: int = 0


def build_uniforms(camera: Camera, img_w: int, img_h: int) -> ProjectUniforms:
    """render.rs:70-99 + camera.rs:200-254."""
    assert img_w > 0 and img_h > 0, "Can't render images with 0 size."
    fx, fy = camera.focal(img_w, img_h)
    cx, cy = camera.center(img_w, img_h)
    wf, hf = F32(img_w), F32(img_h)
    model = camera.camera_model
    if len(camera.model_params) != MODEL_PARAM_COUNT.get(model, -1):
        raise ValueError("camera.model_params does not match the camera model")
    lim_pos_x = (F32(1.15) * wf - cx) / fx
    lim_pos_y = (F32(1.15) * hf - cy) / fy
    lim_neg_x = (F32(-0.15) * wf - cx) / fx
    lim_neg_y = (F32(-0.15) * hf - cy) / fy
    if model == RADIAL_TANGENTIAL_8:        # bound the UNDISTORTED coordinate (camera.rs:229-243)
        p64 = _params64(camera.model_params)
        und = lambda e: F32(rt8_undistort_radius(abs(float(e)), p64)) * F32(np.sign(e))
        lim_pos_x, lim_pos_y, lim_neg_x, lim_neg_y = und(lim_pos_x), und(lim_pos_y), und(lim_neg_x), und(lim_neg_y)
    elif model in (KANNALA_BRANDT_4, THIN_PRISM_FISHEYE):   # fisheye Jacobians are not clamped (camera.rs:244-247)
        lim_pos_x = lim_pos_y = lim_neg_x = lim_neg_y = F32(0.0)
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
        model_params=tuple(float(F32(v)) for v in camera.model_params),
        img_w=img_w, img_h=img_h,
    )
