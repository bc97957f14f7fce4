"""Camera models: host fov<->focal helpers against the reference's own unit tests
(brush-render/src/tests/mod.rs:710-790) and the oracle's distorted-camera projection / Jacobian / VJP against
central finite differences, the way finite_diff.rs:725-790,1084+ exercises KB4 / RT8 / thin-prism fisheye."""
import math

import numpy as np
import pytest

from brush_b200 import camera as cm
from brush_b200.camera import Camera, build_uniforms
from oracle import oracle as orc
from scenes import finite_diff_base_scene, synthetic_scene
from test_oracle_finite_diff import BROAD, _check

KB4 = (cm.KANNALA_BRANDT_4, (-0.05, 0.01, -0.001, 5e-5))
RT8 = (cm.RADIAL_TANGENTIAL_8, (-0.2, 0.05, -0.001, 0.0, 0.0, 0.0, 1e-3, -1e-3))
RT8_RATIONAL = (cm.RADIAL_TANGENTIAL_8, (-0.1, 0.03, -0.002, 0.05, -0.01, 0.001, 5e-3, -4e-3))
TPF = (cm.THIN_PRISM_FISHEYE, (-0.05, 0.01, -0.001, 5e-5, 1e-3, -1e-3, 5e-4, -5e-4))
MODELS = {"kb4": KB4, "rt8": RT8, "rt8_rational": RT8_RATIONAL, "tpf": TPF}


def test_pinhole_focal_to_fov_and_back():
    fov = cm.focal_to_fov(800.0, 1920)
    assert abs(cm.fov_to_focal(fov, 1920) - 800.0) < 1e-9


def test_kb4_focal_to_fov_and_back():
    fov = cm.focal_to_fov(300.0, 1024, cm.KANNALA_BRANDT_4, (0, 0, 0, 0))
    assert abs(fov - 1024 / 300.0) < 1e-9          # zero distortion: r_pix = f * theta
    assert abs(cm.fov_to_focal(fov, 1024, cm.KANNALA_BRANDT_4, (0, 0, 0, 0)) - 300.0) < 1e-9
    p = (-0.01, 0.003, -0.0005, 0.00002)
    fov = cm.focal_to_fov(280.0, 1024, cm.KANNALA_BRANDT_4, p)
    assert abs(cm.fov_to_focal(fov, 1024, cm.KANNALA_BRANDT_4, p) - 280.0) < 1e-6


def test_rt8_and_tpf_focal_to_fov_and_back():
    p = (-0.2, 0.05, -0.001, 0.0, 0.0, 0.0, 0.0, 0.0)
    fov = cm.focal_to_fov(900.0, 1920, cm.RADIAL_TANGENTIAL_8, p)
    assert abs(cm.fov_to_focal(fov, 1920, cm.RADIAL_TANGENTIAL_8, p) - 900.0) < 1e-6
    p = (-0.01, 0.003, -0.0005, 0.00002, 1e-3, -2e-3, 5e-4, -5e-4)
    fov = cm.focal_to_fov(280.0, 1024, cm.THIN_PRISM_FISHEYE, p)
    assert abs(cm.fov_to_focal(fov, 1024, cm.THIN_PRISM_FISHEYE, p) - 280.0) < 1e-6


def test_clamp_limits_per_model():
    base = dict(position=(0, 0, -3.0), fov_x=0.9, fov_y=0.7)
    u = build_uniforms(Camera(**base), 64, 48)
    assert u.lim_pos_x > 0 > u.lim_neg_x
    uk = build_uniforms(Camera(camera_model=KB4[0], model_params=KB4[1], **base), 64, 48)
    assert uk.lim_pos_x == uk.lim_neg_x == uk.lim_pos_y == uk.lim_neg_y == 0.0      # camera.rs:244-247
    ur = build_uniforms(Camera(camera_model=RT8[0], model_params=RT8[1], **base), 64, 48)
    # barrel distortion (k1 < 0): the undistorted bound is wider than the pixel-space one (camera.rs:229-243)
    d_edge = (1.15 * 64 - ur.cx) / ur.fx
    assert ur.lim_pos_x > d_edge > 0
    with pytest.raises(ValueError):
        build_uniforms(Camera(camera_model=cm.KANNALA_BRANDT_4, model_params=(0.1,), **base), 64, 48)


def test_det_atan2_accuracy():
    """orc_atan2f (a fixed IEEE operation sequence shared with the CUDA side) vs libm."""
    import ctypes as C
    lib = orc.lib()
    lib.orc_atan2f_det.argtypes = [C.c_float, C.c_float]
    lib.orc_atan2f_det.restype = C.c_float
    rng = np.random.default_rng(0)
    for _ in range(3000):
        y = float(np.float32(abs(rng.normal()) * 10 ** rng.uniform(-6, 2)))
        x = float(np.float32(rng.normal() * 10 ** rng.uniform(-6, 2)))
        assert abs(lib.orc_atan2f_det(y, x) - math.atan2(y, x)) < 4e-7
    assert lib.orc_atan2f_det(0.0, 1.0) == 0.0 and abs(lib.orc_atan2f_det(1.0, 0.0) - math.pi / 2) < 1e-7


@pytest.mark.parametrize("name", list(MODELS))
def test_render_smoke_finite(name):
    """tests/mod.rs:793-870: every model renders a small scene to finite pixels; something is visible."""
    model, params = MODELS[name]
    rng = np.random.default_rng(3)
    n = 64
    tr = np.zeros((n, 10), np.float32)
    tr[:, 0:3] = rng.uniform(-1, 1, (n, 3))
    tr[:, 3:7] = rng.uniform(-1, 1, (n, 4))
    tr[:, 7:10] = rng.uniform(-3, -1.5, (n, 3))
    sh = rng.uniform(0, 1, (n, 1, 3)).astype(np.float32)
    op = rng.uniform(1, 3, n).astype(np.float32)
    cam = Camera(position=(0, 0, -3.0), fov_x=0.7, fov_y=0.7, camera_model=model, model_params=params)
    r = orc.render_forward(build_uniforms(cam, 48, 48), 48, 48, tr, sh, op)
    assert np.isfinite(r.out_img).all() and r.num_visible > 32 and r.out_img[..., 3].max() > 0.1


@pytest.mark.parametrize("name", list(MODELS))
def test_small_distortion_approaches_pinhole(name):
    """With all coefficients zero RT8 is the pinhole model exactly; KB4 / TPF are the equidistant fisheye, which
    agrees with pinhole near the axis."""
    model, params = MODELS[name]
    cam_p, tr, sh, op = synthetic_scene(3000, 96, 64, k=1, seed=5)
    zero = tuple(0.0 for _ in params)
    cam_m = Camera(position=cam_p.position, rotation=cam_p.rotation, fov_x=0.3, fov_y=0.2, camera_model=model, model_params=zero)
    cam_0 = Camera(position=cam_p.position, rotation=cam_p.rotation, fov_x=0.3, fov_y=0.2)
    a = orc.render_forward(build_uniforms(cam_m, 96, 64), 96, 64, tr, sh, op).out_img
    b = orc.render_forward(build_uniforms(cam_0, 96, 64), 96, 64, tr, sh, op).out_img
    tol = 2e-4 if model == cm.RADIAL_TANGENTIAL_8 else 0.08
    assert np.abs(a - b).mean() < tol


@pytest.mark.parametrize("name", list(MODELS))
def test_finite_difference_camera_models(name):
    model, params = MODELS[name]
    _, tr, sh, op = finite_diff_base_scene()
    cam = Camera(position=(0.1, -0.15, -3.0), rotation=(0.03, -0.05, 0.02, 0.998), fov_x=0.9, fov_y=0.8,
                 camera_model=model, model_params=params)
    _check(cam, 40, 36, tr, sh, op, BROAD, rel=0.02, abs_tol=1e-4)


@pytest.mark.parametrize("name", ["rt8", "rt8_rational"])
def test_finite_difference_rt8_offaxis_wide(name):
    """Splats far off axis with a wide field of view: the full rational distortion + tangential terms matter."""
    model, params = MODELS[name]
    _, tr, sh, op = finite_diff_base_scene()
    tr = tr.copy()
    tr[:, 0] += 0.9
    tr[:, 1] -= 0.6
    cam = Camera(position=(0.0, 0.0, -2.5), fov_x=1.4, fov_y=1.2, camera_model=model, model_params=params)
    _check(cam, 48, 40, tr, sh, op, BROAD, rel=0.02, abs_tol=1e-4)
