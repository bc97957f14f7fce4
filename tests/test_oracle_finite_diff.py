"""Pins the oracle's backward pass the way the reference pins its own: analytical gradients vs central
finite differences with the C^1 smooth alpha cutoff (crates/brush-bench-test/tests/finite_diff.rs).
Scenes, eps and tolerances follow that file (base_scene :43-72, std_cam :74-83, eps 3e-4,
abs 5e-5 + rel 1% :217-219; weighted loss :465-507)."""
import numpy as np
import pytest

from brush_b200.camera import Camera, build_uniforms
from oracle import oracle as orc
from scenes import finite_diff_base_scene, splitmix64, synthetic_scene


def _loss_and_grads(cam, w, h, tr, sh, op, weights=None, mip=False):
    u = build_uniforms(cam, w, h)
    r = orc.render_forward(u, w, h, tr, sh, op, mip=mip, rpass=orc.PASS_BACKWARD_SMOOTH)
    if weights is None:
        v_out = np.full((h, w, 4), 1.0 / (h * w * 4), np.float32)
        loss = float(r.out_img.astype(np.float64).mean())
    else:
        v_out = weights
        loss = float((r.out_img.astype(np.float64) * weights).sum())
    return loss, r, v_out


def _check(cam, w, h, tr, sh, op, cases, weights=None, mip=False, eps=3e-4, rel=0.01, abs_tol=5e-5):
    _, r, v_out = _loss_and_grads(cam, w, h, tr, sh, op, weights, mip)
    _, vt, vsh, vo, _ = orc.render_backward(r, v_out)
    fails = []
    for kind, s, c in cases:
        def pert(d):
            t2, s2, o2 = tr.copy(), sh.copy(), op.copy()
            if kind == "t":
                t2[s, c] += d
            elif kind == "sh":
                s2[s, c // 3, c % 3] += d
            else:
                o2[s] += d
            return _loss_and_grads(cam, w, h, t2, s2, o2, weights, mip)[0]
        num = (pert(eps) - pert(-eps)) / (2 * eps)
        an = {"t": lambda: vt[s, c], "sh": lambda: vsh[s, c // 3, c % 3], "op": lambda: vo[s]}[kind]()
        tol = abs_tol + rel * max(abs(num), abs(an), 1e-8)
        if abs(num - an) > tol:
            fails.append(f"{kind}[{s},{c}] num {num:.6f} an {an:.6f}")
    assert not fails, "\n".join(fails)


BROAD = [("t", 0, 0), ("t", 0, 2), ("t", 1, 1), ("t", 0, 3), ("t", 1, 5), ("t", 0, 7), ("t", 1, 8),
         ("sh", 0, 0), ("sh", 1, 1), ("sh", 2, 2), ("op", 0, 0), ("op", 2, 0)]


def test_finite_difference_gradient_broad():
    cam, tr, sh, op = finite_diff_base_scene()
    _check(cam, 32, 32, tr, sh, op, BROAD)


def test_finite_difference_mip_mode():
    cam, tr, sh, op = finite_diff_base_scene()
    _check(cam, 32, 32, tr, sh, op, BROAD, mip=True, rel=0.02, abs_tol=1e-4)


def test_finite_difference_weighted_loss():
    cam, tr, sh, op = finite_diff_base_scene()
    w = splitmix64(0x51ED, 48 * 40 * 4).reshape(40, 48, 4).astype(np.float32) / (48 * 40)
    _check(cam, 48, 40, tr, sh, op, BROAD, weights=w, rel=0.02, abs_tol=1e-4)


def test_finite_difference_offcentre_rotated_camera():
    _, tr, sh, op = finite_diff_base_scene()
    cam = Camera(position=(0.4, -0.2, -3.2), rotation=(0.05, -0.08, 0.03, 0.995), fov_x=0.7, fov_y=0.5, center_uv=(0.45, 0.55))
    _check(cam, 48, 36, tr, sh, op, BROAD, rel=0.02, abs_tol=1e-4)


def test_finite_difference_sh_degree3_viewdir_path():
    """finite_diff.rs:1242-1389: higher SH bands make the colour depend on the mean through the view direction."""
    cam, tr, sh_dc, op = finite_diff_base_scene()
    sh = np.zeros((4, 16, 3), np.float32)
    sh[:, 0] = sh_dc[:, 0]
    sh[:, 1:] = (splitmix64(77, 4 * 15 * 3).reshape(4, 15, 3) - 0.5).astype(np.float32) * 0.6
    cases = [("t", 0, 0), ("t", 0, 1), ("t", 1, 2), ("t", 2, 0), ("sh", 0, 5), ("sh", 1, 20), ("sh", 3, 44), ("op", 1, 0)]
    _check(cam, 40, 40, tr, sh, op, cases, rel=0.02, abs_tol=1e-4)


def test_backward_gradients_finite_on_random_scenes():
    """fuzz.rs:493-553: finite gradients on random scenes."""
    for i in range(4):
        cam, tr, sh, op = synthetic_scene(2000, 96, 64, k=4, seed=900 + i)
        u = build_uniforms(cam, 96, 64)
        r = orc.render_forward(u, 96, 64, tr, sh, op)
        out = orc.render_backward(r, np.ones((64, 96, 4), np.float32))
        for g in out:
            assert np.isfinite(g).all()
        assert (out[4] >= 0).all()  # refine weight cleaned and clamped to [0, 1e32]
