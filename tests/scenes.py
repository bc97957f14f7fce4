"""Deterministic synthetic scenes shared by the CPU and GPU tests and by bench.py.

`synthetic_scene` is the generator fixed in SURVEY.md section 8(d): camera at the origin looking +z,
pinhole fov_x = 60 deg, means drawn in NDC x log-uniform depth, log-scales U(ln .004, ln .03), quats
U(-1,1)^4 un-normalised, raw opacity U(-2,4), SH DC U(-1,1.5), higher bands U(-.25,.25);
SplitMix64 stream seeded with 0xB2000000 + config index.
"""
from __future__ import annotations

import math

import numpy as np

from brush_b200.camera import Camera, fov_to_focal, focal_to_fov


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n uniform doubles in [0,1) from the SplitMix64 stream (brush-render/src/tests/mod.rs:168-222 style)."""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def synthetic_scene(n: int, w: int, h: int, k: int = 16, seed: int = 0xB2000001, scale_shift: float = 0.0):
    fov_x = math.radians(60.0)
    focal = fov_to_focal(fov_x, w)
    fov_y = focal_to_fov(focal, h)
    cam = Camera(position=(0.0, 0.0, 0.0), rotation=(0.0, 0.0, 0.0, 1.0), fov_x=fov_x, fov_y=fov_y)
    per = 3 + 3 + 4 + 1 + 3 * k
    r = splitmix64(seed, n * per).reshape(n, per)
    u = (r[:, 0] * 2.1 - 1.05)
    v = (r[:, 1] * 2.1 - 1.05)
    z = np.exp(math.log(2.0) + r[:, 2] * (math.log(12.0) - math.log(2.0)))
    x = u * z * math.tan(fov_x / 2)
    y = v * z * math.tan(fov_y / 2)
    means = np.stack([x, y, z], 1)
    lo, hi = math.log(0.004) + scale_shift, math.log(0.03) + scale_shift
    log_scales = lo + r[:, 3:6] * (hi - lo)
    quats = r[:, 6:10] * 2.0 - 1.0
    raw_opac = r[:, 10] * 6.0 - 2.0
    sh = r[:, 11:].reshape(n, k, 3).copy()
    sh[:, 0, :] = sh[:, 0, :] * 2.5 - 1.0
    if k > 1:
        sh[:, 1:, :] = sh[:, 1:, :] * 0.5 - 0.25
    transforms = np.concatenate([means, quats, log_scales], 1).astype(np.float32)
    return cam, transforms, sh.astype(np.float32), raw_opac.astype(np.float32)


def random_v_output(h: int, w: int, seed: int = 0xB2000101) -> np.ndarray:
    """Upstream gradient like finite_diff.rs:465-479: U(0,1) weights on all four channels."""
    return splitmix64(seed, h * w * 4).reshape(h, w, 4).astype(np.float32)


def finite_diff_base_scene():
    """base_scene() / std_cam() of crates/brush-bench-test/tests/finite_diff.rs:43-83."""
    means = np.array([0.20, -0.10, 0.00, -0.30, 0.40, 0.20, 0.10, 0.30, -0.30, -0.20, -0.20, 0.10], np.float32).reshape(4, 3)
    rots = np.array([0.90, 0.10, 0.05, 0.03, 0.70, 0.20, 0.30, 0.10, 0.50, 0.40, 0.30, 0.20, 0.80, 0.10, 0.10, 0.20], np.float32).reshape(4, 4)
    log_scales = np.array([-1.4, -1.5, -1.6, -1.5, -1.4, -1.3, -1.7, -1.5, -1.4, -1.3, -1.6, -1.5], np.float32).reshape(4, 3)
    sh_dc = np.array([0.45, 0.55, 0.50, 0.60, 0.40, 0.30, 0.35, 0.50, 0.65, 0.50, 0.45, 0.55], np.float32).reshape(4, 1, 3)
    raw_opac = np.array([2.5, 2.0, 2.2, 2.4], np.float32)
    cam = Camera(position=(0.0, 0.0, -3.0), rotation=(0.0, 0.0, 0.0, 1.0), fov_x=0.6, fov_y=0.6)
    return cam, np.concatenate([means, rots, log_scales], 1), sh_dc, raw_opac


def golden_case(path: str):
    """Inputs + reference image of one crates/brush-bench-test/test_cases/*.safetensors (reference.rs:79-151)."""
    from safetensors.numpy import load_file

    d = load_file(path)
    ref = d["out_img"]
    h, w, _ = ref.shape
    fov = math.pi * 0.5
    focal = fov_to_focal(fov, w)
    cam = Camera(position=(0.123, 0.456, -8.0), rotation=(0.0, 0.0, 0.0, 1.0), fov_x=focal_to_fov(focal, w), fov_y=focal_to_fov(focal, h))
    transforms = np.concatenate([d["means"], d["quats"], d["scales"]], 1).astype(np.float32)
    return cam, transforms, d["coeffs"].astype(np.float32), d["opacities"].astype(np.float32), ref, (w, h)
