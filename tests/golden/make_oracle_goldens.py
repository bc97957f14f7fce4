#!/usr/bin/env python
"""Writes tests/golden/oracle_camera_models.npz: forward images, counts and a gradient checksum of the ORACLE for one
small scene per camera model.  The reference has no stored vectors for the distorted models; these pin the
oracle against drift (tests/test_oracle_golden.py::test_oracle_matches_its_committed_camera_model_vectors) after it was
validated by finite differences.  Run from the repo root:  python tests/golden/make_oracle_goldens.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from brush_b200.camera import Camera, build_uniforms  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from scenes import random_v_output, synthetic_scene  # noqa: E402

MODELS = {
    "pinhole": (0, ()),
    "kb4": (1, (-0.05, 0.01, -0.001, 5e-5)),
    "rt8": (2, (-0.1, 0.03, -0.002, 0.05, -0.01, 0.001, 5e-3, -4e-3)),
    "tpf": (3, (-0.05, 0.01, -0.001, 5e-5, 1e-3, -1e-3, 5e-4, -5e-4)),
}
N, W, H, K = 1500, 64, 48, 4


def compute():
    out = {}
    cam0, tr, sh, op = synthetic_scene(N, W, H, k=K, seed=0x60DE)
    v_out = random_v_output(H, W, seed=11)
    for name, (model, params) in MODELS.items():
        cam = Camera(position=cam0.position, rotation=cam0.rotation, fov_x=1.1, fov_y=0.9, center_uv=(0.48, 0.53),
                     camera_model=model, model_params=params)
        r = orc.render_forward(build_uniforms(cam, W, H), W, H, tr, sh, op, bg=(0.1, 0.2, 0.3))
        _, vt, vsh, vo, _ = orc.render_backward(r, v_out)
        out[name + "_img"] = r.out_img.astype(np.float32)
        out[name + "_counts"] = np.array([r.num_visible, r.num_intersections], np.int64)
        out[name + "_grad_sums"] = np.array([vt.astype(np.float64).sum(), np.abs(vt.astype(np.float64)).sum(),
                                             vsh.astype(np.float64).sum(), vo.astype(np.float64).sum()])
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_camera_models.npz"), **compute())
    print("written")
