"""Pins the CPU oracle against the reference's own known-answer vectors
(crates/brush-bench-test/src/reference.rs:79-151; test_cases/{tiny,basic,mix}_case.safetensors,
committed unchanged under tests/golden/).  Tolerance = the reference's: 1e-5 + 1e-2*|ref| per element."""
import os

import numpy as np
import pytest

from brush_b200.camera import build_uniforms
from oracle import oracle as orc
from scenes import golden_case


@pytest.mark.parametrize("name,n_expected", [("tiny_case", 4), ("basic_case", 16), ("mix_case", 76873)])
def test_oracle_reproduces_golden(golden_dir, name, n_expected):
    cam, tr, sh, op, ref, (w, h) = golden_case(os.path.join(golden_dir, f"{name}.safetensors"))
    assert tr.shape[0] == n_expected and ref.shape == (82, 123, 4)
    u = build_uniforms(cam, w, h)
    r = orc.render_forward(u, w, h, tr, sh, op, rpass=orc.PASS_BACKWARD)
    assert not np.isnan(r.out_img).any()
    err = np.abs(r.out_img - ref)
    assert (err < 1e-5 + 1e-2 * np.abs(ref)).all(), f"{name}: max err {err.max()}"
    # the restatement is in fact far inside the reference's band
    assert err.max() < 5e-6
    # RenderOutput::validate_counts (render_aux.rs:30-45)
    assert r.num_visible <= tr.shape[0]
    assert r.num_intersections <= r.num_visible * r.tiles_x * r.tiles_y


def test_packed_output_matches_float(golden_dir):
    """rasterize.rs:173-180: packed = trunc(clamp(v*255, 0, 255)) of the float result."""
    cam, tr, sh, op, ref, (w, h) = golden_case(os.path.join(golden_dir, "basic_case.safetensors"))
    u = build_uniforms(cam, w, h)
    f = orc.render_forward(u, w, h, tr, sh, op, rpass=orc.PASS_BACKWARD)
    p = orc.render_forward(u, w, h, tr, sh, op, rpass=orc.PASS_FORWARD)
    exp = np.clip(f.out_img * np.float32(255.0), 0, 255).astype(np.uint32)
    got = np.stack([(p.out_packed >> s) & 0xFF for s in (0, 8, 16, 24)], -1)
    np.testing.assert_array_equal(got, exp)
    # forward pass leaves the tile ranges untrimmed
    np.testing.assert_array_equal(p.tile_offsets, p.tile_offsets_untrimmed)


def test_oracle_matches_its_committed_camera_model_vectors():
    """Drift pin: tests/golden/oracle_camera_models.npz was written by tests/golden/make_oracle_goldens.py after the
    oracle's camera models passed their finite-difference checks."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_oracle_goldens", os.path.join(here, "golden", "make_oracle_goldens.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    got = mod.compute()
    want = np.load(os.path.join(here, "golden", "oracle_camera_models.npz"))
    for name in mod.MODELS:
        assert (got[name + "_counts"] == want[name + "_counts"]).all(), name
        np.testing.assert_allclose(got[name + "_img"], want[name + "_img"], rtol=0, atol=2e-6, err_msg=name)
        np.testing.assert_allclose(got[name + "_grad_sums"], want[name + "_grad_sums"], rtol=1e-5, err_msg=name)
    assert want["kb4_counts"][0] > 500 and not np.array_equal(want["kb4_img"], want["pinhole_img"])
