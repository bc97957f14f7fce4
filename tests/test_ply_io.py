"""PLY export / import, following brush-serde's own tests (export.rs:244-345, import.rs:612-740)."""
import numpy as np
import pytest

from brush_b200 import ply


def _splats(n, degree, seed=0):
    rng = np.random.default_rng(seed)
    k = ply.sh_coeffs_for_degree(degree)
    t = rng.normal(size=(n, 10)).astype(np.float32)
    sh = rng.normal(size=(n, k, 3)).astype(np.float32)
    op = rng.normal(size=n).astype(np.float32)
    return t, sh, op


@pytest.mark.parametrize("degree,rest_fields", [(0, 0), (1, 9), (2, 24), (3, 45)])
def test_ply_field_count_matches_sh_degree(degree, rest_fields):
    t, sh, op = _splats(1, degree)
    data = ply.splat_to_ply(t, sh, op)
    head = data[:data.index(b"end_header")].decode()
    assert head.count("property float f_rest_") == rest_fields
    assert "f_dc_0" in head and f"SH degree: {degree}" in head and "SplatRenderMode: default" in head
    assert ("f_rest_0" in head) == (rest_fields > 0) and f"f_rest_{rest_fields}\n" not in head
    names = [ln.split()[-1] for ln in head.splitlines() if ln.startswith("property")]
    assert names[:14] == ["x", "y", "z", "scale_0", "scale_1", "scale_2", "opacity", "rot_0", "rot_1", "rot_2", "rot_3",
                          "f_dc_0", "f_dc_1", "f_dc_2"]
    assert len(data) - data.index(b"end_header\n") - 11 == 4 * len(names)


@pytest.mark.parametrize("degree", [0, 1, 2, 3])
def test_export_roundtrip(degree):
    t, sh, op = _splats(100, degree, seed=degree)
    data = ply.splat_to_ply(t, sh, op, up_axis=(0.0, -1.0, 0.5), render_mip=True)
    d, meta = ply.load_splat_from_ply(data)
    assert d.num_splats() == 100 and meta.total_splats == 100 and meta.render_mip is True
    assert meta.up_axis == (0.0, -1.0, 0.5)
    t2, sh2, op2 = d.into_arrays()
    assert sh2.shape == sh.shape
    np.testing.assert_array_equal(sh2, sh)                  # coefficient ordering survives the channel-major layout
    np.testing.assert_array_equal(op2, op)
    np.testing.assert_array_equal(t2[:, [0, 1, 2, 7, 8, 9]], t[:, [0, 1, 2, 7, 8, 9]])
    q = t[:, 3:7] / np.linalg.norm(t[:, 3:7], axis=1, keepdims=True)
    np.testing.assert_allclose(t2[:, 3:7], q, rtol=1e-6, atol=1e-7)     # exported normalised


def test_import_positions_and_uchar_colours_with_defaults():
    """The shape of apps/brush-c/tests/data/test_dataset/init.ply: xyz + uchar rgba, nothing else."""
    n = 7
    rng = np.random.default_rng(1)
    xyz = rng.normal(size=(n, 3)).astype("<f4")
    rgba = rng.integers(0, 256, (n, 4), dtype=np.uint8)
    head = ("ply\nformat binary_little_endian 1.0\ncomment made by a test\nelement vertex %d\nproperty float x\n"
            "property float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n"
            "property uchar alpha\nend_header\n" % n).encode()
    body = b"".join(xyz[i].tobytes() + rgba[i].tobytes() for i in range(n))
    d, meta = ply.load_splat_from_ply(head + body)
    assert meta.up_axis is None and meta.render_mip is None
    np.testing.assert_array_equal(d.means, xyz)
    assert d.rotations is None and d.log_scales is None and d.raw_opacities is None
    want = (rgba[:, :3].astype(np.float32) / np.float32(254.0) - np.float32(0.5)) / np.float32(ply.SH_C0)
    np.testing.assert_allclose(d.sh_coeffs[:, 0, :], want, rtol=1e-6)
    t, sh, op = d.into_arrays()
    assert (t[:, 3:7] == [1, 0, 0, 0]).all() and (t[:, 7:10] == -4.0).all() and (op == 0).all()


def test_import_ascii_and_subsample_and_up_axis():
    rows = ["%g %g %g %g" % (i, 2 * i, -i, 0.1 * i) for i in range(10)]
    txt = ("ply\nformat ascii 1.0\ncomment Vertical axis: z\nelement vertex 10\nproperty float x\nproperty float y\n"
           "property float z\nproperty float opacity\nend_header\n" + "\n".join(rows) + "\n").encode()
    d, meta = ply.load_splat_from_ply(txt, subsample_points=3)
    assert meta.up_axis == (0.0, 0.0, -1.0) and meta.total_splats == 3
    np.testing.assert_array_equal(d.means[:, 0], [2, 5, 8])              # rows 3, 6, 9 (1-based multiples of 3)
    np.testing.assert_allclose(d.raw_opacities, [0.2, 0.5, 0.8], rtol=1e-6)
    s = ply.SplatData(means=np.arange(30, dtype=np.float32).reshape(10, 3)).subsample(4)
    assert s.num_splats() == 4 and (s.means[:, 0] == [0, 9, 18, 27]).all()       # step = ceil(10/4) = 3
    assert ply.SplatData(means=np.zeros((5, 3), np.float32)).subsample(0).num_splats() == 5


def test_errors():
    with pytest.raises(ValueError):
        ply.load_splat_from_ply(b"not a ply")
    with pytest.raises(ValueError):
        ply.splat_to_ply(np.zeros((1, 10), np.float32), np.zeros((1, 5, 3), np.float32), np.zeros(1, np.float32))
