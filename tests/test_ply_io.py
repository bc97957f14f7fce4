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


# ---------------------------------------------------------------------------------------------- SuperSplat compressed PLY
def test_quant_decoders_known_answers():
    """brush-serde/src/quant.rs:73-112 restated."""
    u = lambda *v: np.array(v, np.uint32)
    assert np.array_equal(ply.decode_vec_11_10_11(u(0)), np.zeros((1, 3), np.float32))
    np.testing.assert_allclose(ply.decode_vec_11_10_11(u((0x7FF << 21) | (0x3FF << 11) | 0x7FF)), np.ones((1, 3)), atol=1e-6)
    assert np.array_equal(ply.decode_vec_8_8_8_8(u(0)), np.zeros((1, 4), np.float32))
    np.testing.assert_allclose(ply.decode_vec_8_8_8_8(u(0xFFFFFFFF)), np.ones((1, 4)), atol=1e-6)
    vals = u(*[(i * 42949673) & 0xFFFFFFFF for i in range(100)])
    for dec in (ply.decode_vec_11_10_11, ply.decode_vec_8_8_8_8):
        out = dec(vals)
        assert out.min() >= 0.0 and out.max() <= 1.0
    q = ply.decode_quat(u((512 << 20) | (512 << 10) | 512))
    assert abs(np.linalg.norm(q) - 1.0) < 1e-5 and np.isfinite(q).all()
    # field order: bits 24..31 -> x, 16..23 -> y, 8..15 -> z, 0..7 -> w; 11 | 10 | 11 from the top
    np.testing.assert_allclose(ply.decode_vec_8_8_8_8(u(0xFF000000))[0], [1, 0, 0, 0])
    np.testing.assert_allclose(ply.decode_vec_8_8_8_8(u(0x000000FF))[0], [0, 0, 0, 1])
    np.testing.assert_allclose(ply.decode_vec_11_10_11(u(0x7FF << 21))[0], [1, 0, 0])
    np.testing.assert_allclose(ply.decode_vec_11_10_11(u(0x7FF))[0], [0, 0, 1])


def test_decode_quat_places_the_dropped_component():
    """quant.rs:37-71: `largest` indexes (w, x, y, z); the three stored values fill the other slots in order."""
    norm = 0.5 * np.sqrt(2.0)
    enc = lambda f: int(round((f * norm + 0.5) * 1023))
    a, b, c = 0.3, -0.2, 0.1
    for largest in range(4):
        word = np.array([(largest << 30) | (enc(a) << 20) | (enc(b) << 10) | enc(c)], np.uint32)
        q = ply.decode_quat(word)[0]
        others = [q[i] for i in range(4) if i != largest]
        np.testing.assert_allclose(others, [a, b, c], atol=2e-3)
        assert q[largest] > 0.9 and abs(np.linalg.norm(q) - 1.0) < 1e-3


def _encode_compressed(means, log_scales, quats_wxyz, rgb, opacity, sh_rest=None, endian="<"):
    """Test-side encoder of the SuperSplat layout (the reference only decodes): returns the PLY bytes."""
    n = means.shape[0]
    n_chunks = (n + 255) // 256
    meta = np.zeros((n_chunks, 18), np.float32)
    pos_w, scl_w, col_w = (np.zeros(n, np.uint32) for _ in range(3))
    rot_w = np.zeros(n, np.uint32)

    def pack_111011(u01):
        a = np.round(u01[:, 0] * 2047).astype(np.uint32)
        b = np.round(u01[:, 1] * 1023).astype(np.uint32)
        c = np.round(u01[:, 2] * 2047).astype(np.uint32)
        return (a << 21) | (b << 11) | c

    for ci in range(n_chunks):
        sl = slice(ci * 256, min(n, ci * 256 + 256))
        rows = []
        for arr, base in ((means, 0), (log_scales, 6), (rgb, 12)):
            lo, hi = arr[sl].min(0), arr[sl].max(0)
            hi = np.where(hi > lo, hi, lo + 1.0)
            for ax in range(3):
                meta[ci, base + 2 * ax], meta[ci, base + 2 * ax + 1] = lo[ax], hi[ax]
            rows.append((arr[sl] - lo) / (hi - lo))
        pos_w[sl], scl_w[sl] = pack_111011(rows[0]), pack_111011(rows[1])
        c8 = np.round(rows[2] * 255).astype(np.uint32)
        a8 = np.round(opacity[sl] * 255).astype(np.uint32)
        col_w[sl] = (c8[:, 0] << 24) | (c8[:, 1] << 16) | (c8[:, 2] << 8) | a8
    q = quats_wxyz / np.linalg.norm(quats_wxyz, axis=1, keepdims=True)
    largest = np.abs(q).argmax(1)
    q = q * np.sign(q[np.arange(n), largest])[:, None]
    norm = 0.5 * np.sqrt(2.0)
    for i in range(n):
        three = [q[i, j] for j in range(4) if j != largest[i]]
        w = [int(round((v * norm + 0.5) * 1023)) for v in three]
        rot_w[i] = (int(largest[i]) << 30) | (w[0] << 20) | (w[1] << 10) | w[2]
    fmt = "binary_little_endian" if endian == "<" else "binary_big_endian"
    head = ["ply", f"format {fmt} 1.0", "comment Vertical axis: y", f"element chunk {n_chunks}"]
    head += [f"property float {nm}" for nm in ply._QUANT_META_FIELDS]
    head += [f"element vertex {n}"] + [f"property uint packed_{nm}" for nm in ("position", "rotation", "scale", "color")]
    body = meta.astype(endian + "f4").tobytes()
    body += np.stack([pos_w, rot_w, scl_w, col_w], 1).astype(endian + "u4").tobytes()
    if sh_rest is not None:
        head += [f"element sh {n}"] + [f"property uchar f_rest_{i}" for i in range(sh_rest.shape[1])]
        body += np.clip(np.round((sh_rest / 8.0 + 0.5) * 254.0), 0, 255).astype(np.uint8).tobytes()
    head.append("end_header")
    return ("\n".join(head) + "\n").encode() + body, q


@pytest.mark.parametrize("n,per,endian", [(1, 0, "<"), (700, 0, "<"), (513, 3, "<"), (300, 15, ">")])
def test_compressed_ply_import(n, per, endian):
    """import.rs:408-600 through an encoder written for the test: every field comes back within its quantisation step,
    rows use the ranges of chunk i // 256, higher SH bands are de-interleaved from channel-major."""
    rng = np.random.default_rng(n)
    means = rng.normal(size=(n, 3)).astype(np.float32) * 3
    log_scales = rng.uniform(-6, -1, size=(n, 3)).astype(np.float32)
    quats = rng.normal(size=(n, 4))
    rgb = rng.uniform(0, 1, size=(n, 3)).astype(np.float32)
    opac = rng.uniform(0.05, 0.95, size=n).astype(np.float32)
    rest = rng.uniform(-1.5, 1.5, size=(n, 3 * per)).astype(np.float32) if per else None
    data, qn = _encode_compressed(means, log_scales, quats, rgb, opac, rest, endian)
    d, meta = ply.load_splat_from_ply(data)
    assert d.num_splats() == n and meta.total_splats == n and meta.up_axis == (0.0, -1.0, 0.0)
    span = lambda a: np.maximum(a.max(0) - a.min(0), 1e-6)
    assert np.abs(d.means - means).max() <= span(means).max() / 1023 + 1e-5
    assert np.abs(d.log_scales - log_scales).max() <= span(log_scales).max() / 1023 + 1e-5
    np.testing.assert_allclose(d.rotations, qn, atol=2.5e-3)           # (w, x, y, z), unit length up to 10-bit steps
    sig = 1.0 / (1.0 + np.exp(-d.raw_opacities.astype(np.float64)))
    assert np.abs(sig - opac).max() <= 0.5 / 255 + 1e-6                # inverse sigmoid of the 8-bit alpha
    got_rgb = d.sh_coeffs[:, 0, :] * ply.SH_C0 + 0.5                   # rgb_to_sh inverted
    assert np.abs(got_rgb - rgb).max() <= 0.5 / 255 + 1e-5
    assert d.sh_coeffs.shape == (n, 1 + per, 3)
    if per:
        want = rest.reshape(n, 3, per).transpose(0, 2, 1)              # channel-major on disk -> [n, k, 3]
        assert np.abs(d.sh_coeffs[:, 1:, :] - want).max() <= 4.0 / 254 + 1e-5
    t, sh, op = d.into_arrays()
    assert t.shape == (n, 10) and sh.shape[0] == n and op.shape == (n,)


def test_compressed_ply_subsample_and_errors():
    rng = np.random.default_rng(5)
    n = 600
    args = (rng.normal(size=(n, 3)).astype(np.float32), rng.uniform(-5, -1, size=(n, 3)).astype(np.float32), rng.normal(size=(n, 4)),
            rng.uniform(0, 1, size=(n, 3)).astype(np.float32), rng.uniform(0.1, 0.9, size=n).astype(np.float32))
    rest = rng.uniform(-1, 1, size=(n, 9)).astype(np.float32)
    data, _ = _encode_compressed(*args, rest)
    full, _ = ply.load_splat_from_ply(data)
    sub, meta = ply.load_splat_from_ply(data, subsample_points=4)
    assert sub.num_splats() == n // 4 == meta.total_splats
    np.testing.assert_array_equal(sub.means, full.means[3::4])         # rows 4, 8, ... (1-based), each with ITS chunk's ranges
    np.testing.assert_array_equal(sub.sh_coeffs, full.sh_coeffs[3::4])
    # truncated body
    with pytest.raises(ValueError):
        ply.load_splat_from_ply(data[:-100])
    # a chunk element without one of the range fields
    broken = data.replace(b"property float max_b\n", b"property float max_q\n")
    with pytest.raises(ValueError, match="max_b"):
        ply.load_splat_from_ply(broken)
    # chunk element but no vertex element
    head = b"ply\nformat binary_little_endian 1.0\nelement chunk 0\nproperty float min_x\nend_header\n"
    with pytest.raises(ValueError, match="Unknown format"):
        ply.load_splat_from_ply(head)
