"""Cull / count / visibility invariants of the forward path, restated from the reference's fuzz and
render tests and checked on the oracle:
  crates/brush-bench-test/tests/fuzz.rs:61-86,331-487 (poison values, bad geometry fully culled,
  valid-but-extreme not culled), brush-render/src/tests/mod.rs:19-71 (near-plane cull -> black),
  :314-388 (hidden / culled-prefix splats leave the image unchanged), :675-708 (zero quat culled)."""
import numpy as np
import pytest

from brush_b200.camera import Camera, build_uniforms
from oracle import oracle as orc
from scenes import synthetic_scene

W = H = 64
POISON = [np.nan, np.inf, -np.inf, 3.0e38, -3.0e38, 1e-45]


def _cam():
    return Camera(position=(0, 0, -3.0), rotation=(0, 0, 0, 1), fov_x=0.8, fov_y=0.8)


def _one(tr_row, op=2.0, sh=(0.5, 0.5, 0.5)):
    tr = np.array([tr_row], np.float32)
    return tr, np.array([[sh]], np.float32).reshape(1, 1, 3), np.array([op], np.float32)


BASE = [0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, -1.5, -1.5, -1.5]


def _render(tr, sh, op, **kw):
    return orc.render_forward(build_uniforms(_cam(), W, H), W, H, tr, sh, op, **kw)


def test_base_splat_is_visible():
    r = _render(*_one(BASE))
    assert r.num_visible == 1 and r.num_intersections > 0 and r.out_img[..., 3].max() > 0.5
    assert r.visible[0] == 1.0 and r.max_radius[0] > 0


@pytest.mark.parametrize("slot", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("val", [np.nan, np.inf, -np.inf])
def test_nonfinite_geometry_is_culled(slot, val):
    row = list(BASE)
    row[slot] = val
    r = _render(*_one(row))
    assert r.num_visible == 0 and r.num_intersections == 0
    assert np.isfinite(r.out_img).all() and (r.out_img == 0).all()


@pytest.mark.parametrize("val", [np.nan, np.inf])
def test_nonfinite_scale_and_opacity_culled(val):
    row = list(BASE)
    row[8] = val
    assert _render(*_one(row)).num_visible == 0
    assert _render(*_one(BASE, op=val if np.isnan(val) else -val)).num_visible == 0 or np.isinf(val)
    assert _render(*_one(BASE, op=np.nan)).num_visible == 0


def test_zero_quat_culled():
    row = list(BASE)
    row[3:7] = [0, 0, 0, 0]
    assert _render(*_one(row)).num_visible == 0
    row[3:7] = [1e-4, 0, 0, 0]  # |q|^2 = 1e-8 < 1e-6
    assert _render(*_one(row)).num_visible == 0


def test_near_plane_and_behind_camera_culled():
    for z in (-3.0 + 0.005, -4.0):  # z_cam = 0.005 (< 0.01) and -1
        row = list(BASE)
        row[2] = z
        r = _render(*_one(row))
        assert r.num_visible == 0 and (r.out_img == 0).all()


def test_low_opacity_culled():
    r = _render(*_one(BASE, op=-6.0))  # sigmoid(-6) = 0.0025 < 1/255
    assert r.num_visible == 0


def test_valid_but_extreme_not_culled():
    """fuzz.rs:451-487: log_scale in [-30, 40] and huge colours stay visible and finite."""
    for ls in (-30.0, 40.0):
        row = list(BASE)
        row[7:10] = [ls, ls, ls]
        r = _render(*_one(row, sh=(1e37, -1e37, 5.0)))
        assert r.num_visible == 1
        assert np.isfinite(r.out_img).all()
        assert np.isfinite(r.projected).all()
        assert abs(r.projected[0, 6]) <= 100.0  # colour clamp +-100 (project_visible.rs:56-71)


def test_offscreen_culled_and_counts_valid():
    cam, tr, sh, op = synthetic_scene(5000, 128, 96, k=4, seed=5)
    tr[:100, 0] += 1e4  # far off to the side
    u = build_uniforms(cam, 128, 96)
    r = orc.render_forward(u, 128, 96, tr, sh, op)
    assert not np.isin(np.arange(100), r.gid_from_cgid).any()
    assert r.num_intersections == int(r.intersect_counts[r.gid_from_cgid].sum())
    # depth order and stable ties
    d = r.depths_sorted
    assert (np.diff(d) >= 0).all()
    # tile ranges partition the intersection list
    to = r.tile_offsets_untrimmed.reshape(-1, 2).astype(np.int64)
    nz = to[to[:, 1] > to[:, 0]]
    assert (nz[1:, 0] == nz[:-1, 1]).all() and nz[0, 0] == 0 and nz[-1, 1] == r.num_intersections
    # inside a tile, splats are in ascending compact id (= depth) order
    for lo, hi in nz[:50]:
        assert (np.diff(r.cgid_from_isect[lo:hi].astype(np.int64)) > 0).all()
    # trimmed ends never exceed the untrimmed ends
    assert (r.tile_offsets[..., 1] <= r.tile_offsets_untrimmed[..., 1]).all()


def test_hidden_splats_leave_image_unchanged():
    """tests/mod.rs:314-388: splats behind an opaque wall, and culled splats, do not change the image."""
    # two large opaque walls: the first leaves T = 1e-3 (alpha cap 0.999), the second trips the
    # T' <= 1e-4 stop rule, so nothing behind them is ever blended.
    wall = np.array([[0, 0, 0.0, 1, 0, 0, 0, 5.0, 5.0, -4.0], [0, 0, 0.5, 1, 0, 0, 0, 5.0, 5.0, -4.0]], np.float32)
    wall_sh = np.full((2, 1, 3), 1.0, np.float32)
    wall_op = np.array([12.0, 12.0], np.float32)
    a = _render(wall, wall_sh, wall_op)
    behind = np.array([[0.1, 0.1, 2.0, 1, 0, 0, 0, -1.0, -1.0, -1.0], [0, 0, -9.0, 1, 0, 0, 0, -1, -1, -1]], np.float32)
    tr = np.concatenate([behind, wall])
    sh = np.concatenate([np.full((2, 1, 3), -1.0, np.float32), wall_sh])
    op = np.concatenate([np.array([3.0, 3.0], np.float32), wall_op])
    b = _render(tr, sh, op)
    assert b.num_visible == 3 and b.visible[0] == 0.0  # present in the lists but never blended
    centre = (slice(16, 48), slice(16, 48))
    assert np.abs(a.out_img[centre] - b.out_img[centre]).max() < 1e-5


def test_forward_is_deterministic():
    cam, tr, sh, op = synthetic_scene(3000, 96, 64, k=9, seed=11)
    u = build_uniforms(cam, 96, 64)
    a = orc.render_forward(u, 96, 64, tr, sh, op)
    b = orc.render_forward(u, 96, 64, tr, sh, op)
    np.testing.assert_array_equal(a.out_img.view(np.uint32), b.out_img.view(np.uint32))


def test_sort_and_scan_spec():
    """brush-sort/src/lib.rs:147-151 and brush-prefix-sum/src/lib.rs:91-189."""
    rng = np.random.default_rng(0)
    for n, bits in [(15, 32), (1000, 13), (100_000, 32), (5000, 5)]:
        k = rng.integers(0, 2 ** 32 - 1, n, dtype=np.uint64).astype(np.uint32)
        k[: n // 2] %= 50
        v = rng.integers(0, 2 ** 32 - 1, n, dtype=np.uint64).astype(np.uint32)
        ko, vo = orc.radix_argsort(k, v, bits)
        mask = np.uint32((1 << bits) - 1) if bits < 32 else np.uint32(0xFFFFFFFF)
        order = np.argsort(k & mask, kind="stable")
        np.testing.assert_array_equal(ko, k[order])
        np.testing.assert_array_equal(vo, v[order])
    x = rng.integers(0, 1000, 512 * 16 + 123).astype(np.uint32)
    np.testing.assert_array_equal(orc.inclusive_scan(x), np.cumsum(x).astype(np.uint32))


def test_deterministic_exp_log_accuracy():
    """The shared exp/log recipe stays within 1.5 ulp of the correctly rounded result on the ranges the
    path uses (log-scales in [-30,40], sigmoid arguments, ln(255*opacity) in [0, 5.6])."""
    xs = np.concatenate([np.linspace(-30, 40, 20001), np.linspace(-88, 88, 4001)]).astype(np.float32)
    got = np.array([orc.expf_det(float(x)) for x in xs], np.float32)
    ref = np.exp(xs.astype(np.float64))
    ulp = np.abs(got.astype(np.float64) - ref) / np.spacing(ref.astype(np.float32)).astype(np.float64)
    assert ulp.max() < 1.5
    ys = np.linspace(1.0, 255.0, 20001).astype(np.float32)
    gl = np.array([orc.logf_det(float(y)) for y in ys], np.float32)
    rl = np.log(ys.astype(np.float64))
    assert (np.abs(gl - rl) <= 1.5 * np.spacing(np.maximum(np.abs(rl), 1e-3).astype(np.float32))).all()
    assert orc.expf_det(float("inf")) == float("inf") and orc.expf_det(-float("inf")) == 0.0
    assert np.isnan(orc.expf_det(float("nan"))) and orc.expf_det(89.0) == float("inf")
