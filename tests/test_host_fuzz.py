"""Property tests (hypothesis) for the host-side formats and camera helpers: round trips must hold for arbitrary inputs."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from brush_b200 import camera as cm
from brush_b200 import dataset as ds
from brush_b200 import ply

finite32 = st.floats(min_value=-1e6, max_value=1e6, allow_nan=False, allow_infinity=False, width=32)


@settings(max_examples=40, deadline=None)
@given(n=st.integers(1, 40), degree=st.integers(0, 4), seed=st.integers(0, 2 ** 31 - 1), sub=st.integers(1, 5))
def test_ply_round_trip_any_shape(n, degree, seed, sub):
    rng = np.random.default_rng(seed)
    k = (degree + 1) ** 2
    t = rng.normal(size=(n, 10)).astype(np.float32)
    t[:, 3:7] += np.float32(0.1)                      # keep quaternions away from zero length
    sh = rng.normal(size=(n, k, 3)).astype(np.float32)
    op = rng.normal(size=n).astype(np.float32)
    d, meta = ply.load_splat_from_ply(ply.splat_to_ply(t, sh, op), subsample_points=sub)
    keep = np.arange(sub - 1, n, sub)
    assert d.num_splats() == len(keep) == meta.total_splats
    np.testing.assert_array_equal(d.sh_coeffs, sh[keep])
    np.testing.assert_array_equal(d.raw_opacities, op[keep])
    np.testing.assert_array_equal(d.means, t[keep, 0:3])
    np.testing.assert_array_equal(d.log_scales, t[keep, 7:10])
    np.testing.assert_allclose(np.linalg.norm(d.rotations, axis=1), 1.0, atol=1e-5)


@settings(max_examples=60, deadline=None)
@given(fov=st.floats(0.05, 2.6), px=st.integers(8, 8192))
def test_pinhole_fov_focal_round_trip(fov, px):
    f = cm.fov_to_focal(fov, px)
    assert abs(cm.focal_to_fov(f, px) - fov) < 1e-9


@settings(max_examples=60, deadline=None)
@given(fov=st.floats(0.1, 2.0), px=st.integers(16, 4096), k1=st.floats(-0.05, 0.05), k2=st.floats(-0.01, 0.01))
def test_fisheye_and_radial_fov_focal_round_trip(fov, px, k1, k2):
    kb = (k1, k2, 0.0, 0.0)
    f = cm.fov_to_focal(fov, px, cm.KANNALA_BRANDT_4, kb)
    assert abs(cm.focal_to_fov(f, px, cm.KANNALA_BRANDT_4, kb) - fov) < 1e-6
    rt = (k1, k2, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    if fov < 1.4:                                      # fixed-point undistortion converges for moderate fields of view
        f = cm.fov_to_focal(fov, px, cm.RADIAL_TANGENTIAL_8, rt)
        assert abs(cm.focal_to_fov(f, px, cm.RADIAL_TANGENTIAL_8, rt) - fov) < 1e-5


@settings(max_examples=50, deadline=None)
@given(q=st.tuples(finite32, finite32, finite32, finite32).filter(lambda q: 1e-3 < sum(v * v for v in q) < 1e12),
       p=st.tuples(st.floats(-100, 100, width=32), st.floats(-100, 100, width=32), st.floats(-100, 100, width=32)))
def test_world_to_local_maps_camera_centre_to_origin(q, p):
    n = math.sqrt(sum(v * v for v in q))
    cam = cm.Camera(position=p, rotation=tuple(v / n for v in q))
    vm = cam.world_to_local().reshape(4, 3)
    R, t = vm[:3].T.astype(np.float64), vm[3].astype(np.float64)
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=2e-5)
    np.testing.assert_allclose(R @ np.array(p, np.float64) + t, 0, atol=2e-3)


@settings(max_examples=40, deadline=None)
@given(h=st.integers(1, 9), w=st.integers(1, 9), seed=st.integers(0, 2 ** 31 - 1), alpha=st.booleans())
def test_packed_views_round_trip_bytes(h, w, seed, alpha):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 4 if alpha else 3), dtype=np.uint8)
    packed, has_alpha = ds.view_to_packed_data(img, ds.ALPHA_MASKED)
    assert has_alpha == alpha and packed.shape == (h, w)
    back = packed.view(np.uint32)
    un = np.stack([(back >> s) & 0xFF for s in (0, 8, 16, 24)], -1).astype(np.uint8)
    np.testing.assert_array_equal(un[..., :3], img[..., :3])
    np.testing.assert_array_equal(un[..., 3], img[..., 3] if alpha else 255)
    pre, _ = ds.view_to_packed_data(img, ds.ALPHA_TRANSPARENT)
    if alpha:
        a = img[..., 3].astype(np.uint16)
        want = ((img[..., 0].astype(np.uint16) * a + 127) // 255).astype(np.uint32)
        np.testing.assert_array_equal(pre.view(np.uint32) & 0xFF, want)
