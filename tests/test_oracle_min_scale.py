"""Oracle checks for the Mip-Splatting 3D-filter floor (compute_min_scale train.rs:102-125, fold_min_scale
gaussian_splats.rs:86-111).  No reference vector pins these values ("parity unpinned"): the oracle is
checked against a float64 numpy restatement of the same formulas and its backward against central differences."""
import numpy as np

from oracle import oracle as orc


def _scene(n, seed=5):
    rng = np.random.default_rng(seed)
    tr = np.zeros((n, 10), np.float32)
    tr[:, 0:3] = rng.uniform(-3, 3, (n, 3))
    tr[:, 3:7] = rng.uniform(-1, 1, (n, 4))
    tr[:, 7:10] = rng.uniform(np.log(0.004), np.log(0.3), (n, 3))
    op = rng.uniform(-3, 4, n).astype(np.float32)
    cams = np.concatenate([rng.uniform(-5, 5, (7, 3)), rng.uniform(400, 1800, (7, 1))], 1).astype(np.float32)
    return tr, op, cams


def _fold64(ls, raw, f):
    s2 = np.exp(2.0 * ls)
    s2f = s2 + (f * f)[:, None]
    coef = np.sqrt(s2.prod(1) / s2f.prod(1))
    o = np.clip(1.0 / (1.0 + np.exp(-raw)) * coef, 1e-6, 1 - 1e-6)
    return 0.5 * np.log(s2f), np.log(o / (1 - o))


def test_compute_min_scale_matches_definition():
    tr, _, cams = _scene(500)
    f = orc.compute_min_scale(tr, cams, 0.1)
    d = np.linalg.norm(tr[:, None, 0:3].astype(np.float64) - cams[None, :, 0:3], axis=2) / cams[None, :, 3]
    np.testing.assert_allclose(f, np.sqrt(0.1) * d.min(1), rtol=2e-6)


def test_fold_forward_matches_float64():
    tr, op, cams = _scene(2000)
    f = orc.compute_min_scale(tr, cams, 0.1) * 8.0  # exaggerate so the floor matters for many splats
    t2, o2 = orc.fold_min_scale(tr, op, f)
    ls64, raw64 = _fold64(tr[:, 7:10].astype(np.float64), op.astype(np.float64), f.astype(np.float64))
    assert np.array_equal(t2[:, :7], tr[:, :7])
    np.testing.assert_allclose(t2[:, 7:10], ls64, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(o2, raw64, rtol=2e-5, atol=2e-5)
    assert (t2[:, 7:10] >= tr[:, 7:10] - 1e-6).all()          # the floor only ever enlarges a splat
    assert (o2 <= op + 1e-4).all()                             # ... and dims it (energy compensation)
    # f = 0 is the identity (up to exp/log rounding)
    t0, o0 = orc.fold_min_scale(tr, op, np.zeros_like(f))
    np.testing.assert_allclose(t0, tr, rtol=0, atol=2e-6)
    np.testing.assert_allclose(o0, np.clip(op, -13.8, 13.8), rtol=1e-5, atol=1e-5)


def test_fold_backward_matches_central_differences():
    tr, op, cams = _scene(300, seed=9)
    f = orc.compute_min_scale(tr, cams, 0.1) * 8.0
    rng = np.random.default_rng(1)
    vt = rng.normal(size=tr.shape).astype(np.float32)
    vo = rng.normal(size=op.shape).astype(np.float32)
    gt, go = orc.fold_min_scale_backward(tr, op, f, vt, vo)
    assert np.array_equal(gt[:, :7], vt[:, :7])               # means / rotations pass through untouched
    ls, raw, f64 = tr[:, 7:10].astype(np.float64), op.astype(np.float64), f.astype(np.float64)

    def scalar(ls_, raw_):
        a, b = _fold64(ls_, raw_, f64)
        return (a * vt[:, 7:10]).sum(1) + b * vo              # per-splat objective: rows are independent

    eps = 1e-6
    for a in range(3):
        d = np.zeros_like(ls)
        d[:, a] = eps
        fd = (scalar(ls + d, raw) - scalar(ls - d, raw)) / (2 * eps)
        np.testing.assert_allclose(gt[:, 7 + a], fd, rtol=2e-3, atol=2e-4)
    fd = (scalar(ls, raw + eps) - scalar(ls, raw - eps)) / (2 * eps)
    np.testing.assert_allclose(go, fd, rtol=2e-3, atol=2e-4)
