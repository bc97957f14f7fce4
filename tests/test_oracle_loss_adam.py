"""Loss and optimiser oracle: structural properties the reference tests pin
(crates/brush-loss/tests/reference.rs:56-168: SSIM(x,x) ~ 1, SSIM in [-1,1], backward finite and
non-zero, 4-channel alpha path) plus float64 numpy restatements, because the reference stores no
numeric vector for these ("parity unpinned")."""
import numpy as np
import pytest

from oracle import oracle as orc


def _taps():
    x = np.arange(11, dtype=np.float64) - 5
    w = np.exp(-x * x / (2 * 1.5 * 1.5))
    return w / w.sum()


def _blur(img):  # zero padded separable 11-tap blur, float64
    w = _taps()
    h, wd = img.shape
    p = np.pad(img, 5)
    t = sum(w[i] * p[:, i:i + wd] for i in range(11))
    return sum(w[i] * t[i:i + h, :] for i in range(11))


def _loss64(pred, gt, l1_w, ssim_w):
    c1, c2 = 1e-4, 9e-4
    out = np.zeros_like(pred)
    for c in range(3):
        x, y = pred[c], gt[c]
        mu1, mu2 = _blur(x), _blur(y)
        s1 = np.maximum(0, _blur(x * x) - mu1 * mu1)
        s2 = np.maximum(0, _blur(y * y) - mu2 * mu2)
        s12 = _blur(x * y) - mu1 * mu2
        ssim = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2))
        out[c] = l1_w * np.abs(x - y) + ssim_w * np.clip(ssim, -1, 1)
    return out


def _rand_case(h, w, seed):
    rng = np.random.default_rng(seed)
    gt8 = rng.integers(0, 256, (h, w, 4), dtype=np.uint32)
    gt8[..., 3] = 255
    packed = (gt8[..., 0] | (gt8[..., 1] << 8) | (gt8[..., 2] << 16) | (gt8[..., 3] << 24)).astype(np.uint32)
    gt = (gt8[..., :3].astype(np.float64) / 255.0).transpose(2, 0, 1)
    pred = np.clip(gt + rng.normal(0, 0.1, gt.shape), 0, 1)
    return pred, gt, packed


@pytest.mark.parametrize("h,w", [(40, 56), (17, 23), (64, 64)])
def test_loss_forward_matches_float64(h, w):
    pred, gt, packed = _rand_case(h, w, h * w)
    got = orc.image_loss_forward(pred.astype(np.float32), packed, 0.8, -0.2)
    ref = _loss64(pred.astype(np.float32).astype(np.float64), np.round(gt * 255) / 255 * 1.0, 0.8, -0.2)
    assert np.abs(got - ref).max() < 2e-5


def test_ssim_identity_and_range():
    pred, gt, packed = _rand_case(48, 48, 3)
    gtf = (np.round(gt * 255) / 255).astype(np.float32)
    ssim_self = orc.image_loss_forward(gtf, packed, 0.0, 1.0)
    assert np.abs(ssim_self - 1.0).max() < 1e-4
    m = orc.image_loss_forward(pred.astype(np.float32), packed, 0.0, 1.0)
    assert m.min() >= -1.0 and m.max() <= 1.0


def test_loss_backward_matches_numeric_gradient():
    h, w = 24, 28
    pred, gt, packed = _rand_case(h, w, 9)
    pred32 = pred.astype(np.float32)
    rng = np.random.default_rng(1)
    dl = rng.uniform(0.2, 1.0, (3, h, w)).astype(np.float32)
    g = orc.image_loss_backward(pred32, packed, dl, 0.8, -0.2)
    assert np.isfinite(g).all() and np.abs(g).sum() > 0
    gtq = np.round(gt * 255) / 255
    base = pred32.astype(np.float64)
    eps = 1e-5
    for (c, y, x) in [(0, 0, 0), (1, 5, 7), (2, 23, 27), (0, 12, 13), (1, 0, 27), (2, 11, 0)]:
        p1, p2 = base.copy(), base.copy()
        p1[c, y, x] += eps
        p2[c, y, x] -= eps
        num = ((_loss64(p1, gtq, 0.8, -0.2) - _loss64(p2, gtq, 0.8, -0.2)) * dl).sum() / (2 * eps)
        assert abs(num - g[c, y, x]) < 2e-3 * max(1.0, abs(num)), (c, y, x, num, g[c, y, x])


def test_alpha_channel_composite_and_mask():
    h, w = 20, 20
    rng = np.random.default_rng(5)
    gt8 = rng.integers(0, 256, (h, w, 4), dtype=np.uint32)
    packed = (gt8[..., 0] | (gt8[..., 1] << 8) | (gt8[..., 2] << 16) | (gt8[..., 3] << 24)).astype(np.uint32)
    pred = rng.uniform(0, 1, (4, h, w)).astype(np.float32)
    m = orc.image_loss_forward(pred, packed, 1.0, 0.0, bg=(0.2, 0.4, 0.6), mask=False)
    ga = gt8[..., 3] / 255.0
    np.testing.assert_allclose(m[3], np.abs(pred[3] - ga), atol=1e-6)          # lib.rs:215-227
    for c, b in enumerate((0.2, 0.4, 0.6)):                                      # lib.rs:252-256
        eff = gt8[..., c] / 255.0 + (1 - ga) * b
        np.testing.assert_allclose(m[c], np.abs(pred[c] - eff), atol=2e-6)
    mm = orc.image_loss_forward(pred, packed, 1.0, 0.0, bg=None, mask=True)
    np.testing.assert_allclose(mm[0], np.abs(pred[0] - gt8[..., 0] / 255.0) * ga, atol=2e-6)
    g = orc.image_loss_backward(pred, packed, np.ones_like(pred), 1.0, 0.0, bg=None, mask=True)
    np.testing.assert_allclose(g[3], np.sign(pred[3] - ga.astype(np.float32)) * ga, atol=1e-6)


@pytest.mark.parametrize("reduce_v", [False, True])
def test_adam_matches_float64(reduce_v):
    """adam_scaled.rs:75-165 against a float64 restatement over 5 steps."""
    rng = np.random.default_rng(2)
    rows, cols = 300, 12
    p = rng.normal(0, 1, (rows, cols)).astype(np.float32)
    m = np.zeros_like(p)
    v = np.zeros(rows if reduce_v else (rows, cols), np.float32)
    scale = rng.uniform(0.1, 1.0, cols).astype(np.float32)
    p64, m64 = p.astype(np.float64), np.zeros((rows, cols))
    v64 = np.zeros((rows, 1)) if reduce_v else np.zeros((rows, cols))
    b1, b2, eps, lr = 0.9, 0.999, 1e-15, 2e-3
    f1, f2 = 1 - np.float64(np.float32(b1)), 1 - np.float64(np.float32(b2))
    for t in range(1, 6):
        g = rng.normal(0, 1e-3, (rows, cols)).astype(np.float32)
        orc.adam_step(p, g, m, v, lr, t, lr_scale_per_col=scale, reduce_v=reduce_v)
        g64 = g.astype(np.float64)
        gsq = (g64 * g64).mean(1, keepdims=True) if reduce_v else g64 * g64
        m64 = g64 * f1 if t == 1 else m64 * np.float32(b1) + g64 * f1
        v64 = gsq * f2 if t == 1 else v64 * np.float32(b2) + gsq * f2
        upd = (m64 / (1 - np.float64(np.float32(b1)) ** t)) / (np.sqrt(v64 / (1 - np.float64(np.float32(b2)) ** t)) + eps)
        p64 = p64 - upd * (scale.astype(np.float64) * np.float32(lr))
        np.testing.assert_allclose(p, p64, rtol=2e-5, atol=2e-6)
    assert np.isfinite(m).all() and np.isfinite(v).all()
