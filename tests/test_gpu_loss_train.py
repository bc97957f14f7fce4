"""GPU parity for the loss, optimiser and train-step kernels through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from scenes import synthetic_scene  # noqa: E402


@pytest.fixture(scope="module")
def rt():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import brush_b200.loss as L
    import brush_b200.render as R
    import brush_b200.train as T
    from oracle import oracle as orc

    class RT:
        pass

    r = RT()
    r.R, r.L, r.T, r.orc = R, L, T, orc
    r.ctx = R.RenderContext(max_splats=1 << 18, max_w=512, max_h=512)
    yield r
    r.ctx.close()


def _case(h, w, seed, alpha=False):
    rng = np.random.default_rng(seed)
    gt8 = rng.integers(0, 256, (h, w, 4), dtype=np.uint32)
    if not alpha:
        gt8[..., 3] = 255
    packed = (gt8[..., 0] | (gt8[..., 1] << 8) | (gt8[..., 2] << 16) | (gt8[..., 3] << 24)).astype(np.uint32)
    pred = rng.uniform(0, 1, (h, w, 4)).astype(np.float32)
    return pred, packed


@pytest.mark.parametrize("h,w,channels,bg,mask", [(64, 64, 3, None, False), (45, 77, 3, None, False),
                                                  (50, 33, 4, (0.2, 0.4, 0.6), False), (40, 40, 4, None, True),
                                                  (200, 311, 3, (0.1, 0.0, 0.3), True)])
def test_image_loss_forward_backward(rt, h, w, channels, bg, mask):
    pred, packed = _case(h, w, h + w, alpha=channels == 4 or mask or bg is not None)
    d = rt.ctx.device
    cfg = rt.L.ImageLossConfig(0.8, -0.2, bg, mask)
    tp = torch.from_numpy(pred).to(d)
    tg = torch.from_numpy(packed.view(np.int32)).to(d)
    m = rt.L.image_loss_forward(rt.ctx, tp, tg, channels, cfg).cpu().numpy()
    pred_chw = np.ascontiguousarray(pred.transpose(2, 0, 1)[:channels])
    om = rt.orc.image_loss_forward(pred_chw, packed, 0.8, -0.2, bg=bg, mask=mask)
    assert np.abs(m - om).max() < 2e-6
    dl = np.random.default_rng(1).uniform(0.1, 1.0, (channels, h, w)).astype(np.float32)
    g = rt.L.image_loss_backward(rt.ctx, tp, tg, torch.from_numpy(dl).to(d), channels, cfg).cpu().numpy()
    og = rt.orc.image_loss_backward(pred_chw, packed, dl, 0.8, -0.2, bg=bg, mask=mask)
    gg = g.transpose(2, 0, 1)
    assert np.abs(gg[:channels] - og).max() < 2e-5 * max(1.0, np.abs(og).max())
    if channels == 3:
        assert (gg[3] == 0).all()


@pytest.mark.parametrize("h,w,channels,bg,mask", [(64, 64, 3, None, False), (45, 77, 3, None, False), (97, 130, 4, (0.2, 0.4, 0.6), False),
                                                  (33, 31, 4, None, True), (200, 311, 3, None, False)])
def test_image_loss_fused_matches_forward_mean_and_backward(rt, h, w, channels, bg, mask):
    """bg_image_loss_fused == image_loss_forward -> weighted mean -> image_loss_backward (train.rs:238-260)."""
    pred, packed = _case(h, w, 3 * h + w, alpha=channels == 4 or mask or bg is not None)
    d = rt.ctx.device
    cfg = rt.L.ImageLossConfig(0.8, -0.2, bg, mask)
    tp = torch.from_numpy(pred).to(d)
    tg = torch.from_numpy(packed.view(np.int32)).to(d)
    npx = float(h * w)
    chain = [1.0 / (3 * npx)] * 3 + ([0.1 / npx] if channels == 4 else [])
    g, loss = rt.L.image_loss_fused(rt.ctx, tp, tg, channels, cfg, chain)
    pred_chw = np.ascontiguousarray(pred.transpose(2, 0, 1)[:channels])
    om = rt.orc.image_loss_forward(pred_chw, packed, 0.8, -0.2, bg=bg, mask=mask)
    exp_loss = sum(chain[c] * om[c].astype(np.float64).sum() for c in range(channels))
    assert abs(float(loss.item()) - exp_loss) <= 2e-5 * max(1.0, abs(exp_loss))
    dl = np.stack([np.full((h, w), chain[c], np.float32) for c in range(channels)])
    og = rt.orc.image_loss_backward(pred_chw, packed, dl, 0.8, -0.2, bg=bg, mask=mask)
    gg = g.cpu().numpy().transpose(2, 0, 1)
    assert np.abs(gg[:channels] - og).max() < 2e-5 * max(np.abs(og).max(), 1e-12)
    if channels == 3:
        assert (gg[3] == 0).all()


def test_image_loss_chw_layout_equals_hwc(rt):
    """The reference feeds a CHW-permuted tensor (lib.rs:1076); strides make both layouts equivalent."""
    pred, packed = _case(48, 52, 7)
    d = rt.ctx.device
    cfg = rt.L.ImageLossConfig(0.8, -0.2)
    tg = torch.from_numpy(packed.view(np.int32)).to(d)
    hwc = torch.from_numpy(pred[..., :3].copy()).to(d)
    chw_view = hwc.permute(2, 0, 1).contiguous().permute(1, 2, 0)  # CHW memory, HWC indexing
    a = rt.L.image_loss_forward(rt.ctx, hwc, tg, 3, cfg)
    b = rt.L.image_loss_forward(rt.ctx, chw_view, tg, 3, cfg)
    assert torch.equal(a, b)


@pytest.mark.parametrize("rows,cols,reduce_v,scaled", [(1000, 10, False, True), (777, 48, True, True), (5000, 1, False, False),
                                                       (130, 75, True, True)])
def test_adam_vs_oracle(rt, rows, cols, reduce_v, scaled):
    from brush_b200 import _lib
    rng = np.random.default_rng(rows)
    d = rt.ctx.device
    p = rng.normal(0, 1, (rows, cols)).astype(np.float32)
    m = np.zeros_like(p)
    v = np.zeros(rows if reduce_v else (rows, cols), np.float32)
    scale = rng.uniform(0.1, 1.0, cols).astype(np.float32) if scaled else None
    tp, tm, tv = (torch.from_numpy(x.copy()).to(d) for x in (p, m, v))
    ts = torch.from_numpy(scale).to(d) if scaled else None
    lib = _lib.load()
    for t in range(1, 5):
        g = rng.normal(0, 1e-3, (rows, cols)).astype(np.float32)
        rt.orc.adam_step(p, g, m, v, 2e-3, t, lr_scale_per_col=scale, reduce_v=reduce_v)
        tg = torch.from_numpy(g).to(d)
        _lib.check(lib.bg_adam_step(rt.ctx.handle, torch.cuda.current_stream().cuda_stream, tp.data_ptr(), tg.data_ptr(),
                                    tm.data_ptr(), tv.data_ptr(), rows, cols, ts.data_ptr() if scaled else None, 2e-3, 0.9,
                                    0.999, 1e-15, t, int(reduce_v)), "bg_adam_step")
        np.testing.assert_allclose(tp.cpu().numpy(), p, rtol=3e-6, atol=1e-7)
        np.testing.assert_allclose(tm.cpu().numpy(), m, rtol=2e-4, atol=1e-9)  # GPU contracts m*b1+g*f1 into an FMA
        np.testing.assert_allclose(tv.cpu().numpy(), v, rtol=2e-4, atol=1e-13)


def test_train_steps_run_and_reduce_loss(rt):
    """integration.rs:185-312 style smoke + a sanity check that optimisation makes progress."""
    n, w, h = 20_000, 192, 128
    cam, tr, sh, op = synthetic_scene(n, w, h, k=4, seed=99)
    d = rt.ctx.device
    # target = render of the scene; start = perturbed colours/opacities
    tgt = rt.R.render_splats(rt.ctx, cam, (w, h), *(torch.from_numpy(x).to(d) for x in (tr, sh, op)), rpass=0)
    gt_packed = (tgt.out_img | (255 << 24)).clone()
    rng = np.random.default_rng(0)
    sh2 = sh + rng.normal(0, 0.3, sh.shape).astype(np.float32)
    splats = rt.T.Splats(torch.from_numpy(tr).to(d), torch.from_numpy(sh2).to(d), torch.from_numpy(op).to(d))
    bounds = rt.T.bounds_from_pos(0.8, tr[:, :3])
    cfg = rt.T.TrainConfig(background_noise_strength=0.0, total_train_iters=100)
    trainer = rt.T.SplatTrainer(cfg, rt.ctx, bounds)
    batch = rt.T.SceneBatch(img_packed=gt_packed, camera=cam)
    losses = []
    for _ in range(30):
        st = trainer.step(batch, splats)
        losses.append(float(st.loss.item()))
    assert all(np.isfinite(losses))
    assert losses[-1] < losses[0] * 0.9, losses
    for t in (splats.transforms, splats.sh_coeffs, splats.raw_opacities):
        assert torch.isfinite(t).all()
    s = trainer._state
    assert s["vis_weight"].max() > 0 and s["refine_norm"].min() >= 0 and s["max_screen"].max() > 0


def test_train_step_first_update_matches_oracle(rt):
    """One step from zero optimiser state: parameters after the step equal the oracle pipeline
    (render -> loss -> backward -> Adam) within float tolerance; noise disabled for the comparison."""
    from brush_b200.camera import build_uniforms
    n, w, h = 4000, 96, 80
    cam, tr, sh, op = synthetic_scene(n, w, h, k=4, seed=123)
    rng = np.random.default_rng(3)
    gt8 = rng.integers(0, 256, (h, w, 3), dtype=np.uint32)
    packed = (gt8[..., 0] | (gt8[..., 1] << 8) | (gt8[..., 2] << 16) | (255 << 24)).astype(np.uint32)
    d = rt.ctx.device
    splats = rt.T.Splats(*(torch.from_numpy(x.copy()).to(d) for x in (tr, sh, op)))
    bounds = rt.T.bounds_from_pos(0.8, tr[:, :3])
    cfg = rt.T.TrainConfig(background_noise_strength=0.0, mean_noise_weight=0.0)
    trainer = rt.T.SplatTrainer(cfg, rt.ctx, bounds)
    trainer.step(rt.T.SceneBatch(img_packed=torch.from_numpy(packed.view(np.int32)), camera=cam), splats)
    # oracle
    o = rt.orc.render_forward(build_uniforms(cam, w, h), w, h, tr, sh, op)
    pred_chw = np.ascontiguousarray(o.out_img.transpose(2, 0, 1)[:3])
    dl = np.full((3, h, w), 1.0 / (3 * h * w), np.float32)
    gpred = rt.orc.image_loss_backward(pred_chw, packed, dl, 0.8, -0.2)
    v_out = np.zeros((h, w, 4), np.float32)
    v_out[..., :3] = gpred.transpose(1, 2, 0)
    _, vt, vsh, vo, _ = rt.orc.render_backward(o, v_out)
    lr_mean = 2e-5 * bounds.median_size()
    p_t, p_sh, p_o = tr.copy(), sh.reshape(n, -1).copy(), op.copy().reshape(n, 1)
    lr_t = np.array([lr_mean] * 3 + [2e-3] * 4 + [5e-3] * 3, np.float32)
    rt.orc.adam_step(p_t, vt, np.zeros_like(p_t), np.zeros_like(p_t), 1.0, 1, lr_scale_per_col=lr_t)
    sc = np.repeat(np.array([1.0] + [np.float32(1.0) / np.float32(10.0)] * 3, np.float32), 3)
    rt.orc.adam_step(p_sh, vsh.reshape(n, -1).copy(), np.zeros_like(p_sh), np.zeros(n, np.float32), 2e-3, 1, lr_scale_per_col=sc, reduce_v=True)
    rt.orc.adam_step(p_o, vo.reshape(n, 1).copy(), np.zeros_like(p_o), np.zeros_like(p_o), 0.012, 1)
    # Adam's first step moves every touched element by ~lr regardless of gradient size, so compare
    # where the gradient is not vanishing (sign-stable) and allow float slack elsewhere
    def close(a, b, g, lr):
        a, b = a.reshape(-1), b.reshape(-1)
        stable = np.abs(g.reshape(-1)) > 1e-9
        assert np.abs(a - b)[stable].max() <= 1e-3 * lr + 1e-6, np.abs(a - b)[stable].max()
        assert (np.abs(a - b) <= 2.5 * lr + 1e-6).all()
    close(splats.transforms.cpu().numpy()[:, 7:], p_t[:, 7:], vt[:, 7:], 5e-3)
    close(splats.transforms.cpu().numpy()[:, 3:7], p_t[:, 3:7], vt[:, 3:7], 2e-3)
    close(splats.raw_opacities.cpu().numpy(), p_o, vo, 0.012)
    close(splats.sh_coeffs.cpu().numpy()[:, 0], p_sh.reshape(n, 4, 3)[:, 0], vsh[:, 0], 2e-3)


def test_train_refine_train_cycle(rt):
    """integration.rs:185-312 style: steps, a refine that changes N, more steps; everything stays finite and
    the optimizer / refine-record state follows the new size."""
    n, w, h = 20_000, 192, 128
    cam, tr, sh, op = synthetic_scene(n, w, h, k=4, seed=42)
    op[:500] = -9.0  # dead splats: pruned at refine, budget re-used by splits
    d = rt.ctx.device
    tgt = rt.R.render_splats(rt.ctx, cam, (w, h), *(torch.from_numpy(x).to(d) for x in (tr, sh, op)), rpass=0)
    gt = (tgt.out_img | (255 << 24)).clone()
    splats = rt.T.Splats(*(torch.from_numpy(x.copy()).to(d) for x in (tr, sh + 0.1, op)))
    trainer = rt.T.SplatTrainer(rt.T.TrainConfig(total_train_iters=1000, refine_every=5), rt.ctx, rt.T.bounds_from_pos(0.8, tr[:, :3]))
    batch = rt.T.SceneBatch(img_packed=gt, camera=cam)
    for _ in range(5):
        trainer.step(batch, splats)
    stats = trainer.refine(5, splats)
    assert stats.num_pruned >= 500 and stats.total_splats == splats.num_splats()
    assert stats.num_added > 0
    for _ in range(5):
        st = trainer.step(batch, splats)
    assert np.isfinite(float(st.loss.item()))
    for t in (splats.transforms, splats.sh_coeffs, splats.raw_opacities):
        assert torch.isfinite(t).all() and t.shape[0] == stats.total_splats
    assert trainer._state["m_t"].shape[0] == stats.total_splats


def test_min_scale_kernels_vs_oracle(rt):
    """compute_min_scale / fold_min_scale (+ its reverse-mode chain) through the C ABI vs the oracle."""
    n = 50_000
    rng = np.random.default_rng(3)
    tr = np.zeros((n, 10), np.float32)
    tr[:, 0:3] = rng.uniform(-3, 3, (n, 3))
    tr[:, 3:7] = rng.uniform(-1, 1, (n, 4))
    tr[:, 7:10] = rng.uniform(np.log(0.004), np.log(0.3), (n, 3))
    op = rng.uniform(-3, 4, n).astype(np.float32)
    cams = np.concatenate([rng.uniform(-5, 5, (300, 3)), rng.uniform(400, 1800, (300, 1))], 1).astype(np.float32)
    d = rt.ctx.device
    ttr, top, tc = torch.from_numpy(tr).to(d), torch.from_numpy(op).to(d), torch.from_numpy(cams).to(d)
    f = rt.T.compute_min_scale(rt.ctx, ttr, tc, 0.1)
    of = rt.orc.compute_min_scale(tr, cams, 0.1)
    np.testing.assert_allclose(f.cpu().numpy(), of, rtol=1e-6)
    assert rt.T.compute_min_scale(rt.ctx, ttr, tc[:0], 0.1) is None and rt.T.compute_min_scale(rt.ctx, ttr, tc, 0.0) is None
    f8 = (f * 8.0).contiguous()
    t2, o2 = rt.T.fold_min_scale(rt.ctx, ttr, top, f8)
    ot2, oo2 = rt.orc.fold_min_scale(tr, op, of * np.float32(8.0))
    assert torch.equal(t2[:, :7], ttr[:, :7])
    np.testing.assert_allclose(t2.cpu().numpy(), ot2, rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(o2.cpu().numpy(), oo2, rtol=1e-5, atol=1e-5)
    vt = rng.normal(size=tr.shape).astype(np.float32)
    vo = rng.normal(size=op.shape).astype(np.float32)
    gvt, gvo = torch.from_numpy(vt).to(d), torch.from_numpy(vo).to(d)
    rt.T.fold_min_scale_backward(rt.ctx, ttr, top, f8, gvt, gvo)
    ovt, ovo = rt.orc.fold_min_scale_backward(tr, op, of * np.float32(8.0), vt, vo)
    np.testing.assert_allclose(gvt.cpu().numpy(), ovt, rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(gvo.cpu().numpy(), ovo, rtol=2e-4, atol=2e-5)
    # bake: in place, floor dropped
    s = rt.T.Splats(ttr.clone(), torch.zeros((n, 1, 3), device=d), top.clone(), min_scale=f8)
    s.bake_min_scale(rt.ctx)
    assert s.min_scale is None and torch.equal(s.transforms, t2) and torch.equal(s.raw_opacities, o2)


def test_train_with_min_scale_floor(rt):
    """The floor is folded in for the render, its gradient chained back, baked at refine and recomputed after
    (train.rs:432-437, 641-647): a step with a floor of zero equals a step without one."""
    n, w, h = 20_000, 192, 128
    cam, tr, sh, op = synthetic_scene(n, w, h, k=4, seed=7)
    d = rt.ctx.device
    tgt = rt.R.render_splats(rt.ctx, cam, (w, h), *(torch.from_numpy(x).to(d) for x in (tr, sh, op)), rpass=0)
    gt = (tgt.out_img | (255 << 24)).clone()
    batch = rt.T.SceneBatch(img_packed=gt, camera=cam)
    cfg = rt.T.TrainConfig(total_train_iters=1000, background_noise_strength=0.0, mean_noise_weight=0.0)

    def run(min_scale):
        s = rt.T.Splats(*(torch.from_numpy(x.copy()).to(d) for x in (tr, sh + 0.1, op)), min_scale=min_scale)
        grads = []
        t = rt.T.SplatTrainer(cfg, rt.ctx, rt.T.bounds_from_pos(0.8, tr[:, :3]),
                              grad_hook=lambda g: grads.extend(x.clone() for x in g[:3]))
        t.step(batch, s)
        return grads

    g_ref = run(None)
    g_zero = run(torch.zeros(n, dtype=torch.float32, device=d))
    # f = 0: the fold is an exp/log round trip only -> the same gradients up to f32 rounding
    for a, b in zip(g_zero, g_ref):
        rel = (a.double() - b.double()).norm() / b.double().norm()
        assert rel < 1e-3, rel
    # a real floor: trains, refines, floor re-attached with the new N
    s = rt.T.Splats(*(torch.from_numpy(x.copy()).to(d) for x in (tr, sh + 0.1, op)))
    t = rt.T.SplatTrainer(rt.T.TrainConfig(total_train_iters=1000), rt.ctx, rt.T.bounds_from_pos(0.8, tr[:, :3]))
    focal = 0.5 * w / np.tan(0.5 * cam.fov_x)
    t.set_view_cams([(cam.position, focal), ((0.5, 0.0, -1.0), focal)])
    s.min_scale = rt.T.compute_min_scale(rt.ctx, s.transforms, t.view_cams, rt.T.MIN_SCALE_FACTOR)
    for _ in range(4):
        st = t.step(batch, s)
    stats = t.refine(4, s)
    assert s.min_scale is not None and s.min_scale.shape[0] == stats.total_splats == s.num_splats()
    for _ in range(3):
        st = t.step(batch, s)
    assert np.isfinite(float(st.loss.item()))
    assert all(torch.isfinite(x).all() for x in (s.transforms, s.sh_coeffs, s.raw_opacities, s.min_scale))


def test_step_views_equals_sequential_accumulation(rt):
    """SURVEY 8e parity definition: the multi-view step equals the single-GPU step that accumulates the views'
    gradients sequentially (mean over views), up to f32 summation order."""
    import math
    from brush_b200.camera import Camera
    n, w, h = 20_000, 192, 128
    cam0, tr, sh, op = synthetic_scene(n, w, h, k=9, seed=21)
    a = math.radians(4.0) / 2.0
    cam1 = Camera(position=(0.1, -0.05, 0.0), rotation=(0.0, math.sin(a), 0.0, math.cos(a)), fov_x=cam0.fov_x, fov_y=cam0.fov_y,
                  center_uv=cam0.center_uv)
    d = rt.ctx.device
    dev_params = lambda: [torch.from_numpy(x.copy()).to(d) for x in (tr, sh, op)]
    batches = []
    for cam in (cam0, cam1):
        tgt = rt.R.render_splats(rt.ctx, cam, (w, h), *dev_params(), rpass=0)
        batches.append(rt.T.SceneBatch(img_packed=(tgt.out_img | (255 << 24)).clone(), camera=cam))
    cfg = rt.T.TrainConfig(total_train_iters=1000, background_noise_strength=0.0, mean_noise_weight=50.0, seed=11)
    captured = {}

    class Capture(rt.T.SplatTrainer):
        def _apply_updates(self, splats, v_t, v_sh, v_o, v_r, visible, max_radius, median_scale):
            captured.update(v_t=v_t.clone(), v_sh=v_sh.clone(), v_o=v_o.clone(), v_r=v_r.clone(), vis=visible.clone(),
                            rad=max_radius.clone())
            return 0.0   # gradients only: the parameters stay untouched

    bounds = rt.T.bounds_from_pos(0.8, tr[:, :3])
    p = dev_params()
    sh_start = p[1] + 0.1
    fresh = lambda: rt.T.Splats(p[0].clone(), sh_start.clone(), p[2].clone())
    # the step under test: ONE ABI call (bg_train_step_views), SH gradient kept factored, fused update pass
    multi = fresh()
    t_multi = rt.T.SplatTrainer(cfg, rt.ctx, bounds)
    st = t_multi.step_views(batches, multi)
    # the definition: per-view dense gradients through the per-operator entry points, averaged, MAX / SUM statistics,
    # then the same update pass on the dense gradient
    per_view = []
    for b in batches:
        Capture(cfg, rt.ctx, bounds).step(b, fresh())
        per_view.append(dict(captured))
    avg = {k: ((per_view[0][k].double() + per_view[1][k].double()) / 2.0).float() for k in ("v_t", "v_sh", "v_o")}
    ref = fresh()
    t_ref = rt.T.SplatTrainer(cfg, rt.ctx, bounds)
    t_ref._ensure_state(ref)
    t_ref.step_count = 1
    t_ref._apply_updates(ref, avg["v_t"], avg["v_sh"], avg["v_o"], torch.maximum(per_view[0]["v_r"], per_view[1]["v_r"]),
                         per_view[0]["vis"] + per_view[1]["vis"], torch.maximum(per_view[0]["rad"], per_view[1]["rad"]),
                         bounds.median_size())
    torch.cuda.synchronize()
    for name in ("transforms", "sh_coeffs", "raw_opacities"):
        a, b = getattr(ref, name).double(), getattr(multi, name).double()
        assert torch.isfinite(b).all()
        # Adam's first step moves every entry by ~lr * sign(g): entries whose tiny gradient changes sign with the
        # summation order differ by 2 lr; everything else agrees closely
        close = (a - b).abs() <= 1e-6 + 1e-4 * a.abs()
        assert close.double().mean() > 0.995, (name, float(close.double().mean()))
    for key in ("m_t", "m_sh", "m_o", "v_sh"):   # the moments are linear / quadratic in the gradient: direct comparison
        a, b = t_ref._state[key].double(), t_multi._state[key].double()
        assert (a - b).norm() / a.norm() < 1e-4, key
    assert torch.equal(t_multi._state["vis_weight"], per_view[0]["vis"] + per_view[1]["vis"])
    assert torch.equal(t_multi._state["max_screen"], torch.maximum(per_view[0]["rad"], per_view[1]["rad"]))
    torch.testing.assert_close(t_multi._state["refine_norm"], torch.maximum(per_view[0]["v_r"], per_view[1]["v_r"]), rtol=1e-4, atol=1e-7)
    assert np.isfinite(float(st.loss.item()))
    # a second and third step keep working on the moments (t > 1 path) and the chunked update equals the unchunked one
    multi_c = fresh()
    t_c = rt.T.SplatTrainer(cfg, rt.ctx, bounds)
    t_c.step_views(batches, multi_c, chunks=3)   # chunks only matter with a communicator; must be accepted and ignored
    for name in ("transforms", "sh_coeffs", "raw_opacities"):   # (two runs differ by the order of the blend's f32 atomics)
        a, b = getattr(multi_c, name).double(), getattr(multi, name).double()
        assert ((a - b).abs() <= 1e-6 + 1e-4 * a.abs()).double().mean() > 0.995, name
    for _ in range(2):
        st = t_multi.step_views(batches, multi)
    assert np.isfinite(float(st.loss.item()))
    assert all(torch.isfinite(x).all() for x in (multi.transforms, multi.sh_coeffs, multi.raw_opacities))


def test_eval_stats_vs_oracle(rt):
    """eval.rs:22-61: PSNR / SSIM of a render against a ground-truth image, through the product kernels vs the oracle."""
    from brush_b200.eval import eval_stats
    n, w, h = 15_000, 160, 120
    cam, tr, sh, op = synthetic_scene(n, w, h, k=4, seed=31)
    d = rt.ctx.device
    rng = np.random.default_rng(2)
    gt = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    splats = rt.T.Splats(*(torch.from_numpy(x).to(d) for x in (tr, sh, op)))
    s = eval_stats(rt.ctx, splats, cam, gt)
    from brush_b200.camera import build_uniforms
    o = rt.orc.render_forward(build_uniforms(cam, w, h), w, h, tr, sh, op, bg=(0, 0, 0))
    rgb = np.round(o.out_img[..., :3] * np.float32(255.0)) / np.float32(255.0)
    packed = (gt[..., 0].astype(np.uint32) | gt[..., 1].astype(np.uint32) << 8 | gt[..., 2].astype(np.uint32) << 16 | np.uint32(255 << 24))
    chw = np.ascontiguousarray(rgb.transpose(2, 0, 1)).astype(np.float32)
    l1 = rt.orc.image_loss_forward(chw, packed, 1.0, 0.0)
    psnr = 10.0 * np.log10(1.0 / np.mean(l1.astype(np.float64) ** 2))
    ssim = float(rt.orc.image_loss_forward(chw, packed, 0.0, 1.0).astype(np.float64).mean())
    assert abs(float(s.psnr) - psnr) < 2e-3 and abs(float(s.ssim) - ssim) < 2e-4
    # a perfect reconstruction of its own 8-bit image: SSIM 1, PSNR very high
    own = (s.rendered.clamp(0, 1) * 255.0).round().to(torch.uint8).cpu().numpy()
    s2 = eval_stats(rt.ctx, splats, cam, own)
    assert float(s2.ssim) > 0.9999 and float(s2.psnr) > 60.0


def test_train_step_fused_abi_matches_host_orchestrated_step(rt):
    """bg_train_step (one ABI call per SplatTrainer::step, train.rs:176-429) against the step orchestrated from the host
    over the per-operator entry points: same losses, same parameters up to the f32 atomics of the blend backward."""
    n, w, h = 20_000, 192, 128
    cam, tr, sh, op = synthetic_scene(n, w, h, k=4, seed=123)
    d = rt.ctx.device
    tgt = rt.R.render_splats(rt.ctx, cam, (w, h), *(torch.from_numpy(x).to(d) for x in (tr, sh, op)), rpass=0)
    gt = (tgt.out_img | (255 << 24)).clone()
    batch = rt.T.SceneBatch(img_packed=gt, camera=cam)
    cfg = rt.T.TrainConfig(total_train_iters=1000, background_noise_strength=0.0, seed=7)
    bounds = rt.T.bounds_from_pos(0.8, tr[:, :3])

    def run(fused):
        s = rt.T.Splats(*(torch.from_numpy(x.copy()).to(d) for x in (tr, sh + 0.1, op)))
        t = rt.T.SplatTrainer(cfg, rt.ctx, bounds)
        losses = []
        for _ in range(3):
            st = (t.step_fused if fused else t.step)(batch, s)
            losses.append(float(st.loss.item()))
        return s, t, losses

    s_a, t_a, l_a = run(False)
    s_b, t_b, l_b = run(True)
    np.testing.assert_allclose(l_b, l_a, rtol=2e-4)
    for name in ("transforms", "sh_coeffs", "raw_opacities"):
        a, b = getattr(s_a, name).double(), getattr(s_b, name).double()
        assert torch.isfinite(b).all()
        # Adam's first steps move every entry by ~lr * sign(g): entries whose tiny gradient changes sign with the order
        # of the atomics differ by 2 lr; everything else agrees closely
        close = (a - b).abs() <= 1e-6 + 1e-4 * a.abs()
        assert close.double().mean() > 0.995, (name, float(close.double().mean()))
    for key in ("refine_norm", "vis_weight", "max_screen"):
        a, b = t_a._state[key].double(), t_b._state[key].double()
        assert ((a - b).abs() <= 1e-6 + 1e-3 * a.abs()).double().mean() > 0.999, key
    # the draw is a pure function of (seed, offset): the same stream twice, a different one for another seed
    lib = rt.T._lib.load()
    z1, z2, z3 = (torch.empty(10_001, dtype=torch.float32, device=d) for _ in range(3))
    for z, seed in ((z1, 7), (z2, 7), (z3, 8)):
        rt.T._lib.check(lib.bg_normal_noise(rt.ctx.handle, None, seed, 5, z.numel(), z.data_ptr()), "bg_normal_noise")
    torch.cuda.synchronize()
    assert torch.equal(z1, z2) and not torch.equal(z1, z3)
    assert abs(float(z1.mean())) < 0.05 and abs(float(z1.std()) - 1.0) < 0.05 and torch.isfinite(z1).all()
