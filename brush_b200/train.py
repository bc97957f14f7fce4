"""Host mirror of brush-train's per-step path over the C ABI.

  TrainConfig         <- brush-train/src/config.rs:5-132 (the fields the step uses, same defaults)
  Splats              <- brush-render/src/gaussian_splats.rs:57-74 (packed [N,10] / [N,K,3] / [N])
  SceneBatch          <- brush-dataset/src/scene.rs:138-162
  SplatTrainer.step   <- brush-train/src/train.rs:176-429
  bounds_from_pos / BoundingBox.median_size <- splat_init.rs:130-160, bounding_box.rs:23-29

Per step: render forward -> fused L1+SSIM loss -> loss backward -> rasterize/project backward ->
[optional gradient all-reduce hook for view-sharded data parallelism] -> Adam on the three parameter
tensors -> refine statistics + mean noise.  All device work goes through libbrush_b200.so; torch
provides memory, streams and (for N>1) torch.distributed.

  SplatTrainer.refine <- brush-train/src/train.rs:431-893 (prune / resample / force-split / grow / split /
                         opacity decay / bounds), with generic tensor ops exactly as the reference does
                         (it has no custom kernels for refine), but device-resident: weighted sampling
                         without replacement runs on the GPU (seeded Efraimidis-Spirakis keys + top-k)
                         instead of a full host readback + rand::sample_weighted (multinomial.rs:1-26).

  Splats.min_scale / set_view_cams <- the Mip-Splatting 3D-filter floor: compute_min_scale (train.rs:102-125),
                         fold_min_scale / bake_min_scale (gaussian_splats.rs:86-111, 245-252); the floor is folded
                         into scales/opacity for every render, its gradient chained back in place, baked at the
                         start of refine() and recomputed at its end while progress < 0.9 (train.rs:437, 641-647).

Out of scope here: LPIPS.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .camera import Camera
from .loss import ImageLossConfig, image_loss_fused
from .render import (PASS_BACKWARD, RenderContext, _stream_ptr, project_bwd, project_bwd_factored, rasterize_bwd,
                     render_splats, sh_grad_from_views)


@dataclass
class TrainConfig:
    total_train_iters: int = 30000
    lr_mean: float = 2e-5
    lr_mean_end: float = 2e-7
    mean_noise_weight: float = 50.0
    lr_coeffs_dc: float = 2e-3
    lr_coeffs_sh_scale: float = 10.0
    lr_opac: float = 0.012
    lr_scale: float = 5e-3
    lr_rotation: float = 2e-3
    ssim_weight: float = 0.2
    opac_decay: float = 0.004
    max_splats: int = 10_000_000
    refine_every: int = 200
    growth_grad_threshold: float = 0.0025
    growth_select_fraction: float = 0.25
    growth_stop_iter: int = 15000
    split_at_screen_size: float = 0.5
    match_alpha_weight: float = 0.1
    background_color: Sequence[float] = (0.0, 0.0, 0.0)
    background_noise_strength: float = 0.1
    render_mip: bool = False
    seed: int = 0  # the reference uses an unseeded rand::rng(); a shared seed keeps DP ranks identical


@dataclass
class Splats:
    transforms: torch.Tensor      # [N,10]
    sh_coeffs: torch.Tensor       # [N,K,3]
    raw_opacities: torch.Tensor   # [N]
    min_scale: Optional[torch.Tensor] = None   # [N] world-space scale floor (gaussian_splats.rs:73), a constant

    def num_splats(self) -> int:
        return self.transforms.shape[0]

    def folded(self, ctx: RenderContext):
        """(transforms, raw_opacities) as the renderer must see them (gaussian_splats.rs:212-223, 379-384)."""
        if self.min_scale is None:
            return self.transforms, self.raw_opacities
        return fold_min_scale(ctx, self.transforms, self.raw_opacities, self.min_scale)

    def bake_min_scale(self, ctx: RenderContext) -> None:
        """Splats::bake_min_scale (gaussian_splats.rs:245-252): fold permanently, in place, and drop the floor."""
        if self.min_scale is not None:
            fold_min_scale(ctx, self.transforms, self.raw_opacities, self.min_scale, out=(self.transforms, self.raw_opacities))
            self.min_scale = None


def fold_min_scale(ctx: RenderContext, transforms, raw_opac, f, out=None):
    lib = _lib.load()
    n = transforms.shape[0]
    t_out, o_out = out if out is not None else (torch.empty_like(transforms), torch.empty_like(raw_opac))
    _lib.check(lib.bg_fold_min_scale_forward(ctx.handle, _stream_ptr(ctx.device), n, transforms.data_ptr(), raw_opac.data_ptr(),
                                             f.data_ptr(), t_out.data_ptr(), o_out.data_ptr()), "bg_fold_min_scale_forward")
    return t_out, o_out


def fold_min_scale_backward(ctx: RenderContext, transforms, raw_opac, f, v_transforms, v_raw_opac) -> None:
    """In place: gradients w.r.t. the folded values -> w.r.t. the learned ones."""
    lib = _lib.load()
    _lib.check(lib.bg_fold_min_scale_backward(ctx.handle, _stream_ptr(ctx.device), transforms.shape[0], transforms.data_ptr(),
                                              raw_opac.data_ptr(), f.data_ptr(), v_transforms.data_ptr(),
                                              v_raw_opac.data_ptr()), "bg_fold_min_scale_backward")


def compute_min_scale(ctx: RenderContext, transforms, view_cams: torch.Tensor, factor: float) -> Optional[torch.Tensor]:
    """compute_min_scale (train.rs:102-125).  view_cams: device [views,4] = (x, y, z, focal_px)."""
    if factor <= 0.0 or view_cams is None or view_cams.shape[0] == 0:
        return None
    lib = _lib.load()
    f = torch.empty(transforms.shape[0], dtype=torch.float32, device=transforms.device)
    _lib.check(lib.bg_compute_min_scale(ctx.handle, _stream_ptr(ctx.device), transforms.shape[0], transforms.data_ptr(),
                                        view_cams.data_ptr(), view_cams.shape[0], float(factor), f.data_ptr()),
               "bg_compute_min_scale")
    return f


@dataclass
class SceneBatch:
    img_packed: torch.Tensor      # [H,W] int32 (rgba8 little endian); host (pinned) or device
    camera: Camera
    has_alpha: bool = False
    masked_alpha: bool = False    # AlphaMode::Masked

    def img_size(self):
        return int(self.img_packed.shape[0]), int(self.img_packed.shape[1])


@dataclass
class BoundingBox:
    center: np.ndarray
    extent: np.ndarray

    def median_size(self) -> float:
        e = sorted(float(x) for x in self.extent)
        return e[1] * 2.0


def bounds_from_pos(percentile: float, means: np.ndarray) -> BoundingBox:
    """splat_init.rs:130-160."""
    cols = []
    for a in range(3):
        v = means[:, a]
        v = np.sort(v[np.isfinite(v)])
        if v.size == 0:
            return BoundingBox(np.zeros(3, np.float32), np.ones(3, np.float32))
        n = v.size
        lo = int(np.float32((1.0 - percentile) / 2.0) * np.float32(n))
        hi = min(n - 1, int(np.float32((1.0 + percentile) / 2.0) * np.float32(n)))
        cols.append((v[lo], v[hi]))
    mn = np.array([c[0] for c in cols], np.float32)
    mx = np.array([c[1] for c in cols], np.float32)
    return BoundingBox((mx + mn) / 2.0, (mx - mn) / 2.0)


@dataclass
class RefineStats:
    """brush-train/src/msg.rs RefineStats."""
    num_added: int
    num_split_oversized: int
    num_split_high_grad: int
    num_pruned: int
    num_pruned_non_finite: int
    total_splats: int


MIN_SCALE_FREEZE_FRAC = 0.9   # train.rs:37
MIN_SCALE_FACTOR = 0.1        # train.rs:44
MIN_OPACITY = 1.0 / 255.0
BOUND_PERCENTILE = 0.8
FRAC_1_SQRT_2 = 0.7071067811865476


def multinomial_sample(weights: torch.Tensor, n: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Weighted sampling WITHOUT replacement (multinomial.rs:1-26; rand's `sample_weighted` is the
    Efraimidis-Spirakis scheme).  Non-finite or negative weights count as zero; at most
    #(positive weights) indices are returned (all-zero weights -> empty), no duplicates.
    Runs on the weights' device; a seeded generator makes data-parallel ranks agree."""
    w = torch.where(torch.isfinite(weights) & (weights >= 0), weights, torch.zeros_like(weights)).double()
    positive = int((w > 0).sum().item())
    n = min(int(n), positive)
    if n <= 0:
        return torch.empty(0, dtype=torch.long, device=weights.device)
    u = torch.rand(w.shape, dtype=torch.float64, device=w.device, generator=generator).clamp_min(1e-300)
    keys = torch.where(w > 0, torch.log(u) / w, torch.full_like(w, -float("inf")))  # log(u^(1/w))
    return torch.topk(keys, n).indices


def quaternion_vec_multiply(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """quat_vec.rs: rotate v [n,3] by (not necessarily unit) q [n,4] = (w,x,y,z)."""
    qw, qx, qy, qz = q[:, 0:1], q[:, 1:2], q[:, 2:3], q[:, 3:4]
    vx, vy, vz = v[:, 0:1], v[:, 1:2], v[:, 2:3]
    qw2, qx2, qy2, qz2 = qw * qw, qx * qx, qy * qy, qz * qz
    xy, xz, yz, wx, wy, wz = qx * qy, qx * qz, qy * qz, qw * qx, qw * qy, qw * qz
    x = (qw2 + qx2 - qy2 - qz2) * vx + (xy * vy + xz * vz + wy * vz - wz * vy) * 2.0
    y = (qw2 - qx2 + qy2 - qz2) * vy + (xy * vx + yz * vz + wz * vx - wx * vz) * 2.0
    z = (qw2 - qx2 - qy2 + qz2) * vz + (xz * vx + yz * vy + wx * vy - wy * vx) * 2.0
    return torch.cat([x, y, z], 1)


def bounds_from_pos_device(percentile: float, means: torch.Tensor) -> BoundingBox:
    """splat_init.rs:130-160 on the device (one sort per axis, 6 scalars read back)."""
    vals = []
    for a in range(3):
        v = means[:, a]
        v = torch.sort(v[torch.isfinite(v)]).values
        n = v.numel()
        if n == 0:
            return BoundingBox(np.zeros(3, np.float32), np.ones(3, np.float32))
        lo = int(np.float32((1.0 - percentile) / 2.0) * np.float32(n))
        hi = min(n - 1, int(np.float32((1.0 + percentile) / 2.0) * np.float32(n)))
        vals.append(torch.stack([v[lo], v[hi]]))
    mm = torch.stack(vals).cpu().numpy().astype(np.float32)  # [3,2]
    return BoundingBox((mm[:, 1] + mm[:, 0]) / 2.0, (mm[:, 1] - mm[:, 0]) / 2.0)


@dataclass
class TrainStepStats:
    num_visible_event: object
    lr_mean: float
    loss: torch.Tensor  # lazy device scalar (msg.rs:16-27)


class SplatTrainer:
    def __init__(self, config: TrainConfig, ctx: RenderContext, bounds: BoundingBox,
                 grad_hook: Optional[Callable[[Sequence[torch.Tensor]], None]] = None):
        self.config = config
        self.ctx = ctx
        self.bounds = bounds
        self.lr_mean_decay = (config.lr_mean_end / config.lr_mean) ** (1.0 / config.total_train_iters)
        self.ssim_enabled = config.ssim_weight > 0.0
        self.step_count = 0
        self.grad_hook = grad_hook  # called with the gradient tensors before Adam (DP all-reduce)
        self._state = None
        self._v_output = None
        self._v_output_ch = 0
        # ctx may be None for refine()-only use (pure tensor logic, runs on whatever device the splats are on)
        self._gen = torch.Generator(device=ctx.device if ctx is not None else "cpu")
        self._gen.manual_seed(config.seed)
        self._host_rng = np.random.default_rng(config.seed)
        self.view_cams: Optional[torch.Tensor] = None
        self._views_buf = None
        self._fused_ws = None
        self._fused_loss = None

    def set_view_cams(self, view_cams) -> None:
        """train.rs:172-174.  view_cams: sequence of ((x, y, z), focal_px) of the training views."""
        rows = [[float(c[0][0]), float(c[0][1]), float(c[0][2]), float(c[1])] for c in view_cams]
        dev = self.ctx.device if self.ctx is not None else "cpu"
        self.view_cams = torch.tensor(rows, dtype=torch.float32, device=dev).reshape(-1, 4) if rows else None

    # -- optimizer state (train.rs:300-326, adam_scaled.rs)
    def _ensure_state(self, s: Splats):
        if self._state is not None:
            return
        n, k = s.num_splats(), s.sh_coeffs.shape[1]
        dev = s.transforms.device
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        scales = np.ones(k, np.float32)
        scales[1:] = np.float32(1.0) / np.float32(self.config.lr_coeffs_sh_scale)
        self._state = dict(
            m_t=z(n, 10), v_t=z(n, 10), m_sh=z(n, k, 3), v_sh=z(n), m_o=z(n), v_o=z(n),
            sh_lr_scale=torch.from_numpy(np.repeat(scales, 3)).to(dev),
            t_lr=torch.zeros(10, dtype=torch.float32, device=dev),
            refine_norm=z(n), vis_weight=z(n), max_screen=z(n),
        )

    def _adam(self, p, g, m, v, lr, scale, reduce_v):
        lib = _lib.load()
        rows = p.shape[0]
        cols = p.numel() // max(rows, 1)
        _lib.check(lib.bg_adam_step(self.ctx.handle, _stream_ptr(self.ctx.device), p.data_ptr(), g.data_ptr(), m.data_ptr(),
                                    v.data_ptr(), rows, cols, scale.data_ptr() if scale is not None else None,
                                    float(lr), 0.9, 0.999, 1e-15, self.step_count, int(reduce_v)), "bg_adam_step")

    def sample_background(self):
        base = np.asarray(self.config.background_color, np.float32)
        s = self.config.background_noise_strength
        if s <= 0.0:
            return tuple(float(x) for x in np.clip(base, 0.0, 1.0))
        noise = self._host_rng.uniform(-s, s, 3).astype(np.float32)
        return tuple(float(x) for x in np.clip(base + noise, 0.0, 1.0))

    def step(self, batch: SceneBatch, splats: Splats) -> TrainStepStats:
        cfg = self.config
        self._ensure_state(splats)
        st = self._state
        self.step_count += 1
        img_h, img_w = batch.img_size()
        dev = self.ctx.device
        gt_packed = batch.img_packed.to(dev, non_blocking=True)           # H2D upload (train.rs:197-198)
        background = self.sample_background()
        median_scale = self.bounds.median_size()

        r_transforms, r_raw_opac = splats.folded(self.ctx)   # 3D-filter floor folded in (bwd/burn_glue.rs:260-270)
        out = render_splats(self.ctx, batch.camera, (img_w, img_h), r_transforms, splats.sh_coeffs,
                            r_raw_opac, mip=cfg.render_mip, background=background, rpass=PASS_BACKWARD)
        # loss config (train.rs:220-249)
        l1_w, ssim_w = (1.0 - cfg.ssim_weight, -cfg.ssim_weight) if self.ssim_enabled else (1.0, 0.0)
        do_alpha_match = batch.has_alpha and not batch.masked_alpha and cfg.match_alpha_weight > 0.0
        composite = background if (batch.has_alpha and any(b != 0.0 for b in background)) else None
        lcfg = ImageLossConfig(l1_w, ssim_w, composite, batch.masked_alpha)
        channels = 4 if do_alpha_match else 3
        # loss = mean over [h,w,3] (+ alpha mean * weight) (train.rs:254-260): dL/dmap is one constant per
        # channel, so value and gradient come from the fused kernel in one pass.
        npx = float(img_h * img_w)
        chain = [1.0 / (3.0 * npx)] * 3 + ([cfg.match_alpha_weight / npx] if do_alpha_match else [])
        if self._v_output is None or self._v_output.shape != out.out_img.shape or self._v_output_ch != channels:
            self._v_output = torch.zeros_like(out.out_img)   # channel 3 stays zero unless alpha matching
            self._v_output_ch = channels
        v_output, loss = image_loss_fused(self.ctx, out.out_img, gt_packed, channels, lcfg, chain, self._v_output)
        v_combined = rasterize_bwd(out, v_output)
        v_t, v_sh, v_o, v_r = project_bwd(out, r_transforms, splats.sh_coeffs, r_raw_opac, v_combined)
        if splats.min_scale is not None:
            fold_min_scale_backward(self.ctx, splats.transforms, splats.raw_opacities, splats.min_scale, v_t, v_o)
        if self.grad_hook is not None:
            self.grad_hook((v_t, v_sh, v_o, v_r, out.visible, out.max_radius))

        lr_mean = self._apply_updates(splats, v_t, v_sh, v_o, v_r, out.visible, out.max_radius, median_scale)
        return TrainStepStats(num_visible_event=out, lr_mean=lr_mean, loss=loss)


    # ------------------------------------------------------------------------------------------------
    def step_fused(self, batch: SceneBatch, splats: Splats) -> TrainStepStats:
        """The same step through ONE ABI call (bg_train_step): every launch of the step is issued by the library on
        the current stream, scratch comes from a workspace allocated once.  No min-scale floor, no gradient hook."""
        if splats.min_scale is not None or self.grad_hook is not None:
            raise ValueError("step_fused handles the plain single-view step; use step() with a scale floor or a gradient hook")
        cfg = self.config
        self._ensure_state(splats)
        st = self._state
        self.step_count += 1
        img_h, img_w = batch.img_size()
        dev = self.ctx.device
        lib = _lib.load()
        gt_packed = batch.img_packed.to(dev, non_blocking=True)
        background = self.sample_background()
        median_scale = self.bounds.median_size()
        n, k = splats.num_splats(), splats.sh_coeffs.shape[1]
        need = int(lib.bg_train_step_workspace_bytes(n, k, img_w, img_h))
        if self._fused_ws is None or self._fused_ws.numel() < need:
            self._fused_ws = torch.empty(need, dtype=torch.uint8, device=dev)
            self._fused_loss = torch.zeros(1, dtype=torch.float32, device=dev)
        from .camera import build_uniforms
        a = _lib.BgTrainStepArgs()
        a.cam = _lib.camera_struct(build_uniforms(batch.camera, img_w, img_h))
        a.w, a.h, a.n, a.k, a.mip = img_w, img_h, n, k, int(cfg.render_mip)
        for i in range(3):
            a.background[i] = float(background[i])
        a.transforms, a.sh, a.raw_opac = splats.transforms.data_ptr(), splats.sh_coeffs.data_ptr(), splats.raw_opacities.data_ptr()
        a.m_t, a.v_t, a.m_sh, a.v_sh, a.m_o, a.v_o = (st[x].data_ptr() for x in ("m_t", "v_t", "m_sh", "v_sh", "m_o", "v_o"))
        a.refine_norm, a.vis_weight, a.max_screen = (st[x].data_ptr() for x in ("refine_norm", "vis_weight", "max_screen"))
        a.gt_packed = gt_packed.data_ptr()
        a.l1_weight, a.ssim_weight = (1.0 - cfg.ssim_weight, -cfg.ssim_weight) if self.ssim_enabled else (1.0, 0.0)
        do_alpha_match = batch.has_alpha and not batch.masked_alpha and cfg.match_alpha_weight > 0.0
        comp = batch.has_alpha and any(b != 0.0 for b in background)
        a.has_composite_bg = int(comp)
        for i in range(3):
            a.composite_bg[i] = float(background[i])
        a.mask, a.channels, a.alpha_weight = int(batch.masked_alpha), (4 if do_alpha_match else 3), float(cfg.match_alpha_weight)
        lr_mean = cfg.lr_mean * self.lr_mean_decay ** (self.step_count - 1) * float(median_scale)
        a.lr_mean, a.lr_rotation, a.lr_scale = float(np.float32(lr_mean)), cfg.lr_rotation, cfg.lr_scale
        a.lr_coeffs_dc, a.lr_coeffs_sh_scale, a.lr_opac = cfg.lr_coeffs_dc, cfg.lr_coeffs_sh_scale, cfg.lr_opac
        a.noise_scale = float(np.float32(lr_mean) * np.float32(cfg.mean_noise_weight))
        a.median_scale, a.seed, a.step = float(median_scale), int(cfg.seed), self.step_count
        a.workspace, a.workspace_bytes = self._fused_ws.data_ptr(), need
        a.loss_out = self._fused_loss.data_ptr()
        _lib.check(lib.bg_train_step(self.ctx.handle, _stream_ptr(dev), C.byref(a)), "bg_train_step")
        return TrainStepStats(num_visible_event=None, lr_mean=lr_mean, loss=self._fused_loss[0])

    # ------------------------------------------------------------------------------------------------
    def step_views(self, batches: Sequence[SceneBatch], splats: Splats, group=None) -> TrainStepStats:
        """One optimizer step over several views (BASELINE config [4]): the loss is the mean of the
        per-view losses, i.e. the step equals accumulating the views' gradients sequentially on one GPU.
        Under torch.distributed every rank passes ITS views (the same count on every rank); ranks exchange
        the SH-factored gradients (dp.FactoredGradients: all-reduce 44 N B, all-gather 12 N B per view, v_sh
        rebuilt locally in global view order), so all ranks apply bit-identical updates.  At most 16 views
        per step in total."""
        import torch.distributed as dist
        cfg = self.config
        self._ensure_state(splats)
        st = self._state
        self.step_count += 1
        dev = self.ctx.device
        multi = dist.is_initialized() and dist.get_world_size(group) > 1
        world = dist.get_world_size(group) if multi else 1
        local = len(batches)
        views = local * world
        if local == 0 or views > 16:
            raise ValueError("step_views needs 1..16 views per step in total")
        n, k = splats.num_splats(), splats.sh_coeffs.shape[1]
        fb = self._views_buf
        if fb is None or fb["n"] != n or fb["views"] != views or fb["local"] != local:
            z = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
            fb = self._views_buf = dict(n=n, views=views, local=local, small=z(n * 11), tmp=z(n * 11), v_color=z(local, n, 3),
                                        v_color_all=z(views, n, 3), v_sh=z(n, k, 3), v_r=z(n), v_r_acc=z(n), vis=z(n), rad=z(n),
                                        cam_pos=z(views, 3))
        small, tmp = fb["small"], fb["tmp"]
        acc_t, acc_o = small[:n * 10].view(n, 10), small[n * 10:]
        tmp_t, tmp_o = tmp[:n * 10].view(n, 10), tmp[n * 10:]
        background = self.sample_background()          # shared seed: identical on every rank
        median_scale = self.bounds.median_size()
        r_transforms, r_raw_opac = splats.folded(self.ctx)
        l1_w, ssim_w = (1.0 - cfg.ssim_weight, -cfg.ssim_weight) if self.ssim_enabled else (1.0, 0.0)
        loss_sum = None
        out = None
        for i, batch in enumerate(batches):
            img_h, img_w = batch.img_size()
            gt_packed = batch.img_packed.to(dev, non_blocking=True)
            out = render_splats(self.ctx, batch.camera, (img_w, img_h), r_transforms, splats.sh_coeffs, r_raw_opac,
                                mip=cfg.render_mip, background=background, rpass=PASS_BACKWARD)
            do_alpha_match = batch.has_alpha and not batch.masked_alpha and cfg.match_alpha_weight > 0.0
            composite = background if (batch.has_alpha and any(b != 0.0 for b in background)) else None
            lcfg = ImageLossConfig(l1_w, ssim_w, composite, batch.masked_alpha)
            channels = 4 if do_alpha_match else 3
            npx = float(img_h * img_w)
            chain = [1.0 / (3.0 * npx)] * 3 + ([cfg.match_alpha_weight / npx] if do_alpha_match else [])
            if self._v_output is None or self._v_output.shape != out.out_img.shape or self._v_output_ch != channels:
                self._v_output = torch.zeros_like(out.out_img)
                self._v_output_ch = channels
            v_output, loss = image_loss_fused(self.ctx, out.out_img, gt_packed, channels, lcfg, chain, self._v_output)
            loss_sum = loss if loss_sum is None else loss_sum + loss
            v_combined = rasterize_bwd(out, v_output)
            first = i == 0
            project_bwd_factored(out, r_transforms, splats.sh_coeffs, r_raw_opac, v_combined,
                                 outputs=(acc_t if first else tmp_t, fb["v_color"][i], acc_o if first else tmp_o,
                                          fb["v_r_acc"] if first else fb["v_r"]))
            if first:
                fb["vis"].copy_(out.visible)
                fb["rad"].copy_(out.max_radius)
            else:                                       # gather_stats over the local views (stats.rs:40-50)
                small.add_(tmp)
                torch.maximum(fb["v_r_acc"], fb["v_r"], out=fb["v_r_acc"])
                fb["vis"].add_(out.visible)
                torch.maximum(fb["rad"], out.max_radius, out=fb["rad"])
        # ---- exchange: global view index = rank * local + i
        my_pos = torch.tensor([list(b.camera.position) for b in batches], dtype=torch.float32)
        if multi:
            dist.all_gather_into_tensor(fb["v_color_all"].view(-1), fb["v_color"].view(-1), group=group)
            dist.all_reduce(small, op=dist.ReduceOp.SUM, group=group)
            mx = torch.stack([fb["v_r_acc"], fb["rad"]])
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
            fb["v_r_acc"].copy_(mx[0]); fb["rad"].copy_(mx[1])
            dist.all_reduce(fb["vis"], op=dist.ReduceOp.SUM, group=group)
            gathered = fb["cam_pos"]
            dist.all_gather_into_tensor(gathered.view(-1), my_pos.to(dev).view(-1), group=group)
            cam_positions = gathered.cpu().tolist()
        else:
            fb["v_color_all"].copy_(fb["v_color"])
            cam_positions = my_pos.tolist()
        inv = 1.0 / views
        sh_grad_from_views(self.ctx, r_transforms, k, cam_positions, fb["v_color_all"], inv, out=fb["v_sh"])
        if views != 1:
            small.mul_(inv)
        if splats.min_scale is not None:
            fold_min_scale_backward(self.ctx, splats.transforms, splats.raw_opacities, splats.min_scale, acc_t, acc_o)
        loss_mean = loss_sum * (1.0 / local)
        lr_mean = self._apply_updates(splats, acc_t, fb["v_sh"], acc_o, fb["v_r_acc"], fb["vis"], fb["rad"], median_scale)
        return TrainStepStats(num_visible_event=out, lr_mean=lr_mean, loss=loss_mean)

    def _apply_updates(self, splats, v_t, v_sh, v_o, v_r, visible, max_radius, median_scale) -> float:
        """Adam on the three parameter tensors, refine statistics, mean noise (train.rs:300-416)."""
        cfg, st, dev = self.config, self._state, self.ctx.device
        lr_mean = cfg.lr_mean * self.lr_mean_decay ** (self.step_count - 1) * float(median_scale)
        lr_vals = np.array([lr_mean] * 3 + [cfg.lr_rotation] * 4 + [cfg.lr_scale] * 3, np.float32)
        st["t_lr"].copy_(torch.from_numpy(lr_vals), non_blocking=True)
        self._adam(splats.transforms, v_t, st["m_t"], st["v_t"], 1.0, st["t_lr"], False)
        self._adam(splats.sh_coeffs, v_sh, st["m_sh"], st["v_sh"], cfg.lr_coeffs_dc, st["sh_lr_scale"], True)
        self._adam(splats.raw_opacities, v_o, st["m_o"], st["v_o"], cfg.lr_opac, None, False)
        n = splats.num_splats()
        lib = _lib.load()
        # counter-based draw keyed by (seed, step): identical on every data-parallel rank and in bg_train_step
        noise = torch.empty((n, 3), dtype=torch.float32, device=dev)
        _lib.check(lib.bg_normal_noise(self.ctx.handle, _stream_ptr(dev), int(cfg.seed), (self.step_count - 1) * ((3 * n + 3) // 4),
                                       3 * n, noise.data_ptr()), "bg_normal_noise")
        _lib.check(lib.bg_refine_stats_noise(self.ctx.handle, _stream_ptr(dev), n, v_r.data_ptr(), visible.data_ptr(),
                                             max_radius.data_ptr(), st["refine_norm"].data_ptr(),
                                             st["vis_weight"].data_ptr(), st["max_screen"].data_ptr(),
                                             splats.transforms.data_ptr(), splats.raw_opacities.data_ptr(),
                                             noise.data_ptr(), float(np.float32(lr_mean) * np.float32(cfg.mean_noise_weight)),
                                             float(median_scale)), "bg_refine_stats_noise")
        return lr_mean

    # ------------------------------------------------------------------------------------------------
    def refine(self, iteration: int, splats: Splats) -> RefineStats:
        """SplatTrainer::refine + refine_splats + prune_points (train.rs:431-893).  Mutates `splats`
        (tensors are replaced: N changes) and the optimizer / refine-record state."""
        cfg = self.config
        if self._state is None:
            raise RuntimeError("Can only refine after optimizer is initialized")
        st = self._state
        dev = splats.transforms.device
        # refine manipulates the canonical params: bake the current floor first (train.rs:432-437)
        if splats.min_scale is not None:
            splats.bake_min_scale(self.ctx)
        max_allowed = float(np.max(self.bounds.extent)) * 100.0

        # ---- prune mask (train.rs:487-535)
        opac = torch.sigmoid(splats.raw_opacities)
        alpha_mask = opac < MIN_OPACITY
        scale_big = (splats.transforms[:, 7:10].exp() > max_allowed).any(1)
        center = torch.tensor(self.bounds.center, dtype=torch.float32, device=dev).reshape(1, 3)
        bound_mask = ((splats.transforms[:, 0:3] - center).abs() > max_allowed).any(1)
        non_finite = (~torch.isfinite(splats.transforms)).any(1) | (~torch.isfinite(splats.sh_coeffs.flatten(1))).any(1) \
            | ~torch.isfinite(splats.raw_opacities)
        num_non_finite = int(non_finite.sum().item())
        prune = alpha_mask | scale_big | bound_mask | non_finite

        # ---- prune_points (train.rs:848-893)
        keep = (~prune).nonzero(as_tuple=False).squeeze(1)
        n0 = splats.num_splats()
        pruned = 0
        if 0 < keep.numel() < n0:
            pruned = n0 - keep.numel()
            splats.transforms = splats.transforms.index_select(0, keep)
            splats.sh_coeffs = splats.sh_coeffs.index_select(0, keep)
            splats.raw_opacities = splats.raw_opacities.index_select(0, keep)
            for k in ("m_t", "v_t", "m_sh", "v_sh", "m_o", "v_o", "refine_norm", "vis_weight", "max_screen"):
                st[k] = st[k].index_select(0, keep)
        n = splats.num_splats()
        vis_mask = st["vis_weight"] > 0

        split = torch.zeros(n, dtype=torch.bool, device=dev)
        # ---- replace dead gaussians, weighted by opacity x visibility (train.rs:544-556)
        if pruned > 0:
            wts = torch.sigmoid(splats.raw_opacities) * vis_mask.float()
            split[multinomial_sample(wts, pruned, self._gen)] = True
        # ---- force-split splats that are too big on screen (train.rs:562-586), capped by max_splats
        pre = int(split.sum().item())
        if cfg.split_at_screen_size > 0.0:
            cand = ((st["max_screen"] > cfg.split_at_screen_size) & vis_mask & ~split).nonzero(as_tuple=False).squeeze(1)
            budget = max(0, cfg.max_splats - (n + pre))
            split[cand[:budget]] = True
        num_oversized = int(split.sum().item()) - pre
        # ---- growth: sample among splats whose refine weight is above the threshold (train.rs:590-632)
        pre_grad = int(split.sum().item())
        if iteration < min(cfg.growth_stop_iter, cfg.total_train_iters):
            above = (st["refine_norm"] > cfg.growth_grad_threshold) & vis_mask
            threshold_count = int(above.sum().item())
            grow = max(0, int(round(threshold_count * cfg.growth_select_fraction)) - pruned)
            grow = min(grow, max(0, cfg.max_splats - (n + pre_grad)))
            if grow > 0:
                # sampled independently of earlier picks, like the reference's HashSet union
                split[multinomial_sample(above.float() * st["refine_norm"], grow, self._gen)] = True
        num_high_grad = int(split.sum().item()) - pre_grad
        inds = split.nonzero(as_tuple=False).squeeze(1)
        refine_count = inds.numel()

        # ---- refine_splats (train.rs:665-821)
        if refine_count > 0:
            cur = splats.transforms.index_select(0, inds)
            cur_means, rots_raw, cur_log_scale = cur[:, 0:3], cur[:, 3:7], cur[:, 7:10]
            cur_rots = rots_raw / rots_raw.pow(2).sum(1, keepdim=True).sqrt().clamp_min(1e-32)
            cur_sh = splats.sh_coeffs.index_select(0, inds)
            cur_raw_opac = splats.raw_opacities.index_select(0, inds)
            cur_scales = cur_log_scale.exp()
            inv_opac = 1.0 - torch.sigmoid(cur_raw_opac)
            new_opac = (1.0 - inv_opac.pow(FRAC_1_SQRT_2)).clamp(MIN_OPACITY, 1.0 - MIN_OPACITY)
            new_raw_opac = torch.log(new_opac / (1.0 - new_opac))
            sq = cur_scales.pow(2)
            ratio = sq / sq.max(1, keepdim=True).values.clamp_min(1e-30)
            if cfg.split_at_screen_size > 0.0:
                k_max = (st["max_screen"].index_select(0, inds).unsqueeze(1).clamp_min(1e-6).reciprocal()
                         * cfg.split_at_screen_size).clamp_max(FRAC_1_SQRT_2)
                k_axis = -(ratio * (1.0 - k_max)) + 1.0
            else:
                k_axis = -(ratio * (1.0 - FRAC_1_SQRT_2)) + 1.0
            offset_local = (1.0 - k_axis.pow(2)).clamp_min(0.0).sqrt() * cur_scales
            samples = quaternion_vec_multiply(cur_rots, offset_local)
            new_log_scales = cur_log_scale + k_axis.log()
            # parents move to mean - offset and shrink; children sit at mean + offset
            splats.transforms[inds, 0:3] = cur_means - samples
            splats.transforms[inds, 7:10] = new_log_scales
            splats.raw_opacities[inds] = new_raw_opac
            children = torch.cat([cur_means + samples, cur_rots, new_log_scales], 1)
            splats.transforms = torch.cat([splats.transforms, children], 0)
            splats.sh_coeffs = torch.cat([splats.sh_coeffs, cur_sh], 0)
            splats.raw_opacities = torch.cat([splats.raw_opacities, new_raw_opac], 0)
            # both halves of a split restart with zero Adam moments
            for k in ("m_t", "v_t", "m_sh", "v_sh", "m_o", "v_o"):
                st[k][inds] = 0
                st[k] = torch.cat([st[k], torch.zeros((refine_count,) + tuple(st[k].shape[1:]), dtype=torch.float32, device=dev)], 0)

        # ---- opacity decay (train.rs:808-816)
        train_t = min(max(iteration / float(cfg.total_train_iters), 0.0), 1.0)
        minus_opac = cfg.opac_decay * (1.0 - train_t)
        o = (torch.sigmoid(splats.raw_opacities) - minus_opac).clamp(1e-12, 1.0 - 1e-12)
        splats.raw_opacities = torch.log(o / (1.0 - o))

        # ---- bounds, refine record restart (train.rs:637-640, 442-445)
        self.bounds = bounds_from_pos_device(BOUND_PERCENTILE, splats.transforms[:, 0:3])
        n_new = splats.num_splats()
        for k in ("refine_norm", "vis_weight", "max_screen"):
            st[k] = torch.zeros(n_new, dtype=torch.float32, device=dev)
        splats.transforms = splats.transforms.contiguous()
        splats.sh_coeffs = splats.sh_coeffs.contiguous()
        splats.raw_opacities = splats.raw_opacities.contiguous()
        # fresh 3D-filter floor against the new positions / count (train.rs:641-647)
        progress = iteration / float(max(cfg.total_train_iters, 1))
        if progress < MIN_SCALE_FREEZE_FRAC and self.view_cams is not None and self.ctx is not None:
            splats.min_scale = compute_min_scale(self.ctx, splats.transforms, self.view_cams, MIN_SCALE_FACTOR)
        return RefineStats(num_added=refine_count, num_split_oversized=num_oversized, num_split_high_grad=num_high_grad,
                           num_pruned=pruned, num_pruned_non_finite=num_non_finite, total_splats=n_new)
