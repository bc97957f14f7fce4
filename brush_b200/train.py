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

Out of scope here (SURVEY.md 8f "next" row N1): refine() (densify/prune), LPIPS.
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
from .render import PASS_BACKWARD, RenderContext, _stream_ptr, project_bwd, rasterize_bwd, render_splats


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

    def num_splats(self) -> int:
        return self.transforms.shape[0]


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
        self._gen = torch.Generator(device=ctx.device)
        self._gen.manual_seed(config.seed)
        self._host_rng = np.random.default_rng(config.seed)

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

        out = render_splats(self.ctx, batch.camera, (img_w, img_h), splats.transforms, splats.sh_coeffs,
                            splats.raw_opacities, mip=cfg.render_mip, background=background, rpass=PASS_BACKWARD)
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
        v_t, v_sh, v_o, v_r = project_bwd(out, splats.transforms, splats.sh_coeffs, splats.raw_opacities, v_combined)
        if self.grad_hook is not None:
            self.grad_hook((v_t, v_sh, v_o, v_r, out.visible, out.max_radius))

        # learning rates (train.rs:328-350)
        lr_mean = cfg.lr_mean * self.lr_mean_decay ** (self.step_count - 1) * float(median_scale)
        lr_vals = np.array([lr_mean] * 3 + [cfg.lr_rotation] * 4 + [cfg.lr_scale] * 3, np.float32)
        st["t_lr"].copy_(torch.from_numpy(lr_vals), non_blocking=True)
        self._adam(splats.transforms, v_t, st["m_t"], st["v_t"], 1.0, st["t_lr"], False)
        self._adam(splats.sh_coeffs, v_sh, st["m_sh"], st["v_sh"], cfg.lr_coeffs_dc, st["sh_lr_scale"], True)
        self._adam(splats.raw_opacities, v_o, st["m_o"], st["v_o"], cfg.lr_opac, None, False)

        # refine stats + noise on the updated opacities (train.rs:280-298, 389-416)
        n = splats.num_splats()
        noise = torch.randn((n, 3), dtype=torch.float32, device=dev, generator=self._gen)
        lib = _lib.load()
        _lib.check(lib.bg_refine_stats_noise(self.ctx.handle, _stream_ptr(dev), n, v_r.data_ptr(), out.visible.data_ptr(),
                                             out.max_radius.data_ptr(), st["refine_norm"].data_ptr(),
                                             st["vis_weight"].data_ptr(), st["max_screen"].data_ptr(),
                                             splats.transforms.data_ptr(), splats.raw_opacities.data_ptr(),
                                             noise.data_ptr(), float(np.float32(lr_mean) * np.float32(cfg.mean_noise_weight)),
                                             float(median_scale)), "bg_refine_stats_noise")
        return TrainStepStats(num_visible_event=out, lr_mean=lr_mean, loss=loss)
