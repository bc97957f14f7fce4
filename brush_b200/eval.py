"""Evaluation metrics: eval_stats <- brush-train/src/eval.rs:22-61.

Render on black with the float output, round-trip through 8 bit, PSNR from the L1 map squared
(|a-b|^2 == (a-b)^2), SSIM as the mean of the SSIM map -- both through bg_image_loss_forward."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .dataset import ALPHA_MASKED, view_to_packed_data
from .loss import ImageLossConfig, image_loss_forward
from .render import PASS_BACKWARD, RenderContext, RenderOutput, render_splats


@dataclass
class EvalSample:
    rendered: torch.Tensor     # [H,W,3] after the 8-bit round trip
    psnr: torch.Tensor         # scalar
    ssim: torch.Tensor         # scalar
    render_aux: RenderOutput

    def save_to_disk(self, path: str) -> None:
        """eval.rs:66-81: the rendered image (already on the 8-bit grid) as an 8-bit RGB file; parent directories are
        created."""
        import os
        from PIL import Image
        img = (self.rendered.detach().clamp(0.0, 1.0) * 255.0).round().to(torch.uint8).cpu().numpy()
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        Image.fromarray(img, "RGB").save(path)


def eval_stats(ctx: RenderContext, splats, camera, gt_image: np.ndarray, alpha_mode: str = ALPHA_MASKED,
               render_mip: bool = False) -> EvalSample:
    """splats: train.Splats (a min-scale floor is folded in, as in render_splats, gaussian_splats.rs:379-384).
    render_mip: the splats' render mode -- evaluation renders with the filter the model is trained with
    (gaussian_splats.rs:395: `splats.render_mip`).  gt_image: [H,W,3] or [H,W,4] u8; an alpha channel goes through
    view_to_packed_data(alpha_mode) like a training view (eval.rs:31)."""
    h, w = gt_image.shape[0], gt_image.shape[1]
    packed, _ = view_to_packed_data(gt_image, alpha_mode)
    gt = torch.from_numpy(packed).to(ctx.device)
    transforms, raw_opac = splats.folded(ctx)
    out = render_splats(ctx, camera, (w, h), transforms, splats.sh_coeffs, raw_opac, mip=render_mip,
                        background=(0.0, 0.0, 0.0), rpass=PASS_BACKWARD)
    rgb = torch.round(out.out_img[..., 0:3] * 255.0) / 255.0            # eval.rs:41-42
    rgb = rgb.contiguous()
    l1 = image_loss_forward(ctx, rgb, gt, 3, ImageLossConfig(1.0, 0.0, None, False))
    mse = l1.pow(2).mean()
    psnr = torch.log(1.0 / mse) * 10.0 / float(np.log(10.0))
    ssim = image_loss_forward(ctx, rgb, gt, 3, ImageLossConfig(0.0, 1.0, None, False)).mean()
    return EvalSample(rendered=rgb, psnr=psnr, ssim=ssim, render_aux=out)
