"""Training loop driver: the schedule of brush-process/src/train_stream.rs:150-500 without its app plumbing
(message emitter, viewer slot, rerun, LOD decimation phases).

  schedule predicates  <- train_stream.rs:318-326 (refine gating), :350-353 (eval cadence), :377-383 (export cadence)
  train_loop           <- train_stream.rs:176-497: loader -> step -> refine -> eval -> export
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np


@dataclass
class ProcessConfig:                     # brush-process/src/config.rs (the fields the loop reads)
    eval_every: int = 1000
    export_every: int = 5000
    export_path: str = "."
    export_name: str = "export_{iter}.ply"
    start_iter: int = 0
    seed: int = 42
    eval_save_to_disk: bool = False     # the rendered eval images go to <export_path>/eval_<iter>/<image name>.png


def should_refine(it: int, refine_every: int, total_iters: int) -> bool:
    """train_stream.rs:318-326, for the step that just ran with 0-based index `it`."""
    progress = min(max(it / float(max(total_iters, 1)), 0.0), 1.0)
    return it > 0 and it % refine_every == 0 and progress <= 0.95


def should_eval(done: int, eval_every: int, total_iters: int) -> bool:
    """train_stream.rs:350-353, `done` = number of finished iterations."""
    return done % eval_every == 0 or done == total_iters


def should_export(done: int, export_every: int, total_iters: int) -> bool:
    """train_stream.rs:377-383 (no LOD phases)."""
    return done % export_every == 0 or done == total_iters


def train_loop(ctx, splats, train_views: Sequence, eval_views: Sequence, config, process: Optional[ProcessConfig] = None,
               on_step: Optional[Callable] = None, alpha_mode: str = "masked") -> List[dict]:
    """Runs config.total_train_iters steps; returns the evaluation records.  `splats`: train.Splats on ctx's device;
    views: dataset.SceneView lists."""
    import torch
    from . import ply
    from .dataset import SceneLoader
    from .eval import eval_stats
    from .train import BOUND_PERCENTILE, SplatTrainer, bounds_from_pos_device
    process = process or ProcessConfig()
    from PIL import Image
    loader = SceneLoader(train_views, alpha_mode, seed=process.seed)
    trainer = SplatTrainer(config, ctx, bounds_from_pos_device(ctx, BOUND_PERCENTILE, splats.transforms))
    # refine grows the model up to config.max_splats: the context must have been created for it (the reference sizes
    # its buffers per render; here capacity is fixed at bg_ctx_create)
    cap = min(int(config.max_splats), max(int(ctx.max_splats), 0))
    if ctx.max_splats < min(config.max_splats, splats.num_splats()):
        raise ValueError(f"render context holds {ctx.max_splats} splats, the model already has {splats.num_splats()}")
    if cap < config.max_splats:
        import dataclasses
        config = dataclasses.replace(config, max_splats=cap)     # never grow past what the context can render
        trainer.config = config
    view_cams = []
    for v in train_views:                                   # (position, focal in px at native resolution): the 3D filter
        with Image.open(v.image_path) as im:                # header only: no decode
            iw, ih = im.size
        view_cams.append((v.camera.position, float(v.camera.focal(iw, ih)[0])))
    trainer.set_view_cams(view_cams)
    total = config.total_train_iters
    evals: List[dict] = []
    last_out = None

    def check_overflow():
        # a view whose tile list exceeds the arena is rendered with the overflowing intersections dropped: never silently
        if last_out is not None and last_out.intersection_overflow:
            raise RuntimeError("intersection arena overflow: create the RenderContext with a larger max_intersections "
                               f"(num_intersections {last_out.num_intersections})")

    for it in range(process.start_iter, total):
        stats = trainer.step(loader.next_batch(), splats)
        if stats.num_visible_event is not None:
            last_out = stats.num_visible_event
        refine = None
        if should_refine(it, config.refine_every, total):
            check_overflow()                                # refine synchronises anyway
            refine = trainer.refine(it, splats)
        done = it + 1
        if on_step is not None:
            on_step(done, stats, refine)
        if eval_views and should_eval(done, process.eval_every, total):
            check_overflow()
            psnr, ssim = [], []
            for v in eval_views:
                gt = v.load_image()                        # view.image.load(): a mask file, if any, is the alpha channel
                s = eval_stats(ctx, splats, v.camera, gt, alpha_mode, render_mip=config.render_mip)
                psnr.append(float(s.psnr)); ssim.append(float(s.ssim))
                if process.eval_save_to_disk:              # train_stream.rs:543-550
                    s.save_to_disk(os.path.join(process.export_path, f"eval_{done}", f"{v.img_name()}.png"))
            evals.append({"iter": done, "psnr": float(np.mean(psnr)), "ssim": float(np.mean(ssim)), "splats": splats.num_splats()})
        if should_export(done, process.export_every, total):
            t_fold, o_fold = splats.folded(ctx)            # export.rs:183: the floor is folded into a COPY, never stored
            data = ply.splat_to_ply(t_fold.cpu().numpy(), splats.sh_coeffs.cpu().numpy(), o_fold.cpu().numpy(),
                                    render_mip=config.render_mip)
            os.makedirs(process.export_path, exist_ok=True)
            with open(os.path.join(process.export_path, process.export_name.replace("{iter}", str(done))), "wb") as f:
                f.write(data)
    torch.cuda.synchronize(ctx.device)
    check_overflow()
    loader.close()
    return evals
