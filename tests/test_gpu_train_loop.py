"""End-to-end training on a small synthetic COLMAP-format set (BASELINE config [2] in miniature):
scripts/train_colmap.py writes the set with the product forward kernels, the oracle vouches for the images, and
loop.train_loop (train_stream.rs:220-497: loader -> step -> refine -> eval) must fit the held-out views."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


@pytest.mark.gpu
def test_colmap_set_trains_and_ground_truth_matches_oracle(tmp_path):
    import torch
    from PIL import Image
    import train_colmap
    import brush_b200.render as R
    import brush_b200.train as T
    from brush_b200 import dataset as ds
    from brush_b200 import splat_init
    from brush_b200.camera import build_uniforms
    from brush_b200.loop import ProcessConfig, train_loop
    from oracle import oracle as orc

    w, h, views, hidden_n, init_points = 256, 160, 24, 6_000, 3_000
    ctx = R.RenderContext(40_000, w, h, 0, device=0)
    (base, tr, sh, op), cams = train_colmap.make_dataset(str(tmp_path), ctx, views, w, h, hidden_n, init_points, seed=0xB2000002)
    # --- the written ground truth is what the oracle renders (8-bit quantisation: one level)
    loaded = ds.load_colmap(str(tmp_path), eval_split_every=8)
    assert len(loaded.train) + len(loaded.eval) == views and len(loaded.eval) == 3 and not loaded.warnings
    assert loaded.init_splat.num_splats() == init_points
    for idx in (0, 5):
        view = (loaded.eval + loaded.train)[0] if idx == 0 else loaded.train[idx]
        o = orc.render_forward(build_uniforms(view.camera, w, h), w, h, tr, sh, op, bg=(0, 0, 0))
        want = np.clip(o.out_img[..., :3] * 255.0, 0, 255)
        got = np.asarray(Image.open(view.image_path).convert("RGB"), np.float32)
        # the loader's camera went through COLMAP's pose convention and f64 text: sub-pixel identical geometry
        assert np.mean(np.abs(got - want) > 1.5) < 2e-3, float(np.mean(np.abs(got - want) > 1.5))
    # --- train from the COLMAP points
    tr0, sh0, op0 = splat_init.to_init_splats(loaded.init_splat)
    sh0 = splat_init.with_sh_degree(sh0, 3)
    d = ctx.device
    splats = T.Splats(*(torch.from_numpy(np.ascontiguousarray(x)).to(d) for x in (tr0, sh0, op0)))
    cfg = T.TrainConfig(total_train_iters=300, max_splats=30_000, refine_every=60, growth_stop_iter=250, seed=1)
    seen = []
    evals = train_loop(ctx, splats, loaded.train, loaded.eval, cfg, ProcessConfig(eval_every=150, export_every=10 ** 9, seed=7),
                       on_step=lambda done, st, rf: seen.append((done, rf.total_splats if rf is not None else None)))
    assert [e["iter"] for e in evals] == [150, 300]
    assert evals[-1]["psnr"] > evals[0]["psnr"] + 0.5 and evals[-1]["psnr"] > 15.5, evals     # keeps improving
    assert evals[-1]["ssim"] > evals[0]["ssim"] + 0.1 and evals[-1]["ssim"] <= 1.0
    refined = [c for _, c in seen if c is not None]
    assert len(refined) == 4 and all(0 < c <= 30_000 for c in refined)
    assert splats.num_splats() == refined[-1]
    assert all(torch.isfinite(x).all() for x in (splats.transforms, splats.sh_coeffs, splats.raw_opacities))
    ctx.close()
