"""Initial splats and the loop schedule (host logic, CPU): brush-train/src/splat_init.rs, brush-process/src/train_stream.rs."""
import math

import numpy as np

from brush_b200 import loop, splat_init
from brush_b200.camera import Camera
from brush_b200.ply import SplatData


def test_estimate_scene_scale():
    assert splat_init.estimate_scene_scale([]) == 1.0 and splat_init.estimate_scene_scale([Camera()]) == 1.0
    cams = [Camera(position=(float(i) * 2.0, 0.0, 0.0)) for i in range(5)]
    assert abs(splat_init.estimate_scene_scale(cams) - 6.0) < 1e-6          # 3 x spacing 2
    near = [Camera(position=(0.01 * i, 0.0, 0.0)) for i in range(5)]
    assert splat_init.estimate_scene_scale(near) == 1.0                     # 1 m floor


def test_create_random_splats_shapes_and_ranges():
    rng = np.random.default_rng(0)
    cams = [Camera(position=(0, 0, 0), fov_x=0.8, fov_y=0.6), Camera(position=(4, 0, 0), rotation=(0, 0.3, 0, 0.954), fov_x=0.8, fov_y=0.6)]
    t, sh, op = splat_init.create_random_splats(1000, cams, rng)
    assert t.shape == (1000, 10) and sh.shape == (1000, 1, 3) and op.shape == (1000,) and t.dtype == np.float32
    scale = splat_init.estimate_scene_scale(cams)                           # 12
    np.testing.assert_allclose(np.linalg.norm(t[:, 3:7], axis=1), 1.0, atol=1e-5)
    assert np.allclose(t[:, 7:10], math.log(scale / 10.0), atol=1e-6)       # cbrt(1000) = 10
    assert (op >= splat_init.inverse_sigmoid(0.1) - 1e-6).all() and (op <= splat_init.inverse_sigmoid(0.25) + 1e-6).all()
    assert (sh >= 0).all() and (sh <= 1).all()
    d = np.minimum(np.linalg.norm(t[:, :3], axis=1), np.linalg.norm(t[:, :3] - [4, 0, 0], axis=1))
    assert d.min() >= scale * 0.05 - 1e-4 and d.min() < scale * 0.2          # depths between near and far


def test_knn_scales_on_a_grid():
    g = np.stack(np.meshgrid(np.arange(6), np.arange(6), np.arange(6), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.2
    ls = splat_init.compute_knn_scales(g)
    assert ls.shape == (216, 3) and (ls[:, 0] == ls[:, 1]).all()
    np.testing.assert_allclose(ls, math.log(0.2 / 2.0), atol=1e-5)          # (0.2 + 0.2) / 4, within the clamp
    assert (splat_init.compute_knn_scales(g[:2]) == 0).all()                # fewer than 3 points
    far = np.array([[0, 0, 0], [100, 0, 0], [0, 100, 0], [0, 0, 100]], np.float32)
    from brush_b200.train import bounds_from_pos
    cap = math.log(max(bounds_from_pos(0.75, far).median_size(), 0.01) * 0.1)
    assert np.allclose(splat_init.compute_knn_scales(far), cap, atol=1e-5)  # clamped to 10 % of the bounds


def test_to_init_splats_defaults_and_sh_degree():
    pts = np.random.default_rng(1).normal(size=(50, 3)).astype(np.float32)
    t, sh, op = splat_init.to_init_splats(SplatData(means=pts))
    assert (t[:, 3:7] == [1, 0, 0, 0]).all() and (op == 0).all() and (sh == 0.5).all() and sh.shape == (50, 1, 3)
    np.testing.assert_array_equal(t[:, 7:10], splat_init.compute_knn_scales(pts))
    up = splat_init.with_sh_degree(sh, 3)
    assert up.shape == (50, 16, 3) and (up[:, 0] == 0.5).all() and (up[:, 1:] == 0).all()
    assert splat_init.with_sh_degree(up, 1).shape == (50, 4, 3)


def test_loop_schedule_predicates():
    total, every = 30000, 200
    refines = [it for it in range(total) if loop.should_refine(it, every, total)]
    assert refines[0] == 200 and refines[-1] == 28400 and 0 not in refines            # stops past 95 %
    assert all(it % every == 0 for it in refines)
    evals = [d for d in range(1, total + 1) if loop.should_eval(d, 1000, total)]
    assert evals[0] == 1000 and evals[-1] == total and len(evals) == 30
    assert loop.should_eval(777, 1000, 777) and not loop.should_eval(776, 1000, 777)
    exports = [d for d in range(1, total + 1) if loop.should_export(d, 5000, total)]
    assert exports == [5000, 10000, 15000, 20000, 25000, 30000]
