"""N > 1 host logic on CPU: two gloo ranks, each computes the (oracle) gradients of its own view; the
ViewShardedReducer must give every rank the single-process result that accumulates the views'
gradients sequentially and scales by 1/views, and the refine statistics must be
combined with MAX / SUM / MAX (stats.rs:40-50)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _view_grads(rank):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import math
    from brush_b200.camera import Camera, build_uniforms
    from oracle import oracle as orc
    from scenes import random_v_output, synthetic_scene
    n, w, h = 3000, 96, 64
    cam0, tr, sh, op = synthetic_scene(n, w, h, k=4, seed=321)
    a = math.radians(3.0 * rank) / 2
    cam = Camera(position=cam0.position, rotation=(0.0, math.sin(a), 0.0, math.cos(a)), fov_x=cam0.fov_x, fov_y=cam0.fov_y)
    r = orc.render_forward(build_uniforms(cam, w, h), w, h, tr, sh, op)
    _, vt, vsh, vo, vr = orc.render_backward(r, random_v_output(h, w, seed=5 + rank))
    return [torch.from_numpy(x.copy()) for x in (vt, vsh, vo, vr, r.visible, r.max_radius)]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from brush_b200.dp import ViewShardedReducer
    g = _view_grads(rank)
    ViewShardedReducer(num_views_total=world).hook(g)
    torch.save(g, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_view_sharded_reduction_matches_sequential(tmp_path):
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    per_view = [_view_grads(r) for r in range(world)]
    exp = [(per_view[0][i] + per_view[1][i]) / world for i in range(3)]
    exp_ref = torch.maximum(per_view[0][3], per_view[1][3])
    exp_vis = per_view[0][4] + per_view[1][4]
    exp_rad = torch.maximum(per_view[0][5], per_view[1][5])
    outs = [torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in range(world)]
    for i in range(6):
        assert torch.equal(outs[0][i], outs[1][i]), "ranks must hold identical reduced tensors"
    for i in range(3):
        torch.testing.assert_close(outs[0][i], exp[i], rtol=1e-6, atol=1e-9)
    assert torch.equal(outs[0][3], exp_ref) and torch.equal(outs[0][4], exp_vis) and torch.equal(outs[0][5], exp_rad)
    assert outs[0][0].abs().sum() > 0


def test_reducer_single_process_is_scale_only():
    from brush_b200.dp import ViewShardedReducer
    g = [torch.ones(4, 10), torch.ones(4, 2, 3) * 2, torch.ones(4) * 3]
    ViewShardedReducer(num_views_total=4).reduce_gradients(g)
    assert torch.allclose(g[0], torch.full((4, 10), 0.25)) and torch.allclose(g[2], torch.full((4,), 0.75))


def test_flat_gradients_views_alias_one_buffer():
    from brush_b200.dp import FlatGradients, ViewShardedReducer
    fg = FlatGradients(5, 4, "cpu")
    v_t, v_sh, v_o, v_r = fg.outputs()
    v_t.fill_(1.0); v_sh.fill_(2.0); v_o.fill_(3.0)
    # segments start on 16-byte boundaries (n*10 = 50 floats is padded to 52)
    assert fg.flat.numel() == 52 + 60 + 5
    assert fg.flat[:50].eq(1).all() and fg.flat[50:52].eq(0).all() and fg.flat[52:112].eq(2).all() and fg.flat[112:].eq(3).all()
    assert all(t.data_ptr() % 16 == 0 for t in (v_t, v_sh, v_o))
    ViewShardedReducer(num_views_total=2).reduce_flat(fg)
    assert torch.allclose(v_sh, torch.full_like(v_sh, 1.0)) and torch.allclose(v_o, torch.full_like(v_o, 1.5))
