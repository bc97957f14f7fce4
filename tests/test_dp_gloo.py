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


# ---- SH-factored exchange: collectives and buffer layout on CPU (the rebuild kernel is replaced by a torch restatement)
def _factored_inputs(rank):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import math
    from brush_b200.camera import Camera, build_uniforms
    from oracle import oracle as orc
    from scenes import random_v_output, synthetic_scene
    n, w, h = 3000, 96, 64
    cam0, tr, sh, op = synthetic_scene(n, w, h, k=4, seed=321)
    a = math.radians(3.0 * rank) / 2
    cam = Camera(position=(0.1 * rank, -0.05 * rank, 0.0), rotation=(0.0, math.sin(a), 0.0, math.cos(a)), fov_x=cam0.fov_x,
                 fov_y=cam0.fov_y)
    r = orc.render_forward(build_uniforms(cam, w, h), w, h, tr, sh, op)
    vc, vt, vsh, vo, vr = orc.render_backward(r, random_v_output(h, w, seed=5 + rank))
    v_color = np.zeros((n, 3), np.float32)
    v_color[r.gid_from_cgid] = vc[:, 5:8]
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x).copy())
    return dict(tr=t(tr), cam_pos=cam.position, vt=t(vt), vsh=t(vsh), vo=t(vo), vr=t(vr), v_color=t(v_color),
                visible=t(r.visible), max_radius=t(r.max_radius))


def _sh_grad_from_views_torch(ctx, transforms, k, cam_positions, v_all, out_scale=1.0, out=None, view_stride=0):
    """kernels/sh.rs:265-355 for degree <= 1: v_sh = out_scale * sum_v Y(dir_v) (x) v_color_v."""
    assert k == 4
    n = transforms.shape[0]
    acc = torch.zeros((n, 4, 3), dtype=torch.float32)
    for v, pos in enumerate(cam_positions):
        row = v_all[v].reshape(-1)
        col = row[:3 * n].reshape(n, 3)
        d = transforms[:, 0:3] - torch.tensor(pos, dtype=torch.float32)
        d = d / d.norm(dim=1, keepdim=True)
        y = torch.stack([torch.full((n,), 0.2820948), -0.4886025 * d[:, 1], 0.4886025 * d[:, 2], -0.4886025 * d[:, 0]], 1)
        acc += y[:, :, None] * col[:, None, :]
    out.copy_(acc * out_scale)
    return out


def _factored_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import brush_b200.render as R
    from brush_b200.dp import FactoredGradients, ShFactoredReducer
    R.sh_grad_from_views = _sh_grad_from_views_torch
    d = _factored_inputs(rank)
    n = d["tr"].shape[0]
    fg = FactoredGradients(n, 4, world, "cpu")
    v_t, v_color, v_o, v_r = fg.outputs()         # where project_bwd_factored writes
    v_t.copy_(d["vt"]); v_color.copy_(d["v_color"]); v_o.copy_(d["vo"]); v_r.copy_(d["vr"])
    visible, max_radius = d["visible"].clone(), d["max_radius"].clone()
    cam_positions = [_factored_inputs(r)["cam_pos"] for r in range(world)]
    ShFactoredReducer(None, world).reduce(fg, d["tr"], cam_positions, visible, max_radius)
    g = fg.gradients()
    torch.save([g[0].clone(), g[1].clone(), g[2].clone(), g[3].clone(), visible, max_radius], os.path.join(out_dir, f"f{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sh_factored_exchange_matches_dense_mean(tmp_path):
    """Two collectives (all-reduce of v_transforms | v_raw_opac | visible, all-gather of v_color | v_refine | radius) plus
    the local rebuild give every rank the mean of the views' DENSE gradients and the MAX / SUM / MAX refine statistics."""
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_factored_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    views = [_factored_inputs(r) for r in range(world)]
    outs = [torch.load(os.path.join(tmp_path, f"f{r}.pt")) for r in range(world)]
    for i in range(6):
        assert torch.equal(outs[0][i], outs[1][i]), "ranks must hold identical reduced tensors"
    v_t, v_sh, v_o, v_r, vis, rad = outs[0]
    torch.testing.assert_close(v_t, (views[0]["vt"] + views[1]["vt"]) / 2, rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(v_o, (views[0]["vo"] + views[1]["vo"]) / 2, rtol=1e-6, atol=1e-9)
    want_sh = (views[0]["vsh"] + views[1]["vsh"]) / 2          # the oracle's dense per-view SH gradients
    assert (v_sh - want_sh).abs().max() <= 2e-5 * want_sh.abs().max()
    assert want_sh.abs().max() > 0
    assert torch.equal(v_r, torch.maximum(views[0]["vr"], views[1]["vr"]))
    assert torch.equal(vis, views[0]["visible"] + views[1]["visible"])
    assert torch.equal(rad, torch.maximum(views[0]["max_radius"], views[1]["max_radius"]))
