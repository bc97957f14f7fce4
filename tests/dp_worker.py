"""Worker of tests/test_gpu_nccl.py (one process per GPU, launched by torch.distributed.run).

Checks the view-sharded step under NCCL on real devices:
  1. after bg_train_step_views every rank holds BIT-IDENTICAL parameters, moments and refine statistics;
  2. the result equals the same step run on ONE device over all the views (sequential accumulation), up to f32
     summation order;
  3. the pipelined (chunked) exchange gives the same result as the unchunked one (and is bit-identical across ranks);
  4. bg_dp_exchange on its own: all-reduce / all-gather land where csrc/bg_dp.cuh says.
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    import brush_b200.render as R
    import brush_b200.train as T
    from brush_b200.camera import Camera
    from brush_b200.dp import DpComm
    from scenes import synthetic_scene

    n, w, h, k = 20_000, 192, 128, 9
    local = 2
    views = local * world
    cam0, tr, sh, op = synthetic_scene(n, w, h, k=k, seed=21)
    ctx = R.RenderContext(n, w, h, 0, device=local_rank)

    def cam(v):
        a = math.radians(3.0 * v) / 2.0
        return Camera(position=(0.05 * v, -0.02 * v, 0.0), rotation=(0.0, math.sin(a), 0.0, math.cos(a)), fov_x=cam0.fov_x,
                      fov_y=cam0.fov_y, center_uv=cam0.center_uv)

    params = lambda: [torch.from_numpy(x.copy()).to(dev) for x in (tr, sh, op)]
    batches_all = []
    for v in range(views):
        tgt = R.render_splats(ctx, cam(v), (w, h), *params(), rpass=0)
        batches_all.append(T.SceneBatch(img_packed=(tgt.out_img | (255 << 24)).clone(), camera=cam(v)))
    mine = batches_all[rank * local:(rank + 1) * local]
    cfg = T.TrainConfig(total_train_iters=1000, background_noise_strength=0.0, mean_noise_weight=50.0, seed=5)
    bounds = T.bounds_from_pos(0.8, tr[:, :3])

    def run(batches, group_on, chunks, steps=3):
        p = params()
        s = T.Splats(p[0], p[1] + 0.1, p[2])
        t = T.SplatTrainer(cfg, ctx, bounds)
        losses = []
        for _ in range(steps):
            st = t.step_views(batches, s, chunks=chunks, distributed=group_on)
            losses.append(float(st.loss.item()))
        torch.cuda.synchronize()
        return s, t, losses

    s_dp, t_dp, l_dp = run(mine, True, 1)
    s_ch, t_ch, l_ch = run(mine, True, 4)
    # 1. bit-identical across ranks; 3. chunked == unchunked
    def flat(s, t):
        return torch.cat([s.transforms.reshape(-1), s.sh_coeffs.reshape(-1), s.raw_opacities.reshape(-1)] +
                         [t._state[x].reshape(-1) for x in ("m_t", "v_t", "m_sh", "v_sh", "m_o", "v_o", "refine_norm", "vis_weight", "max_screen")])
    mineflat = flat(s_dp, t_dp)
    gathered = [torch.empty_like(mineflat) for _ in range(world)]
    dist.all_gather(gathered, mineflat)
    for r in range(world):
        assert torch.equal(gathered[r].view(torch.int32), gathered[0].view(torch.int32)), f"rank {r} differs from rank 0"
    # (two executions of the same step differ by the order of the f32 atomics in the blend backward: compare like 2.)
    for name in ("transforms", "sh_coeffs", "raw_opacities"):
        a, b = getattr(s_ch, name).double(), getattr(s_dp, name).double()
        assert (((a - b).abs() <= 1e-6 + 1e-3 * a.abs()).double().mean()) > 0.99, ("chunked", name)
    assert torch.equal(t_ch._state["vis_weight"], t_dp._state["vis_weight"])
    chflat = flat(s_ch, t_ch)
    gathered_c = [torch.empty_like(chflat) for _ in range(world)]
    dist.all_gather(gathered_c, chflat)
    for r in range(world):
        assert torch.equal(gathered_c[r].view(torch.int32), gathered_c[0].view(torch.int32)), f"chunked: rank {r} differs from rank 0"
    # 2. equals the one-device step over all views
    s_one, t_one, l_one = run(batches_all, False, 1)
    for name in ("transforms", "sh_coeffs", "raw_opacities"):
        a, b = getattr(s_one, name).double(), getattr(s_dp, name).double()
        close = (a - b).abs() <= 1e-6 + 1e-3 * a.abs()
        assert close.double().mean() > 0.99, (name, float(close.double().mean()))
    for key in ("m_t", "m_sh", "m_o"):
        a, b = t_one._state[key].double(), t_dp._state[key].double()
        assert (a - b).norm() / a.norm() < 1e-3, (key, float((a - b).norm() / a.norm()))
    # after three steps the two runs' parameters differ by rounding, so their statistics agree closely, not bitwise
    for key in ("vis_weight", "max_screen"):
        a, b = t_one._state[key].double(), t_dp._state[key].double()
        assert ((a - b).abs() <= 1e-6 + 1e-3 * a.abs()).double().mean() > 0.999, key
    # the loss each rank reports is the mean over ITS views; the mean over ranks is the one-device loss
    lt = torch.tensor(l_dp, device=dev, dtype=torch.float64)
    dist.all_reduce(lt)
    np.testing.assert_allclose((lt / world).cpu().numpy(), np.array(l_one), rtol=1e-4)
    # 4. bg_dp_exchange on its own: the all-reduce sums in place, the all-gather lands rank r's rows in block r of each slice
    comm = DpComm(ctx)
    R_ = 3 * local
    for chunks in (1, 3):
        small = torch.full((12 * n,), float(rank + 1), device=dev)
        stat = torch.arange(2 * n, device=dev, dtype=torch.float32) * (1.0 if rank == world - 1 else 0.5)
        stat_max = torch.arange(2 * n, device=dev, dtype=torch.float32)
        record = torch.arange(R_ * n, device=dev, dtype=torch.float32) + 1.0e7 * rank
        recv = torch.zeros(world * record.numel(), device=dev)
        comm.exchange(n, local, small, stat, record, recv, chunks=chunks)
        torch.cuda.synchronize()
        assert torch.equal(small, torch.full_like(small, float(sum(range(1, world + 1)))))
        assert torch.equal(stat, stat_max)
        per = ((n + chunks - 1) // chunks + 63) // 64 * 64
        for c in range(chunks):
            g0, g1 = min(per * c, n), min(per * (c + 1), n)
            blk = recv[R_ * world * g0: R_ * world * g1].view(world, -1)
            for r in range(world):
                assert torch.equal(blk[r], record[R_ * g0: R_ * g1] - 1.0e7 * rank + 1.0e7 * r), (chunks, c, r)
    # bg_dp_pack_view: assign on the first view, accumulate / MAX on the next
    vt, vo, vc = torch.rand(n, 10, device=dev), torch.rand(n, device=dev), torch.rand(n, 3, device=dev)
    vr, vis, rad = torch.rand(n, device=dev), (torch.rand(n, device=dev) > 0.5).float(), torch.rand(n, device=dev)
    small, record = torch.full((12 * n,), 7.0, device=dev), torch.full((R_ * n,), 7.0, device=dev)
    stat = torch.full((2 * n,), 7.0, device=dev)
    comm.pack_view(n, local, 0, True, vt, vo, vc, vr, vis, rad, small, stat, record)
    comm.pack_view(n, local, 1, False, 2 * vt, 2 * vo, 3 * vc, vr * 0.5, vis, rad * 2, small, stat, record)
    torch.cuda.synchronize()
    sm, rc, stt = small.view(n, 12), record.view(n, R_), stat.view(n, 2)
    assert torch.equal(sm[:, :10], vt + 2 * vt) and torch.equal(sm[:, 10], vo + 2 * vo) and torch.equal(sm[:, 11], vis + vis)
    assert torch.equal(rc[:, 0:3], vc) and torch.equal(rc[:, 3:6], 3 * vc) and torch.equal(stt[:, 0], vr) and torch.equal(stt[:, 1], rad * 2)
    comm.close()
    ctx.close()
    dist.barrier()
    if rank == 0:
        print(f"DP_WORKER_OK world={world} losses={l_dp}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
