#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: task statement, section 4).

Workload (BASELINE.json configs[1]): 1M synthetic Gaussians (SURVEY.md 8d generator, K=16),
1920x1080, one "step" = render forward (RasterPass::Backward) + rasterize backward + project
backward of one view.  metric = forward+backward Mpix/s.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port of the reference
                                                           # kernels on the host cores (the reference
                                                           # itself needs cargo+wgpu: not buildable here)
N > 1 (torchrun, one rank per GPU): the step is view-sharded -- every rank renders its own view of the
replicated scene and the dense per-Gaussian gradients are summed with one NCCL all-reduce.
Weak scaling: per-GPU work is fixed, value = N * pixels / max-over-ranks step time.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

N_SPLATS, IMG_W, IMG_H, SH_K = 1_000_000, 1920, 1080, 16
WORKLOAD = "configs[1]: 1M synthetic Gaussians (K=16), 1920x1080, render fwd + rasterize bwd + project bwd, 1 view/GPU"
METRIC, UNIT = "fwd+bwd Mpix/s @1M Gaussians 1080p", "Mpix/s"
KERNELS_PER_STEP = 14  # fwd: epoch bump, cull, 4 sort, scan, visible+emit, 2 sort, offsets, blend = 12; bwd: blend, project = 2


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Streams `nvidia-smi -lms 100` for one GPU; keeps the samples that fall inside marked timed regions."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.windows, self.proc = gpu_index, [], [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        time.sleep(0.15)
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=5)
        inside = [r for (t, r) in self.rows if any(a <= t <= b + 0.1 for a, b in self.windows)] or [r for _, r in self.rows[-3:]]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[0]) for r in inside if len(r) >= 7 and num(r[0]) is not None]
        mx = [num(r[1]) for r in inside if len(r) >= 7 and num(r[1]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in inside if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(inside)}


def scene_np():
    from scenes import random_v_output, synthetic_scene
    cam, tr, sh, op = synthetic_scene(N_SPLATS, IMG_W, IMG_H, k=SH_K, seed=0xB2000001)
    return cam, tr, sh, op, random_v_output(IMG_H, IMG_W)


def rank_camera(cam, rank):
    """View `rank` of the step's batch: the base camera yawed by 2 degrees per rank."""
    from brush_b200.camera import Camera
    a = math.radians(2.0 * rank) / 2.0
    return Camera(position=cam.position, rotation=(0.0, math.sin(a), 0.0, math.cos(a)), fov_x=cam.fov_x, fov_y=cam.fov_y,
                  center_uv=cam.center_uv)


def cpu_oracle_pass(u, tr, sh, op, v_out):
    from oracle import oracle as orc
    t0 = time.perf_counter()
    r = orc.render_forward(u, IMG_W, IMG_H, tr, sh, op, rpass=orc.PASS_BACKWARD)
    orc.render_backward(r, v_out)
    dt = time.perf_counter() - t0
    r.close()
    return dt


def _oracle_on_all_cores():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU legs are the oracle on ALL the host's cores.  The
    OpenMP default (one thread per core the runtime may use) is what a plain `python bench.py` gets; it is restored
    by dropping the variable before the oracle's OpenMP runtime starts, with an explicit count as a fallback."""
    os.environ.pop("OMP_NUM_THREADS", None)
    from oracle import oracle as orc
    if orc.num_threads() <= 1:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        orc.set_num_threads(max(1, n // 2) if n >= 4 else n)   # physical cores: SMT siblings slow this code down
    return orc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from brush_b200.camera import build_uniforms
    orc = _oracle_on_all_cores()
    cam, tr, sh, op, v_out = scene_np()
    u = build_uniforms(cam, IMG_W, IMG_H)
    # One "step" of this arm is one full fwd+bwd pass of the workload on the host cores (seconds each).
    # The number of timed passes is bounded so that the whole run stays within a few minutes whatever
    # --steps says; ms_per_step is the mean over the passes actually timed.
    t_first = cpu_oracle_pass(u, tr, sh, op, v_out)  # warm-up (page-in, thread pool)
    n_pass = max(1, min(args.steps, int(150.0 / max(t_first, 1e-3))))
    t = [cpu_oracle_pass(u, tr, sh, op, v_out) for _ in range(n_pass)]
    ms = sum(t) / n_pass * 1e3
    val = IMG_W * IMG_H / (ms * 1e-3) / 1e6
    cores = orc.num_threads()
    sample = (f"{n_pass} full fwd+bwd passes of the same 1M-Gaussian 1080p scene timed (of {args.steps} steps requested; "
              f"bounded to ~150 s), OpenMP, {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU restatement of Brush's kernels (oracle/), not the Brush binary: "
                       "the reference needs cargo + wgpu and has no CPU path (SURVEY.md F3/F4)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)
    return 0


_REAL_STDOUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner), so the
    process-wide fd 1 is pointed at stderr for the whole run and the line goes to a private copy of the original."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import brush_b200.render as R
    from brush_b200.camera import build_uniforms

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cam0, tr, sh, op, v_out_np = scene_np()
    cam = rank_camera(cam0, rank)
    ctx = R.RenderContext(N_SPLATS, IMG_W, IMG_H, 0, device=local_rank)
    ttr, tsh, top = (torch.from_numpy(x).to(dev) for x in (tr, sh, op))
    v_out = torch.from_numpy(v_out_np).to(dev)
    v_out_host = torch.from_numpy(v_out_np).pin_memory()
    P = IMG_W * IMG_H

    from brush_b200.dp import FactoredGradients, FlatGradients, ShFactoredReducer, ViewShardedReducer
    # N>1: view-sharded DP.  Default exchange is SH-factored (all-reduce 44 N B + all-gather 12 N B per rank and a
    # local rebuild of v_sh); BG_DP_DENSE=1 selects the plain all-reduce of the dense (44+12K) N B gradient.
    factored = world > 1 and os.environ.get("BG_DP_DENSE") is None
    reducer = ViewShardedReducer(num_views_total=world)
    if factored:
        fg = FactoredGradients(N_SPLATS, SH_K, world, dev)
        fred = ShFactoredReducer(ctx, world)
        cam_positions = [rank_camera(cam0, r).position for r in range(world)]
        project_bwd = R.project_bwd_factored
    else:
        fg = FlatGradients(N_SPLATS, SH_K, dev)      # gradients live in one flat buffer: one collective per step
        project_bwd = R.project_bwd

    def allreduce(g):
        if world > 1:
            if factored:   # gradients AND refine statistics in two collectives
                fred.reduce(fg, ttr, cam_positions, last_out[0].visible, last_out[0].max_radius)
            else:
                reducer.reduce_flat(fg)               # SUM over ranks, 1/views scaling
                reducer.reduce_stats(g[3], last_out[0].visible, last_out[0].max_radius)

    last_out = [None]

    def step_compute():   # this rank's kernels (what the CUDA graph holds)
        out = R.render_splats(ctx, cam, (IMG_W, IMG_H), ttr, tsh, top)
        last_out[0] = out
        vc = R.rasterize_bwd(out, v_out)
        g = project_bwd(out, ttr, tsh, top, vc, outputs=fg.outputs())
        return out, g

    def step_device():
        out, g = step_compute()
        allreduce(g)
        return out, g

    copy_stream = torch.cuda.Stream(dev)
    staged = [torch.empty_like(v_out), torch.empty_like(v_out)]
    staged_ev = [torch.cuda.Event(), torch.cuda.Event()]
    result_host = torch.empty(8, dtype=torch.float32).pin_memory()

    def stage(i):  # H2D of the step's upstream gradient image from pinned host memory
        with torch.cuda.stream(copy_stream):
            staged[i & 1].copy_(v_out_host, non_blocking=True)
            staged_ev[i & 1].record(copy_stream)

    def step_e2e(i, last):
        torch.cuda.current_stream(dev).wait_event(staged_ev[i & 1])
        vo = staged[i & 1]
        out = R.render_splats(ctx, cam, (IMG_W, IMG_H), ttr, tsh, top)
        if not last:
            stage(i + 1)  # next step's upload overlaps this step's kernels
        last_out[0] = out
        vc = R.rasterize_bwd(out, vo)
        g = project_bwd(out, ttr, tsh, top, vc, outputs=fg.outputs())
        allreduce(g)
        g = fg.gradients() if factored else g
        res = torch.stack([g[0].sum(), g[1].sum(), g[2].sum(), g[3].sum()])
        result_host[:4].copy_(res, non_blocking=True)  # D2H of the step's result
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        w0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        sampler.mark(w0, time.time())
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- warm-up
    for _ in range(args.warmup):
        out, g = step_device()
    torch.cuda.synchronize(dev)
    V, I = out.num_visible, out.num_intersections
    toff = out.tile_offsets().cpu().numpy().astype(np.int64)
    per_tile = toff[..., 1] - toff[..., 0]
    T = per_tile.size

    # The step is ~20 short launches; replaying it as one CUDA graph removes the host launch gaps
    # (the library keeps nothing launch-specific on the host: counters and look-back epochs live on the device).
    # N>1: the graph holds this rank's kernels only; the NCCL collectives of the gradient exchange are issued eagerly
    # after each replay (capturing them too hung at N=2: two collectives on NCCL's internal stream inside one capture).
    use_graph = os.environ.get("BG_BENCH_NO_GRAPH") is None
    graph = None
    if use_graph:
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    step_compute()
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out, g = step_compute()
            for _ in range(3):
                graph.replay()
                allreduce(g)
            torch.cuda.synchronize(dev)
            assert out.num_visible == V and out.num_intersections == I
        except Exception as e:  # capture not possible: measure the eager loop instead
            sys.stderr.write(f"[bench] CUDA graph capture failed ({e}); timing eager launches\n")
            graph = None
    def step_graph(i):
        graph.replay()
        allreduce(g)     # no-op at N=1

    ms_dev = timed(step_graph if graph is not None else (lambda i: step_device()), args.steps)
    # ---- e2e: host input, copies inside the timed region
    stage(0)
    for i in range(2):
        step_e2e(i, False)
    torch.cuda.synchronize(dev)
    stage(0)
    ms_e2e = timed(lambda i: step_e2e(i, i == args.steps - 1), args.steps)
    _ = float(result_host[0])
    # ---- dominant kernel alone (blend backward) for the roofline figure
    out = R.render_splats(ctx, cam, (IMG_W, IMG_H), ttr, tsh, top)
    for _ in range(3):
        R.rasterize_bwd(out, v_out)
    ms_bwd = timed(lambda i: R.rasterize_bwd(out, v_out), args.steps) / args.steps
    ms_fwd_all = timed(lambda i: R.render_splats(ctx, cam, (IMG_W, IMG_H), ttr, tsh, top), args.steps) / args.steps

    ms_step = ms_dev / args.steps
    value = world * P / (ms_step * 1e-3) / 1e6
    e2e_value = world * P / (ms_e2e / args.steps * 1e-3) / 1e6
    peak, peak_src = measured_peak_gbs()
    algo_bytes = 40 * I + 32 * P + 80 * V
    achieved = algo_bytes / (ms_bwd * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("rasterize_bwd_kernel")
    except Exception:
        pass
    k16 = SH_K
    b_fwd = 52 * N_SPLATS + (200 + 12 * k16) * V + 88 * I + 8 * T + 16 * P
    b_bwd = 40 * I + 32 * P + (168 + 12 * k16) * V + (48 + 12 * k16) * N_SPLATS
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_gaussians": N_SPLATS, "width": IMG_W, "height": IMG_H, "sh_k": SH_K,
                   "num_visible": V, "num_intersections": I, "splats_per_tile_mean": float(per_tile.mean()),
                   "splats_per_tile_max": int(per_tile.max()),
                   "parallelism": "single GPU" if world == 1 else (f"view-sharded dp{world}, SH-factored exchange: all-reduce 48N B + all-gather 20N B/rank (gradients and refine statistics), v_sh rebuilt locally"
                                                                           if factored else f"view-sharded dp{world}, one NCCL all-reduce of the dense gradients per step"),
                   "launch": ("one CUDA graph replay per step" + ("" if world == 1 else " + eager NCCL exchange")) if graph is not None else "eager launches",
                   "l2": "inputs larger than L2 (236 MB of Gaussian parameters + 33 MB images per step vs 126 MB L2)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(v_out_host.numel() * 4), "d2h_bytes_per_step": 16 + 16,
                "note": "upstream-gradient image uploaded from pinned host memory every step (double-buffered on a copy stream), "
                        "gradient checksums + counters read back; Gaussian parameters stay resident as in the reference trainer"},
        "gpu_launches": (KERNELS_PER_STEP + (1 if factored else 0)) * args.steps,
        "clocks": None,
        "roofline": {"bound": "hbm", "kernel": "rasterize_bwd_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes": algo_bytes, "kernel_ms": ms_bwd,
                     "note": "blend kernels are FP32/SFU-issue bound, not HBM bound (SURVEY.md H5); frac is reported on the "
                             "mandated HBM basis", "pairs_upper_bound": int(I) * 256,
                     "pipeline_fwd_bwd": {"algorithmic_bytes": b_fwd + b_bwd, "achieved": (b_fwd + b_bwd) / (ms_step * 1e-3) / 1e9,
                                          "frac": (b_fwd + b_bwd) / (ms_step * 1e-3) / 1e9 / peak, "forward_ms": ms_fwd_all}},
    }
    if world == 1 and not args.no_train:
        # secondary figure of the metric: full train step (render + L1/SSIM loss + backward + Adam x3 + stats/noise)
        import brush_b200.train as T
        gt = torch.randint(0, 2 ** 31 - 1, (IMG_H, IMG_W), dtype=torch.int32, device=dev) | (255 << 24)
        splats = T.Splats(ttr.clone(), tsh.clone(), top.clone())
        trainer = T.SplatTrainer(T.TrainConfig(), ctx, T.bounds_from_pos(0.8, tr[:, :3]))
        batch = T.SceneBatch(img_packed=gt, camera=cam)
        for _ in range(3):
            trainer.step(batch, splats)
        ms_train = timed(lambda i: trainer.step(batch, splats), args.steps) / args.steps
        line["train"] = {"iters_per_s": 1e3 / ms_train, "ms_per_iter": ms_train,
                         "note": "SplatTrainer.step, 1 view/step, GT resident on the device, refine() not included"}
    if not args.no_train and 8 % world == 0:
        # BASELINE config [4] shape (at this scene's 1M Gaussians): ONE optimizer step over 8 views, views sharded over
        # the ranks, SH-factored exchange, identical Adam update on every rank (SplatTrainer.step_views)
        import brush_b200.train as T
        local = 8 // world
        batches = []
        g = torch.Generator(device=dev)
        g.manual_seed(1234)
        for v in range(local):
            gtv = torch.randint(0, 2 ** 31 - 1, (IMG_H, IMG_W), dtype=torch.int32, device=dev, generator=g) | (255 << 24)
            batches.append(T.SceneBatch(img_packed=gtv, camera=rank_camera(cam0, rank * local + v)))
        splats8 = T.Splats(ttr.clone(), tsh.clone(), top.clone())
        trainer8 = T.SplatTrainer(T.TrainConfig(), ctx, T.bounds_from_pos(0.8, tr[:, :3]))
        for _ in range(2):
            trainer8.step_views(batches, splats8)
        k8 = max(10, args.steps // 5)
        ms8 = timed(lambda i: trainer8.step_views(batches, splats8), k8) / k8
        line["train_8_views"] = {"iters_per_s": 1e3 / ms8, "ms_per_iter": ms8, "views_per_step": 8, "views_per_rank": local,
                                 "note": "SplatTrainer.step_views: 8 views per optimizer step sharded over the ranks "
                                         "(BASELINE config [4] at 1M Gaussians), SH-factored gradient exchange"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        orc = _oracle_on_all_cores()
        u = build_uniforms(cam0, IMG_W, IMG_H)
        cpu_oracle_pass(u, tr, sh, op, v_out_np)
        reps = 3
        dt = sum(cpu_oracle_pass(u, tr, sh, op, v_out_np) for _ in range(reps)) / reps
        line["cpu_baseline"] = {"value": P / dt / 1e6, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                "sample": f"{reps} full fwd+bwd passes of the same scene on the host cores (oracle/, OpenMP)"}
    line["clocks"] = sampler.stop()
    if rank == 0:
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
