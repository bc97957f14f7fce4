#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (contract: task statement, section 4).

Headline workload (BASELINE.json configs[1]): 1M synthetic Gaussians (SURVEY.md 8d generator, K=16), 1920x1080,
one "step" = render forward (RasterPass::Backward) + rasterize backward + project backward of one view.
metric = forward+backward Mpix/s.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # CPU arm: the restatement of the reference kernels (oracle/)
                                                           # on the host's physical cores (the reference itself needs
                                                           # cargo + wgpu: not buildable here)
  python bench.py --configs 1,3,4,2                        # which BASELINE configs ride in the line (default: all)

N > 1 (torchrun, one rank per GPU): the step is view-sharded -- every rank renders its own view of the replicated scene
and the ranks exchange the gradients over the library's NCCL communicator (bg_dp_exchange: all-reduce SUM 48 N B,
all-reduce MAX 8 N B, all-gather 12 N B per rank, interleaved per-Gaussian rows; the SH gradient stays in its per-view
rank-one form, which is what the optimiser pass consumes).  Weak scaling: per-GPU work is fixed, value = N * pixels / max-over-ranks step time.

The other configs ride in the same JSON line under "configs":
  "3": 4M Gaussians at 3840x2160 (sort / scan / blend stress), forward+backward, N=1
  "4": 2M Gaussians, ONE optimizer step over 8 views sharded over the N ranks (SplatTrainer.step_views)
  "2": a short end-to-end training run on a synthetic COLMAP-format 200-view set (loader -> step -> refine -> eval),
       N=1 (scripts/train_colmap.py holds the full-length version)
  "matrix": the sizes of the reference's own bench definitions (crates/brush-bench-test/src/benches.rs:222-287: SH degree 0,
       {0.5, 1, 2.5} M splats at 1080p and 2 M at four resolutions), forward (packed output) and forward+backward, N=1, on
       the SURVEY 8d scene generator (the reference's bench scene passes degrees where radians are expected)
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

SH_K = 16
CONFIGS = {   # BASELINE.json configs by index; seeds follow tests/scenes.py (0xB2000000 + index)
    1: dict(n=1_000_000, w=1920, h=1080, seed=0xB2000001, shift=0.0),
    3: dict(n=4_000_000, w=3840, h=2160, seed=0xB2000003, shift=-math.log(2.0)),   # focal doubles at 4K: scales halve
    4: dict(n=2_000_000, w=1920, h=1080, seed=0xB2000004, shift=0.0),
}
# benches.rs:222-287 (SPLAT_COUNTS x 1080p, 2M x RESOLUTIONS; with_sh_degree(0) :99).  Keys 100+ keep them apart from the
# BASELINE configs; the intersection arena is sized for the larger images (the same scene covers more tiles there).
MATRIX = [("0.5M@1080p", 500_000, 1920, 1080), ("1M@1080p", 1_000_000, 1920, 1080), ("2.5M@1080p", 2_500_000, 1920, 1080),
          ("2M@1024x1024", 2_000_000, 1024, 1024), ("2M@1080p", 2_000_000, 1920, 1080), ("2M@1440p", 2_000_000, 2560, 1440),
          ("2M@3200x1800", 2_000_000, 3200, 1800)]
for _i, (_name, _n, _w, _h) in enumerate(MATRIX):
    CONFIGS[100 + _i] = dict(n=_n, w=_w, h=_h, seed=0xB2000100 + _i, shift=0.0, k=1, forward_only=True,
                             isect_cap=_n * (32 if _w > 1920 else 16))
N_SPLATS, IMG_W, IMG_H = CONFIGS[1]["n"], CONFIGS[1]["w"], CONFIGS[1]["h"]
WORKLOAD = "configs[1]: 1M synthetic Gaussians (K=16), 1920x1080, render fwd + rasterize bwd + project bwd, 1 view/GPU"
METRIC, UNIT = "fwd+bwd Mpix/s @1M Gaussians 1080p", "Mpix/s"
KERNELS_PER_STEP = 14  # fwd: epoch bump, cull, 4 sort, scan, visible+emit, 2 sort, offsets, blend = 12; bwd: blend, project = 2


def base_config():
    """The keys both arms print, so that the driver can tell they ran the same workload."""
    return {"workload": WORKLOAD, "n_gaussians": N_SPLATS, "width": IMG_W, "height": IMG_H, "sh_k": SH_K}


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


def scene_np(cfg_idx=1):
    from scenes import random_v_output, synthetic_scene
    c = CONFIGS[cfg_idx]
    cam, tr, sh, op = synthetic_scene(c["n"], c["w"], c["h"], k=c.get("k", SH_K), seed=c["seed"], scale_shift=c["shift"])
    return cam, tr, sh, op, random_v_output(c["h"], c["w"])


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


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _oracle_on_physical_cores():
    """Both CPU legs (cpu_baseline of the default run, and --impl reference) run the restatement on the host's PHYSICAL
    cores: one OpenMP thread per core, SMT siblings left idle (they slow this code down).  torchrun exports
    OMP_NUM_THREADS=1 to its workers, so the count is always set explicitly."""
    os.environ.pop("OMP_NUM_THREADS", None)
    from oracle import oracle as orc
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    orc.set_num_threads(max(1, n // 2) if n >= 4 else n)
    return orc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from brush_b200.camera import build_uniforms
    orc = _oracle_on_physical_cores()
    cam, tr, sh, op, v_out = scene_np(1)
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
              f"bounded to ~150 s), OpenMP, {cores} threads = physical cores of {cpu_model()}")
    cfg = base_config()
    cfg["note"] = ("CPU restatement of Brush's kernels (oracle/), not the Brush binary: the reference needs cargo + wgpu and "
                   "has no CPU path (SURVEY.md F3/F4)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
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


def log(msg):
    sys.stderr.write(f"[bench] {msg}\n")
    sys.stderr.flush()


class Bench:
    """Shared timing helpers of the GPU legs (device timing with CUDA events, max over ranks)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.sampler = ClockSampler(self.local_rank)
        self.sampler.start()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps):
        """Total milliseconds of `steps` calls of fn(i): barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        self.barrier()
        w0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        self.sampler.mark(w0, time.time())
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def capture(self, fn, replays=3):
        """fn as ONE CUDA graph (None if capture is not possible)."""
        torch = self.torch
        try:
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn()
            torch.cuda.current_stream(self.dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            for _ in range(replays):
                g.replay()
            torch.cuda.synchronize(self.dev)
            return g
        except Exception as e:  # capture not possible: measure the eager loop instead
            log(f"CUDA graph capture failed ({e}); timing eager launches")
            return None


def fwd_bwd_leg(B: Bench, cfg_idx: int, steps: int, warmup: int, headline: bool):
    """forward + rasterize backward + project backward of one view per rank; returns the measurements as a dict."""
    import brush_b200.render as R
    torch = B.torch
    c = CONFIGS[cfg_idx]
    n, w, h = c["n"], c["w"], c["h"]
    world, rank, dev = B.world, B.rank, B.dev
    cam0, tr, sh, op, v_out_np = scene_np(cfg_idx)
    cam = rank_camera(cam0, rank)
    ctx = R.RenderContext(n, w, h, int(c.get("isect_cap", 0)), device=B.local_rank)
    ttr, tsh, top = (torch.from_numpy(x).to(dev) for x in (tr, sh, op))
    v_out = torch.from_numpy(v_out_np).to(dev)
    sh_k = int(c.get("k", SH_K))
    P = w * h
    dp = world > 1
    comm = None
    if dp:
        from brush_b200.dp import DpComm
        comm = DpComm(ctx)
        small = torch.zeros(12 * n, dtype=torch.float32, device=dev)       # rows: v_transforms | v_raw_opac | visible
        stat = torch.zeros(2 * n, dtype=torch.float32, device=dev)         # rows: v_refine | max_radius (MAX)
        record = torch.zeros(3 * n, dtype=torch.float32, device=dev)       # rows: v_color of the rank's view
        recv = torch.zeros(world * 3 * n, dtype=torch.float32, device=dev)
        outs = (torch.empty((n, 10), device=dev), torch.empty((n, 3), device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev))
        dense = None
    else:
        dense = (torch.empty((n, 10), device=dev), torch.empty((n, sh_k, 3), device=dev), torch.empty(n, device=dev),
                 torch.empty(n, device=dev))
    last = [None]
    staged = [v_out]     # upstream-gradient input of the backward (static buffer)

    def compute_fwd():
        last[0] = R.render_splats(ctx, cam, (w, h), ttr, tsh, top)
        return last[0]

    def compute_bwd(vo):
        out = last[0]
        vc = R.rasterize_bwd(out, vo)
        if dp:
            g = R.project_bwd_factored(out, ttr, tsh, top, vc, outputs=outs)
            comm.pack_view(n, 1, 0, True, outs[0], outs[2], outs[1], outs[3], out.visible, out.max_radius, small, stat, record)
        else:
            g = R.project_bwd(out, ttr, tsh, top, vc, outputs=dense)
        return g

    def exchange():
        if dp:
            comm.exchange(n, 1, small, stat, record, recv, chunks=1)

    def step():
        out = compute_fwd()
        g = compute_bwd(staged[0])
        exchange()
        return out, g

    for _ in range(warmup):
        out, g = step()
    torch.cuda.synchronize(dev)
    V, I = out.num_visible, out.num_intersections
    overflow = out.intersection_overflow
    toff = out.tile_offsets().cpu().numpy().astype(np.int64)
    per_tile = toff[..., 1] - toff[..., 0]
    T = per_tile.size
    stats = R.blend_stats(out, v_out)
    # The step is ~20 short launches; replaying it from CUDA graphs removes the host launch gaps (the library keeps
    # nothing launch-specific on the host: counters and look-back epochs live on the device).  Two graphs -- the forward,
    # and the backward once per input buffer -- so that the e2e leg can upload the next step's input under the forward.
    # N>1: the graphs hold this rank's kernels; the exchange (NCCL on the communicator's own stream) is issued eagerly.
    log(f"config [{cfg_idx}]: warm-up done (V={V}, I={I}); capturing the step")
    g_fwd = g_bwd = None
    if not os.environ.get("BG_BENCH_NO_GRAPH"):
        g_fwd = B.capture(compute_fwd)
        if g_fwd is not None:
            g_bwd = [B.capture(lambda: compute_bwd(staged[0]))]
            if any(x is None for x in g_bwd):
                g_fwd = g_bwd = None
    if g_fwd is not None:
        launch = "two CUDA graph replays per step (forward, backward)" + (" + eager exchange" if dp else "")
        run_compute = lambda: (g_fwd.replay(), g_bwd[0].replay())
    else:
        launch = "eager launches"
        run_compute = lambda: (compute_fwd(), compute_bwd(staged[0]))
    run_step = lambda i: (run_compute(), exchange())
    for _ in range(3):
        run_step(0)
    torch.cuda.synchronize(dev)
    assert last[0].num_visible == V and last[0].num_intersections == I
    log(f"config [{cfg_idx}]: timing {steps} steps ({launch})")
    ms_step = B.timed(run_step, steps) / steps
    res = {"n": n, "w": w, "h": h, "V": V, "I": I, "T": T, "P": P, "ms_step": ms_step, "launch": launch, "overflow": overflow,
           "per_tile_mean": float(per_tile.mean()), "per_tile_max": int(per_tile.max()), "stats": stats}
    if c.get("forward_only"):   # RasterPass::Forward (packed rgba8 output), as the reference's forward_rendering group renders
        f_only = lambda: R.render_splats(ctx, cam, (w, h), ttr, tsh, top, rpass=R.PASS_FORWARD)
        for _ in range(3):
            f_only()
        torch.cuda.synchronize(dev)
        res["ms_forward_packed"] = B.timed(lambda i: f_only(), steps) / steps
    if dp:   # phases: this rank's kernels alone (the same graphs), the exchange alone
        res["ms_compute"] = B.timed(lambda i: run_compute(), steps) / steps
        res["ms_exchange"] = B.timed(lambda i: exchange(), steps) / steps
    if headline:
        # ---- e2e: host input, copies inside the timed region.  The per-step host input of this path in the reference's
        # trainer is the ground-truth image (train.rs:197-198 uploads the batch image; the upstream gradient is born on the
        # device).  Step i: the forward runs while the copy stream uploads image i+1 (packed rgba8, pinned) into buffer
        # (i+1)&1; then the fused L1+SSIM kernel turns render + image into the upstream gradient, the backward follows,
        # and the loss is read back.  So this leg does MORE device work per step than `value`
        # (the loss kernel) -- its copies are what it is about.
        from brush_b200.loss import ImageLossConfig, image_loss_fused
        gen = np.random.default_rng(99)
        gt_host = torch.from_numpy((gen.integers(0, 2 ** 24, size=(h, w), dtype=np.int64) | (255 << 24)).astype(np.uint32)
                                   .view(np.int32)).pin_memory()
        gt_dev = [torch.empty((h, w), dtype=torch.int32, device=dev) for _ in range(2)]
        lcfg = ImageLossConfig(0.8, -0.2, None, False)     # train.rs:220-249 defaults: (1 - 0.2) L1 - 0.2 SSIM
        chain = [1.0 / (3.0 * P)] * 3
        chain_dev = torch.tensor(chain, dtype=torch.float32, device=dev)
        v_img = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
        loss_out = [None]
        copy_stream = torch.cuda.Stream(dev)
        staged_ev = [torch.cuda.Event(), torch.cuda.Event()]
        bwd_done = [torch.cuda.Event(), torch.cuda.Event()]
        result_host = torch.empty(8, dtype=torch.float32).pin_memory()
        for e in bwd_done:
            e.record()

        def loss_bwd(b):
            vo, loss = image_loss_fused(ctx, last[0].out_img, gt_dev[b], 3, lcfg, chain, v_img, weights=chain_dev)
            g = compute_bwd(vo)
            loss_out[0] = loss
            return g

        g_lb = None
        if g_fwd is not None:
            g_lb = [B.capture(lambda b=b: loss_bwd(b)) for b in range(2)]
            if any(x is None for x in g_lb):
                g_lb = None

        def stage(i):  # H2D of step i's image; must not overwrite the buffer before the loss kernel of step i-2 has read it
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(bwd_done[i & 1])
                gt_dev[i & 1].copy_(gt_host, non_blocking=True)
                staged_ev[i & 1].record(copy_stream)

        def step_e2e(i, is_last):
            cur = torch.cuda.current_stream(dev)
            if not is_last:
                stage(i + 1)  # next step's upload overlaps this step's kernels
            if g_lb is not None:
                g_fwd.replay()
            else:
                compute_fwd()
            cur.wait_event(staged_ev[i & 1])
            if g_lb is not None:
                g_lb[i & 1].replay()
                g = outs if dp else dense
            else:
                g = loss_bwd(i & 1)
            bwd_done[i & 1].record(cur)
            exchange()
            result_host[:1].copy_(loss_out[0].reshape(1), non_blocking=True)  # D2H of the step's result: the loss

        stage(0)
        for i in range(2):
            step_e2e(i, False)
        torch.cuda.synchronize(dev)
        stage(0)
        res["ms_e2e"] = B.timed(lambda i: step_e2e(i, i == steps - 1), steps) / steps
        res["e2e_loss"] = float(result_host[0])
        res["h2d_bytes"] = int(gt_host.numel() * 4)
        # ---- dominant kernel alone (blend backward) for the roofline figure, and the forward alone
        out = R.render_splats(ctx, cam, (w, h), ttr, tsh, top)
        for _ in range(3):
            R.rasterize_bwd(out, v_out)
        res["ms_bwd_kernel"] = B.timed(lambda i: R.rasterize_bwd(out, v_out), steps) / steps
        res["ms_forward"] = B.timed(lambda i: R.render_splats(ctx, cam, (w, h), ttr, tsh, top), steps) / steps
        res["scene"] = (cam0, tr, sh, op, v_out_np)
        res["ctx"], res["dev_params"] = ctx, (ttr, tsh, top)
    if comm is not None:
        comm.close()
    if not headline:
        ctx.close()
    return res


def train_step_leg(B: Bench, ctx, dev_params, tr_np, cam, steps):
    """Secondary figure of the metric at N=1: SplatTrainer.step (render + L1/SSIM loss + backward + the update pass)."""
    import brush_b200.train as T
    torch = B.torch
    ttr, tsh, top = dev_params
    gt = torch.randint(0, 2 ** 31 - 1, (IMG_H, IMG_W), dtype=torch.int32, device=B.dev) | (255 << 24)
    splats = T.Splats(ttr.clone(), tsh.clone(), top.clone())
    trainer = T.SplatTrainer(T.TrainConfig(), ctx, T.bounds_from_pos(0.8, tr_np[:, :3]))
    batch = T.SceneBatch(img_packed=gt, camera=cam)
    for _ in range(3):
        trainer.step(batch, splats)
    ms = B.timed(lambda i: trainer.step(batch, splats), steps) / steps
    for _ in range(3):
        trainer.step_fused(batch, splats)
    ms_abi = B.timed(lambda i: trainer.step_fused(batch, splats), steps) / steps
    return {"iters_per_s": 1e3 / min(ms, ms_abi), "ms_per_iter": min(ms, ms_abi), "ms_host_orchestrated": ms, "ms_one_abi_call": ms_abi,
            "note": "SplatTrainer.step, 1 view/step, 1M Gaussians 1080p, GT resident on the device, refine() not included"}


def views_leg(B: Bench, steps: int):
    """BASELINE config [4]: ONE optimizer step over 8 views, 2M Gaussians, views sharded over the ranks."""
    import brush_b200.render as R
    import brush_b200.train as T
    torch = B.torch
    world, rank, dev = B.world, B.rank, B.dev
    c = CONFIGS[4]
    n, w, h = c["n"], c["w"], c["h"]
    cam0, tr, sh, op, _ = scene_np(4)
    ctx = R.RenderContext(n, w, h, 0, device=B.local_rank)
    local = 8 // world
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    batches = []
    for v in range(local):
        gtv = torch.randint(0, 2 ** 31 - 1, (h, w), dtype=torch.int32, device=dev, generator=gen) | (255 << 24)
        batches.append(T.SceneBatch(img_packed=gtv, camera=rank_camera(cam0, rank * local + v)))
    params = [torch.from_numpy(x).to(dev) for x in (tr, sh, op)]
    splats = T.Splats(*params)
    trainer = T.SplatTrainer(T.TrainConfig(), ctx, T.bounds_from_pos(0.8, tr[:, :3]))
    log(f"config [4]: warm-up ({local} views per rank)")
    for _ in range(2):
        trainer.step_views(batches, splats)
    torch.cuda.synchronize(dev)
    log("config [4]: timing")
    k8 = max(10, steps // 5)
    ms = B.timed(lambda i: trainer.step_views(batches, splats), k8) / k8
    res = {"iters_per_s": 1e3 / ms, "ms_per_iter": ms, "views_per_step": 8, "views_per_rank": local, "n_gaussians": n,
           "width": w, "height": h,
           "note": "bg_train_step_views: 8 views per optimizer step sharded over the ranks, SH-factored gradient exchange, "
                   "SH part of the update pass under the all-reduce"}
    if world > 1:   # phase: the same step without the exchange (this rank's views only)
        s2 = T.Splats(*(p.clone() for p in params))
        t2 = T.SplatTrainer(T.TrainConfig(), ctx, T.bounds_from_pos(0.8, tr[:, :3]))
        for _ in range(2):
            t2.step_views(batches, s2, distributed=False)
        res["ms_local_views_no_exchange"] = B.timed(lambda i: t2.step_views(batches, s2, distributed=False), k8) / k8
    ctx.close()
    return res


def train_run_leg(B: Bench, iters: int):
    """BASELINE config [2], shortened: synthetic COLMAP-format 200-view set -> loader -> step -> refine -> eval."""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import train_colmap
    return train_colmap.run(device=B.local_rank, iters=iters, views=200, width=IMG_W, height=IMG_H, init_points=500_000,
                            max_splats=2_000_000, quiet=True)


def matrix_leg(B: Bench, steps: int):
    """The reference's bench matrix (benches.rs:222-287) on this path: per point forward-only and forward+backward."""
    out = {}
    for i, (name, n, w, h) in enumerate(MATRIX):
        try:
            r = fwd_bwd_leg(B, 100 + i, steps, 3, headline=False)
            P = r["P"]
            out[name] = {"n_gaussians": n, "width": w, "height": h, "sh_k": 1,
                         "fwd_bwd_ms": r["ms_step"], "fwd_bwd_mpix_per_s": P / (r["ms_step"] * 1e-3) / 1e6,
                         "forward_ms": r.get("ms_forward_packed"),
                         "forward_mpix_per_s": (P / (r["ms_forward_packed"] * 1e-3) / 1e6) if r.get("ms_forward_packed") else None,
                         "num_visible": r["V"], "num_intersections": r["I"], "overflow": r["overflow"],
                         "pairs_live": r["stats"]["pairs_live"], "launch": r["launch"]}
        except Exception as e:   # one point must not take the line down
            out[name] = {"error": repr(e)}
        B.torch.cuda.empty_cache()
    return out


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--configs", default="1,3,4,2,matrix",
                    help="BASELINE configs to run (1 is always run); `matrix` = the reference's bench sizes (benches.rs:222-287)")
    ap.add_argument("--train-iters", type=int, default=600, help="length of the config [2] run inside the bench line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    toks = [x.strip() for x in args.configs.split(",") if x.strip()]
    want = {int(x) for x in toks if x.isdigit()} | {1}
    want_matrix = any(x in ("matrix", "m") for x in toks)

    from brush_b200.camera import build_uniforms
    B = Bench(args)
    world, rank = B.world, B.rank

    log(f"rank {rank}/{world}: config [1]")
    h1 = fwd_bwd_leg(B, 1, args.steps, args.warmup, headline=True)
    log("config [1] done")
    V, I, T, P = h1["V"], h1["I"], h1["T"], h1["P"]
    ms_step = h1["ms_step"]
    value = world * P / (ms_step * 1e-3) / 1e6
    e2e_value = world * P / (h1["ms_e2e"] * 1e-3) / 1e6
    peak, peak_src = measured_peak_gbs()
    algo_bytes = 40 * I + 32 * P + 80 * V
    ms_bwd = h1["ms_bwd_kernel"]
    achieved = algo_bytes / (ms_bwd * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic, traffic_src = tj.get("blend_bwd_kernel"), tj.get("source")
    except Exception:
        pass
    b_fwd = 52 * N_SPLATS + (200 + 12 * SH_K) * V + 88 * I + 8 * T + 16 * P
    b_bwd = 40 * I + 32 * P + (168 + 12 * SH_K) * V + (48 + 12 * SH_K) * N_SPLATS
    st = h1["stats"]
    cfg = base_config()
    cfg.update({
        "num_visible": V, "num_intersections": I, "splats_per_tile_mean": h1["per_tile_mean"], "splats_per_tile_max": h1["per_tile_max"],
        "pairs": {"tile_list_entries": st["tile_list_entries"], "evaluated": st["pairs_evaluated"], "live": st["pairs_live"],
                  "stopping": st["pairs_stopping"], "lane_utilisation": st["lane_utilisation"],
                  "note": "blend loop of one view: pixel-splat pairs evaluated by the backward walk (64 per warp-splat iteration) / "
                          "pairs that blended; the forward hands the backward the exact splat sets, so untouched tile-list "
                          "entries cost nothing in the backward"},
        "parallelism": "single GPU" if world == 1 else
                       f"view-sharded dp{world}: one exchange per step on the library's NCCL communicator (all-reduce SUM 48N B + "
                       "all-reduce MAX 8N B + all-gather 12N B per rank: gradients and refine statistics; the SH gradient "
                       "stays per-view rank one)",
        "launch": h1["launch"],
        "l2": "inputs larger than L2 (236 MB of Gaussian parameters + 33 MB images per step vs 126 MB L2)"})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h1["h2d_bytes"], "d2h_bytes_per_step": 4,
                "loss": h1["e2e_loss"],
                "note": "per step: the ground-truth image (packed rgba8, pinned host memory) is uploaded on a copy stream under the "
                        "forward, the fused L1+SSIM kernel turns render + image into the upstream gradient, the backward follows; "
                        "the loss is read back.  One kernel MORE per step than `value` (the loss); Gaussian "
                        "parameters stay resident as in the reference trainer (train.rs:197-198 uploads the batch image only)"},
        "gpu_launches": (KERNELS_PER_STEP + (1 if world > 1 else 0)) * args.steps,
        "clocks": None,
        "roofline": {"bound": "hbm", "kernel": "blend_bwd_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes": algo_bytes, "kernel_ms": ms_bwd,
                     "note": "the blend kernels are instruction-issue bound (FP32 + MUFU + select), not HBM bound (SURVEY.md H5): "
                             "frac is reported on the mandated HBM basis; instructions per live pair are in profiles/",
                     "pairs_upper_bound": int(I) * 256, "pairs_live": st["pairs_live"],
                     "ns_per_live_pair_bwd": ms_bwd * 1e6 / max(st["pairs_live"], 1),
                     "pipeline_fwd_bwd": {"algorithmic_bytes": b_fwd + b_bwd, "achieved": (b_fwd + b_bwd) / (ms_step * 1e-3) / 1e9,
                                          "frac": (b_fwd + b_bwd) / (ms_step * 1e-3) / 1e9 / peak, "forward_ms": h1["ms_forward"]}},
    }
    if world > 1:
        line["phases"] = {"compute_ms": h1["ms_compute"], "exchange_ms": h1["ms_exchange"], "step_ms": ms_step}
    cam0, tr, sh, op, v_out_np = h1["scene"]
    if world == 1 and not args.no_train:
        line["train"] = train_step_leg(B, h1["ctx"], h1["dev_params"], tr, rank_camera(cam0, 0), args.steps)
    h1["ctx"].close()
    del h1["dev_params"]
    B.torch.cuda.empty_cache()
    extra = {}
    if 3 in want and world == 1:
        log("config [3]: 4M Gaussians at 3840x2160")
        r3 = fwd_bwd_leg(B, 3, max(10, args.steps // 5), 3, headline=False)
        extra["3"] = {"workload": "configs[3]: 4M synthetic Gaussians, 3840x2160, fwd+bwd, 1 view", "mpix_per_s": r3["P"] / (r3["ms_step"] * 1e-3) / 1e6,
                      "ms_per_step": r3["ms_step"], "num_visible": r3["V"], "num_intersections": r3["I"], "overflow": r3["overflow"],
                      "pairs_live": r3["stats"]["pairs_live"], "pairs_evaluated": r3["stats"]["pairs_evaluated"], "launch": r3["launch"]}
        B.torch.cuda.empty_cache()
    if 4 in want and not args.no_train and 8 % world == 0:
        log("config [4]: 8 views per optimizer step, 2M Gaussians")
        extra["4"] = views_leg(B, args.steps)
        line["train_8_views"] = extra["4"]
        B.torch.cuda.empty_cache()
    if 2 in want and world == 1 and not args.no_train:
        log("config [2]: end-to-end training run on a synthetic COLMAP set")
        try:
            extra["2"] = train_run_leg(B, args.train_iters)
        except Exception as e:   # the headline must survive a failure of this leg
            extra["2"] = {"error": repr(e)}
    if want_matrix and world == 1:   # last GPU leg: nothing measured above depends on it
        log("matrix: the reference's bench sizes (SH degree 0)")
        try:
            extra["matrix"] = matrix_leg(B, max(10, args.steps // 5))
        except Exception as e:
            extra["matrix"] = {"error": repr(e)}
    line["configs"] = extra
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        orc = _oracle_on_physical_cores()
        u = build_uniforms(cam0, IMG_W, IMG_H)
        cpu_oracle_pass(u, tr, sh, op, v_out_np)
        reps = 3
        dt = sum(cpu_oracle_pass(u, tr, sh, op, v_out_np) for _ in range(reps)) / reps
        line["cpu_baseline"] = {"value": P / dt / 1e6, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                "sample": f"{reps} full fwd+bwd passes of the same scene on the host's physical cores "
                                          f"({cpu_model()}; oracle/, OpenMP)"}
    line["clocks"] = B.sampler.stop()
    if rank == 0:
        _emit(line)
    if world > 1:
        B.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
