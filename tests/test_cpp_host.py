"""C++ host layer (include/brush_b200.hpp): compiled with g++ against the C ABI.  CPU: the camera uniforms must be
the bits brush_b200.camera.build_uniforms produces (the oracle and the Python path use those), fov<->focal round trips,
error behaviour.  GPU: a forward+backward through the C++ operators equals the same calls through the Python mirror."""
import math
import os
import struct
import subprocess

import numpy as np
import pytest

from brush_b200 import camera as cm
from brush_b200.camera import Camera, build_uniforms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "host_check")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


@pytest.fixture(scope="module")
def exe():
    from brush_b200 import build
    build.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "host_check.cpp")
    hdrs = [os.path.join(ROOT, "include", h) for h in ("brush_b200.hpp", "brush_b200.h")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(p) for p in [src] + hdrs):
        lib = os.path.join(ROOT, "brush_b200")
        cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"), src,
               "-o", EXE, "-L", lib, "-lbrush_b200", "-L", os.path.join(CUDA, "lib64"), "-lcudart",
               f"-Wl,-rpath,{lib}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    return EXE


def _cam_line(c: Camera, w, h):
    p = list(c.model_params) + [0.0] * (8 - len(c.model_params))
    vals = [*c.position, *c.rotation, c.fov_x, c.fov_y, *c.center_uv]
    return " ".join(repr(float(v)) for v in vals) + f" {c.camera_model} " + " ".join(repr(float(v)) for v in p) + f" {w} {h}"


def _cameras():
    rng = np.random.default_rng(4)
    cams = []
    models = [(cm.PINHOLE, ()), (cm.KANNALA_BRANDT_4, (-0.05, 0.01, -0.001, 5e-5)),
              (cm.RADIAL_TANGENTIAL_8, (-0.2, 0.05, -0.001, 0.01, 0.0, 0.0, 1e-3, -1e-3)),
              (cm.THIN_PRISM_FISHEYE, (-0.05, 0.01, -0.001, 5e-5, 1e-3, -1e-3, 5e-4, -5e-4))]
    for i in range(24):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        model, params = models[i % 4]
        cams.append((Camera(position=tuple(float(np.float32(v)) for v in rng.normal(size=3) * 3),
                            rotation=tuple(float(np.float32(v)) for v in q), fov_x=float(rng.uniform(0.3, 1.6)),
                            fov_y=float(rng.uniform(0.3, 1.4)),
                            center_uv=(float(np.float32(rng.uniform(0.4, 0.6))), float(np.float32(rng.uniform(0.4, 0.6)))),
                            camera_model=model, model_params=params), int(rng.integers(16, 4000)), int(rng.integers(16, 2200))))
    return cams


def test_uniforms_match_the_python_host_bit_for_bit(exe):
    cams = _cameras()
    inp = "\n".join(_cam_line(c, w, h) for c, w, h in cams) + "\n"
    r = subprocess.run([exe, "uniforms"], input=inp, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == len(cams)
    for (c, w, h), ln in zip(cams, lines):
        tok = ln.split()
        got = np.array([float.fromhex(t) for t in tok[:24]], np.float32)
        u = build_uniforms(c, w, h)
        want = np.array([*u.viewmat, u.fx, u.fy, u.cx, u.cy, *u.cam_pos, u.lim_pos_x, u.lim_pos_y, u.lim_neg_x, u.lim_neg_y,
                         u.half_max_render_fov], np.float32)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (c, w, h, got, want)
        assert int(tok[24]) == c.camera_model
        params = np.array([float.fromhex(t) for t in tok[25:33]], np.float32)
        assert np.array_equal(params[:len(c.model_params)], np.array(c.model_params, np.float32))


def test_fov_focal_round_trip_cpp(exe):
    """brush-render/src/tests/mod.rs:710-790 through the C++ functions."""
    cases = [(800.0, 1920, 0, [0] * 8), (300.0, 1024, 1, [0] * 8), (280.0, 1024, 1, [-0.01, 0.003, -0.0005, 0.00002, 0, 0, 0, 0]),
             (900.0, 1920, 2, [-0.2, 0.05, -0.001, 0, 0, 0, 0, 0]), (280.0, 1024, 3, [-0.01, 0.003, -0.0005, 0.00002, 1e-3, -2e-3, 5e-4, -5e-4])]
    inp = "".join(f"{f} {px} {m} " + " ".join(repr(float(v)) for v in p) + "\n" for f, px, m, p in cases)
    r = subprocess.run([exe, "fov"], input=inp, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for (f, px, m, p), ln in zip(cases, r.stdout.strip().splitlines()):
        fov, back = (float(t) for t in ln.split())
        assert abs(back - f) < 1e-6
        assert abs(fov - cm.focal_to_fov(f, px, m, tuple(p[:cm.MODEL_PARAM_COUNT[m]]))) < 1e-12
    assert abs(float(r.stdout.splitlines()[1].split()[0]) - 1024 / 300.0) < 1e-9   # zero-distortion KB4: fov = pixels / f


def test_errors_are_exceptions(exe):
    r = subprocess.run([exe, "errors"], capture_output=True, text=True)
    assert r.returncode == 0 and "caught 3" in r.stdout and "caught_more 3" in r.stdout, (r.stdout, r.stderr)


def test_bounding_box_median_size_cpp(exe):
    """brush-render/src/bounding_box.rs:36-58: NaN extents do not break the ordering (f32::total_cmp), normal case = 4."""
    r = subprocess.run([exe, "bounds"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with_nan, all_nan, normal = (float(t) for t in r.stdout.split())
    assert math.isfinite(with_nan) and with_nan == 6.0      # total order: [2, 3, NaN] -> 3 * 2
    assert math.isnan(all_nan)
    assert abs(normal - 4.0) < 1e-6


@pytest.mark.gpu
def test_cpp_operators_match_python_mirror(exe, tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import brush_b200.render as R
    from scenes import random_v_output, synthetic_scene
    n, w, h, k = 20_000, 320, 240, 4
    cam0, tr, sh, op = synthetic_scene(n, w, h, k=k, seed=77)
    cam = Camera(position=(0.05, -0.02, 0.1), rotation=(0.01, 0.03, -0.02, 0.999), fov_x=cam0.fov_x, fov_y=cam0.fov_y,
                 center_uv=(0.49, 0.52))
    v_out = random_v_output(h, w)
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    line = _cam_line(cam, w, h).encode()
    scene = tmp_path / "scene.bin"
    with open(scene, "wb") as f:
        f.write(struct.pack("<6I", n, k, w, h, 0, 1))
        f.write(struct.pack("<I", len(line)) + line)
        f.write(bg.tobytes() + tr.tobytes() + sh.tobytes() + op.tobytes() + v_out.tobytes())
    outp = tmp_path / "out.bin"
    r = subprocess.run([exe, "render", str(scene), str(outp)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(outp, "rb").read()
    V, I = struct.unpack_from("<2I", raw, 0)
    off = 8
    def take(count):
        nonlocal off
        a = np.frombuffer(raw, np.float32, count, off)
        off += 4 * count
        return a
    img, vt, vsh, vo, vis = take(w * h * 4).reshape(h, w, 4), take(n * 10).reshape(n, 10), take(n * k * 3).reshape(n, k, 3), take(n), take(n)
    ctx = R.RenderContext(n, w, h)
    d = ctx.device
    ttr, tsh, top = (torch.from_numpy(x).to(d) for x in (tr, sh, op))
    out = R.render_splats(ctx, cam, (w, h), ttr, tsh, top, background=tuple(float(b) for b in bg))
    assert (out.num_visible, out.num_intersections) == (V, I)
    assert np.array_equal(out.out_img.cpu().numpy(), img) and np.array_equal(out.visible.cpu().numpy(), vis)
    vc = R.rasterize_bwd(out, torch.from_numpy(v_out).to(d))
    pvt, pvsh, pvo, _ = R.project_bwd(out, ttr, tsh, top, vc)
    for a, b in ((vt, pvt), (vsh, pvsh), (vo, pvo)):   # f32 atomics in the blend backward: equal up to summation order
        b = b.cpu().numpy()
        assert np.linalg.norm(a.astype(np.float64) - b) <= 1e-4 * np.linalg.norm(b) + 1e-12
    ctx.close()


@pytest.mark.gpu
def test_cpp_splat_trainer_matches_python_fused_step(exe, tmp_path):
    """brush_b200::SplatTrainer (C++) and SplatTrainer.step_fused (Python) drive the same bg_train_step: same losses."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import brush_b200.render as R
    import brush_b200.train as T
    from scenes import synthetic_scene
    n, w, h, k = 15_000, 192, 128, 4
    cam, tr, sh, op = synthetic_scene(n, w, h, k=k, seed=55)
    ctx = R.RenderContext(n, w, h)
    d = ctx.device
    tgt = R.render_splats(ctx, cam, (w, h), *(torch.from_numpy(x).to(d) for x in (tr, sh, op)), rpass=0)
    gt = (tgt.out_img | (255 << 24)).clone()
    sh0 = (sh + np.float32(0.1)).astype(np.float32)
    bounds = T.bounds_from_pos(0.8, tr[:, :3])
    line = _cam_line(cam, w, h).encode()
    scene = tmp_path / "train.bin"
    with open(scene, "wb") as f:
        f.write(struct.pack("<6I", n, k, w, h, 0, 1))
        f.write(struct.pack("<I", len(line)) + line)
        f.write(np.zeros(3, np.float32).tobytes() + tr.tobytes() + sh0.tobytes() + op.tobytes())
        f.write(np.zeros((h, w, 4), np.float32).tobytes())
        f.write(gt.cpu().numpy().astype(np.int32).tobytes())
        f.write(struct.pack("<f", bounds.median_size()))
    r = subprocess.run([exe, "train", str(scene), "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cpp_losses = [float(ln.split()[1]) for ln in r.stdout.strip().splitlines() if ln.startswith("loss")]
    cfg = T.TrainConfig(total_train_iters=1000, background_noise_strength=0.0, seed=7)
    splats = T.Splats(*(torch.from_numpy(x.copy()).to(d) for x in (tr, sh0, op)))
    trainer = T.SplatTrainer(cfg, ctx, bounds)
    batch = T.SceneBatch(img_packed=gt, camera=cam)
    py_losses = [float(trainer.step_fused(batch, splats).loss.item()) for _ in range(3)]
    assert len(cpp_losses) == 3 and all(math.isfinite(x) for x in cpp_losses)
    np.testing.assert_allclose(cpp_losses, py_losses, rtol=1e-3)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="written after the round's GPU budget was spent: compiled and error-path checked on the CPU, "
                                        "not yet run on hardware")
def test_cpp_refine_and_step_views_match_python_mirror(exe, tmp_path):
    """brush_b200::SplatTrainer::step_views / refine (C++) against SplatTrainer.step_views / refine (Python): both drive
    bg_train_step_views and bg_refine with the same seed, so the refine counts agree (up to the blend backward's f32
    atomics moving a value across a threshold) and the losses match."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import brush_b200.render as R
    import brush_b200.train as T
    from scenes import synthetic_scene
    n, w, h, k = 15_000, 192, 128, 4
    cam, tr, sh, op = synthetic_scene(n, w, h, k=k, seed=56)
    op = op.copy()
    op[:500] = np.float32(-8.0)        # opacity < 1/255: pruned and replaced (train.rs:455-520)
    ctx = R.RenderContext(2 * n, w, h)
    d = ctx.device
    tgt = R.render_splats(ctx, cam, (w, h), *(torch.from_numpy(x).to(d) for x in (tr, sh, op)), rpass=0)
    gt = (tgt.out_img | (255 << 24)).clone()
    sh0 = (sh + np.float32(0.1)).astype(np.float32)
    line = _cam_line(cam, w, h).encode()
    scene = tmp_path / "refine.bin"
    with open(scene, "wb") as f:
        f.write(struct.pack("<6I", n, k, w, h, 0, 1))
        f.write(struct.pack("<I", len(line)) + line)
        f.write(np.zeros(3, np.float32).tobytes() + tr.tobytes() + sh0.tobytes() + op.tobytes())
        f.write(np.zeros((h, w, 4), np.float32).tobytes())
        f.write(gt.cpu().numpy().astype(np.int32).tobytes())
    steps = 4
    r = subprocess.run([exe, "refine", str(scene), str(steps)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    lines = r.stdout.strip().splitlines()
    cpp_losses = [float(ln.split()[1]) for ln in lines if ln.startswith("loss")]
    cpp_bounds = [float(t) for t in next(ln for ln in lines if ln.startswith("bounds")).split()[1:]]
    tok = next(ln for ln in lines if ln.startswith("refine")).split()
    cpp_stats = {tok[i]: int(tok[i + 1]) for i in range(1, len(tok), 2)}
    cfg = T.TrainConfig(total_train_iters=1000, background_noise_strength=0.0, seed=7, max_splats=2 * n)
    splats = T.Splats(*(torch.from_numpy(x.copy()).to(d) for x in (tr, sh0, op)))
    bounds = T.bounds_from_pos_device(ctx, T.BOUND_PERCENTILE, splats.transforms)
    np.testing.assert_allclose(cpp_bounds, [*bounds.center, *bounds.extent], rtol=1e-6)
    trainer = T.SplatTrainer(cfg, ctx, bounds)
    batch = T.SceneBatch(img_packed=gt, camera=cam)
    py_losses = [float(trainer.step_views([batch], splats, distributed=False).loss.item()) for _ in range(steps)]
    rs = trainer.refine(steps, splats)
    py_losses.append(float(trainer.step_views([batch], splats, distributed=False).loss.item()))
    assert len(cpp_losses) == steps + 1 and all(math.isfinite(x) for x in cpp_losses)
    np.testing.assert_allclose(cpp_losses, py_losses, rtol=2e-3)
    assert cpp_stats["pruned"] >= 500 and rs.num_pruned >= 500
    for name, want in (("added", rs.num_added), ("pruned", rs.num_pruned), ("total", rs.total_splats)):
        assert abs(cpp_stats[name] - want) <= max(3, want // 500), (name, cpp_stats, rs)
    ctx.close()
