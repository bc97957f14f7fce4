"""Poison / robustness fuzz on the CUDA path (crates/brush-bench-test/tests/fuzz.rs:61-86, 269-553) and the scale
stress the reference pins (brush-render/src/tests/mod.rs:74-120, 394-450; brush-sort/src/lib.rs:290-339).

Every forward case is rendered by the kernels AND by the oracle on the same poisoned input: the positive-phrased cull
guards (project_forward.rs:43-111) must take the same side for every NaN / Inf / denormal / huge value, i.e.
num_visible, num_intersections and the depth-ordered visible set are bit-equal, and the image agrees and is finite."""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

F32_MAX = float(np.finfo(np.float32).max)
F32_TINY = float(np.finfo(np.float32).tiny)
POISON = [float("nan"), -float("nan"), float("inf"), -float("inf"), 0.0, -0.0, F32_TINY, F32_TINY / 2.0, 1e-40, float(np.finfo(np.float32).eps),
          1e38, -1e38, F32_MAX, -F32_MAX, 1e20, -1e20, 1.0, -1.0, 0.01, 1e10, 1.0 / 255.0, 16.0]


class Sm64:
    """The SplitMix64 stream of fuzz.rs:30-58."""

    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def u64(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def f01(self):
        return float(np.float32(self.u64() / float(0xFFFFFFFFFFFFFFFF)))

    def uniform(self, lo, hi):
        return float(np.float32(lo + self.f01() * (hi - lo)))

    def choice(self, items):
        return items[self.u64() % len(items)]

    def usize_in(self, lo, hi):
        return lo + self.u64() % (hi - lo)


@pytest.fixture(scope="module")
def rt():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import brush_b200.render as R
    from brush_b200.camera import Camera, build_uniforms
    from oracle import oracle as orc
    from types import SimpleNamespace
    ctx = R.RenderContext(max_splats=1 << 17, max_w=512, max_h=512, max_intersections=140_000_000 // 8)
    yield SimpleNamespace(R=R, orc=orc, build_uniforms=build_uniforms, Camera=Camera, ctx=ctx)
    ctx.close()


def std_cam(rt):
    return rt.Camera(position=(0.0, 0.0, -3.0), rotation=(0.0, 0.0, 0.0, 1.0), fov_x=0.5, fov_y=0.5)


def rand_cam(rt, rng):
    pos = (rng.uniform(-10, 10), rng.uniform(-10, 10), rng.uniform(-30.0, -0.1))
    ax = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1)], np.float64)
    angle = rng.uniform(0.0, 2 * math.pi)
    nrm = np.linalg.norm(ax)
    if nrm == 0:
        q = (0.0, 0.0, 0.0, 1.0)
    else:
        ax = ax / nrm
        q = tuple(float(v) for v in ax * math.sin(angle / 2)) + (math.cos(angle / 2),)
    fov = rng.uniform(0.3, 1.2)
    return rt.Camera(position=pos, rotation=q, fov_x=fov, fov_y=fov)


IMG_SIZES = [(1, 1), (1, 17), (17, 1), (15, 15), (16, 16), (17, 17), (33, 47), (64, 64), (97, 129), (128, 128), (257, 257)]


def arrays(means, rots, ls, dc, opac):
    n = len(opac)
    tr = np.concatenate([np.asarray(means, np.float32).reshape(n, 3), np.asarray(rots, np.float32).reshape(n, 4),
                         np.asarray(ls, np.float32).reshape(n, 3)], 1)
    return np.ascontiguousarray(tr), np.asarray(dc, np.float32).reshape(n, 1, 3).copy(), np.asarray(opac, np.float32).copy()


def render_both(rt, cam, w, h, tr, sh, op, mip=False, tag=""):
    d = rt.ctx.device
    out = rt.R.render_splats(rt.ctx, cam, (w, h), torch.from_numpy(tr).to(d), torch.from_numpy(sh).to(d), torch.from_numpy(op).to(d), mip=mip)
    o = rt.orc.render_forward(rt.build_uniforms(cam, w, h), w, h, tr, sh, op, mip=mip)
    n = tr.shape[0]
    # assert_basic_counts + validate (render_aux.rs:30-45)
    assert out.num_visible <= n and out.num_intersections <= out.num_visible * (-(-w // 16)) * (-(-h // 16)), tag
    assert (out.num_visible, out.num_intersections) == (o.num_visible, o.num_intersections), tag
    assert out.intersection_overflow == 0, tag
    np.testing.assert_array_equal(out.global_from_compact_gid().cpu().numpy().view(np.uint32), o.gid_from_cgid, err_msg=tag)
    img = out.out_img.cpu().numpy()
    assert np.isfinite(img).all(), tag
    err = np.abs(img - o.out_img)
    bad = err > 1e-5 + 1e-4 * np.abs(o.out_img)
    assert bad.mean() <= 2e-4 and (not bad.any() or err[bad].max() <= 1.5 / 255 * 1.5), (tag, int(bad.sum()), float(err.max()))
    vis = out.visible.cpu().numpy()
    assert set(np.unique(vis)) <= {0.0, 1.0} and np.isfinite(out.max_radius.cpu().numpy()).all(), tag
    return out, o


def test_single_bad_slot_combinations(rt):
    """fuzz.rs:268-303: every (slot, poison) combination, one bad slot per scene."""
    n, w = 120, 48
    for slot in range(14):
        for poison in POISON:
            means, rots, ls = [0.0, 0.0, 3.0] * n, [1.0, 0.0, 0.0, 0.0] * n, [-1.0] * (3 * n)
            dc, opac = [0.5] * (3 * n), [2.0] * n
            if slot < 3:
                means[slot] = poison
            elif slot < 7:
                rots[slot - 3] = poison
            elif slot < 10:
                ls[slot - 7] = poison
            elif slot < 13:
                dc[slot - 10] = poison
            else:
                opac[0] = poison
            render_both(rt, std_cam(rt), w, w, *arrays(means, rots, ls, dc, opac), tag=f"slot={slot} poison={poison!r}")


def poisoned_scene(seed, n, rate):
    rng = Sm64(seed)

    def pick(lo, hi):
        fb = rng.uniform(lo, hi)
        return rng.choice(POISON) if rng.f01() < rate else fb
    means, rots, ls, dc, opac = [], [], [], [], []
    for _ in range(n):
        means += [pick(-3, 3) for _ in range(3)]
        rots += [pick(-1, 1) for _ in range(4)]
        ls += [pick(-4, 2) for _ in range(3)]
        dc += [pick(0, 1) for _ in range(3)]
        opac.append(pick(-2, 2))
    return means, rots, ls, dc, opac


def test_random_poisoned_scenes(rt):
    """fuzz.rs:305-329: scene, camera, image size, poison rate and render mode all random."""
    for seed in range(100):
        rng = Sm64((seed * 0xA5A5CAFE) & 0xFFFFFFFFFFFFFFFF)
        n = rng.usize_in(1, 256)
        rate = rng.uniform(0.0, 0.95)
        w, h = IMG_SIZES[rng.u64() % len(IMG_SIZES)]
        cam = rand_cam(rt, rng)
        mip = rng.f01() < 0.3
        render_both(rt, cam, w, h, *arrays(*poisoned_scene(seed, n, rate)), mip=mip, tag=f"seed={seed} n={n} rate={rate:.2f} img={w}x{h} mip={mip}")


def test_bad_geometry_is_fully_culled(rt):
    """fuzz.rs:331-447."""
    n, w = 16, 64
    rng = Sm64(0xDEADBEEF)
    ident, centre = [1.0, 0.0, 0.0, 0.0] * n, [0.0, 0.0, 3.0] * n
    nan = float("nan")

    def one_bad(count, lo, hi):
        vals = []
        for _ in range(n):
            bad = rng.usize_in(0, count)
            vals += [nan if s == bad else rng.uniform(lo, hi) for s in range(count)]
        return vals
    cases = {
        "nan_positions": ([nan] * (3 * n), ident, [0.0] * (3 * n), [0.0] * n),
        "inf_positions": ([float("inf") if i % 2 == 0 else -float("inf") for i in range(3 * n)], ident, [0.0] * (3 * n), [0.0] * n),
        "nan_quats": (centre, one_bad(4, -1, 1), [0.0] * (3 * n), [0.0] * n),
        "zero_quats": (centre, [0.0] * (4 * n), [0.0] * (3 * n), [0.0] * n),
        "nan_scales": (centre, ident, one_bad(3, -4, 4), [0.0] * n),
        "inf_scales": (centre, ident, [120.0] * (3 * n), [0.0] * n),
        "nan_opac": (centre, ident, [0.0] * (3 * n), [nan] * n),
    }
    for tag, (means, rots, ls, opac) in cases.items():
        out, _ = render_both(rt, std_cam(rt), w, w, *arrays(means, rots, ls, [0.5] * (3 * n), opac), tag=tag)
        assert out.num_visible == 0 and out.num_intersections == 0, tag
        assert float(out.out_img.abs().max()) == 0.0, tag


def test_valid_but_extreme_stays_visible(rt):
    """fuzz.rs:451-491: huge log-scales and huge finite colours are legitimate training states."""
    n, w = 4, 64
    for ls_val in (-30.0, -10.0, 0.0, 10.0, 20.0, 30.0, 40.0):
        out, _ = render_both(rt, std_cam(rt), w, w, *arrays([0.0, 0.0, 3.0] * n, [1.0, 0, 0, 0] * n, [ls_val] * (3 * n), [0.5] * (3 * n), [2.0] * n),
                             tag=f"log_scale={ls_val}")
        assert out.num_visible == n, f"log_scale={ls_val} over-culled"
    for mag in (1e10, 1e25, F32_MAX / 2.0):
        out, _ = render_both(rt, std_cam(rt), w, w, *arrays([0.0, 0.0, 3.0] * n, [1.0, 0, 0, 0] * n, [-1.0] * (3 * n), [mag, -mag, mag] * n, [3.0] * n),
                             tag=f"colour {mag}")
        assert out.num_visible == n, f"colour mag={mag} over-culled"


def _backward_finite(rt, cam, w, h, tr, sh, op, mip, tag):
    d = rt.ctx.device
    ttr, tsh, top = (torch.from_numpy(x).to(d) for x in (tr, sh, op))
    out = rt.R.render_splats(rt.ctx, cam, (w, h), ttr, tsh, top, mip=mip)
    v_out = torch.full((h, w, 4), 1.0 / (h * w * 4), device=d)      # d(mean of the image)
    vc = rt.R.rasterize_bwd(out, v_out)
    grads = rt.R.project_bwd(out, ttr, tsh, top, vc)
    for g, name in zip(grads, ("v_transforms", "v_sh", "v_raw_opac", "v_refine")):
        assert torch.isfinite(g).all(), f"{tag}: {name} has non-finite entries"
    o = rt.orc.render_forward(rt.build_uniforms(cam, w, h), w, h, tr, sh, op, mip=mip)
    _, ovt, ovsh, ovo, _ = rt.orc.render_backward(o, v_out.cpu().numpy())
    for g, r, name in ((grads[0], ovt, "v_transforms"), (grads[1], ovsh, "v_sh"), (grads[2], ovo, "v_raw_opac")):
        g = g.cpu().numpy().astype(np.float64)
        scale = max(np.abs(r).max(), 1e-30)
        assert np.abs(g - r).max() <= 2e-3 * scale + 1e-12, (tag, name, float(np.abs(g - r).max()), float(scale))


def test_bwd_random_scenes_gradients_are_finite(rt):
    """fuzz.rs:493-520 (and equal to the oracle's adjoint)."""
    for seed in range(60):
        rng = Sm64(0xBDBDBDBD ^ ((seed * 0xA5A5CAFE) & 0xFFFFFFFFFFFFFFFF))
        n = rng.usize_in(4, 256)
        w, h = rng.usize_in(16, 128), rng.usize_in(16, 128)
        cam = rand_cam(rt, rng)
        mip = rng.f01() < 0.3
        r2 = Sm64(seed)   # finite_scene (fuzz.rs:125-148)
        means, rots, ls, dc, opac = [], [], [], [], []
        for _ in range(n):
            means += [r2.uniform(-3, 3) for _ in range(3)]
            rots += [r2.uniform(-1, 1) for _ in range(4)]
            ls += [r2.uniform(-4, 1) for _ in range(3)]
            dc += [r2.uniform(0, 1) for _ in range(3)]
            opac.append(r2.uniform(-2, 2))
        _backward_finite(rt, cam, w, h, *arrays(means, rots, ls, dc, opac), mip, f"seed={seed}")


def test_bwd_extreme_inputs_stay_finite(rt):
    """fuzz.rs:522-553."""
    n, w = 8, 64
    for ls_val in (-20.0, -5.0, 0.0, 5.0, 15.0, 30.0, 40.0):
        for mag in (0.1, 10.0, 1e6, F32_MAX / 2.0):
            tr, sh, op = arrays([0.0, 0.0, 3.0] * n, [1.0, 0, 0, 0] * n, [ls_val] * (3 * n), [mag, -mag, mag] * n, [2.0] * n)
            d = rt.ctx.device
            ttr, tsh, top = (torch.from_numpy(x).to(d) for x in (tr, sh, op))
            out = rt.R.render_splats(rt.ctx, std_cam(rt), (w, w), ttr, tsh, top)
            vc = rt.R.rasterize_bwd(out, torch.full((w, w, 4), 1.0 / (w * w * 4), device=d))
            for g in rt.R.project_bwd(out, ttr, tsh, top, vc):
                assert torch.isfinite(g).all(), (ls_val, mag)


# ------------------------------------------------------------------------------------------ scale stress
def rng_scene(n, mean_range, ls_range, op_range, seed):
    """brush-render/src/tests/mod.rs:168-222 (vectorised SplitMix64, same draw order per splat)."""
    from scenes import splitmix64
    r = splitmix64(seed, n * 14).astype(np.float32).reshape(n, 14)   # (z / 2^64) vs the reference's z / u64::MAX: same to 1 ulp
    u = lambda c, lo, hi: lo + r[:, c] * (hi - lo)
    means = np.stack([u(0, -mean_range, mean_range), u(1, -mean_range, mean_range), u(2, -mean_range, mean_range)], 1)
    quats = np.stack([u(3, -1, 1), u(4, -1, 1), u(5, -1, 1), u(6, -1, 1)], 1)
    ls = np.stack([u(7, *ls_range), u(8, *ls_range), u(9, *ls_range)], 1)
    sh = np.stack([u(10, 0, 1), u(11, 0, 1), u(12, 0, 1)], 1).reshape(n, 1, 3)
    op = u(13, *op_range)
    return np.ascontiguousarray(np.concatenate([means, quats, ls], 1).astype(np.float32)), sh.astype(np.float32).copy(), op.astype(np.float32).copy()


def test_mega_stress_fullscreen_splats_no_dropped_tile(rt):
    """tests/mod.rs:394-450: 120k splats that each cover the whole 512x512 image (32 x 32 tiles x 120k = 123M
    intersections).  Deterministic here (index-order compaction), every tile receives contributions, and the tile
    lists hold exactly visible x tiles entries."""
    R = rt.R
    cam = rt.Camera(position=(0.0, 0.0, -5.0), rotation=(0.0, 0.0, 0.0, 1.0), fov_x=0.5, fov_y=0.5)
    w = 512
    tr, sh, op = rng_scene(120_000, 0.1, (3.5, 4.0), (-3.0, -1.5), 0x5EED)
    ctx = R.RenderContext(max_splats=120_000, max_w=w, max_h=w, max_intersections=126_000_000)
    d = ctx.device
    p = [torch.from_numpy(x).to(d) for x in (tr, sh, op)]
    a = R.render_splats(ctx, cam, (w, w), *p)
    img_a = a.out_img.clone()
    assert a.intersection_overflow == 0
    assert a.num_intersections == a.num_visible * 32 * 32 and a.num_visible > 100_000
    ftr, fsh, fop = rng_scene(100, 0.5, (-1.0, 0.5), (0.0, 1.0), 0xFACE)
    R.render_splats(ctx, cam, (w, w), *(torch.from_numpy(x).to(d) for x in (ftr, fsh, fop)))   # unrelated render in between
    b = R.render_splats(ctx, cam, (w, w), *p)
    assert torch.equal(b.out_img, img_a)
    tiles = img_a[..., 3].reshape(32, 16, 32, 16).sum(dim=(1, 3))
    assert torch.isfinite(img_a).all() and float(tiles.min()) > 1e-3
    # an arena that is too small reports the overflow instead of writing out of bounds
    small = R.RenderContext(max_splats=120_000, max_w=w, max_h=w, max_intersections=50_000_000)
    c = R.render_splats(small, cam, (w, w), *p)
    assert c.intersection_overflow != 0 and torch.isfinite(c.out_img).all()
    small.close()
    ctx.close()


def test_renders_thirty_million_splats(rt):
    """tests/mod.rs:74-120: far more Gaussians than one 1-D dispatch of the reference could address."""
    R = rt.R
    n = 30_000_000
    cam = rt.Camera(position=(0.0, 0.0, -5.0), rotation=(0.0, 0.0, 0.0, 1.0), fov_x=0.5, fov_y=0.5)
    ctx = R.RenderContext(max_splats=n, max_w=64, max_h=64, max_intersections=64_000_000)
    d = ctx.device
    g = torch.Generator(device=d)
    g.manual_seed(1)
    u = lambda shape, lo, hi: torch.rand(shape, device=d, generator=g) * (hi - lo) + lo
    tr = torch.cat([u((n, 3), -2, 2), u((n, 4), -1, 1), u((n, 3), -4, -2)], 1).contiguous()
    out = R.render_splats(ctx, cam, (64, 64), tr, u((n, 1, 3), 0, 1), u((n,), -2, 2))
    assert out.num_visible > 0 and out.intersection_overflow == 0
    assert out.num_intersections >= out.num_visible
    assert torch.isfinite(out.out_img).all() and float(out.out_img[..., 3].max()) > 1e-3
    ctx.close()


def test_sorting_seventy_million_keys(rt):
    """brush-sort/src/lib.rs:290-339 (the >= 67M-key regression): a permutation of 0..70M sorted by key must come out
    as the identity with the inverse permutation as values; then a tie-heavy 12-bit key set checks stability."""
    import ctypes as C
    from brush_b200 import _lib
    R = rt.R
    n = 70_000_000
    ctx = R.RenderContext(max_splats=1024, max_w=64, max_h=64, max_intersections=n)
    d = ctx.device
    lib = _lib.load()
    g = torch.Generator(device=d)
    g.manual_seed(0xD15EA5E)
    perm = torch.randperm(n, device=d, generator=g).to(torch.int32)
    vals = torch.arange(n, device=d, dtype=torch.int32)
    ko, vo = torch.empty_like(perm), torch.empty_like(vals)
    _lib.check(lib.bg_radix_argsort_u32(ctx.handle, torch.cuda.current_stream().cuda_stream, perm.data_ptr(), vals.data_ptr(), n, None, 32,
                                        ko.data_ptr(), vo.data_ptr()), "bg_radix_argsort_u32")
    torch.cuda.synchronize()
    assert torch.equal(ko, vals)
    inv = torch.empty_like(vals)
    inv[perm.long()] = vals
    assert torch.equal(vo, inv)
    del inv
    keys = (torch.randint(0, 1 << 12, (n,), device=d, generator=g, dtype=torch.int32))
    _lib.check(lib.bg_radix_argsort_u32(ctx.handle, torch.cuda.current_stream().cuda_stream, keys.data_ptr(), vals.data_ptr(), n, None, 12,
                                        ko.data_ptr(), vo.data_ptr()), "bg_radix_argsort_u32")
    torch.cuda.synchronize()
    assert bool((ko[1:] >= ko[:-1]).all())
    assert torch.equal(keys[vo.long()], ko)
    same = ko[1:] == ko[:-1]
    assert bool((vo[1:][same] > vo[:-1][same]).all())            # stable: ties keep their input order
    ctx.close()
