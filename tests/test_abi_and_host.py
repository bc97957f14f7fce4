"""CPU-side checks of the product: the C-ABI library loads and exports every symbol declared in
include/brush_b200.h (no compute calls without a GPU), and the host-side mirrors of the reference's
camera / bounds logic behave as the reference's unit tests require."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "brush_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from brush_b200 import _lib, build
    lib_path = build.build()
    h = ctypes.CDLL(lib_path)
    declared = _declared_functions()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(h, name), f"{name} declared in brush_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(declared)
    lib = _lib.load()
    assert lib.bg_abi_version() == _lib.ABI_VERSION


def test_integration_doc_binds_every_declared_symbol():
    """INTEGRATION.md is the reference-side binding a maintainer would add: every entry point of the header appears in its
    `extern "C"` block with the same number of parameters as the C declaration."""
    hdr = open(os.path.join(ROOT, "include", "brush_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    doc_nc = re.sub(r"/\*.*?\*/", "", re.sub(r"//[^\n]*", "", doc), flags=re.S)

    def n_params(text, name, opener):
        m = re.search(opener % re.escape(name), text)
        assert m, f"{name} not found"
        depth, i, args = 1, m.end(), ""
        while depth:
            c = text[i]
            depth += c == "("
            depth -= c == ")"
            if depth:
                args += c
            i += 1
        args = args.strip()
        return 0 if args in ("", "void") else args.count(",") + 1

    for name in _declared_functions():
        assert f"fn {name}(" in doc_nc, f"{name} is declared in brush_b200.h but INTEGRATION.md does not bind it"
        c_n = n_params(hdr, name, r"\b%s\s*\(")
        r_n = n_params(doc_nc, name, r"fn %s\(")
        assert c_n == r_n, f"{name}: {c_n} parameters in the header, {r_n} in INTEGRATION.md"


def test_null_and_invalid_arguments_return_status_codes():
    """apps/brush-c/src/lib.rs:119-121 style: null -> error code, never a crash (no GPU needed)."""
    from brush_b200 import _lib
    lib = _lib.load()
    assert lib.bg_ctx_create(0, 0, 64, 64, 0, ctypes.byref(ctypes.c_void_p())) == _lib.BG_ERR_INVALID
    assert lib.bg_ctx_create(0, 10, 64, 64, 0, None) == _lib.BG_ERR_NULL
    assert lib.bg_ctx_destroy(None) == _lib.BG_ERR_NULL
    assert lib.bg_render_forward(None, None, None, 1, 1, 0, 1, None, None, None, 0, None, 0, None, None, None, None) == _lib.BG_ERR_NULL
    assert lib.bg_adam_step(None, None, None, None, None, None, 1, 1, None, 0.0, 0.9, 0.999, 1e-15, 1, 0) == _lib.BG_ERR_NULL
    assert lib.bg_ctx_arena_bytes(None) == 0


def test_struct_layouts_match_header():
    from brush_b200 import _lib
    assert ctypes.sizeof(_lib.BgCamera) == 4 * (12 + 4 + 3 + 4 + 1 + 1 + 8)
    assert _lib.BgRenderState.n.offset == 9 * 8


def test_no_product_module_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "brush_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.replace("no CPU fallback", ""), f"{f} mentions the oracle"


def test_fov_focal_round_trip():
    """brush-render/src/tests/mod.rs:710-789."""
    from brush_b200.camera import focal_to_fov, fov_to_focal
    for px in (64, 123, 1920):
        for fov in (0.2, 0.6, math.pi / 2, 2.5):
            f = fov_to_focal(fov, px)
            assert abs(focal_to_fov(f, px) - fov) < 1e-12
    assert abs(fov_to_focal(math.pi / 2, 256) - 128.0) < 1e-9


def test_world_to_local_inverts_pose():
    from brush_b200.camera import Camera
    rng = np.random.default_rng(0)
    for _ in range(10):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        p = rng.normal(size=3) * 3
        cam = Camera(position=tuple(p), rotation=tuple(q))
        vm = cam.world_to_local().reshape(4, 3)  # rows = columns c0,c1,c2,t
        R = vm[:3].T
        t = vm[3]
        x, y, z, w = q
        Rl2w = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        np.testing.assert_allclose(R, Rl2w.T, atol=2e-6)
        np.testing.assert_allclose(R @ p + t, 0, atol=1e-5)  # the camera centre maps to the origin


def test_uniforms_match_reference_formulas():
    from brush_b200.camera import Camera, build_uniforms
    cam = Camera(position=(0.123, 0.456, -8.0), rotation=(0, 0, 0, 1), fov_x=math.pi / 2, fov_y=1.2, center_uv=(0.5, 0.4))
    u = build_uniforms(cam, 123, 82)
    assert abs(u.fx - 61.5) < 1e-4 and abs(u.cx - 61.5) < 1e-5 and abs(u.cy - 32.8) < 1e-5
    assert abs(u.lim_pos_x - (1.15 * 123 - u.cx) / u.fx) < 1e-6 and abs(u.lim_neg_y - (-0.15 * 82 - u.cy) / u.fy) < 1e-6
    np.testing.assert_allclose(u.viewmat, [1, 0, 0, 0, 1, 0, 0, 0, 1, -0.123, -0.456, 8.0], atol=1e-7)


def test_bounds_and_median_size():
    """splat_init.rs:130-160 / bounding_box.rs:23-29."""
    from brush_b200.train import bounds_from_pos
    pts = np.stack([np.linspace(-10, 10, 1001), np.linspace(0, 2, 1001), np.linspace(-1, 5, 1001)], 1).astype(np.float32)
    pts[0] = np.nan
    b = bounds_from_pos(0.8, pts)
    assert abs(b.extent[0] - 8.0) < 0.05 and abs(b.extent[1] - 0.8) < 0.01 and abs(b.extent[2] - 2.4) < 0.02
    assert abs(b.median_size() - 2 * 2.4) < 0.05


def test_synthetic_scene_is_deterministic():
    from scenes import synthetic_scene
    a = synthetic_scene(1000, 64, 64)
    b = synthetic_scene(1000, 64, 64)
    for x, y in zip(a[1:], b[1:]):
        np.testing.assert_array_equal(x, y)


def test_ctypes_structs_match_the_c_header(tmp_path):
    """sizeof / offsetof of every ABI struct as the C compiler lays it out vs the ctypes mirrors."""
    import subprocess
    from brush_b200 import _lib
    fields = {"BgCamera": ["viewmat", "fx", "cam_pos", "half_max_render_fov", "camera_model", "model_params"],
              "BgRenderState": ["projected", "counters_host", "n", "tiles_y", "mip"],
              "BgTrainStepArgs": ["cam", "w", "background", "transforms", "gt_packed", "l1_weight", "composite_bg", "channels",
                                  "lr_mean", "median_scale", "seed", "step", "workspace", "loss_out", "state_out"]}
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "brush_b200.h"', 'int main(void){']
    for st, fs in fields.items():
        prog.append(f'printf("{st} %zu", sizeof({st}));')
        for f in fs:
            prog.append(f'printf(" %zu", offsetof({st}, {f}));')
        prog.append('printf("\\n");')
    prog.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    for line in out:
        tok = line.split()
        cls = getattr(_lib, tok[0])
        assert ctypes.sizeof(cls) == int(tok[1]), tok[0]
        for f, off in zip(fields[tok[0]], tok[2:]):
            name = "pass_" if f == "pass" else f
            assert getattr(cls, name).offset == int(off), (tok[0], f)


def test_median_size_orders_like_total_cmp():
    """brush-render/src/bounding_box.rs:36-58: a NaN extent must not break the ordering; normal case = 4."""
    from brush_b200.train import BoundingBox
    nan = float("nan")
    z = np.zeros(3, np.float32)
    assert BoundingBox(z, np.array([nan, 2.0, 3.0], np.float32)).median_size() == 6.0      # [2, 3, NaN] -> 3 * 2
    assert math.isnan(BoundingBox(z, np.array([nan, nan, nan], np.float32)).median_size())
    assert abs(BoundingBox(z, np.array([1.0, 2.0, 3.0], np.float32)).median_size() - 4.0) < 1e-6   # from_min_max(-1, (1, 3, 5))
    assert BoundingBox(z, np.array([3.0, 1.0, 2.0], np.float32)).median_size() == 4.0
    assert BoundingBox(z, np.array([-1.0, 5.0, 0.5], np.float32)).median_size() == 1.0
