"""ctypes wrapper around oracle/liborc.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this
module.  Nothing under brush_b200/ imports it; the product fails loudly when
its CUDA library is missing instead of falling back to the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from types import SimpleNamespace

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")

PASS_FORWARD, PASS_BACKWARD, PASS_BACKWARD_SMOOTH = 0, 1, 2


class OrcCamera(C.Structure):
    _fields_ = [
        ("viewmat", C.c_float * 12),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("cam_pos", C.c_float * 3),
        ("lim_pos_x", C.c_float), ("lim_pos_y", C.c_float), ("lim_neg_x", C.c_float), ("lim_neg_y", C.c_float),
        ("half_max_render_fov", C.c_float),
        ("camera_model", C.c_uint32),
        ("model_params", C.c_float * 8),
    ]


_FP = C.POINTER(C.c_float)
_UP = C.POINTER(C.c_uint32)


class OrcRender(C.Structure):
    _fields_ = [
        ("n", C.c_uint32), ("k", C.c_uint32), ("w", C.c_uint32), ("h", C.c_uint32),
        ("tiles_x", C.c_uint32), ("tiles_y", C.c_uint32),
        ("num_visible", C.c_uint32), ("num_intersections", C.c_uint32),
        ("pass_", C.c_int), ("mip", C.c_int),
        ("out_img", _FP), ("out_packed", _UP), ("visible", _FP), ("max_radius", _FP),
        ("intersect_counts", _UP), ("depths_sorted", _FP), ("gid_from_cgid", _UP), ("cum_tiles_hit", _UP),
        ("projected", _FP), ("tile_id_from_isect", _UP), ("cgid_from_isect", _UP),
        ("tile_offsets", _UP), ("tile_offsets_untrimmed", _UP),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/liborc.so with oracle/Makefile (gcc).  Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_render_forward.restype = C.POINTER(OrcRender)
        L.orc_render_forward.argtypes = [C.POINTER(OrcCamera), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_render_free.argtypes = [C.POINTER(OrcRender)]
        L.orc_render_free.restype = None
        L.orc_rasterize_backward.argtypes = [C.POINTER(OrcRender), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_rasterize_backward.restype = None
        L.orc_project_backward.argtypes = [C.POINTER(OrcCamera), C.POINTER(OrcRender)] + [C.c_void_p] * 8
        L.orc_project_backward.restype = None
        L.orc_radix_argsort_u32.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_radix_argsort_u32.restype = None
        L.orc_inclusive_scan_u32.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_inclusive_scan_u32.restype = None
        L.orc_image_loss_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float,
                                             C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_image_loss_forward.restype = None
        L.orc_image_loss_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                              C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_image_loss_backward.restype = None
        L.orc_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p,
                                    C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]
        L.orc_adam_step.restype = None
        L.orc_compute_min_scale.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_float, C.c_void_p]
        L.orc_compute_min_scale.restype = None
        L.orc_fold_min_scale_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_fold_min_scale_fwd.restype = None
        L.orc_fold_min_scale_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_fold_min_scale_bwd.restype = None
        L.orc_expf_det.argtypes = [C.c_float]
        L.orc_expf_det.restype = C.c_float
        L.orc_logf_det.argtypes = [C.c_float]
        L.orc_logf_det.restype = C.c_float
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_set_num_threads.restype = None
        _lib = L
    return _lib


def camera_struct(u) -> OrcCamera:
    """u: brush_b200.camera.ProjectUniforms (host mirror of the reference uniforms)."""
    c = OrcCamera()
    for i in range(12):
        c.viewmat[i] = float(u.viewmat[i])
    c.fx, c.fy, c.cx, c.cy = u.fx, u.fy, u.cx, u.cy
    for i in range(3):
        c.cam_pos[i] = float(u.cam_pos[i])
    c.lim_pos_x, c.lim_pos_y, c.lim_neg_x, c.lim_neg_y = u.lim_pos_x, u.lim_pos_y, u.lim_neg_x, u.lim_neg_y
    c.half_max_render_fov = u.half_max_render_fov
    c.camera_model = u.camera_model
    for i, v in enumerate(getattr(u, "model_params", ())):
        c.model_params[i] = float(v)
    return c


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _arr(p, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True).reshape(shape)


class RenderResult(SimpleNamespace):
    """Owns the C-side OrcRender (needed by the backward calls) plus numpy copies of every output."""

    def close(self):
        if getattr(self, "_handle", None) is not None:
            lib().orc_render_free(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render_forward(uniforms, img_w, img_h, transforms, sh, raw_opac, mip=False, bg=(0.0, 0.0, 0.0),
                   rpass=PASS_BACKWARD) -> RenderResult:
    transforms, sh, raw_opac = _f32(transforms), _f32(sh), _f32(raw_opac)
    n, k = transforms.shape[0], sh.shape[1]
    assert transforms.shape == (n, 10) and sh.shape == (n, k, 3) and raw_opac.shape == (n,)
    cam = camera_struct(uniforms)
    bg_a = _f32(bg)
    h = lib().orc_render_forward(C.byref(cam), img_w, img_h, n, k, _ptr(transforms), _ptr(sh), _ptr(raw_opac),
                                 int(bool(mip)), _ptr(bg_a), int(rpass))
    if not h:
        raise ValueError("orc_render_forward rejected its arguments")
    r = h.contents
    V, I = r.num_visible, r.num_intersections
    T = r.tiles_x * r.tiles_y
    out = RenderResult(
        _handle=h, _cam=cam, _inputs=(transforms, sh, raw_opac), _bg=bg_a,
        n=n, k=k, w=img_w, h=img_h, tiles_x=r.tiles_x, tiles_y=r.tiles_y,
        num_visible=V, num_intersections=I, mip=bool(mip), rpass=rpass,
        out_img=_arr(r.out_img, (img_h, img_w, 4), np.float32) if rpass != PASS_FORWARD else None,
        out_packed=_arr(r.out_packed, (img_h, img_w), np.uint32) if rpass == PASS_FORWARD else None,
        visible=_arr(r.visible, (n,), np.float32) if rpass != PASS_FORWARD else None,
        max_radius=_arr(r.max_radius, (n,), np.float32),
        intersect_counts=_arr(r.intersect_counts, (n,), np.uint32),
        depths_sorted=_arr(r.depths_sorted, (V,), np.float32),
        gid_from_cgid=_arr(r.gid_from_cgid, (V,), np.uint32),
        cum_tiles_hit=_arr(r.cum_tiles_hit, (V,), np.uint32),
        projected=_arr(r.projected, (V, 9), np.float32),
        tile_id_from_isect=_arr(r.tile_id_from_isect, (I,), np.uint32),
        cgid_from_isect=_arr(r.cgid_from_isect, (I,), np.uint32),
        tile_offsets=_arr(r.tile_offsets, (r.tiles_y, r.tiles_x, 2), np.uint32),
        tile_offsets_untrimmed=_arr(r.tile_offsets_untrimmed, (r.tiles_y, r.tiles_x, 2), np.uint32),
    )
    assert T == r.tiles_x * r.tiles_y
    return out


def rasterize_backward(res: RenderResult, v_output, smooth=None):
    v_output = _f32(v_output)
    assert v_output.shape == (res.h, res.w, 4)
    if smooth is None:
        smooth = res.rpass == PASS_BACKWARD_SMOOTH
    v_combined = np.zeros((max(res.num_visible, 1), 10), np.float32)
    lib().orc_rasterize_backward(res._handle, _ptr(res._bg), _ptr(v_output), int(bool(smooth)), _ptr(v_combined))
    return v_combined[: res.num_visible]


def project_backward(res: RenderResult, v_combined):
    transforms, sh, raw_opac = res._inputs
    v_combined = _f32(v_combined)
    n, k = res.n, res.k
    v_t = np.zeros((n, 10), np.float32)
    v_sh = np.zeros((n, k, 3), np.float32)
    v_o = np.zeros((n,), np.float32)
    v_r = np.zeros((n,), np.float32)
    vc = v_combined if v_combined.size else np.zeros((1, 10), np.float32)
    lib().orc_project_backward(C.byref(res._cam), res._handle, _ptr(transforms), _ptr(sh), _ptr(raw_opac), _ptr(vc),
                               _ptr(v_t), _ptr(v_sh), _ptr(v_o), _ptr(v_r))
    return v_t, v_sh, v_o, v_r


def render_backward(res: RenderResult, v_output, smooth=None):
    """rasterize_bwd followed by project_bwd (bwd/burn_glue.rs:121-182)."""
    vc = rasterize_backward(res, v_output, smooth)
    return (vc,) + project_backward(res, vc)


def radix_argsort(keys, vals, bits):
    keys, vals = _u32(keys), _u32(vals)
    ko, vo = np.empty_like(keys), np.empty_like(vals)
    lib().orc_radix_argsort_u32(_ptr(keys), _ptr(vals), keys.shape[0], bits, _ptr(ko), _ptr(vo))
    return ko, vo


def inclusive_scan(x):
    x = _u32(x)
    o = np.empty_like(x)
    lib().orc_inclusive_scan_u32(_ptr(x), x.shape[0], _ptr(o))
    return o


def image_loss_forward(pred_chw, gt_packed, l1_w, ssim_w, bg=None, mask=False):
    pred_chw, gt_packed = _f32(pred_chw), _u32(gt_packed)
    c, h, w = pred_chw.shape
    out = np.zeros_like(pred_chw)
    bg_a = _f32(bg) if bg is not None else None
    lib().orc_image_loss_forward(_ptr(pred_chw), _ptr(gt_packed), c, h, w, l1_w, ssim_w, _ptr(bg_a), int(mask), _ptr(out))
    return out


def image_loss_backward(pred_chw, gt_packed, dl_dmap, l1_w, ssim_w, bg=None, mask=False):
    pred_chw, gt_packed, dl_dmap = _f32(pred_chw), _u32(gt_packed), _f32(dl_dmap)
    c, h, w = pred_chw.shape
    out = np.zeros_like(pred_chw)
    bg_a = _f32(bg) if bg is not None else None
    lib().orc_image_loss_backward(_ptr(pred_chw), _ptr(gt_packed), _ptr(dl_dmap), c, h, w, l1_w, ssim_w, _ptr(bg_a),
                                  int(mask), _ptr(out))
    return out


def adam_step(p, g, m, v, lr, t, lr_scale_per_col=None, beta1=0.9, beta2=0.999, eps=1e-15, reduce_v=False):
    """In place on p, m, v (float32 C-contiguous, p/g/m: [rows, cols], v: [rows, cols] or [rows])."""
    rows = p.shape[0]
    cols = int(np.prod(p.shape[1:])) if p.ndim > 1 else 1
    for a in (p, g, m, v):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    s = _f32(lr_scale_per_col) if lr_scale_per_col is not None else None
    lib().orc_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), rows, cols, _ptr(s), lr, beta1, beta2, eps, t, int(reduce_v))


def compute_min_scale(transforms, view_cams, factor):
    """compute_min_scale (train.rs:102-125).  view_cams: [views,4] = x, y, z, focal_px."""
    t, c = _f32(transforms), _f32(view_cams).reshape(-1, 4)
    out = np.empty(t.shape[0], dtype=np.float32)
    lib().orc_compute_min_scale(_ptr(t), t.shape[0], _ptr(c), c.shape[0], factor, _ptr(out))
    return out


def fold_min_scale(transforms, raw_opac, f):
    """fold_min_scale (gaussian_splats.rs:86-111) -> (transforms', raw_opac')."""
    t, o, ff = _f32(transforms), _f32(raw_opac), _f32(f)
    to, oo = np.empty_like(t), np.empty_like(o)
    lib().orc_fold_min_scale_fwd(_ptr(t), _ptr(o), _ptr(ff), t.shape[0], _ptr(to), _ptr(oo))
    return to, oo


def fold_min_scale_backward(transforms, raw_opac, f, v_transforms_folded, v_raw_opac_folded):
    """Reverse-mode chain of fold_min_scale -> (v_transforms, v_raw_opac) w.r.t. the learned parameters."""
    t, o, ff = _f32(transforms), _f32(raw_opac), _f32(f)
    vt, vo = _f32(v_transforms_folded).copy(), _f32(v_raw_opac_folded).copy()
    lib().orc_fold_min_scale_bwd(_ptr(t), _ptr(o), _ptr(ff), t.shape[0], _ptr(vt), _ptr(vo))
    return vt, vo


def expf_det(x: float) -> float:
    return lib().orc_expf_det(x)


def logf_det(x: float) -> float:
    return lib().orc_logf_det(x)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))
