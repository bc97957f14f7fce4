// bg_project.cuh -- per-Gaussian projection math shared by the forward projection kernels and
// the projection backward kernel.  Include only from translation units built with -fmad=false.
//
// Reference semantics:
//   kernels/helpers.rs:84-264   compute_bbox_extent, get_tile_bbox, calc_cov2d, compensate_cov2d,
//                               count_contributing_tiles, will_primitive_contribute
//   kernels/camera_model/pinhole.rs:24-56   project_pinhole, calculate_project_jacobian_pinhole
//   kernels/sh.rs:41-136        sh_coeffs_to_color
#pragma once
#include "bg_common.cuh"
#include "bg_math.cuh"
#include "bg_camera.cuh"

namespace bg {

__device__ __forceinline__ M3 view_rotation(const BgCamera &u) {
    M3 m;
    m.c0 = mk3(u.viewmat[0], u.viewmat[1], u.viewmat[2]);
    m.c1 = mk3(u.viewmat[3], u.viewmat[4], u.viewmat[5]);
    m.c2 = mk3(u.viewmat[6], u.viewmat[7], u.viewmat[8]);
    return m;
}
__device__ __forceinline__ V3 world_to_cam(V3 mean, const BgCamera &u) {
    return add(mul(view_rotation(u), mean), mk3(u.viewmat[9], u.viewmat[10], u.viewmat[11]));
}
__device__ __forceinline__ void project_pinhole(V3 p, const BgCamera &u, float &ox, float &oy) {
    float inv_z = 1.0f / p.z;
    ox = u.fx * p.x * inv_z + u.cx;
    oy = u.fy * p.y * inv_z + u.cy;
}
__device__ __forceinline__ M23 jacobian_pinhole(V3 p, const BgCamera &u) {
    float inv_z = 1.0f / p.z;
    float dx = u.fx * inv_z;
    float dy = u.fy * inv_z;
    float cxn = clampf(p.x * inv_z, u.lim_neg_x, u.lim_pos_x);
    float cyn = clampf(p.y * inv_z, u.lim_neg_y, u.lim_pos_y);
    M23 j;
    j.c0 = mk2(dx, 0.0f);
    j.c1 = mk2(0.0f, dy);
    j.c2 = mk2(-dx * cxn, -dy * cyn);
    return j;
}
// DISTORTED = false: pinhole, resolved at compile time (all BASELINE configs); true: the three distorted models,
// selected by the uniform u.camera_model (the reference specialises per model AND per coefficient set).
template <bool DISTORTED>
__device__ __forceinline__ M23 project_jacobian(V3 p, const BgCamera &u) {
    if (DISTORTED) return jacobian_distorted(p, u);
    return jacobian_pinhole(p, u);
}
template <bool DISTORTED>
__device__ __forceinline__ void project_mean(V3 p, const BgCamera &u, float &ox, float &oy) {
    if (DISTORTED) project_distorted(p, u, ox, oy);
    else project_pinhole(p, u, ox, oy);
}
template <bool DISTORTED>
__device__ __forceinline__ bool in_front(V3 mean_c, const BgCamera &u) {  // project_forward.rs:47-61
    if (DISTORTED) return in_front_distorted(mean_c, u);
    return !(mean_c.z < 0.01f);
}
template <bool DISTORTED>
__device__ __forceinline__ S2 calc_cov2d(V3 scl, Q4 quat, V3 mean_c, const BgCamera &u) {
    M3 ns = mul_diag(mul(view_rotation(u), quat_to_mat3(quat)), scl);
    M23 v = mul(project_jacobian<DISTORTED>(mean_c, u), ns);
    S2 raw = gram(v);
    const float lim = 1.0e18f;
    float ma = max_abs(raw);
    float sd = (ma > lim) ? lim / ma : 1.0f;
    return scale(raw, sd);
}
template <bool MIP>
__device__ __forceinline__ S2 compensate_cov2d(S2 c, float &filter_comp) {
    const float blur = MIP ? 0.1f : 0.3f;
    S2 b; b.c00 = c.c00 + blur; b.c01 = c.c01; b.c11 = c.c11 + blur;
    filter_comp = 1.0f;
    if (MIP) {
        float det_raw = fmaxf(det2_strict(c), 0.0f);
        float det_blur = det2_strict(b);
        filter_comp = sqrtf(det_raw / det_blur);
    }
    return b;
}
__device__ __forceinline__ void bbox_extent(S2 conic, float pt, float &ex, float &ey) {
    float det = conic.c00 * conic.c11 - conic.c01 * conic.c01;
    bool degenerate = det <= 0.0f;
    float inv_det = degenerate ? 0.0f : 1.0f / det;
    float x = sqrtf(2.0f * pt * conic.c11 * inv_det);
    float y = sqrtf(2.0f * pt * conic.c00 * inv_det);
    ex = degenerate ? -1.0f : x;
    ey = degenerate ? -1.0f : y;
}
struct TileBox { uint32_t min_x, min_y, max_x, max_y; };
__device__ __forceinline__ TileBox tile_bbox(float px, float py, float ex, float ey, uint32_t bw, uint32_t bh) {
    const float tw = (float)TILE_W;
    float cx = px / tw, cy = py / tw, dx = ex / tw, dy = ey / tw;
    float bwf = (float)bw, bhf = (float)bh;
    TileBox b;
    b.min_x = (uint32_t)clampf(cx - dx, 0.0f, bwf);
    b.min_y = (uint32_t)clampf(cy - dy, 0.0f, bhf);
    b.max_x = (uint32_t)clampf(cx + dx + 1.0f, 0.0f, bwf);
    b.max_y = (uint32_t)clampf(cy + dy + 1.0f, 0.0f, bhf);
    return b;
}
// StopThePop tile test.  One definition, used by both the counting and the emitting kernel, so
// the two walks cannot disagree (the reference pads with sentinel tiles because its two WGSL
// compilations can: map_gaussians.rs:43-79).
__device__ __forceinline__ bool tile_hit(uint32_t tx, uint32_t ty, float mx, float my, S2 conic, float pt) {
    float rmin_x = (float)(tx * TILE_W), rmin_y = (float)(ty * TILE_W);
    float rmax_x = rmin_x + (float)TILE_W, rmax_y = rmin_y + (float)TILE_W;
    bool x_left = mx < rmin_x, x_right = mx > rmax_x;
    bool in_x = !(x_left || x_right);
    bool y_above = my < rmin_y, y_below = my > rmax_y;
    bool in_y = !(y_above || y_below);
    bool hit = in_x && in_y;
    if (!hit) {
        float corner_x = x_left ? rmin_x : rmax_x;
        float corner_y = y_above ? rmin_y : rmax_y;
        float width = rmax_x - rmin_x, height = rmax_y - rmin_y;
        float dxf = x_left ? width : -width;
        float dyf = y_above ? height : -height;
        float diff_x = mx - corner_x, diff_y = my - corner_y;
        // helpers.rs:251-256: t = select(in_range, 0, clamp(raw)).  The quotient is only evaluated
        // where the select keeps it; the value is identical to the reference's unconditional form.
        float tx_ = 0.0f, ty_ = 0.0f;
        if (!in_y) {
            float tx_raw = (dxf * conic.c00 * diff_x + dxf * conic.c01 * diff_y) / (dxf * conic.c00 * dxf);
            tx_ = clampf(tx_raw, 0.0f, 1.0f);
        }
        if (!in_x) {
            float ty_raw = (dyf * conic.c01 * diff_x + dyf * conic.c11 * diff_y) / (dyf * conic.c11 * dyf);
            ty_ = clampf(ty_raw, 0.0f, 1.0f);
        }
        float qx = corner_x + tx_ * dxf;
        float qy = corner_y + ty_ * dyf;
        hit = calc_sigma(qx, qy, conic, mx, my) <= pt;
    }
    return hit;
}

// SH basis -> colour, coefficients read through `ld(i)` (i = float index inside the row).
template <int DEG, typename LD>
__device__ __forceinline__ V3 sh_to_color(LD ld, V3 v) {
    auto c3 = [&](int b) { return mk3(ld(b), ld(b + 1), ld(b + 2)); };
    V3 color = scale(c3(0), 0.2820948f);
    if (DEG >= 1) {
        const float f0a = 0.4886025f;
        color = add(color, scale(c3(3), -f0a * v.y));
        color = add(color, scale(c3(6), f0a * v.z));
        color = add(color, scale(c3(9), -f0a * v.x));
    }
    float z2 = v.z * v.z;
    float fc1 = v.x * v.x - v.y * v.y;
    float fs1 = 2.0f * v.x * v.y;
    float p6 = 0.9461747f * z2 - 0.31539157f;
    if (DEG >= 2) {
        float f0b = -1.0925485f * v.z;
        const float f1a = 0.54627424f;
        color = add(color, scale(c3(12), f1a * fs1));
        color = add(color, scale(c3(15), f0b * v.y));
        color = add(color, scale(c3(18), p6));
        color = add(color, scale(c3(21), f0b * v.x));
        color = add(color, scale(c3(24), f1a * fc1));
    }
    float fc2 = v.x * fc1 - v.y * fs1;
    float fs2 = v.x * fs1 + v.y * fc1;
    float p12 = v.z * (1.8658817f * z2 - 1.119529f);
    if (DEG >= 3) {
        float f0c = -2.285229f * z2 + 0.4570458f;
        float f1b = 1.4453057f * v.z;
        const float f2a = -0.5900436f;
        color = add(color, scale(c3(27), f2a * fs2));
        color = add(color, scale(c3(30), f1b * fs1));
        color = add(color, scale(c3(33), f0c * v.y));
        color = add(color, scale(c3(36), p12));
        color = add(color, scale(c3(39), f0c * v.x));
        color = add(color, scale(c3(42), f1b * fc1));
        color = add(color, scale(c3(45), f2a * fc2));
    }
    if (DEG >= 4) {
        float f0d = v.z * (-4.683326f * z2 + 2.0071396f);
        float f1c = 3.3116114f * z2 - 0.47308735f;
        float f2b = -1.7701308f * v.z;
        const float f3a = 0.62583575f;
        float fc3 = v.x * fc2 - v.y * fs2;
        float fs3 = v.x * fs2 + v.y * fc2;
        float p20 = 1.9843135f * v.z * p12 - 1.0062306f * p6;
        color = add(color, scale(c3(48), f3a * fs3));
        color = add(color, scale(c3(51), f2b * fs2));
        color = add(color, scale(c3(54), f1c * fs1));
        color = add(color, scale(c3(57), f0d * v.y));
        color = add(color, scale(c3(60), p20));
        color = add(color, scale(c3(63), f0d * v.x));
        color = add(color, scale(c3(66), f1c * fc1));
        color = add(color, scale(c3(69), f2b * fc2));
        color = add(color, scale(c3(72), f3a * fc3));
    }
    return color;
}

__host__ __device__ __forceinline__ int sh_degree_from_k(uint32_t k) {
    return k == 1 ? 0 : k == 4 ? 1 : k == 9 ? 2 : k == 16 ? 3 : k == 25 ? 4 : -1;
}

}  // namespace bg
