/*
 * oracle/orc_backward.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restatement of the backward pass of the reference rasterizer:
 *   rasterize_bwd : /root/reference/crates/brush-render/src/bwd/kernels/rasterize_backwards.rs:100-391
 *   project_bwd   : /root/reference/crates/brush-render/src/bwd/kernels/project_backwards.rs:17-254
 *   SH VJPs       : /root/reference/crates/brush-render/src/kernels/sh.rs:138-355
 *   pinhole VJP   : /root/reference/crates/brush-render/src/kernels/camera_model/pinhole.rs:58-123
 *
 * The reference accumulates v_combined with f32 atomics (order across tiles
 * is nondeterministic).  The oracle computes the per-(tile,splat) partial sums
 * exactly as one reference thread does (pixels in rank order 0..255), then adds
 * the partials in intersection order -- one legal atomic order, reproducible.
 */
#include "orc_api.h"
#include "orc_math.h"
#include "orc_camera.h"

#include <stdlib.h>

#define TILE_WIDTH 16u
#define TILE_SIZE 256u
#define ALPHA_CUTOFF_MID (1.0f / 255.0f)

float orc_alpha_cutoff_weight(float alpha);
float orc_alpha_cutoff_weight_deriv(float alpha);
osym2 orc_calc_cov2d(ovec3 scale, oquat quat, ovec3 mean_c, const OrcCamera *u);
osym2 orc_compensate_cov2d(osym2 c, int mip, float *filter_comp);

void orc_rasterize_backward(const OrcRender *r, const float *bg3, const float *v_output, int smooth,
                            float *v_combined) {
    const uint32_t w = r->w, h = r->h, tiles_x = r->tiles_x, num_tiles = r->tiles_x * r->tiles_y;
    const uint32_t V = r->num_visible, I = r->num_intersections;
    const float bg_r = bg3 ? bg3[0] : 0.0f, bg_g = bg3 ? bg3[1] : 0.0f, bg_b = bg3 ? bg3[2] : 0.0f;
    memset(v_combined, 0, sizeof(float) * 10 * (size_t)(V ? V : 1));
    float *partial = (float *)calloc((size_t)(I ? I : 1) * 10, sizeof(float));

#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t tile = 0; tile < (int64_t)num_tiles; tile++) {
        const uint32_t range_lo = r->tile_offsets[tile * 2], range_hi = r->tile_offsets[tile * 2 + 1];
        if (range_hi <= range_lo) continue;
        const uint32_t ox = ((uint32_t)tile % tiles_x) * TILE_WIDTH, oy = ((uint32_t)tile / tiles_x) * TILE_WIDTH;
        /* load_pixel_state (rasterize_backwards.rs:186-228) */
        float st[TILE_SIZE][4];
        for (uint32_t rank = 0; rank < TILE_SIZE; rank++) {
            uint32_t px = ox + rank % TILE_WIDTH, py = oy + rank / TILE_WIDTH;
            if (px < w && py < h) {
                const float *o = r->out_img + ((size_t)px + (size_t)py * w) * 4;
                float t_final = 1.0f - o[3];
                st[rank][0] = o[0] - t_final * bg_r;
                st[rank][1] = o[1] - t_final * bg_g;
                st[rank][2] = o[2] - t_final * bg_b;
                st[rank][3] = 1.0f;
            } else {
                st[rank][0] = st[rank][1] = st[rank][2] = st[rank][3] = 0.0f;
            }
        }
        for (uint32_t is = range_lo; is < range_hi; is++) {
            const uint32_t cg = r->cgid_from_isect[is];
            const float *sp = r->projected + (size_t)cg * 9;
            const float xy_x = sp[0], xy_y = sp[1], color_a = sp[5];
            const osym2 conic = {sp[2], sp[3], sp[4]};
            const float cr = sp[6], cgn = sp[7], cb = sp[8];
            const float clamped_r = orc_max(cr, 0.0f), clamped_g = orc_max(cgn, 0.0f), clamped_b = orc_max(cb, 0.0f);
            float g_xy_x = 0, g_xy_y = 0, g_cx = 0, g_cy = 0, g_cz = 0, g_r = 0, g_g = 0, g_b = 0, g_a = 0, g_ref = 0;
            /* accumulate_grads_for_batch (rasterize_backwards.rs:249-391), pixels in rank order */
            for (uint32_t rank = 0; rank < TILE_SIZE; rank++) {
                float state_x = st[rank][0], state_y = st[rank][1], state_z = st[rank][2], state_w = st[rank][3];
                if (!(state_w > 1.0e-4f)) continue;
                uint32_t px = ox + rank % TILE_WIDTH, py = oy + rank / TILE_WIDTH;
                float pcx = (float)px + 0.5f, pcy = (float)py + 0.5f;
                float dx = xy_x - pcx, dy = xy_y - pcy;
                float sigma = 0.5f * (conic.c00 * dx * dx + conic.c11 * dy * dy) + conic.c01 * dx * dy;
                float gaussian = orc_expf(-sigma);
                float alpha = orc_min(0.999f, color_a * gaussian);
                float w_cut = smooth ? orc_alpha_cutoff_weight(alpha) : (alpha >= ALPHA_CUTOFF_MID ? 1.0f : 0.0f);
                if (!(sigma >= 0.0f && w_cut > 0.0f)) continue;
                float alpha_eff = alpha * w_cut;
                float next_t = state_w * (1.0f - alpha_eff);
                if (next_t <= 1.0e-4f) { st[rank][3] = 0.0f; continue; }
                float vis = alpha_eff * state_w;
                size_t pb = ((size_t)px + (size_t)py * w) * 4;
                float v_o_x = v_output[pb], v_o_y = v_output[pb + 1], v_o_z = v_output[pb + 2], v_a = v_output[pb + 3];
                float final_a = r->out_img[pb + 3];
                float t_final = 1.0f - final_a;
                float v_o_w = (v_a - (bg_r * v_o_x + bg_g * v_o_y + bg_b * v_o_z)) * t_final;
                g_r += (cr >= 0.0f) ? vis * v_o_x : 0.0f;
                g_g += (cgn >= 0.0f) ? vis * v_o_y : 0.0f;
                g_b += (cb >= 0.0f) ? vis * v_o_z : 0.0f;
                float ra = 1.0f / (1.0f - alpha_eff);
                float dot_rgb = ((state_w * clamped_r - state_x) * v_o_x + (state_w * clamped_g - state_y) * v_o_y +
                                 (state_w * clamped_b - state_z) * v_o_z) * ra;
                float nrx = state_x - vis * clamped_r, nry = state_y - vis * clamped_g, nrz = state_z - vis * clamped_b;
                float v_alpha_eff = dot_rgb + v_o_w * ra;
                float dw_dalpha = smooth ? orc_alpha_cutoff_weight_deriv(alpha) : 0.0f * alpha;
                float v_alpha = v_alpha_eff * (w_cut + alpha * dw_dalpha);
                float v_sigma = -alpha * v_alpha;
                float vxy_x = v_sigma * (conic.c00 * dx + conic.c01 * dy);
                float vxy_y = v_sigma * (conic.c01 * dx + conic.c11 * dy);
                if (color_a * gaussian <= 0.999f) {
                    g_cx += 0.5f * v_sigma * dx * dx;
                    g_cy += v_sigma * dx * dy;
                    g_cz += 0.5f * v_sigma * dy * dy;
                    g_xy_x += vxy_x;
                    g_xy_y += vxy_y;
                    g_a += v_alpha * gaussian;
                    float isx = (float)w, isy = (float)h;
                    float len = sqrtf(vxy_x * isx * vxy_x * isx + vxy_y * isy * vxy_y * isy);
                    g_ref += len / orc_max(final_a, 1.0e-5f);
                }
                st[rank][0] = nrx; st[rank][1] = nry; st[rank][2] = nrz; st[rank][3] = next_t;
            }
            float *p = partial + (size_t)is * 10;
            p[0] = g_xy_x; p[1] = g_xy_y; p[2] = g_cx; p[3] = g_cy; p[4] = g_cz;
            p[5] = g_r; p[6] = g_g; p[7] = g_b; p[8] = g_a; p[9] = g_ref;
        }
    }
    /* the "atomics": add partials in intersection order, tiles ascending.  Only intersections
     * inside a tile's (trimmed) range were visited; the others have zero partials. */
    for (uint32_t is = 0; is < I; is++) {
        const float *p = partial + (size_t)is * 10;
        float *d = v_combined + (size_t)r->cgid_from_isect[is] * 10;
        for (int k = 0; k < 10; k++) d[k] += p[k];
    }
    free(partial);
}

/* ---- project backward helpers ---- */
static inline ovec3 rd3(const float *c, uint32_t base) { return v3(c[base], c[base + 1], c[base + 2]); }
static inline void wr3(float *c, uint32_t base, ovec3 v) { c[base] = v.x; c[base + 1] = v.y; c[base + 2] = v.z; }

/* kernels/sh.rs:265-355 */
static void sh_coeffs_to_color_vjp(float *vc_out, uint32_t degree, ovec3 v, ovec3 vc) {
    wr3(vc_out, 0, v3_scale(vc, 0.2820948f));
    if (degree >= 1) {
        float f0a = 0.4886025f;
        wr3(vc_out, 3, v3_scale(vc, -f0a * v.y));
        wr3(vc_out, 6, v3_scale(vc, f0a * v.z));
        wr3(vc_out, 9, v3_scale(vc, -f0a * v.x));
        if (degree >= 2) {
            float z2 = v.z * v.z;
            float f0b = -1.0925485f * v.z;
            float f1a = 0.54627424f;
            float fc1 = v.x * v.x - v.y * v.y;
            float fs1 = 2.0f * v.x * v.y;
            float p4 = f1a * fs1, p5 = f0b * v.y, p6 = 0.9461747f * z2 - 0.31539157f, p7 = f0b * v.x, p8 = f1a * fc1;
            wr3(vc_out, 12, v3_scale(vc, p4));
            wr3(vc_out, 15, v3_scale(vc, p5));
            wr3(vc_out, 18, v3_scale(vc, p6));
            wr3(vc_out, 21, v3_scale(vc, p7));
            wr3(vc_out, 24, v3_scale(vc, p8));
            if (degree >= 3) {
                float f0c = -2.285229f * z2 + 0.4570458f;
                float f1b = 1.4453057f * v.z;
                float f2a = -0.5900436f;
                float fc2 = v.x * fc1 - v.y * fs1;
                float fs2 = v.x * fs1 + v.y * fc1;
                float p12 = v.z * (1.8658817f * z2 - 1.119529f);
                float p9 = f2a * fs2, p10 = f1b * fs1, p11 = f0c * v.y, p13 = f0c * v.x, p14 = f1b * fc1, p15 = f2a * fc2;
                wr3(vc_out, 27, v3_scale(vc, p9));
                wr3(vc_out, 30, v3_scale(vc, p10));
                wr3(vc_out, 33, v3_scale(vc, p11));
                wr3(vc_out, 36, v3_scale(vc, p12));
                wr3(vc_out, 39, v3_scale(vc, p13));
                wr3(vc_out, 42, v3_scale(vc, p14));
                wr3(vc_out, 45, v3_scale(vc, p15));
                if (degree >= 4) {
                    float f0d = v.z * (-4.683326f * z2 + 2.0071396f);
                    float f1c = 3.3116114f * z2 - 0.47308735f;
                    float f2b = -1.7701308f * v.z;
                    float f3a = 0.62583575f;
                    float fc3 = v.x * fc2 - v.y * fs2;
                    float fs3 = v.x * fs2 + v.y * fc2;
                    float p20 = 1.9843135f * v.z * p12 + -1.0062306f * p6;
                    float p16 = f3a * fs3, p17 = f2b * fs2, p18 = f1c * fs1, p19 = f0d * v.y, p21 = f0d * v.x,
                          p22 = f1c * fc1, p23 = f2b * fc2, p24 = f3a * fc3;
                    wr3(vc_out, 48, v3_scale(vc, p16));
                    wr3(vc_out, 51, v3_scale(vc, p17));
                    wr3(vc_out, 54, v3_scale(vc, p18));
                    wr3(vc_out, 57, v3_scale(vc, p19));
                    wr3(vc_out, 60, v3_scale(vc, p20));
                    wr3(vc_out, 63, v3_scale(vc, p21));
                    wr3(vc_out, 66, v3_scale(vc, p22));
                    wr3(vc_out, 69, v3_scale(vc, p23));
                    wr3(vc_out, 72, v3_scale(vc, p24));
                }
            }
        }
    }
}

/* kernels/sh.rs:138-259 */
static ovec3 sh_color_viewdir_vjp(const float *c, uint32_t degree, ovec3 v, ovec3 vc) {
    float gx = 0.0f, gy = 0.0f, gz = 0.0f;
    if (degree >= 1) {
        float f0a = 0.4886025f;
        float s_n1 = v3_dot(rd3(c, 3), vc), s_z0 = v3_dot(rd3(c, 6), vc), s_p1 = v3_dot(rd3(c, 9), vc);
        gx += -f0a * s_p1;
        gy += -f0a * s_n1;
        gz += f0a * s_z0;
        if (degree >= 2) {
            float z = v.z, x = v.x, y = v.y;
            float c2 = -1.0925485f, f1a = 0.54627424f;
            float s_n2 = v3_dot(rd3(c, 12), vc);
            s_n1 = v3_dot(rd3(c, 15), vc);
            s_z0 = v3_dot(rd3(c, 18), vc);
            s_p1 = v3_dot(rd3(c, 21), vc);
            float s_p2 = v3_dot(rd3(c, 24), vc);
            gx += 2.0f * f1a * y * s_n2 + c2 * z * s_p1 + 2.0f * f1a * x * s_p2;
            gy += 2.0f * f1a * x * s_n2 + c2 * z * s_n1 - 2.0f * f1a * y * s_p2;
            gz += c2 * y * s_n1 + 2.0f * 0.9461747f * z * s_z0 + c2 * x * s_p1;
            if (degree >= 3) {
                float z2 = z * z, x2 = x * x, y2 = y * y;
                float f2a = -0.5900436f, c1b = 1.4453057f;
                float f1b = c1b * z;
                float c0c = -2.285229f;
                float f0c = c0c * z2 + 0.4570458f;
                float f0c_dz = 2.0f * c0c * z;
                float s_n3 = v3_dot(rd3(c, 27), vc);
                s_n2 = v3_dot(rd3(c, 30), vc);
                s_n1 = v3_dot(rd3(c, 33), vc);
                s_z0 = v3_dot(rd3(c, 36), vc);
                s_p1 = v3_dot(rd3(c, 39), vc);
                s_p2 = v3_dot(rd3(c, 42), vc);
                float s_p3 = v3_dot(rd3(c, 45), vc);
                float d12_z = 3.0f * 1.8658817f * z2 - 1.119529f;
                gx += f2a * 6.0f * x * y * s_n3 + 2.0f * f1b * y * s_n2 + f0c * s_p1 + 2.0f * f1b * x * s_p2 +
                      f2a * 3.0f * (x2 - y2) * s_p3;
                gy += f2a * 3.0f * (x2 - y2) * s_n3 + 2.0f * f1b * x * s_n2 + f0c * s_n1 + (-2.0f) * f1b * y * s_p2 +
                      f2a * (-6.0f) * x * y * s_p3;
                gz += 2.0f * c1b * x * y * s_n2 + f0c_dz * y * s_n1 + d12_z * s_z0 + f0c_dz * x * s_p1 +
                      c1b * (x2 - y2) * s_p2;
                if (degree >= 4) {
                    float fc1 = x2 - y2, fs1 = 2.0f * x * y;
                    float fc2 = x * fc1 - y * fs1, fs2 = x * fs1 + y * fc1;
                    float f0d = z * (-4.683326f * z2 + 2.0071396f);
                    float f0d_dz = -14.049978f * z2 + 2.0071396f;
                    float f1c = 3.3116114f * z2 - 0.47308735f;
                    float f1c_dz = 2.0f * 3.3116114f * z;
                    float f2b_dz_const = -1.7701308f;
                    float f2b = f2b_dz_const * z;
                    float f3a = 0.62583575f;
                    float p_sh12 = z * (1.8658817f * z2 - 1.119529f);
                    float dp_sh12_dz = 3.0f * 1.8658817f * z2 - 1.119529f;
                    float dp_sh6_dz = 2.0f * 0.9461747f * z;
                    float dp_sh20_dz = 1.9843135f * (p_sh12 + z * dp_sh12_dz) - 1.0062306f * dp_sh6_dz;
                    float s_n4 = v3_dot(rd3(c, 48), vc);
                    s_n3 = v3_dot(rd3(c, 51), vc);
                    s_n2 = v3_dot(rd3(c, 54), vc);
                    s_n1 = v3_dot(rd3(c, 57), vc);
                    s_z0 = v3_dot(rd3(c, 60), vc);
                    s_p1 = v3_dot(rd3(c, 63), vc);
                    s_p2 = v3_dot(rd3(c, 66), vc);
                    s_p3 = v3_dot(rd3(c, 69), vc);
                    float s_p4 = v3_dot(rd3(c, 72), vc);
                    gx += f3a * 4.0f * fs2 * s_n4 + f2b * 3.0f * fs1 * s_n3 + f1c * 2.0f * y * s_n2 + f0d * s_p1 +
                          f1c * 2.0f * x * s_p2 + f2b * 3.0f * fc1 * s_p3 + f3a * 4.0f * fc2 * s_p4;
                    gy += f3a * 4.0f * fc2 * s_n4 + f2b * 3.0f * fc1 * s_n3 + f1c * 2.0f * x * s_n2 + f0d * s_n1 +
                          f1c * (-2.0f) * y * s_p2 + f2b * (-3.0f) * fs1 * s_p3 + f3a * (-4.0f) * fs2 * s_p4;
                    gz += f2b_dz_const * fs2 * s_n3 + f1c_dz * fs1 * s_n2 + f0d_dz * y * s_n1 + dp_sh20_dz * s_z0 +
                          f0d_dz * x * s_p1 + f1c_dz * fc1 * s_p2 + f2b_dz_const * fc2 * s_p3;
                }
            }
        }
    }
    return v3(gx, gy, gz);
}

/* project_backwards.rs:18-50 */
static oquat apply_normalize_vjp(oquat q, oquat g) {
    float lsq = q_dot(q, q);
    float l = sqrtf(lsq);
    float inv = 1.0f / (l * lsq);
    float qw = q.w, qx = q.x, qy = q.y, qz = q.z, gw = g.w, gx = g.x, gy = g.y, gz = g.z;
    float cc0 = -qw * qx, cc1 = -qx * qy, cc2 = -qy * qw;
    float cs0 = -qw * qz, cs1 = -qx * qz, cs2 = -qy * qz;
    float sw = qw * qw, sx = qx * qx, sy = qy * qy, sz = qz * qz;
    oquat r;
    r.w = ((lsq - sw) * gw + cc0 * gx + cc2 * gy + cs0 * gz) * inv;
    r.x = (cc0 * gw + (lsq - sx) * gx + cc1 * gy + cs1 * gz) * inv;
    r.y = (cc2 * gw + cc1 * gx + (lsq - sy) * gy + cs2 * gz) * inv;
    r.z = (cs0 * gw + cs1 * gx + cs2 * gy + (lsq - sz) * gz) * inv;
    return r;
}

/* project_backwards.rs:53-77.  v_r column major: c{i}_{x,y,z} */
static oquat quat_to_mat_vjp(oquat q, omat3 v) {
    float qw = q.w, qx = q.x, qy = q.y, qz = q.z;
    float w_grad = qx * (v.c1.z - v.c2.y) + qy * (v.c2.x - v.c0.z) + qz * (v.c0.y - v.c1.x);
    float x_grad = -2.0f * qx * (v.c1.y + v.c2.z) + qy * (v.c0.y + v.c1.x) + qz * (v.c0.z + v.c2.x) +
                   qw * (v.c1.z - v.c2.y);
    float y_grad = qx * (v.c0.y + v.c1.x) - 2.0f * qy * (v.c0.x + v.c2.z) + qz * (v.c1.z + v.c2.y) +
                   qw * (v.c2.x - v.c0.z);
    float z_grad = qx * (v.c0.z + v.c2.x) + qy * (v.c1.z + v.c2.y) - 2.0f * qz * (v.c0.x + v.c1.y) +
                   qw * (v.c0.y - v.c1.x);
    oquat r = {2.0f * w_grad, 2.0f * x_grad, 2.0f * y_grad, 2.0f * z_grad};
    return r;
}

/* project_backwards.rs:84-97 */
static osym2 inverse2x2_vjp(osym2 minv, osym2 v) {
    float tmp00 = -minv.c00 * v.c00 + -minv.c01 * v.c01;
    float tmp01 = -minv.c01 * v.c00 + -minv.c11 * v.c01;
    float tmp10 = -minv.c00 * v.c01 + -minv.c01 * v.c11;
    float tmp11 = -minv.c01 * v.c01 + -minv.c11 * v.c11;
    osym2 r = {tmp00 * minv.c00 + tmp10 * minv.c01, tmp01 * minv.c00 + tmp11 * minv.c01,
               tmp01 * minv.c01 + tmp11 * minv.c11};
    return r;
}

static inline omat3 view_rotation(const OrcCamera *u) {
    omat3 m;
    m.c0 = v3(u->viewmat[0], u->viewmat[1], u->viewmat[2]);
    m.c1 = v3(u->viewmat[3], u->viewmat[4], u->viewmat[5]);
    m.c2 = v3(u->viewmat[6], u->viewmat[7], u->viewmat[8]);
    return m;
}

/* kernels/camera_model/pinhole.rs:58-123 */
static ovec3 projection_vjp_pinhole(omat2x3 jac, ovec3 mean_c, osym3 cov_c, const OrcCamera *u, osym2 v_cov2d,
                                    ovec2 v_mean2d) {
    float fx = u->fx, fy = u->fy;
    float mx = mean_c.x, my = mean_c.y, mz = mean_c.z;
    float inv_z = 1.0f / mz;
    float mx_rz_raw = mx * inv_z, my_rz_raw = my * inv_z;
    float mx_rz = orc_clamp(mx_rz_raw, u->lim_neg_x, u->lim_pos_x);
    float my_rz = orc_clamp(my_rz_raw, u->lim_neg_y, u->lim_pos_y);
    int in_x = mx_rz_raw <= u->lim_pos_x && mx_rz_raw >= u->lim_neg_x;
    int in_y = my_rz_raw <= u->lim_pos_y && my_rz_raw >= u->lim_neg_y;
    float inv_z2 = inv_z * inv_z;
    float inv_z3 = inv_z2 * inv_z;
    float v_mx = fx * inv_z * v_mean2d.x;
    float v_my = fy * inv_z * v_mean2d.y;
    float v_mz = -(fx * mx * v_mean2d.x + fy * my * v_mean2d.y) * inv_z2;
    omat2x3 tmp = s2_mul_mat2x3(v_cov2d, jac);
    float vj00 = 2.0f * v3_dot(m23_row0(tmp), s3_row0(cov_c));
    float vj11 = 2.0f * v3_dot(m23_row1(tmp), s3_row1(cov_c));
    float vj20 = 2.0f * v3_dot(m23_row0(tmp), s3_row2(cov_c));
    float vj21 = 2.0f * v3_dot(m23_row1(tmp), s3_row2(cov_c));
    float tx = mz * mx_rz, ty = mz * my_rz;
    if (in_x) v_mx += -fx * inv_z2 * vj20; else v_mz += -fx * inv_z3 * vj20 * tx;
    if (in_y) v_my += -fy * inv_z2 * vj21; else v_mz += -fy * inv_z3 * vj21 * ty;
    v_mz += -fx * inv_z2 * vj00 - fy * inv_z2 * vj11 + 2.0f * fx * tx * inv_z3 * vj20 + 2.0f * fy * ty * inv_z3 * vj21;
    return v3(v_mx, v_my, v_mz);
}

static uint32_t sh_degree_from_coeffs(uint32_t k) {
    switch (k) { case 1: return 0; case 4: return 1; case 9: return 2; case 16: return 3; case 25: return 4; default: return 0; }
}

void orc_project_backward(const OrcCamera *cam, const OrcRender *r, const float *transforms, const float *sh,
                          const float *raw_opac, const float *v_combined, float *v_transforms, float *v_sh,
                          float *v_raw_opac, float *v_refine) {
    const uint32_t n = r->n, k = r->k, V = r->num_visible;
    const uint32_t degree = sh_degree_from_coeffs(k);
    const int mip = r->mip;
    memset(v_transforms, 0, sizeof(float) * 10 * (size_t)n);
    memset(v_sh, 0, sizeof(float) * 3 * (size_t)k * n);
    memset(v_raw_opac, 0, sizeof(float) * (size_t)n);
    memset(v_refine, 0, sizeof(float) * (size_t)n);
    const ovec3 cam_pos = v3(cam->cam_pos[0], cam->cam_pos[1], cam->cam_pos[2]);
    const omat3 view_rot = view_rotation(cam);
    const ovec3 view_t = v3(cam->viewmat[9], cam->viewmat[10], cam->viewmat[11]);

#pragma omp parallel for schedule(static)
    for (int64_t ci = 0; ci < (int64_t)V; ci++) {
        const uint32_t gi = r->gid_from_cgid[ci];
        const float *rg = v_combined + (size_t)ci * 10;
        float v_mean2d_x = rg[0], v_mean2d_y = rg[1], v_conics_x = rg[2], v_conics_y = rg[3], v_conics_z = rg[4];
        float v_color_r = rg[5], v_color_g = rg[6], v_color_b = rg[7], v_alpha_in = rg[8], v_refine_in = rg[9];
        int any = v_mean2d_x != 0.0f || v_mean2d_y != 0.0f || v_conics_x != 0.0f || v_conics_y != 0.0f ||
                  v_conics_z != 0.0f || v_color_r != 0.0f || v_color_g != 0.0f || v_color_b != 0.0f ||
                  v_alpha_in != 0.0f || v_refine_in != 0.0f;
        if (!any) continue;
        const float *t = transforms + (size_t)gi * 10;
        ovec3 mean = v3(t[0], t[1], t[2]);
        ovec3 scale = v3(orc_expf(t[7]), orc_expf(t[8]), orc_expf(t[9]));
        oquat quat_unorm = {t[3], t[4], t[5], t[6]};
        oquat quat = q_normalize(quat_unorm);

        ovec3 u_world = v3_sub(mean, cam_pos);
        float u_len = v3_length(u_world);
        ovec3 v = v3_scale(u_world, 1.0f / u_len);
        const float *coeffs = sh + (size_t)gi * k * 3;
        ovec3 v_color = v3(v_color_r, v_color_g, v_color_b);
        sh_coeffs_to_color_vjp(v_sh + (size_t)gi * k * 3, degree, v, v_color);
        ovec3 v_v_sh = sh_color_viewdir_vjp(coeffs, degree, v, v_color);
        float v_dot_vv = v3_dot(v, v_v_sh);
        ovec3 v_mean_from_sh = v3_scale(v3_sub(v_v_sh, v3_scale(v, v_dot_vv)), 1.0f / u_len);

        ovec3 mean_c = v3_add(m3_mul_vec3(view_rot, mean), view_t);
        omat3 rm = q_to_mat3(quat);
        omat3 m = m3_mul_diag(rm, scale);
        osym2 raw_cov = orc_calc_cov2d(scale, quat, mean_c, cam);
        float filter_comp;
        osym2 cov = orc_compensate_cov2d(raw_cov, mip, &filter_comp);
        float opac_sig = orc_sigmoid(raw_opac[gi]);
        v_raw_opac[gi] = filter_comp * v_alpha_in * opac_sig * (1.0f - opac_sig);
        float refine_clean = orc_is_finite(v_refine_in) ? v_refine_in : 0.0f;
        v_refine[gi] = orc_clamp(refine_clean, 0.0f, 1.0e32f);

        osym2 conic_inv = s2_inverse(cov);
        osym2 v_inv = {v_conics_x, v_conics_y * 0.5f, v_conics_z};
        osym2 v_cov2d = inverse2x2_vjp(conic_inv, v_inv);
        osym3 covar = m3_outer_product_self(m);
        osym3 cov_c = s3_congruence(covar, view_rot);
        omat2x3 jac = orc_cam_jacobian(mean_c, cam);
        ovec3 v_mean_c; /* camera_model/mod.rs:84-136 */
        switch (cam->camera_model) {
            case ORC_CAM_KB4: v_mean_c = cam_vjp_kb4(jac, mean_c, cov_c, cam, v_cov2d, v2(v_mean2d_x, v_mean2d_y), cam->model_params); break;
            case ORC_CAM_RT8: v_mean_c = cam_vjp_rt8(mean_c, cov_c, cam, v_cov2d, v2(v_mean2d_x, v_mean2d_y), cam->model_params); break;
            case ORC_CAM_TPF: v_mean_c = cam_vjp_tpf(jac, mean_c, cov_c, cam, v_cov2d, v2(v_mean2d_x, v_mean2d_y), cam->model_params); break;
            default: v_mean_c = projection_vjp_pinhole(jac, mean_c, cov_c, cam, v_cov2d, v2(v_mean2d_x, v_mean2d_y)); break;
        }
        osym3 vcc = m23_transpose_congruence_sym2(jac, v_cov2d);
        ovec3 v_mean = v3_add(m3_transpose_mul_vec3(view_rot, v_mean_c), v_mean_from_sh);
        omat3 v_m = s3_mul_mat3(s3_scale(s3_transpose_congruence(vcc, view_rot), 2.0f), m);
        ovec3 v_scale_exp = v3(v3_dot(rm.c0, v_m.c0) * scale.x, v3_dot(rm.c1, v_m.c1) * scale.y,
                               v3_dot(rm.c2, v_m.c2) * scale.z);
        oquat q_grad = quat_to_mat_vjp(quat, m3_mul_diag(v_m, scale));
        oquat v_q = apply_normalize_vjp(quat_unorm, q_grad);

        float *o = v_transforms + (size_t)gi * 10;
        o[0] = v_mean.x; o[1] = v_mean.y; o[2] = v_mean.z;
        o[3] = v_q.w; o[4] = v_q.x; o[5] = v_q.y; o[6] = v_q.z;
        o[7] = v_scale_exp.x; o[8] = v_scale_exp.y; o[9] = v_scale_exp.z;
    }
}
