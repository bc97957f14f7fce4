/* orc_camera.h -- CPU restatement of the reference's camera models (TEST INFRASTRUCTURE ONLY).
 *
 *   dispatch            : brush-render/src/kernels/camera_model/mod.rs:41-136
 *   pinhole             : kernels/camera_model/pinhole.rs:24-123
 *   Kannala-Brandt 4    : kernels/camera_model/kannala_brandt_4.rs:18-337
 *   radial-tangential 8 : kernels/camera_model/radial_tangential_8.rs:23-377
 *   thin-prism fisheye  : kernels/camera_model/thin_prism_fisheye.rs:37-203
 *
 * Each function follows the reference statement by statement (the hand-derived Hessian contractions
 * included), so that the CUDA side -- which gets the same second derivatives from forward-mode dual
 * numbers instead -- is checked by an independent derivation.  The reference bakes the distortion
 * coefficients into the kernel at compile time; here they are OrcCamera.model_params:
 *   KB4: k1 k2 k3 k4 | RT8: k1 k2 k3 k4 k5 k6 p1 p2 | TPF: k1 k2 k3 k4 p1 p2 sx1 sy1.
 * atan2 is orc_atan2f (a fixed sequence of IEEE operations, mirrored by det_atan2f on the CUDA side) so
 * that projected positions and tile lists can be compared bit for bit. */
#ifndef ORC_CAMERA_H
#define ORC_CAMERA_H

#include "orc_api.h"
#include "orc_math.h"

enum { ORC_CAM_PINHOLE = 0, ORC_CAM_KB4 = 1, ORC_CAM_RT8 = 2, ORC_CAM_TPF = 3 };

/* atan on [0, inf) by the Cephes single-precision scheme: two range reductions + a degree-4 polynomial in x^2. */
static inline float orc_atanf_pos(float x) {
    float y0 = 0.0f;
    if (x > 2.414213562373095f) { y0 = 1.5707963267948966f; x = -(1.0f / x); }
    else if (x > 0.4142135623730950f) { y0 = 0.7853981633974483f; x = (x - 1.0f) / (x + 1.0f); }
    float z = x * x;
    float p = 8.05374449538e-2f * z - 1.38776856032e-1f;
    p = p * z + 1.99777106478e-1f;
    p = p * z - 3.33329491539e-1f;
    p = p * z;
    return y0 + (p * x + x);
}
/* atan2(y, x) for y >= 0 (y is a radius here); x may have either sign. */
static inline float orc_atan2f(float y, float x) {
    if (x > 0.0f) return orc_atanf_pos(y / x);
    if (x < 0.0f) return 3.14159265358979f - orc_atanf_pos(y / -x);
    return (y > 0.0f) ? 1.5707963267948966f : 0.0f;
}

/* ---- pinhole (pinhole.rs:24-56) */
static inline void cam_project_pinhole(ovec3 p, const OrcCamera *u, float *ox, float *oy) {
    float inv_z = 1.0f / p.z;
    *ox = u->fx * p.x * inv_z + u->cx;
    *oy = u->fy * p.y * inv_z + u->cy;
}
static inline omat2x3 cam_jacobian_pinhole(ovec3 p, const OrcCamera *u) {
    float inv_z = 1.0f / p.z;
    float dx = u->fx * inv_z, dy = u->fy * inv_z;
    float clamped_x = orc_clamp(p.x * inv_z, u->lim_neg_x, u->lim_pos_x);
    float clamped_y = orc_clamp(p.y * inv_z, u->lim_neg_y, u->lim_pos_y);
    omat2x3 j;
    j.c0 = v2(dx, 0.0f);
    j.c1 = v2(0.0f, dy);
    j.c2 = v2(-dx * clamped_x, -dy * clamped_y);
    return j;
}

/* ---- KB4 (kannala_brandt_4.rs:18-54) */
static inline void cam_project_kb4(ovec3 p, const OrcCamera *u, const float *k, float *ox, float *oy) {
    float x = p.x, y = p.y, z = p.z;
    float inv_z = 1.0f / z;
    float pinhole_u = u->fx * x * inv_z + u->cx;
    float pinhole_v = u->fy * y * inv_z + u->cy;
    float r = sqrtf(x * x + y * y);
    float theta = orc_atan2f(r, z);
    float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta2 * theta4, theta8 = theta4 * theta4;
    float d = theta * (1.0f + k[0] * theta2 + k[1] * theta4 + k[2] * theta6 + k[3] * theta8);
    float inv_r = 1.0f / r;
    float fisheye_u = u->fx * (d * x * inv_r) + u->cx;
    float fisheye_v = u->fy * (d * y * inv_r) + u->cy;
    int near_axis = r < 1e-6f;
    *ox = near_axis ? pinhole_u : fisheye_u;
    *oy = near_axis ? pinhole_v : fisheye_v;
}
/* kannala_brandt_4.rs:56-152 */
static inline omat2x3 cam_jacobian_kb4(ovec3 p, const OrcCamera *u, const float *k) {
    float fx = u->fx, fy = u->fy;
    float x = p.x, y = p.y, z = p.z;
    float inv_z = 1.0f / z;
    float x2 = x * x, y2 = y * y, xy = x * y;
    float r2 = x2 + y2;
    float r = sqrtf(r2);
    float inv_r = 1.0f / r;
    float inv_r3 = inv_r * inv_r * inv_r;
    float rho2 = r2 + z * z;
    float inv_rho2 = 1.0f / rho2;
    float inv_rho2_r = inv_rho2 * inv_r;
    float theta = orc_atan2f(r, z);
    float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
    float d = theta * (1.0f + k[0] * theta2 + k[1] * theta4 + k[2] * theta6 + k[3] * theta8);
    float dd_dtheta = 1.0f + 3.0f * k[0] * theta2 + 5.0f * k[1] * theta4 + 7.0f * k[2] * theta6 + 9.0f * k[3] * theta8;
    float dth_dx = x * z * inv_rho2_r, dth_dy = y * z * inv_rho2_r, dth_dz = -r * inv_rho2;
    float dd_dx = dd_dtheta * dth_dx, dd_dy = dd_dtheta * dth_dy, dd_dz = dd_dtheta * dth_dz;
    float xr = x * inv_r;
    float dxr_dx = y2 * inv_r3, dxr_dy = -xy * inv_r3;
    float du_dx = fx * (dd_dx * xr + d * dxr_dx);
    float du_dy = fx * (dd_dy * xr + d * dxr_dy);
    float du_dz = fx * (dd_dz * xr);
    float yr = y * inv_r;
    float dyr_dx = -xy * inv_r3, dyr_dy = x2 * inv_r3;
    float dv_dx = fy * (dd_dx * yr + d * dyr_dx);
    float dv_dy = fy * (dd_dy * yr + d * dyr_dy);
    float dv_dz = fy * (dd_dz * yr);
    int near_axis = r < 1e-6f;
    float dx = fx * inv_z, dy = fy * inv_z;
    omat2x3 j;
    j.c0 = v2(near_axis ? dx : du_dx, near_axis ? 0.0f : dv_dx);
    j.c1 = v2(near_axis ? 0.0f : du_dy, near_axis ? dy : dv_dy);
    j.c2 = v2(near_axis ? -dx * x * inv_z : du_dz, near_axis ? -dy * y * inv_z : dv_dz);
    return j;
}
/* kannala_brandt_4.rs:154-337 */
static inline ovec3 cam_vjp_kb4(omat2x3 jac, ovec3 mean_c, osym3 cov_c, const OrcCamera *u, osym2 v_cov2d, ovec2 v_mean2d,
                                const float *k) {
    float fx = u->fx, fy = u->fy;
    float k1 = k[0], k2 = k[1], k3 = k[2], k4 = k[3];
    float mx = mean_c.x, my = mean_c.y, mz = mean_c.z;
    float r2 = mx * mx + my * my;
    float r = orc_max(sqrtf(r2), 1.0e-8f);
    float rho2 = r2 + mz * mz;
    float theta = orc_atan2f(r, mz);
    float th2 = theta * theta, th4 = th2 * th2, th6 = th4 * th2, th8 = th4 * th4;
    float theta_d = theta * (1.0f + k1 * th2 + k2 * th4 + k3 * th6 + k4 * th8);
    float p1 = 1.0f + 3.0f * k1 * th2 + 5.0f * k2 * th4 + 7.0f * k3 * th6 + 9.0f * k4 * th8;
    float p2 = 6.0f * k1 * theta + 20.0f * k2 * theta * th2 + 42.0f * k3 * theta * th4 + 72.0f * k4 * theta * th6;
    float inv_r = 1.0f / r;
    float inv_r3 = inv_r * inv_r * inv_r;
    float inv_r5 = inv_r3 * inv_r * inv_r;
    float inv_rho2 = 1.0f / rho2;
    float inv_rho2_sq = inv_rho2 * inv_rho2;
    float inv_rho2_r = inv_rho2 * inv_r;
    float dth_x = mx * mz * inv_rho2_r, dth_y = my * mz * inv_rho2_r, dth_z = -r * inv_rho2;
    float xr = mx * inv_r, yr = my * inv_r;
    float dxr_x = my * my * inv_r3, dxr_y = -mx * my * inv_r3;
    float dyr_x = dxr_y, dyr_y = mx * mx * inv_r3;
    float dg_x = p1 * dth_x, dg_y = p1 * dth_y, dg_z = p1 * dth_z;
    float v_mx = v2_dot(v_mean2d, jac.c0);
    float v_my = v2_dot(v_mean2d, jac.c1);
    float v_mz = v2_dot(v_mean2d, jac.c2);
    omat2x3 tmp = s2_mul_mat2x3(v_cov2d, jac);
    float vj_u0 = 2.0f * v3_dot(m23_row0(tmp), s3_row0(cov_c));
    float vj_u1 = 2.0f * v3_dot(m23_row0(tmp), s3_row1(cov_c));
    float vj_u2 = 2.0f * v3_dot(m23_row0(tmp), s3_row2(cov_c));
    float vj_v0 = 2.0f * v3_dot(m23_row1(tmp), s3_row0(cov_c));
    float vj_v1 = 2.0f * v3_dot(m23_row1(tmp), s3_row1(cov_c));
    float vj_v2 = 2.0f * v3_dot(m23_row1(tmp), s3_row2(cov_c));
    float three_r2_z2 = 3.0f * r2 + mz * mz;
    float r2_minus_z2 = r2 - mz * mz;
    float h_th_00 = mz * (r2 * rho2 - mx * mx * three_r2_z2) * inv_r3 * inv_rho2_sq;
    float h_th_11 = mz * (r2 * rho2 - my * my * three_r2_z2) * inv_r3 * inv_rho2_sq;
    float h_th_01 = -mx * my * mz * three_r2_z2 * inv_r3 * inv_rho2_sq;
    float h_th_02 = mx * r2_minus_z2 * inv_r * inv_rho2_sq;
    float h_th_12 = my * r2_minus_z2 * inv_r * inv_rho2_sq;
    float h_th_22 = 2.0f * mz * r * inv_rho2_sq;
    float two_x2_my2 = 2.0f * mx * mx - my * my;
    float two_y2_mx2 = 2.0f * my * my - mx * mx;
    float h_xr_00 = -3.0f * mx * my * my * inv_r5;
    float h_xr_01 = my * two_x2_my2 * inv_r5;
    float h_xr_11 = mx * two_y2_mx2 * inv_r5;
    float h_yr_00 = my * two_x2_my2 * inv_r5;
    float h_yr_01 = mx * two_y2_mx2 * inv_r5;
    float h_yr_11 = -3.0f * mx * mx * my * inv_r5;
    { /* (0,0) */
        float d2g = p2 * dth_x * dth_x + p1 * h_th_00;
        float d_ju = fx * (d2g * xr + dg_x * dxr_x + dg_x * dxr_x + theta_d * h_xr_00);
        float d_jv = fy * (d2g * yr + dg_x * dyr_x + dg_x * dyr_x + theta_d * h_yr_00);
        v_mx += vj_u0 * d_ju + vj_v0 * d_jv;
    }
    { /* (1,0) */
        float d2g = p2 * dth_y * dth_x + p1 * h_th_01;
        float d_ju = fx * (d2g * xr + dg_y * dxr_x + dg_x * dxr_y + theta_d * h_xr_01);
        float d_jv = fy * (d2g * yr + dg_y * dyr_x + dg_x * dyr_y + theta_d * h_yr_01);
        v_mx += vj_u1 * d_ju + vj_v1 * d_jv;
    }
    { /* (2,0) */
        float d2g = p2 * dth_z * dth_x + p1 * h_th_02;
        float d_ju = fx * (d2g * xr + dg_x * 0.0f + dg_z * dxr_x);
        float d_jv = fy * (d2g * yr + dg_x * 0.0f + dg_z * dyr_x);
        v_mx += vj_u2 * d_ju + vj_v2 * d_jv;
    }
    { /* (0,1) */
        float d2g = p2 * dth_x * dth_y + p1 * h_th_01;
        float d_ju = fx * (d2g * xr + dg_x * dxr_y + dg_y * dxr_x + theta_d * h_xr_01);
        float d_jv = fy * (d2g * yr + dg_x * dyr_y + dg_y * dyr_x + theta_d * h_yr_01);
        v_my += vj_u0 * d_ju + vj_v0 * d_jv;
    }
    { /* (1,1) */
        float d2g = p2 * dth_y * dth_y + p1 * h_th_11;
        float d_ju = fx * (d2g * xr + dg_y * dxr_y + dg_y * dxr_y + theta_d * h_xr_11);
        float d_jv = fy * (d2g * yr + dg_y * dyr_y + dg_y * dyr_y + theta_d * h_yr_11);
        v_my += vj_u1 * d_ju + vj_v1 * d_jv;
    }
    { /* (2,1) */
        float d2g = p2 * dth_z * dth_y + p1 * h_th_12;
        float d_ju = fx * (d2g * xr + dg_y * 0.0f + dg_z * dxr_y);
        float d_jv = fy * (d2g * yr + dg_y * 0.0f + dg_z * dyr_y);
        v_my += vj_u2 * d_ju + vj_v2 * d_jv;
    }
    { /* (0,2) */
        float d2g = p2 * dth_x * dth_z + p1 * h_th_02;
        float d_ju = fx * (d2g * xr + dg_z * dxr_x + dg_x * 0.0f);
        float d_jv = fy * (d2g * yr + dg_z * dyr_x + dg_x * 0.0f);
        v_mz += vj_u0 * d_ju + vj_v0 * d_jv;
    }
    { /* (1,2) */
        float d2g = p2 * dth_y * dth_z + p1 * h_th_12;
        float d_ju = fx * (d2g * xr + dg_z * dxr_y + dg_y * 0.0f);
        float d_jv = fy * (d2g * yr + dg_z * dyr_y + dg_y * 0.0f);
        v_mz += vj_u1 * d_ju + vj_v1 * d_jv;
    }
    { /* (2,2) */
        float d2g = p2 * dth_z * dth_z + p1 * h_th_22;
        float d_ju = fx * (d2g * xr);
        float d_jv = fy * (d2g * yr);
        v_mz += vj_u2 * d_ju + vj_v2 * d_jv;
    }
    return v3(v_mx, v_my, v_mz);
}

/* ---- RT8 (radial_tangential_8.rs:23-64) */
static inline void cam_project_rt8(ovec3 p, const OrcCamera *u, const float *c, float *ox, float *oy) {
    float k1 = c[0], k2 = c[1], k3 = c[2], k4 = c[3], k5 = c[4], k6 = c[5], p1 = c[6], p2 = c[7];
    float x_ = p.x / p.z, y_ = p.y / p.z;
    float x_2 = x_ * x_, y_2 = y_ * y_;
    float r2 = x_2 + y_2, r4 = r2 * r2, r6 = r4 * r2;
    float d = (1.0f + k1 * r2 + k2 * r4 + k3 * r6) / (1.0f + k4 * r2 + k5 * r4 + k6 * r6);
    float x_y_ = x_ * y_;
    float x__ = x_ * d + 2.0f * p1 * x_y_ + p2 * (r2 + 2.0f * x_2);
    float y__ = y_ * d + 2.0f * p2 * x_y_ + p1 * (r2 + 2.0f * y_2);
    *ox = u->fx * x__ + u->cx;
    *oy = u->fy * y__ + u->cy;
}
/* radial_tangential_8.rs:66-142 */
static inline omat2x3 cam_jacobian_rt8(ovec3 p, const OrcCamera *u, const float *c) {
    float fx = u->fx, fy = u->fy;
    float k1 = c[0], k2 = c[1], k3 = c[2], k4 = c[3], k5 = c[4], k6 = c[5], p1 = c[6], p2 = c[7];
    float x = p.x, y = p.y, z = p.z;
    float inv_z = 1.0f / z;
    float inv_z2 = inv_z * inv_z;
    float x_n = orc_clamp(x * inv_z, u->lim_neg_x, u->lim_pos_x);
    float y_n = orc_clamp(y * inv_z, u->lim_neg_y, u->lim_pos_y);
    float xc = x_n * z, yc = y_n * z;
    float r2 = x_n * x_n + y_n * y_n, r4 = r2 * r2, r6 = r4 * r2;
    float n_poly = 1.0f + k1 * r2 + k2 * r4 + k3 * r6;
    float dn_poly = 1.0f + k4 * r2 + k5 * r4 + k6 * r6;
    float np_poly = k1 + 2.0f * k2 * r2 + 3.0f * k3 * r4;
    float dnp_poly = k4 + 2.0f * k5 * r2 + 3.0f * k6 * r4;
    float inv_dn = 1.0f / dn_poly;
    float inv_dn2 = inv_dn * inv_dn;
    float r_val = n_poly * inv_dn;
    float rp_val = (np_poly * dn_poly - n_poly * dnp_poly) * inv_dn2;
    float d00 = r_val + 2.0f * x_n * x_n * rp_val + 2.0f * p1 * y_n + 6.0f * p2 * x_n;
    float d01 = 2.0f * x_n * y_n * rp_val + 2.0f * p1 * x_n + 2.0f * p2 * y_n;
    float d10 = d01;
    float d11 = r_val + 2.0f * y_n * y_n * rp_val + 6.0f * p1 * y_n + 2.0f * p2 * x_n;
    omat2x3 j;
    j.c0 = v2(fx * d00 * inv_z, fy * d10 * inv_z);
    j.c1 = v2(fx * d01 * inv_z, fy * d11 * inv_z);
    j.c2 = v2(-fx * (d00 * xc + d01 * yc) * inv_z2, -fy * (d10 * xc + d11 * yc) * inv_z2);
    return j;
}
/* radial_tangential_8.rs:144-377 */
static inline ovec3 cam_vjp_rt8(ovec3 mean_c, osym3 cov_c, const OrcCamera *u, osym2 v_cov2d, ovec2 v_mean2d, const float *c) {
    float fx = u->fx, fy = u->fy;
    float k1 = c[0], k2 = c[1], k3 = c[2], k4 = c[3], k5 = c[4], k6 = c[5], p1 = c[6], p2 = c[7];
    float mx = mean_c.x, my = mean_c.y, mz = mean_c.z;
    float inv_z = 1.0f / mz;
    float mx_rz_raw = mx * inv_z, my_rz_raw = my * inv_z;
    float mx_rz = orc_clamp(mx_rz_raw, u->lim_neg_x, u->lim_pos_x);
    float my_rz = orc_clamp(my_rz_raw, u->lim_neg_y, u->lim_pos_y);
    int in_x = mx_rz_raw <= u->lim_pos_x && mx_rz_raw >= u->lim_neg_x;
    int in_y = my_rz_raw <= u->lim_pos_y && my_rz_raw >= u->lim_neg_y;
    float xc = mx_rz * mz, yc = my_rz * mz;
    float inv_z2 = inv_z * inv_z, inv_z3 = inv_z2 * inv_z;
    float x = xc * inv_z, y = yc * inv_z;
    float r2 = x * x + y * y, r4 = r2 * r2;
    float n_poly = 1.0f + k1 * r2 + k2 * r4 + k3 * r2 * r4;
    float dn_poly = 1.0f + k4 * r2 + k5 * r4 + k6 * r2 * r4;
    float np_poly = k1 + 2.0f * k2 * r2 + 3.0f * k3 * r4;
    float dnp_poly = k4 + 2.0f * k5 * r2 + 3.0f * k6 * r4;
    float npp_poly = 2.0f * k2 + 6.0f * k3 * r2;
    float dnpp_poly = 2.0f * k5 + 6.0f * k6 * r2;
    float inv_dn = 1.0f / dn_poly;
    float inv_dn2 = inv_dn * inv_dn, inv_dn3 = inv_dn2 * inv_dn;
    float rr = n_poly * inv_dn;
    float rrp = (np_poly * dn_poly - n_poly * dnp_poly) * inv_dn2;
    float rrpp = (npp_poly * dn_poly * dn_poly - 2.0f * np_poly * dn_poly * dnp_poly - n_poly * dnpp_poly * dn_poly +
                  2.0f * n_poly * dnp_poly * dnp_poly) * inv_dn3;
    float rx = 2.0f * x * rrp, ry = 2.0f * y * rrp;
    float rpx = 2.0f * x * rrpp, rpy = 2.0f * y * rrpp;
    float d00 = rr + 2.0f * x * x * rrp + 2.0f * p1 * y + 6.0f * p2 * x;
    float d01 = 2.0f * x * y * rrp + 2.0f * p1 * x + 2.0f * p2 * y;
    float d10 = d01;
    float d11 = rr + 2.0f * y * y * rrp + 6.0f * p1 * y + 2.0f * p2 * x;
    float js00 = fx * d00 * inv_z, js01 = fx * d01 * inv_z, js02 = -fx * (d00 * xc + d01 * yc) * inv_z2;
    float js10 = fy * d10 * inv_z, js11 = fy * d11 * inv_z, js12 = -fy * (d10 * xc + d11 * yc) * inv_z2;
    float je00 = in_x ? js00 : 0.0f, je10 = in_x ? js10 : 0.0f;
    float je01 = in_y ? js01 : 0.0f, je11 = in_y ? js11 : 0.0f;
    float je02 = (in_x ? 0.0f : mx_rz * js00) + (in_y ? 0.0f : my_rz * js01) + js02;
    float je12 = (in_x ? 0.0f : mx_rz * js10) + (in_y ? 0.0f : my_rz * js11) + js12;
    float v_mx = je00 * v_mean2d.x + je10 * v_mean2d.y;
    float v_my = je01 * v_mean2d.x + je11 * v_mean2d.y;
    float v_mz = je02 * v_mean2d.x + je12 * v_mean2d.y;
    omat2x3 je;
    je.c0 = v2(je00, je10); je.c1 = v2(je01, je11); je.c2 = v2(je02, je12);
    omat2x3 tmp = s2_mul_mat2x3(v_cov2d, je);
    float ve_u0 = 2.0f * v3_dot(m23_row0(tmp), s3_row0(cov_c));
    float ve_u1 = 2.0f * v3_dot(m23_row0(tmp), s3_row1(cov_c));
    float ve_u2 = 2.0f * v3_dot(m23_row0(tmp), s3_row2(cov_c));
    float ve_v0 = 2.0f * v3_dot(m23_row1(tmp), s3_row0(cov_c));
    float ve_v1 = 2.0f * v3_dot(m23_row1(tmp), s3_row1(cov_c));
    float ve_v2 = 2.0f * v3_dot(m23_row1(tmp), s3_row2(cov_c));
    float vs_u0 = in_x ? ve_u0 : mx_rz * ve_u2, vs_v0 = in_x ? ve_v0 : mx_rz * ve_v2;
    float vs_u1 = in_y ? ve_u1 : my_rz * ve_u2, vs_v1 = in_y ? ve_v1 : my_rz * ve_v2;
    float vs_u2 = ve_u2, vs_v2 = ve_v2;
    float dd00_dx = rx + 4.0f * x * rrp + 2.0f * x * x * rpx + 6.0f * p2;
    float dd00_dy = ry + 2.0f * x * x * rpy + 2.0f * p1;
    float dd01_dx = 2.0f * y * rrp + 2.0f * x * y * rpx + 2.0f * p1;
    float dd01_dy = 2.0f * x * rrp + 2.0f * x * y * rpy + 2.0f * p2;
    float dd10_dx = dd01_dx, dd10_dy = dd01_dy;
    float dd11_dx = rx + 2.0f * y * y * rpx + 2.0f * p2;
    float dd11_dy = ry + 4.0f * y * rrp + 2.0f * y * y * rpy + 6.0f * p1;
    float dd00_dxc = dd00_dx * inv_z, dd00_dyc = dd00_dy * inv_z, dd00_dz = -(xc * dd00_dx + yc * dd00_dy) * inv_z2;
    float dd01_dxc = dd01_dx * inv_z, dd01_dyc = dd01_dy * inv_z, dd01_dz = -(xc * dd01_dx + yc * dd01_dy) * inv_z2;
    float dd10_dxc = dd10_dx * inv_z, dd10_dyc = dd10_dy * inv_z, dd10_dz = -(xc * dd10_dx + yc * dd10_dy) * inv_z2;
    float dd11_dxc = dd11_dx * inv_z, dd11_dyc = dd11_dy * inv_z, dd11_dz = -(xc * dd11_dx + yc * dd11_dy) * inv_z2;
    float djs00_dxc = fx * dd00_dxc * inv_z, djs00_dyc = fx * dd00_dyc * inv_z, djs00_dz = fx * (dd00_dz * inv_z - d00 * inv_z2);
    float djs01_dxc = fx * dd01_dxc * inv_z, djs01_dyc = fx * dd01_dyc * inv_z, djs01_dz = fx * (dd01_dz * inv_z - d01 * inv_z2);
    float djs10_dxc = fy * dd10_dxc * inv_z, djs10_dyc = fy * dd10_dyc * inv_z, djs10_dz = fy * (dd10_dz * inv_z - d10 * inv_z2);
    float djs11_dxc = fy * dd11_dxc * inv_z, djs11_dyc = fy * dd11_dyc * inv_z, djs11_dz = fy * (dd11_dz * inv_z - d11 * inv_z2);
    float djs02_dxc = -fx * (dd00_dxc * xc + d00 + dd01_dxc * yc) * inv_z2;
    float djs02_dyc = -fx * (dd00_dyc * xc + dd01_dyc * yc + d01) * inv_z2;
    float djs02_dz = -fx * ((dd00_dz * xc + dd01_dz * yc) * inv_z2 - 2.0f * (d00 * xc + d01 * yc) * inv_z3);
    float djs12_dxc = -fy * (dd10_dxc * xc + d10 + dd11_dxc * yc) * inv_z2;
    float djs12_dyc = -fy * (dd10_dyc * xc + dd11_dyc * yc + d11) * inv_z2;
    float djs12_dz = -fy * ((dd10_dz * xc + dd11_dz * yc) * inv_z2 - 2.0f * (d10 * xc + d11 * yc) * inv_z3);
    float c_xc = vs_u0 * djs00_dxc + vs_u1 * djs01_dxc + vs_u2 * djs02_dxc + vs_v0 * djs10_dxc + vs_v1 * djs11_dxc + vs_v2 * djs12_dxc;
    float c_yc = vs_u0 * djs00_dyc + vs_u1 * djs01_dyc + vs_u2 * djs02_dyc + vs_v0 * djs10_dyc + vs_v1 * djs11_dyc + vs_v2 * djs12_dyc;
    float c_z = vs_u0 * djs00_dz + vs_u1 * djs01_dz + vs_u2 * djs02_dz + vs_v0 * djs10_dz + vs_v1 * djs11_dz + vs_v2 * djs12_dz;
    if (in_x) v_mx += c_xc;
    if (in_y) v_my += c_yc;
    v_mz += c_z;
    if (!in_x) v_mz += mx_rz * c_xc;
    if (!in_y) v_mz += my_rz * c_yc;
    return v3(v_mx, v_my, v_mz);
}

/* ---- thin-prism fisheye (thin_prism_fisheye.rs:37-203); coefficients: k1..k4, p1, p2, sx1, sy1 */
static inline void tpf_polys(float x, float y, const float *c, float *nu, float *nv, float *dnu_dx, float *dnu_dy,
                             float *dnv_dx, float *dnv_dy) {
    float p1 = c[4], p2 = c[5], sx1 = c[6], sy1 = c[7];
    float x2 = x * x, y2 = y * y, xy = x * y, r2 = x2 + y2;
    *nu = 2.0f * p1 * xy + p2 * (3.0f * x2 + y2) + sx1 * r2;
    *nv = 2.0f * p2 * xy + p1 * (x2 + 3.0f * y2) + sy1 * r2;
    *dnu_dx = 2.0f * (p1 * y + (3.0f * p2 + sx1) * x);
    *dnu_dy = 2.0f * (p1 * x + (p2 + sx1) * y);
    *dnv_dx = 2.0f * (p2 * y + (p1 + sy1) * x);
    *dnv_dy = 2.0f * (p2 * x + (3.0f * p1 + sy1) * y);
}
static inline void cam_project_tpf(ovec3 p, const OrcCamera *u, const float *c, float *ox, float *oy) {
    float uk, vk;
    cam_project_kb4(p, u, c, &uk, &vk);
    float inv_z = 1.0f / p.z;
    float inv_z2 = inv_z * inv_z;
    float nu, nv, a, b, cc, d;
    tpf_polys(p.x, p.y, c, &nu, &nv, &a, &b, &cc, &d);
    *ox = uk + u->fx * nu * inv_z2;
    *oy = vk + u->fy * nv * inv_z2;
}
static inline omat2x3 cam_jacobian_tpf(ovec3 p, const OrcCamera *u, const float *c) {
    omat2x3 kj = cam_jacobian_kb4(p, u, c);
    float fx = u->fx, fy = u->fy;
    float inv_z = 1.0f / p.z;
    float inv_z2 = inv_z * inv_z, inv_z3 = inv_z2 * inv_z;
    float nu, nv, dnu_dx, dnu_dy, dnv_dx, dnv_dy;
    tpf_polys(p.x, p.y, c, &nu, &nv, &dnu_dx, &dnu_dy, &dnv_dx, &dnv_dy);
    omat2x3 j;
    j.c0 = v2(kj.c0.x + fx * dnu_dx * inv_z2, kj.c0.y + fy * dnv_dx * inv_z2);
    j.c1 = v2(kj.c1.x + fx * dnu_dy * inv_z2, kj.c1.y + fy * dnv_dy * inv_z2);
    j.c2 = v2(kj.c2.x + -2.0f * fx * nu * inv_z3, kj.c2.y + -2.0f * fy * nv * inv_z3);
    return j;
}
static inline ovec3 cam_vjp_tpf(omat2x3 jac, ovec3 mean_c, osym3 cov_c, const OrcCamera *u, osym2 v_cov2d, ovec2 v_mean2d,
                                const float *c) {
    ovec3 kg = cam_vjp_kb4(jac, mean_c, cov_c, u, v_cov2d, v_mean2d, c);
    float fx = u->fx, fy = u->fy;
    float p1 = c[4], p2 = c[5], sx1 = c[6], sy1 = c[7];
    float inv_z = 1.0f / mean_c.z;
    float inv_z2 = inv_z * inv_z, inv_z3 = inv_z2 * inv_z, inv_z4 = inv_z2 * inv_z2;
    float nu, nv, dnu_dx, dnu_dy, dnv_dx, dnv_dy;
    tpf_polys(mean_c.x, mean_c.y, c, &nu, &nv, &dnu_dx, &dnu_dy, &dnv_dx, &dnv_dy);
    float d2nu_dxx = 6.0f * p2 + 2.0f * sx1, d2nu_dyy = 2.0f * p2 + 2.0f * sx1, d2nu_dxy = 2.0f * p1;
    float d2nv_dxx = 2.0f * p1 + 2.0f * sy1, d2nv_dyy = 6.0f * p1 + 2.0f * sy1, d2nv_dxy = 2.0f * p2;
    float h_u_00 = d2nu_dxx * inv_z2, h_u_01 = d2nu_dxy * inv_z2, h_u_11 = d2nu_dyy * inv_z2;
    float h_u_02 = -2.0f * dnu_dx * inv_z3, h_u_12 = -2.0f * dnu_dy * inv_z3, h_u_22 = 6.0f * nu * inv_z4;
    float h_v_00 = d2nv_dxx * inv_z2, h_v_01 = d2nv_dxy * inv_z2, h_v_11 = d2nv_dyy * inv_z2;
    float h_v_02 = -2.0f * dnv_dx * inv_z3, h_v_12 = -2.0f * dnv_dy * inv_z3, h_v_22 = 6.0f * nv * inv_z4;
    omat2x3 tmp = s2_mul_mat2x3(v_cov2d, jac);
    float vj_u0 = 2.0f * v3_dot(m23_row0(tmp), s3_row0(cov_c));
    float vj_u1 = 2.0f * v3_dot(m23_row0(tmp), s3_row1(cov_c));
    float vj_u2 = 2.0f * v3_dot(m23_row0(tmp), s3_row2(cov_c));
    float vj_v0 = 2.0f * v3_dot(m23_row1(tmp), s3_row0(cov_c));
    float vj_v1 = 2.0f * v3_dot(m23_row1(tmp), s3_row1(cov_c));
    float vj_v2 = 2.0f * v3_dot(m23_row1(tmp), s3_row2(cov_c));
    float v_mx = fx * (vj_u0 * h_u_00 + vj_u1 * h_u_01 + vj_u2 * h_u_02) + fy * (vj_v0 * h_v_00 + vj_v1 * h_v_01 + vj_v2 * h_v_02);
    float v_my = fx * (vj_u0 * h_u_01 + vj_u1 * h_u_11 + vj_u2 * h_u_12) + fy * (vj_v0 * h_v_01 + vj_v1 * h_v_11 + vj_v2 * h_v_12);
    float v_mz = fx * (vj_u0 * h_u_02 + vj_u1 * h_u_12 + vj_u2 * h_u_22) + fy * (vj_v0 * h_v_02 + vj_v1 * h_v_12 + vj_v2 * h_v_22);
    return v3(kg.x + v_mx, kg.y + v_my, kg.z + v_mz);
}

/* ---- dispatch (camera_model/mod.rs:41-136) */
static inline void orc_cam_project(ovec3 p, const OrcCamera *u, float *ox, float *oy) {
    switch (u->camera_model) {
        case ORC_CAM_KB4: cam_project_kb4(p, u, u->model_params, ox, oy); break;
        case ORC_CAM_RT8: cam_project_rt8(p, u, u->model_params, ox, oy); break;
        case ORC_CAM_TPF: cam_project_tpf(p, u, u->model_params, ox, oy); break;
        default: cam_project_pinhole(p, u, ox, oy); break;
    }
}
static inline omat2x3 orc_cam_jacobian(ovec3 p, const OrcCamera *u) {
    switch (u->camera_model) {
        case ORC_CAM_KB4: return cam_jacobian_kb4(p, u, u->model_params);
        case ORC_CAM_RT8: return cam_jacobian_rt8(p, u, u->model_params);
        case ORC_CAM_TPF: return cam_jacobian_tpf(p, u, u->model_params);
        default: return cam_jacobian_pinhole(p, u);
    }
}
/* project_forward.rs:47-61: near-plane cull for pinhole, angular cull for the distorted models */
static inline int orc_cam_in_front(ovec3 mean_c, const OrcCamera *u) {
    if (u->camera_model == ORC_CAM_PINHOLE) return !(mean_c.z < 0.01f);
    float r = sqrtf(mean_c.x * mean_c.x + mean_c.y * mean_c.y);
    float theta = orc_atan2f(r, mean_c.z);
    return !(theta > u->half_max_render_fov);
}

#endif
