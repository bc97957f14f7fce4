// bg_camera.cuh -- the distorted camera models (Kannala-Brandt 4, radial-tangential 8, thin-prism fisheye).
// Include only from translation units built with -fmad=false (bit-reproducible forward projection).
//
// Reference semantics:
//   kernels/camera_model/mod.rs:41-136            dispatch of project / Jacobian / VJP
//   kernels/camera_model/kannala_brandt_4.rs      project_kb4 :18-54, Jacobian :56-152, VJP :154-337
//   kernels/camera_model/radial_tangential_8.rs   project_rt8 :23-64, Jacobian :66-142, VJP :144-377
//   kernels/camera_model/thin_prism_fisheye.rs    polys :37-61, project :63-80, Jacobian :82-118, VJP :120-203
//
// The reference specialises its kernels on the distortion coefficients at compile time and carries a
// hand-derived Hessian contraction per model.  Here the coefficients are uniforms (BgCamera.model_params)
// and each Jacobian is written ONCE as a template over the scalar type: instantiated with float it is the
// forward Jacobian (the reference's operation order, so the CPU checker reproduces projected splats bit for bit);
// instantiated with a 3-partial dual number it yields dJ/d(point), which is all the VJP's second-order path
// needs.  The pinhole model keeps its dedicated closed forms (bg_project.cuh / project_bwd.cu).
#pragma once
#include "bg_common.cuh"
#include "bg_math.cuh"

namespace bg {

// ---- deterministic atan2 for y >= 0 (Cephes atanf scheme): a fixed sequence of IEEE operations, so that the
// CPU checker in tests/ can mirror it exactly
__device__ __forceinline__ float det_atanf_pos(float x) {
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
__device__ __forceinline__ float det_atan2f(float y, float x) {
    if (x > 0.0f) return det_atanf_pos(y / x);
    if (x < 0.0f) return 3.14159265358979f - det_atanf_pos(y / -x);
    return (y > 0.0f) ? 1.5707963267948966f : 0.0f;
}

// ---- forward-mode dual number: value + partials w.r.t. three seeds
struct D3 { float v, d0, d1, d2; };
__device__ __forceinline__ D3 mkd(float v, float d0, float d1, float d2) { D3 r; r.v = v; r.d0 = d0; r.d1 = d1; r.d2 = d2; return r; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return mkd(a.v + b.v, a.d0 + b.d0, a.d1 + b.d1, a.d2 + b.d2); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return mkd(a.v - b.v, a.d0 - b.d0, a.d1 - b.d1, a.d2 - b.d2); }
__device__ __forceinline__ D3 operator-(D3 a) { return mkd(-a.v, -a.d0, -a.d1, -a.d2); }
__device__ __forceinline__ D3 operator*(D3 a, D3 b) {
    return mkd(a.v * b.v, a.d0 * b.v + a.v * b.d0, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + a.v * b.d2);
}
__device__ __forceinline__ D3 operator/(D3 a, D3 b) {
    float inv = 1.0f / b.v, q = a.v * inv;
    return mkd(q, (a.d0 - q * b.d0) * inv, (a.d1 - q * b.d1) * inv, (a.d2 - q * b.d2) * inv);
}
__device__ __forceinline__ D3 operator+(D3 a, float b) { return mkd(a.v + b, a.d0, a.d1, a.d2); }
__device__ __forceinline__ D3 operator+(float a, D3 b) { return mkd(a + b.v, b.d0, b.d1, b.d2); }
__device__ __forceinline__ D3 operator-(D3 a, float b) { return mkd(a.v - b, a.d0, a.d1, a.d2); }
__device__ __forceinline__ D3 operator-(float a, D3 b) { return mkd(a - b.v, -b.d0, -b.d1, -b.d2); }
__device__ __forceinline__ D3 operator*(D3 a, float b) { return mkd(a.v * b, a.d0 * b, a.d1 * b, a.d2 * b); }
__device__ __forceinline__ D3 operator*(float a, D3 b) { return mkd(a * b.v, a * b.d0, a * b.d1, a * b.d2); }
__device__ __forceinline__ D3 operator/(float a, D3 b) {
    float inv = 1.0f / b.v, q = a * inv;
    return mkd(q, -q * b.d0 * inv, -q * b.d1 * inv, -q * b.d2 * inv);
}
// scalar helpers overloaded for float / D3
__device__ __forceinline__ float t_val(float a) { return a; }
__device__ __forceinline__ float t_val(D3 a) { return a.v; }
__device__ __forceinline__ float t_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ D3 t_sqrt(D3 a) {
    float s = sqrtf(a.v), h = 0.5f / s;
    return mkd(s, a.d0 * h, a.d1 * h, a.d2 * h);
}
__device__ __forceinline__ float t_atan2(float y, float x) { return det_atan2f(y, x); }
__device__ __forceinline__ D3 t_atan2(D3 y, D3 x) {
    float inv = 1.0f / (x.v * x.v + y.v * y.v);
    return mkd(det_atan2f(y.v, x.v), (x.v * y.d0 - y.v * x.d0) * inv, (x.v * y.d1 - y.v * x.d1) * inv,
               (x.v * y.d2 - y.v * x.d2) * inv);
}
template <typename T> __device__ __forceinline__ T t_select(bool c, T a, T b) { return c ? a : b; }
__device__ __forceinline__ D3 t_const(D3, float c) { return mkd(c, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float t_const(float, float c) { return c; }

template <typename T> struct J6 { T ux, uy, uz, vx, vy, vz; };   // rows (u, v) x columns (x, y, z)

// ---- KB4 Jacobian (kannala_brandt_4.rs:56-152)
template <typename T>
__device__ __forceinline__ J6<T> jac_kb4(T x, T y, T z, float fx, float fy, const float *k) {
    T inv_z = 1.0f / z;
    T x2 = x * x, y2 = y * y, xy = x * y;
    T r2 = x2 + y2;
    T r = t_sqrt(r2);
    T inv_r = 1.0f / r;
    T inv_r3 = inv_r * inv_r * inv_r;
    T rho2 = r2 + z * z;
    T inv_rho2 = 1.0f / rho2;
    T inv_rho2_r = inv_rho2 * inv_r;
    T theta = t_atan2(r, z);
    T theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta4 * theta4;
    T d = theta * (1.0f + k[0] * theta2 + k[1] * theta4 + k[2] * theta6 + k[3] * theta8);
    T dd_dtheta = 1.0f + 3.0f * k[0] * theta2 + 5.0f * k[1] * theta4 + 7.0f * k[2] * theta6 + 9.0f * k[3] * theta8;
    T dth_dx = x * z * inv_rho2_r, dth_dy = y * z * inv_rho2_r, dth_dz = -r * inv_rho2;
    T dd_dx = dd_dtheta * dth_dx, dd_dy = dd_dtheta * dth_dy, dd_dz = dd_dtheta * dth_dz;
    T xr = x * inv_r;
    T dxr_dx = y2 * inv_r3, dxr_dy = -xy * inv_r3;
    T du_dx = fx * (dd_dx * xr + d * dxr_dx);
    T du_dy = fx * (dd_dy * xr + d * dxr_dy);
    T du_dz = fx * (dd_dz * xr);
    T yr = y * inv_r;
    T dyr_dx = -xy * inv_r3, dyr_dy = x2 * inv_r3;
    T dv_dx = fy * (dd_dx * yr + d * dyr_dx);
    T dv_dy = fy * (dd_dy * yr + d * dyr_dy);
    T dv_dz = fy * (dd_dz * yr);
    const bool near_axis = t_val(r) < 1e-6f;
    T dx = fx * inv_z, dy = fy * inv_z;
    T zero = t_const(x, 0.0f);
    J6<T> j;
    j.ux = t_select(near_axis, dx, du_dx);       j.vx = t_select(near_axis, zero, dv_dx);
    j.uy = t_select(near_axis, zero, du_dy);     j.vy = t_select(near_axis, dy, dv_dy);
    j.uz = t_select(near_axis, -dx * x * inv_z, du_dz);
    j.vz = t_select(near_axis, -dy * y * inv_z, dv_dz);
    return j;
}

// thin-prism polynomials (thin_prism_fisheye.rs:37-61); c: k1..k4, p1, p2, sx1, sy1
template <typename T>
__device__ __forceinline__ void tpf_polys(T x, T y, const float *c, T &nu, T &nv, T &dnu_dx, T &dnu_dy, T &dnv_dx, T &dnv_dy) {
    const float p1 = c[4], p2 = c[5], sx1 = c[6], sy1 = c[7];
    T x2 = x * x, y2 = y * y, xy = x * y, r2 = x2 + y2;
    nu = 2.0f * p1 * xy + p2 * (3.0f * x2 + y2) + sx1 * r2;
    nv = 2.0f * p2 * xy + p1 * (x2 + 3.0f * y2) + sy1 * r2;
    dnu_dx = 2.0f * (p1 * y + (3.0f * p2 + sx1) * x);
    dnu_dy = 2.0f * (p1 * x + (p2 + sx1) * y);
    dnv_dx = 2.0f * (p2 * y + (p1 + sy1) * x);
    dnv_dy = 2.0f * (p2 * x + (3.0f * p1 + sy1) * y);
}
// thin_prism_fisheye.rs:82-118
template <typename T>
__device__ __forceinline__ J6<T> jac_tpf(T x, T y, T z, float fx, float fy, const float *c) {
    J6<T> kj = jac_kb4<T>(x, y, z, fx, fy, c);
    T inv_z = 1.0f / z;
    T inv_z2 = inv_z * inv_z, inv_z3 = inv_z2 * inv_z;
    T nu, nv, dnu_dx, dnu_dy, dnv_dx, dnv_dy;
    tpf_polys<T>(x, y, c, nu, nv, dnu_dx, dnu_dy, dnv_dx, dnv_dy);
    J6<T> j;
    j.ux = kj.ux + fx * dnu_dx * inv_z2;  j.vx = kj.vx + fy * dnv_dx * inv_z2;
    j.uy = kj.uy + fx * dnu_dy * inv_z2;  j.vy = kj.vy + fy * dnv_dy * inv_z2;
    j.uz = kj.uz + -2.0f * fx * nu * inv_z3;
    j.vz = kj.vz + -2.0f * fy * nv * inv_z3;
    return j;
}

// RT8 Jacobian at the clamp surrogate point (radial_tangential_8.rs:66-142): x_n, y_n are the (clamped)
// normalised coordinates fed to the distortion, (xc, yc) = (x_n, y_n) z.  The VJP differentiates the same
// expression w.r.t. (xc, yc, z) with x_n = xc / z (:186-188).
template <typename T>
__device__ __forceinline__ J6<T> jac_rt8_core(T x_n, T y_n, T xc, T yc, T inv_z, float fx, float fy, const float *c) {
    const float k1 = c[0], k2 = c[1], k3 = c[2], k4 = c[3], k5 = c[4], k6 = c[5], p1 = c[6], p2 = c[7];
    T inv_z2 = inv_z * inv_z;
    T r2 = x_n * x_n + y_n * y_n, r4 = r2 * r2, r6 = r4 * r2;
    T n_poly = 1.0f + k1 * r2 + k2 * r4 + k3 * r6;
    T dn_poly = 1.0f + k4 * r2 + k5 * r4 + k6 * r6;
    T np_poly = k1 + 2.0f * k2 * r2 + 3.0f * k3 * r4;
    T dnp_poly = k4 + 2.0f * k5 * r2 + 3.0f * k6 * r4;
    T inv_dn = 1.0f / dn_poly;
    T inv_dn2 = inv_dn * inv_dn;
    T r_val = n_poly * inv_dn;
    T rp_val = (np_poly * dn_poly - n_poly * dnp_poly) * inv_dn2;
    T d00 = r_val + 2.0f * x_n * x_n * rp_val + 2.0f * p1 * y_n + 6.0f * p2 * x_n;
    T d01 = 2.0f * x_n * y_n * rp_val + 2.0f * p1 * x_n + 2.0f * p2 * y_n;
    T d10 = d01;
    T d11 = r_val + 2.0f * y_n * y_n * rp_val + 6.0f * p1 * y_n + 2.0f * p2 * x_n;
    J6<T> j;
    j.ux = fx * d00 * inv_z;  j.vx = fy * d10 * inv_z;
    j.uy = fx * d01 * inv_z;  j.vy = fy * d11 * inv_z;
    j.uz = -fx * (d00 * xc + d01 * yc) * inv_z2;
    j.vz = -fy * (d10 * xc + d11 * yc) * inv_z2;
    return j;
}

__device__ __forceinline__ M23 to_m23(const J6<float> &j) {
    M23 m;
    m.c0 = mk2(j.ux, j.vx); m.c1 = mk2(j.uy, j.vy); m.c2 = mk2(j.uz, j.vz);
    return m;
}

// calculate_project_jacobian for the distorted models (mod.rs:59-78)
__device__ __forceinline__ M23 jacobian_distorted(V3 p, const BgCamera &u) {
    switch (u.camera_model) {
        case BG_CAMERA_KANNALA_BRANDT_4: return to_m23(jac_kb4<float>(p.x, p.y, p.z, u.fx, u.fy, u.model_params));
        case BG_CAMERA_THIN_PRISM_FISHEYE: return to_m23(jac_tpf<float>(p.x, p.y, p.z, u.fx, u.fy, u.model_params));
        default: {  // radial-tangential 8
            float inv_z = 1.0f / p.z;
            float x_n = clampf(p.x * inv_z, u.lim_neg_x, u.lim_pos_x);
            float y_n = clampf(p.y * inv_z, u.lim_neg_y, u.lim_pos_y);
            return to_m23(jac_rt8_core<float>(x_n, y_n, x_n * p.z, y_n * p.z, inv_z, u.fx, u.fy, u.model_params));
        }
    }
}

// project for the distorted models (mod.rs:41-56)
__device__ __forceinline__ void project_kb4(V3 p, const BgCamera &u, float &ox, float &oy) {
    const float *k = u.model_params;
    float x = p.x, y = p.y, z = p.z;
    float inv_z = 1.0f / z;
    float pinhole_u = u.fx * x * inv_z + u.cx;
    float pinhole_v = u.fy * y * inv_z + u.cy;
    float r = sqrtf(x * x + y * y);
    float theta = det_atan2f(r, z);
    float theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta2 * theta4, theta8 = theta4 * theta4;
    float d = theta * (1.0f + k[0] * theta2 + k[1] * theta4 + k[2] * theta6 + k[3] * theta8);
    float inv_r = 1.0f / r;
    float fisheye_u = u.fx * (d * x * inv_r) + u.cx;
    float fisheye_v = u.fy * (d * y * inv_r) + u.cy;
    bool near_axis = r < 1e-6f;
    ox = near_axis ? pinhole_u : fisheye_u;
    oy = near_axis ? pinhole_v : fisheye_v;
}
__device__ __forceinline__ void project_distorted(V3 p, const BgCamera &u, float &ox, float &oy) {
    const float *c = u.model_params;
    switch (u.camera_model) {
        case BG_CAMERA_KANNALA_BRANDT_4: project_kb4(p, u, ox, oy); break;
        case BG_CAMERA_THIN_PRISM_FISHEYE: {
            float uk, vk;
            project_kb4(p, u, uk, vk);
            float inv_z = 1.0f / p.z;
            float inv_z2 = inv_z * inv_z;
            float nu, nv, a, b, cc, d;
            tpf_polys<float>(p.x, p.y, c, nu, nv, a, b, cc, d);
            ox = uk + u.fx * nu * inv_z2;
            oy = vk + u.fy * nv * inv_z2;
            break;
        }
        default: {  // radial_tangential_8.rs:23-64
            const float k1 = c[0], k2 = c[1], k3 = c[2], k4 = c[3], k5 = c[4], k6 = c[5], p1 = c[6], p2 = c[7];
            float x_ = p.x / p.z, y_ = p.y / p.z;
            float x_2 = x_ * x_, y_2 = y_ * y_;
            float r2 = x_2 + y_2, r4 = r2 * r2, r6 = r4 * r2;
            float d = (1.0f + k1 * r2 + k2 * r4 + k3 * r6) / (1.0f + k4 * r2 + k5 * r4 + k6 * r6);
            float x_y_ = x_ * y_;
            float x__ = x_ * d + 2.0f * p1 * x_y_ + p2 * (r2 + 2.0f * x_2);
            float y__ = y_ * d + 2.0f * p2 * x_y_ + p1 * (r2 + 2.0f * y_2);
            ox = u.fx * x__ + u.cx;
            oy = u.fy * y__ + u.cy;
        }
    }
}

// Angular cull of the distorted models (project_forward.rs:53-61)
__device__ __forceinline__ bool in_front_distorted(V3 mean_c, const BgCamera &u) {
    float r = sqrtf(mean_c.x * mean_c.x + mean_c.y * mean_c.y);
    float theta = det_atan2f(r, mean_c.z);
    return !(theta > u.half_max_render_fov);
}

// calculate_projection_vjp for the distorted models (mod.rs:84-136): gradient w.r.t. the camera-space mean
// given the gradients w.r.t. mean2d and cov2d.
//   path 1: J^T v_mean2d;  path 2: sum_ij vJ[i][j] dJ[i][j]/d mean, vJ = 2 v_cov2d J cov_c.
// RT8 routes both paths through the Jacobian clamp exactly as the reference does (:244-262, :360-374):
// S = d(xc, yc, z)/d(mean), J_eff = J_surr S, vJ_surr = vJ_eff S^T.
__device__ __forceinline__ V3 projection_vjp_distorted(M23 jac, V3 mean_c, S3 cov_c, const BgCamera &u, S2 v_cov2d, V2 v_mean2d) {
    const float *c = u.model_params;
    const float mx = mean_c.x, my = mean_c.y, mz = mean_c.z;
    if (u.camera_model != BG_CAMERA_RADIAL_TANGENTIAL_8) {
        const D3 X = mkd(mx, 1.f, 0.f, 0.f), Y = mkd(my, 0.f, 1.f, 0.f), Z = mkd(mz, 0.f, 0.f, 1.f);
        const J6<D3> dj = (u.camera_model == BG_CAMERA_KANNALA_BRANDT_4) ? jac_kb4<D3>(X, Y, Z, u.fx, u.fy, c)
                                                                          : jac_tpf<D3>(X, Y, Z, u.fx, u.fy, c);
        M23 tmp = mul(v_cov2d, jac);
        const V3 vju = scale(mk3(dot(row0(tmp), row0(cov_c)), dot(row0(tmp), row1(cov_c)), dot(row0(tmp), row2(cov_c))), 2.0f);
        const V3 vjv = scale(mk3(dot(row1(tmp), row0(cov_c)), dot(row1(tmp), row1(cov_c)), dot(row1(tmp), row2(cov_c))), 2.0f);
        float v_mx = dot(v_mean2d, jac.c0), v_my = dot(v_mean2d, jac.c1), v_mz = dot(v_mean2d, jac.c2);
        v_mx += vju.x * dj.ux.d0 + vju.y * dj.uy.d0 + vju.z * dj.uz.d0 + vjv.x * dj.vx.d0 + vjv.y * dj.vy.d0 + vjv.z * dj.vz.d0;
        v_my += vju.x * dj.ux.d1 + vju.y * dj.uy.d1 + vju.z * dj.uz.d1 + vjv.x * dj.vx.d1 + vjv.y * dj.vy.d1 + vjv.z * dj.vz.d1;
        v_mz += vju.x * dj.ux.d2 + vju.y * dj.uy.d2 + vju.z * dj.uz.d2 + vjv.x * dj.vx.d2 + vjv.y * dj.vy.d2 + vjv.z * dj.vz.d2;
        return mk3(v_mx, v_my, v_mz);
    }
    const float inv_z = 1.0f / mz;
    const float mx_raw = mx * inv_z, my_raw = my * inv_z;
    const float mx_rz = clampf(mx_raw, u.lim_neg_x, u.lim_pos_x);
    const float my_rz = clampf(my_raw, u.lim_neg_y, u.lim_pos_y);
    const bool in_x = mx_raw <= u.lim_pos_x && mx_raw >= u.lim_neg_x;
    const bool in_y = my_raw <= u.lim_pos_y && my_raw >= u.lim_neg_y;
    const D3 XC = mkd(mx_rz * mz, 1.f, 0.f, 0.f), YC = mkd(my_rz * mz, 0.f, 1.f, 0.f), Z = mkd(mz, 0.f, 0.f, 1.f);
    const D3 IZ = 1.0f / Z;
    const J6<D3> js = jac_rt8_core<D3>(XC * IZ, YC * IZ, XC, YC, IZ, u.fx, u.fy, c);
    // J_eff = J_surr S
    const float je00 = in_x ? js.ux.v : 0.0f, je10 = in_x ? js.vx.v : 0.0f;
    const float je01 = in_y ? js.uy.v : 0.0f, je11 = in_y ? js.vy.v : 0.0f;
    const float je02 = (in_x ? 0.0f : mx_rz * js.ux.v) + (in_y ? 0.0f : my_rz * js.uy.v) + js.uz.v;
    const float je12 = (in_x ? 0.0f : mx_rz * js.vx.v) + (in_y ? 0.0f : my_rz * js.vy.v) + js.vz.v;
    float v_mx = je00 * v_mean2d.x + je10 * v_mean2d.y;
    float v_my = je01 * v_mean2d.x + je11 * v_mean2d.y;
    float v_mz = je02 * v_mean2d.x + je12 * v_mean2d.y;
    M23 je;
    je.c0 = mk2(je00, je10); je.c1 = mk2(je01, je11); je.c2 = mk2(je02, je12);
    M23 tmp = mul(v_cov2d, je);
    const V3 veu = scale(mk3(dot(row0(tmp), row0(cov_c)), dot(row0(tmp), row1(cov_c)), dot(row0(tmp), row2(cov_c))), 2.0f);
    const V3 vev = scale(mk3(dot(row1(tmp), row0(cov_c)), dot(row1(tmp), row1(cov_c)), dot(row1(tmp), row2(cov_c))), 2.0f);
    const float vs_u0 = in_x ? veu.x : mx_rz * veu.z, vs_v0 = in_x ? vev.x : mx_rz * vev.z;
    const float vs_u1 = in_y ? veu.y : my_rz * veu.z, vs_v1 = in_y ? vev.y : my_rz * vev.z;
    const float vs_u2 = veu.z, vs_v2 = vev.z;
    const float c_xc = vs_u0 * js.ux.d0 + vs_u1 * js.uy.d0 + vs_u2 * js.uz.d0 + vs_v0 * js.vx.d0 + vs_v1 * js.vy.d0 + vs_v2 * js.vz.d0;
    const float c_yc = vs_u0 * js.ux.d1 + vs_u1 * js.uy.d1 + vs_u2 * js.uz.d1 + vs_v0 * js.vx.d1 + vs_v1 * js.vy.d1 + vs_v2 * js.vz.d1;
    const float c_z = vs_u0 * js.ux.d2 + vs_u1 * js.uy.d2 + vs_u2 * js.uz.d2 + vs_v0 * js.vx.d2 + vs_v1 * js.vy.d2 + vs_v2 * js.vz.d2;
    if (in_x) v_mx += c_xc;
    if (in_y) v_my += c_yc;
    v_mz += c_z;
    if (!in_x) v_mz += mx_rz * c_xc;
    if (!in_y) v_mz += my_rz * c_yc;
    return mk3(v_mx, v_my, v_mz);
}

}  // namespace bg
