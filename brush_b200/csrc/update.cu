// update.cu -- the parameter update of one training step as ONE pass over the Gaussians.
// Replaces, per step (brush-train/src/train.rs:280-416):
//   AdamScaled::step on transforms [N,10] (per-column LR), SH coefficients [N,K,3] (per-band LR, second moment =
//   row mean of g^2) and raw opacity [N]                               (adam_scaled.rs:75-165, train.rs:328-381)
//   RefineRecord::gather_stats (MAX refine weight, SUM visible, MAX radius)            (stats.rs:40-50)
//   the mean noise  means += clamp(N(0,1) (1-sigmoid(opac'))^150 vis lr 50, +-median)    (train.rs:389-416)
// The reference issues ~80 generic tensor ops for this; round 1 used five kernels (3 Adam, noise draw, stats+noise).
// One thread owns one Gaussian: every row is read and written once with 128/256-bit accesses (consecutive threads own
// consecutive rows, so a warp's accesses cover one contiguous span and every fetched sector is used), the normal
// draws are evaluated in registers (counter-based Philox, bg_rng.cuh; only Gaussians whose noise weight is non-zero
// draw at all), nothing but the parameters, moments and the refine record touches HBM:
//   (316 + 48 K + g) N bytes, g = 44 + 12 K (dense gradient) or 48 + 12 views (factored).
//
// Multi-view steps: the SH gradient of one view is rank one per Gaussian,
// v_sh[g,k,:] = Y_k(dir(mean_g, camera_v)) v_color_v[g,:]  (kernels/sh.rs:265-355), so with FACTORED the kernel takes
// the views' colour gradients (all-gathered records) and forms (1/views) sum_v Y(dir_v) v_color_v in registers -- the
// dense [N,K,3] gradient is never written or read -- and reduces the MAX statistics over the records on the way.
// View order is the global view index, so every data-parallel rank computes bit-identical updates.
//
// Compiled with -fmad=false: plain IEEE multiply / add / divide / sqrt in the order written here (the order of
// AdamScaled::step), so the update is a pure function of its inputs on any IEEE machine.
#include <algorithm>

#include "bg_common.cuh"
#include "bg_rng.cuh"
#include "bg_sh.cuh"
#include "bg_update.cuh"

namespace bg {

__device__ __forceinline__ float adam_m(float m, float g, const UpdateParams &P) { return P.first ? g * P.f1 : m * P.beta1 + g * P.f1; }
__device__ __forceinline__ float adam_v(float v, float gsq, const UpdateParams &P) { return P.first ? gsq * P.f2 : v * P.beta2 + gsq * P.f2; }
__device__ __forceinline__ float adam_p(float p, float m, float v, float step, const UpdateParams &P) {
    const float m_hat = m / P.bc1, v_hat = v / P.bc2;
    return p - (m_hat / (sqrtf(v_hat) + P.eps)) * step;
}

__device__ __forceinline__ void ld256(const float *p, float *o) {   // plain (read-write data) 256-bit load
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7])
                 : "l"(p) : "memory");
}

constexpr int UP_THREADS = 128;

template <int DEG, bool FACTORED>
__global__ void __launch_bounds__(UP_THREADS)
train_update_kernel(const UpdateParams P) {
    constexpr int K = (DEG + 1) * (DEG + 1);
    constexpr int KF = K * 3;
    const uint32_t j = blockIdx.x * UP_THREADS + threadIdx.x;   // index inside this launch's slice
    if (j >= P.count) return;
    const uint32_t i = P.g_begin + j;

    // ---- transforms row: Adam with per-column learning rates (train.rs:328-350)
    float p[10], old_mean[3];
    {
        float g[10], m[10], v[10];
        const float2 *p2 = reinterpret_cast<const float2 *>(P.transforms + (size_t)i * 10);
        const float2 *g2 = reinterpret_cast<const float2 *>(P.g_t + (size_t)i * 10);
        float2 *m2 = reinterpret_cast<float2 *>(P.m_t + (size_t)i * 10);
        float2 *v2 = reinterpret_cast<float2 *>(P.v_t + (size_t)i * 10);
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const float2 a = p2[q], b = __ldg(g2 + q);
            p[2 * q] = a.x; p[2 * q + 1] = a.y; g[2 * q] = b.x; g[2 * q + 1] = b.y;
            if (!P.first) { const float2 c = m2[q], d = v2[q]; m[2 * q] = c.x; m[2 * q + 1] = c.y; v[2 * q] = d.x; v[2 * q + 1] = d.y; }
            else { m[2 * q] = m[2 * q + 1] = v[2 * q] = v[2 * q + 1] = 0.0f; }
        }
        old_mean[0] = p[0]; old_mean[1] = p[1]; old_mean[2] = p[2];
#pragma unroll
        for (int c = 0; c < 10; c++) {
            const float gg = P.grad_scale == 1.0f ? g[c] : g[c] * P.grad_scale;
            m[c] = adam_m(m[c], gg, P);
            v[c] = adam_v(v[c], gg * gg, P);
            p[c] = adam_p(p[c], m[c], v[c], P.lr_t[c], P);
        }
#pragma unroll
        for (int q = 0; q < 5; q++) { m2[q] = make_float2(m[2 * q], m[2 * q + 1]); v2[q] = make_float2(v[2 * q], v[2 * q + 1]); }
    }
    // ---- raw opacity
    float raw;
    {
        const float g0 = __ldg(P.g_o + i);
        const float gg = P.grad_scale == 1.0f ? g0 : g0 * P.grad_scale;
        const float mm = adam_m(P.first ? 0.0f : P.m_o[i], gg, P);
        const float vv = adam_v(P.first ? 0.0f : P.v_o[i], gg * gg, P);
        raw = adam_p(P.raw_opac[i], mm, vv, P.lr_opac, P);
        P.m_o[i] = mm; P.v_o[i] = vv; P.raw_opac[i] = raw;
    }
    // ---- refine statistics of the step (stats.rs:40-50): MAX over the views, SUM of the visibility counts
    const float vis = __ldg(P.visible + i);
    {
        float vr, rad;
        if (FACTORED) {
            vr = __ldg(P.refine_all + j); rad = __ldg(P.radius_all + j);
            for (uint32_t r = 1; r < P.world; r++) {
                vr = fmaxf(vr, __ldg(P.refine_all + (size_t)r * P.count + j));
                rad = fmaxf(rad, __ldg(P.radius_all + (size_t)r * P.count + j));
            }
        } else {
            vr = __ldg(P.v_refine + i); rad = __ldg(P.max_radius + i);
        }
        P.refine_norm[i] = fmaxf(vr, P.refine_norm[i]);
        P.vis_weight[i] = P.vis_weight[i] + vis;
        P.max_screen[i] = fmaxf(rad, P.max_screen[i]);
    }
    // ---- mean noise on the updated means, gated by the updated opacity (train.rs:389-416)
    if (P.noisy) {
        const float opac = 1.0f / (1.0f + expf(-raw));
        const float wgt = fminf(fmaxf(powf(1.0f - opac, 150.0f), 0.0f), 1.0f) * (vis > 0.0f ? 1.0f : 0.0f);
        const float wm = wgt * P.noise_scale;
        if (wm != 0.0f) {
            const unsigned long long e0 = 3ull * i;
            const uint32_t off = (uint32_t)(e0 & 3ull);   // elements 3i..3i+2 of the stream: quad e0/4, spilling into the next
            float z[8];
            normal_quad(P.seed, P.noise_offset + e0 / 4, z);
            if (off > 1u) normal_quad(P.seed, P.noise_offset + e0 / 4 + 1, z + 4);
            else z[4] = z[5] = z[6] = z[7] = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float zc = off == 0u ? z[c] : (off == 1u ? z[c + 1] : (off == 2u ? z[c + 2] : z[c + 3]));
                p[c] += fminf(fmaxf(zc * wm, -P.median_scale), P.median_scale);
            }
        }
    }
    {
        float2 *p2 = reinterpret_cast<float2 *>(P.transforms + (size_t)i * 10);
#pragma unroll
        for (int q = 0; q < 5; q++) p2[q] = make_float2(p[2 * q], p[2 * q + 1]);
    }

    // ---- SH coefficients: gradient row (dense, or rebuilt from the views), row-mean second moment, per-band LR
    float g[KF];
    if (FACTORED) {
#pragma unroll
        for (int c = 0; c < KF; c++) g[c] = 0.0f;
        for (uint32_t v = 0; v < P.views; v++) {
            const uint32_t r = v / P.local, li = v - r * P.local;
            const float *vc = P.colours + (((size_t)li * P.world + r) * P.count + j) * 3;
            const float cr = __ldg(vc), cg = __ldg(vc + 1), cb = __ldg(vc + 2);
            if (cr == 0.0f && cg == 0.0f && cb == 0.0f) continue;
            const float4 cp = __ldg(reinterpret_cast<const float4 *>(P.cam_all) + v);
            const V3 u_world = sub(mk3(old_mean[0], old_mean[1], old_mean[2]), mk3(cp.x, cp.y, cp.z));
            const V3 vdir = scale(u_world, 1.0f / length(u_world));
            float Y[K];
            sh_basis<DEG>(vdir, Y);
#pragma unroll
            for (int k = 0; k < K; k++) {
                g[3 * k] += cr * Y[k];
                g[3 * k + 1] += cg * Y[k];
                g[3 * k + 2] += cb * Y[k];
            }
        }
#pragma unroll
        for (int c = 0; c < KF; c++) g[c] = g[c] * P.sh_grad_scale;
    } else {
        const float *src = P.g_sh + (size_t)i * KF;
        if ((KF % 8) == 0 && (reinterpret_cast<uintptr_t>(P.g_sh) & 31u) == 0) {
#pragma unroll
            for (int q = 0; q < KF / 8; q++) ldg256(src + 8 * q, g + 8 * q);
        } else if ((KF % 4) == 0) {
#pragma unroll
            for (int q = 0; q < KF / 4; q++) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(src) + q);
                g[4 * q] = t.x; g[4 * q + 1] = t.y; g[4 * q + 2] = t.z; g[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < KF; c++) g[c] = __ldg(src + c);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < KF; c++) s += g[c] * g[c];
    const float mean_sq = s / (float)KF;
    const float vv = adam_v(P.first ? 0.0f : P.v_sh[i], mean_sq, P);
    P.v_sh[i] = vv;
    float *ps = P.sh + (size_t)i * KF, *ms = P.m_sh + (size_t)i * KF;
    const bool a32 = ((reinterpret_cast<uintptr_t>(P.sh) | reinterpret_cast<uintptr_t>(P.m_sh)) & 31u) == 0;
    if ((KF % 8) == 0 && a32) {
#pragma unroll
        for (int q = 0; q < KF / 8; q++) {
            float pp[8], mm[8];
            ld256(ps + 8 * q, pp);
            if (!P.first) ld256(ms + 8 * q, mm);
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int c = 8 * q + e;
                mm[e] = adam_m(P.first ? 0.0f : mm[e], g[c], P);
                pp[e] = adam_p(pp[e], mm[e], vv, c < 3 ? P.lr_sh_dc : P.lr_sh_rest, P);
            }
            stg256(ms + 8 * q, mm);
            stg256(ps + 8 * q, pp);
        }
    } else if ((KF % 4) == 0) {
#pragma unroll
        for (int q = 0; q < KF / 4; q++) {
            float4 p4 = reinterpret_cast<float4 *>(ps)[q];
            float4 m4 = P.first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4 *>(ms)[q];
            float *pp = &p4.x, *mm = &m4.x;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int c = 4 * q + e;
                mm[e] = adam_m(mm[e], g[c], P);
                pp[e] = adam_p(pp[e], mm[e], vv, c < 3 ? P.lr_sh_dc : P.lr_sh_rest, P);
            }
            reinterpret_cast<float4 *>(ms)[q] = m4;
            reinterpret_cast<float4 *>(ps)[q] = p4;
        }
    } else {
#pragma unroll
        for (int c = 0; c < KF; c++) {
            const float mm = adam_m(P.first ? 0.0f : ms[c], g[c], P);
            ms[c] = mm;
            ps[c] = adam_p(ps[c], mm, vv, c < 3 ? P.lr_sh_dc : P.lr_sh_rest, P);
        }
    }
}

template <int DEG>
static cudaError_t launch_deg(cudaStream_t s, const UpdateParams &P, bool factored) {
    const unsigned grid = (P.count + UP_THREADS - 1) / UP_THREADS;
    if (factored) train_update_kernel<DEG, true><<<grid, UP_THREADS, 0, s>>>(P);
    else train_update_kernel<DEG, false><<<grid, UP_THREADS, 0, s>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_train_update(cudaStream_t s, int deg, const UpdateParams &P, bool factored) {
    if (P.count == 0) return cudaSuccess;
    switch (deg) {
        case 0: return launch_deg<0>(s, P, factored);
        case 1: return launch_deg<1>(s, P, factored);
        case 2: return launch_deg<2>(s, P, factored);
        case 3: return launch_deg<3>(s, P, factored);
        case 4: return launch_deg<4>(s, P, factored);
    }
    return cudaErrorInvalidValue;
}

}  // namespace bg
