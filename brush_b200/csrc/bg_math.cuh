// bg_math.cuh -- device-side f32 math for the per-Gaussian (projection) kernels.
//
// Semantics follow the reference's device math types,
//   crates/brush-cube/src/lib.rs:34-578 (Vec3A, Quat, Mat3, Mat2x3, Sym2, Sym3, sigmoid, calc_sigma)
// with the reference's operation order.  Translation units that include this header for the
// projection stage are compiled with -fmad=false: the only fused operations are the explicit
// __fmaf_rn calls in det_expf / det_logf, whose recipe is fixed in DESIGN.md ("Deterministic
// exp/log").  With IEEE div/sqrt (nvcc defaults) the whole per-Gaussian stage is then a pure
// function of its inputs, identical on any IEEE-754 machine -- which is what lets the parity
// tests demand bit-exact culling, tile counts and projected rows.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bg {

__device__ __forceinline__ bool is_finite(float x) { return ((__float_as_uint(x) >> 23) & 0xFFu) != 0xFFu; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float pow2i(int k) { return __uint_as_float((uint32_t)(k + 127) << 23); }

// exp: n = rint(x*log2e); r = x - n*ln2 (two fused steps); e^r = 1 + r + r^2 P(r); scale in two halves.
__device__ __forceinline__ float det_expf(float x) {
    if (x != x) return x;
    if (x > 88.72283935546875f) return __int_as_float(0x7f800000);
    if (x < -103.97208404541015625f) return 0.0f;
    float t = x * 1.44269502162933349609375f;
    float n = rintf(t);
    float r = __fmaf_rn(n, -0.693359375f, x);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    float r2 = __fmul_rn(r, r);
    float e = __fadd_rn(__fmaf_rn(p, r2, r), 1.0f);
    int ni = (int)n;
    int n1 = ni / 2;
    int n2 = ni - n1;
    return __fmul_rn(__fmul_rn(e, pow2i(n1)), pow2i(n2));
}

__device__ __forceinline__ float det_logf(float x) {
    if (x != x || x < 0.0f) return __int_as_float(0x7fc00000);
    if (x == 0.0f) return __int_as_float(0xff800000);
    if (!is_finite(x)) return x;
    uint32_t u = __float_as_uint(x);
    int e = 0;
    if ((u >> 23) == 0u) {
        x = __fmul_rn(x, 8388608.0f);
        u = __float_as_uint(x);
        e = -23;
    }
    e += (int)(u >> 23) - 126;
    float m = __uint_as_float((u & 0x007FFFFFu) | 0x3F000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = __fadd_rn(__fadd_rn(m, m), -1.0f); } else { m = __fadd_rn(m, -1.0f); }
    float z = __fmul_rn(m, m);
    float p = 7.0376836292e-2f;
    p = __fmaf_rn(p, m, -1.1514610310e-1f);
    p = __fmaf_rn(p, m, 1.1676998740e-1f);
    p = __fmaf_rn(p, m, -1.2420140846e-1f);
    p = __fmaf_rn(p, m, 1.4249322787e-1f);
    p = __fmaf_rn(p, m, -1.6668057665e-1f);
    p = __fmaf_rn(p, m, 2.0000714765e-1f);
    p = __fmaf_rn(p, m, -2.4999993993e-1f);
    p = __fmaf_rn(p, m, 3.3333331174e-1f);
    float y = __fmul_rn(__fmul_rn(m, z), p);
    float fe = (float)e;
    y = __fmaf_rn(fe, -2.12194440e-4f, y);
    y = __fmaf_rn(z, -0.5f, y);
    float r = __fadd_rn(m, y);
    r = __fmaf_rn(fe, 0.693359375f, r);
    return r;
}

__device__ __forceinline__ float det_sigmoid(float x) { return 1.0f / (1.0f + det_expf(-x)); }

struct V3 { float x, y, z; };
struct V2 { float x, y; };
struct Q4 { float w, x, y, z; };
struct M3 { V3 c0, c1, c2; };          // column major
struct M23 { V2 c0, c1, c2; };         // 2x3, column major
struct S2 { float c00, c01, c11; };    // symmetric 2x2
struct S3 { float c00, c01, c02, c11, c12, c22; };

__device__ __forceinline__ V3 mk3(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ V2 mk2(float x, float y) { V2 v; v.x = x; v.y = y; return v; }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 scale(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + 0.0f; }
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 normalize(V3 a) { return scale(a, 1.0f / length(a)); }
__device__ __forceinline__ bool is_finite(V3 a) { return is_finite(a.x) && is_finite(a.y) && is_finite(a.z); }
__device__ __forceinline__ V2 add(V2 a, V2 b) { return mk2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 scale(V2 a, float s) { return mk2(a.x * s, a.y * s); }
__device__ __forceinline__ float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float dot(Q4 a, Q4 b) { return ((a.w * b.w + a.x * b.x) + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ Q4 scale(Q4 a, float s) { Q4 q; q.w = a.w * s; q.x = a.x * s; q.y = a.y * s; q.z = a.z * s; return q; }
__device__ __forceinline__ Q4 normalize(Q4 a) { return scale(a, 1.0f / sqrtf(dot(a, a))); }

__device__ __forceinline__ M3 quat_to_mat3(Q4 q) {
    float w = q.w, qx = q.x, qy = q.y, qz = q.z;
    float x2 = qx * qx, y2 = qy * qy, z2 = qz * qz;
    float xy = qx * qy, xz = qx * qz, yz = qy * qz;
    float wx = w * qx, wy = w * qy, wz = w * qz;
    M3 m;
    m.c0 = mk3(1.0f - 2.0f * (y2 + z2), 2.0f * (xy + wz), 2.0f * (xz - wy));
    m.c1 = mk3(2.0f * (xy - wz), 1.0f - 2.0f * (x2 + z2), 2.0f * (yz + wx));
    m.c2 = mk3(2.0f * (xz + wy), 2.0f * (yz - wx), 1.0f - 2.0f * (x2 + y2));
    return m;
}
__device__ __forceinline__ V3 mul(M3 m, V3 v) { return add(add(scale(m.c0, v.x), scale(m.c1, v.y)), scale(m.c2, v.z)); }
__device__ __forceinline__ V3 tmul(M3 m, V3 v) { return mk3(dot(m.c0, v), dot(m.c1, v), dot(m.c2, v)); }
__device__ __forceinline__ M3 mul(M3 m, M3 n) { M3 r; r.c0 = mul(m, n.c0); r.c1 = mul(m, n.c1); r.c2 = mul(m, n.c2); return r; }
__device__ __forceinline__ M3 mul_diag(M3 m, V3 s) { M3 r; r.c0 = scale(m.c0, s.x); r.c1 = scale(m.c1, s.y); r.c2 = scale(m.c2, s.z); return r; }
__device__ __forceinline__ V3 row0(M3 m) { return mk3(m.c0.x, m.c1.x, m.c2.x); }
__device__ __forceinline__ V3 row1(M3 m) { return mk3(m.c0.y, m.c1.y, m.c2.y); }
__device__ __forceinline__ V3 row2(M3 m) { return mk3(m.c0.z, m.c1.z, m.c2.z); }
__device__ __forceinline__ S3 outer_self(M3 m) {
    V3 r0 = row0(m), r1 = row1(m), r2 = row2(m);
    S3 s; s.c00 = dot(r0, r0); s.c01 = dot(r0, r1); s.c02 = dot(r0, r2); s.c11 = dot(r1, r1); s.c12 = dot(r1, r2); s.c22 = dot(r2, r2);
    return s;
}
__device__ __forceinline__ V2 mul(M23 m, V3 v) { return add(add(scale(m.c0, v.x), scale(m.c1, v.y)), scale(m.c2, v.z)); }
__device__ __forceinline__ M23 mul(M23 m, M3 n) { M23 r; r.c0 = mul(m, n.c0); r.c1 = mul(m, n.c1); r.c2 = mul(m, n.c2); return r; }
__device__ __forceinline__ V3 row0(M23 m) { return mk3(m.c0.x, m.c1.x, m.c2.x); }
__device__ __forceinline__ V3 row1(M23 m) { return mk3(m.c0.y, m.c1.y, m.c2.y); }
__device__ __forceinline__ S2 gram(M23 m) {
    S2 s;
    s.c00 = m.c0.x * m.c0.x + m.c1.x * m.c1.x + m.c2.x * m.c2.x;
    s.c01 = m.c0.x * m.c0.y + m.c1.x * m.c1.y + m.c2.x * m.c2.y;
    s.c11 = m.c0.y * m.c0.y + m.c1.y * m.c1.y + m.c2.y * m.c2.y;
    return s;
}
__device__ __forceinline__ V2 mul(S2 s, V2 v) { return add(scale(mk2(s.c00, s.c01), v.x), scale(mk2(s.c01, s.c11), v.y)); }
__device__ __forceinline__ S2 scale(S2 s, float k) { S2 r; r.c00 = s.c00 * k; r.c01 = s.c01 * k; r.c11 = s.c11 * k; return r; }
__device__ __forceinline__ float max_abs(S2 s) { return fmaxf(fmaxf(fabsf(s.c00), fabsf(s.c11)), fabsf(s.c01)); }
__device__ __forceinline__ M23 mul(S2 s, M23 n) { M23 r; r.c0 = mul(s, n.c0); r.c1 = mul(s, n.c1); r.c2 = mul(s, n.c2); return r; }
__device__ __forceinline__ S2 inverse(S2 s) {
    float det = s.c00 * s.c11 - s.c01 * s.c01;
    float inv_det = (det > 0.0f) ? 1.0f / det : 0.0f;
    S2 r; r.c00 = s.c11 * inv_det; r.c01 = -s.c01 * inv_det; r.c11 = s.c00 * inv_det;
    return r;
}
__device__ __forceinline__ float det2_strict(S2 s) { float ad = s.c00 * s.c11; float bc = s.c01 * s.c01; return ad - bc; }
__device__ __forceinline__ bool is_finite(S2 s) { return is_finite(s.c00) && is_finite(s.c11) && is_finite(s.c01); }
__device__ __forceinline__ S3 tcongruence(M23 m, S2 sym) {  // M^T * sym * M
    V2 sc0 = mul(sym, m.c0), sc1 = mul(sym, m.c1), sc2 = mul(sym, m.c2);
    S3 r; r.c00 = dot(m.c0, sc0); r.c01 = dot(m.c0, sc1); r.c02 = dot(m.c0, sc2); r.c11 = dot(m.c1, sc1); r.c12 = dot(m.c1, sc2); r.c22 = dot(m.c2, sc2);
    return r;
}
__device__ __forceinline__ V3 row0(S3 s) { return mk3(s.c00, s.c01, s.c02); }
__device__ __forceinline__ V3 row1(S3 s) { return mk3(s.c01, s.c11, s.c12); }
__device__ __forceinline__ V3 row2(S3 s) { return mk3(s.c02, s.c12, s.c22); }
__device__ __forceinline__ V3 mul(S3 s, V3 v) { return add(add(scale(row0(s), v.x), scale(row1(s), v.y)), scale(row2(s), v.z)); }
__device__ __forceinline__ S3 scale(S3 s, float k) { S3 r; r.c00 = s.c00 * k; r.c01 = s.c01 * k; r.c02 = s.c02 * k; r.c11 = s.c11 * k; r.c12 = s.c12 * k; r.c22 = s.c22 * k; return r; }
__device__ __forceinline__ M3 mul(S3 s, M3 m) { M3 r; r.c0 = mul(s, m.c0); r.c1 = mul(s, m.c1); r.c2 = mul(s, m.c2); return r; }
__device__ __forceinline__ S3 congruence(S3 s, M3 m) {  // m * s * m^T
    V3 r0 = row0(m), r1 = row1(m), r2 = row2(m);
    V3 s0 = mul(s, r0), s1 = mul(s, r1), s2 = mul(s, r2);
    S3 r; r.c00 = dot(r0, s0); r.c01 = dot(r0, s1); r.c02 = dot(r0, s2); r.c11 = dot(r1, s1); r.c12 = dot(r1, s2); r.c22 = dot(r2, s2);
    return r;
}
__device__ __forceinline__ S3 tcongruence(S3 s, M3 m) {  // m^T * s * m
    V3 s0 = mul(s, m.c0), s1 = mul(s, m.c1), s2 = mul(s, m.c2);
    S3 r; r.c00 = dot(m.c0, s0); r.c01 = dot(m.c0, s1); r.c02 = dot(m.c0, s2); r.c11 = dot(m.c1, s1); r.c12 = dot(m.c1, s2); r.c22 = dot(m.c2, s2);
    return r;
}
__device__ __forceinline__ float calc_sigma(float px, float py, S2 conic, float mx, float my) {
    float dx = px - mx;
    float dy = py - my;
    return 0.5f * (conic.c00 * dx * dx + conic.c11 * dy * dy) + conic.c01 * dx * dy;
}

}  // namespace bg
