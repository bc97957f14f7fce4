/*
 * oracle/orc_math.h -- TEST INFRASTRUCTURE ONLY (CPU oracle, never shipped).
 *
 * Scalar f32 restatement of the device-side math types of the reference:
 *   /root/reference/crates/brush-cube/src/lib.rs:34-578
 *     (Vec3A, Vec2, Quat, Mat3, Mat2x3, Sym2, Sym3, sigmoid, is_finite_f32,
 *      calc_sigma)
 *
 * Every helper keeps the reference's operation order (which operand is
 * multiplied first, which sums are formed first).  The file must be compiled
 * with -ffp-contract=off so that gcc never fuses a*b+c: the only fused
 * operations are the explicit fmaf() calls inside orc_expf / orc_logf.
 *
 * Transcendentals.  The reference calls WGSL exp()/log(), whose results are
 * implementation defined to a few ulp.  The oracle uses its own exp/log
 * (Cephes-style range reduction + polynomial, <= 1 ulp) built only from
 * IEEE-754 +,*,fma and integer bit operations.  The CUDA projection kernels
 * use the same recipe, which makes the whole per-Gaussian stage (cull,
 * conic, extents, tile hits, SH colour) reproducible bit for bit on CPU and
 * GPU.  Spec of the recipe: DESIGN.md "Deterministic exp/log".
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t orc_f2u(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static inline float orc_u2f(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }

/* brush-cube/src/lib.rs:561-565 : exponent bits all ones <=> NaN / Inf. */
static inline int orc_is_finite(float x) { return ((orc_f2u(x) >> 23) & 0xFFu) != 0xFFu; }

/* WGSL-style min/max/clamp: clamp(x,lo,hi) = min(max(x,lo),hi). */
static inline float orc_min(float a, float b) { return fminf(a, b); }
static inline float orc_max(float a, float b) { return fmaxf(a, b); }
static inline float orc_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* 2^k for k in [-126,127] by building the exponent field. */
static inline float orc_pow2i(int k) { return orc_u2f((uint32_t)(k + 127) << 23); }

/* Deterministic expf.  n = rint(x*log2e); r = x - n*ln2 (two-term Cody-Waite,
 * fused); e^r = 1 + r + r^2*P(r) (Cephes expf polynomial, Horner with fmaf);
 * result = (poly * 2^(n/2)) * 2^(n - n/2). */
static inline float orc_expf(float x) {
    if (x != x) return x;
    if (x > 88.72283935546875f) return INFINITY;
    if (x < -103.97208404541015625f) return 0.0f;
    float t = x * 1.44269502162933349609375f;
    float n = rintf(t);
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float e = fmaf(p, r2, r) + 1.0f;
    int ni = (int)n;
    int n1 = ni / 2;
    int n2 = ni - n1;
    return (e * orc_pow2i(n1)) * orc_pow2i(n2);
}

/* Deterministic logf for finite x > 0 (Cephes logf).  Returns -inf for 0,
 * NaN for x < 0 or NaN, +inf for +inf. */
static inline float orc_logf(float x) {
    if (x != x || x < 0.0f) return NAN;
    if (x == 0.0f) return -INFINITY;
    if (!orc_is_finite(x)) return x;
    uint32_t u = orc_f2u(x);
    int e = 0;
    if ((u >> 23) == 0u) { /* denormal: scale up by 2^23 */
        x = x * 8388608.0f;
        u = orc_f2u(x);
        e = -23;
    }
    e += (int)(u >> 23) - 126;                       /* x = m * 2^e, m in [0.5,1) */
    float m = orc_u2f((u & 0x007FFFFFu) | 0x3F000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float p = 7.0376836292e-2f;
    p = fmaf(p, m, -1.1514610310e-1f);
    p = fmaf(p, m, 1.1676998740e-1f);
    p = fmaf(p, m, -1.2420140846e-1f);
    p = fmaf(p, m, 1.4249322787e-1f);
    p = fmaf(p, m, -1.6668057665e-1f);
    p = fmaf(p, m, 2.0000714765e-1f);
    p = fmaf(p, m, -2.4999993993e-1f);
    p = fmaf(p, m, 3.3333331174e-1f);
    float y = (m * z) * p;
    float fe = (float)e;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(z, -0.5f, y);
    float r = m + y;
    r = fmaf(fe, 0.693359375f, r);
    return r;
}

/* brush-cube/src/lib.rs:555-558 */
static inline float orc_sigmoid(float x) { return 1.0f / (1.0f + orc_expf(-x)); }

typedef struct { float x, y, z; } ovec3;
typedef struct { float x, y; } ovec2;
typedef struct { float w, x, y, z; } oquat;
/* column major, c{i} = column i (brush-cube/src/lib.rs:219-231) */
typedef struct { ovec3 c0, c1, c2; } omat3;
typedef struct { ovec2 c0, c1, c2; } omat2x3;
typedef struct { float c00, c01, c11; } osym2;
typedef struct { float c00, c01, c02, c11, c12, c22; } osym3;

static inline ovec3 v3(float x, float y, float z) { ovec3 v = {x, y, z}; return v; }
static inline ovec2 v2(float x, float y) { ovec2 v = {x, y}; return v; }
static inline ovec3 v3_add(ovec3 a, ovec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline ovec3 v3_sub(ovec3 a, ovec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline ovec3 v3_scale(ovec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
/* lib.rs:83-87 : p0+p1+p2+p3 with the padding lane p3 == 0 */
static inline float v3_dot(ovec3 a, ovec3 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + 0.0f; }
static inline float v3_length(ovec3 a) { return sqrtf(v3_dot(a, a)); }
static inline ovec3 v3_normalize(ovec3 a) { return v3_scale(a, 1.0f / v3_length(a)); }
static inline int v3_is_finite(ovec3 a) { return orc_is_finite(a.x) && orc_is_finite(a.y) && orc_is_finite(a.z); }

static inline ovec2 v2_add(ovec2 a, ovec2 b) { return v2(a.x + b.x, a.y + b.y); }
static inline ovec2 v2_scale(ovec2 a, float s) { return v2(a.x * s, a.y * s); }
static inline float v2_dot(ovec2 a, ovec2 b) { return a.x * b.x + a.y * b.y; }

static inline float q_dot(oquat a, oquat b) { return ((a.w * b.w + a.x * b.x) + a.y * b.y) + a.z * b.z; }
static inline oquat q_scale(oquat a, float s) { oquat q = {a.w * s, a.x * s, a.y * s, a.z * s}; return q; }
static inline oquat q_normalize(oquat a) { return q_scale(a, 1.0f / sqrtf(q_dot(a, a))); }

/* lib.rs:190-216 */
static inline omat3 q_to_mat3(oquat q) {
    float w = q.w, qx = q.x, qy = q.y, qz = q.z;
    float x2 = qx * qx, y2 = qy * qy, z2 = qz * qz;
    float xy = qx * qy, xz = qx * qz, yz = qy * qz;
    float wx = w * qx, wy = w * qy, wz = w * qz;
    omat3 m;
    m.c0 = v3(1.0f - 2.0f * (y2 + z2), 2.0f * (xy + wz), 2.0f * (xz - wy));
    m.c1 = v3(2.0f * (xy - wz), 1.0f - 2.0f * (x2 + z2), 2.0f * (yz + wx));
    m.c2 = v3(2.0f * (xz + wy), 2.0f * (yz - wx), 1.0f - 2.0f * (x2 + y2));
    return m;
}

/* lib.rs:262-268 */
static inline ovec3 m3_mul_vec3(omat3 m, ovec3 v) {
    return v3_add(v3_add(v3_scale(m.c0, v.x), v3_scale(m.c1, v.y)), v3_scale(m.c2, v.z));
}
static inline ovec3 m3_transpose_mul_vec3(omat3 m, ovec3 v) {
    return v3(v3_dot(m.c0, v), v3_dot(m.c1, v), v3_dot(m.c2, v));
}
static inline omat3 m3_mul_mat3(omat3 m, omat3 n) {
    omat3 r = {m3_mul_vec3(m, n.c0), m3_mul_vec3(m, n.c1), m3_mul_vec3(m, n.c2)};
    return r;
}
static inline omat3 m3_mul_diag(omat3 m, ovec3 s) {
    omat3 r = {v3_scale(m.c0, s.x), v3_scale(m.c1, s.y), v3_scale(m.c2, s.z)};
    return r;
}
static inline ovec3 m3_row0(omat3 m) { return v3(m.c0.x, m.c1.x, m.c2.x); }
static inline ovec3 m3_row1(omat3 m) { return v3(m.c0.y, m.c1.y, m.c2.y); }
static inline ovec3 m3_row2(omat3 m) { return v3(m.c0.z, m.c1.z, m.c2.z); }
/* lib.rs:309-322 : M * M^T */
static inline osym3 m3_outer_product_self(omat3 m) {
    ovec3 r0 = m3_row0(m), r1 = m3_row1(m), r2 = m3_row2(m);
    osym3 s = {v3_dot(r0, r0), v3_dot(r0, r1), v3_dot(r0, r2), v3_dot(r1, r1), v3_dot(r1, r2), v3_dot(r2, r2)};
    return s;
}

/* lib.rs:346-352 */
static inline ovec2 m23_mul_vec3(omat2x3 m, ovec3 v) {
    return v2_add(v2_add(v2_scale(m.c0, v.x), v2_scale(m.c1, v.y)), v2_scale(m.c2, v.z));
}
static inline omat2x3 m23_mul_mat3(omat2x3 m, omat3 n) {
    omat2x3 r = {m23_mul_vec3(m, n.c0), m23_mul_vec3(m, n.c1), m23_mul_vec3(m, n.c2)};
    return r;
}
static inline ovec3 m23_row0(omat2x3 m) { return v3(m.c0.x, m.c1.x, m.c2.x); }
static inline ovec3 m23_row1(omat2x3 m) { return v3(m.c0.y, m.c1.y, m.c2.y); }
/* lib.rs:378-385 */
static inline osym2 m23_gram(omat2x3 m) {
    osym2 s;
    s.c00 = m.c0.x * m.c0.x + m.c1.x * m.c1.x + m.c2.x * m.c2.x;
    s.c01 = m.c0.x * m.c0.y + m.c1.x * m.c1.y + m.c2.x * m.c2.y;
    s.c11 = m.c0.y * m.c0.y + m.c1.y * m.c1.y + m.c2.y * m.c2.y;
    return s;
}

static inline ovec2 s2_col0(osym2 s) { return v2(s.c00, s.c01); }
static inline ovec2 s2_col1(osym2 s) { return v2(s.c01, s.c11); }
static inline ovec2 s2_mul_vec2(osym2 s, ovec2 v) { return v2_add(v2_scale(s2_col0(s), v.x), v2_scale(s2_col1(s), v.y)); }
static inline osym2 s2_scale(osym2 s, float k) { osym2 r = {s.c00 * k, s.c01 * k, s.c11 * k}; return r; }
static inline float s2_max_abs(osym2 s) { return orc_max(orc_max(fabsf(s.c00), fabsf(s.c11)), fabsf(s.c01)); }
static inline omat2x3 s2_mul_mat2x3(osym2 s, omat2x3 n) {
    omat2x3 r = {s2_mul_vec2(s, n.c0), s2_mul_vec2(s, n.c1), s2_mul_vec2(s, n.c2)};
    return r;
}
/* lib.rs:431-440 */
static inline osym2 s2_inverse(osym2 s) {
    float det = s.c00 * s.c11 - s.c01 * s.c01;
    float inv_det = (det > 0.0f) ? 1.0f / det : 0.0f;
    osym2 r = {s.c11 * inv_det, -s.c01 * inv_det, s.c00 * inv_det};
    return r;
}
/* lib.rs:444-448 */
static inline float s2_det2_strict(osym2 s) { float ad = s.c00 * s.c11; float bc = s.c01 * s.c01; return ad - bc; }
static inline int s2_is_finite(osym2 s) { return orc_is_finite(s.c00) && orc_is_finite(s.c11) && orc_is_finite(s.c01); }

/* lib.rs:363-376 : M^T * sym * M */
static inline osym3 m23_transpose_congruence_sym2(omat2x3 m, osym2 sym) {
    ovec2 sc0 = s2_mul_vec2(sym, m.c0), sc1 = s2_mul_vec2(sym, m.c1), sc2 = s2_mul_vec2(sym, m.c2);
    osym3 r = {v2_dot(m.c0, sc0), v2_dot(m.c0, sc1), v2_dot(m.c0, sc2), v2_dot(m.c1, sc1), v2_dot(m.c1, sc2), v2_dot(m.c2, sc2)};
    return r;
}

static inline ovec3 s3_row0(osym3 s) { return v3(s.c00, s.c01, s.c02); }
static inline ovec3 s3_row1(osym3 s) { return v3(s.c01, s.c11, s.c12); }
static inline ovec3 s3_row2(osym3 s) { return v3(s.c02, s.c12, s.c22); }
static inline ovec3 s3_mul_vec3(osym3 s, ovec3 v) {
    return v3_add(v3_add(v3_scale(s3_row0(s), v.x), v3_scale(s3_row1(s), v.y)), v3_scale(s3_row2(s), v.z));
}
static inline osym3 s3_scale(osym3 s, float k) {
    osym3 r = {s.c00 * k, s.c01 * k, s.c02 * k, s.c11 * k, s.c12 * k, s.c22 * k};
    return r;
}
static inline omat3 s3_mul_mat3(osym3 s, omat3 m) {
    omat3 r = {s3_mul_vec3(s, m.c0), s3_mul_vec3(s, m.c1), s3_mul_vec3(s, m.c2)};
    return r;
}
/* lib.rs:502-514 : m * self * m^T */
static inline osym3 s3_congruence(osym3 s, omat3 m) {
    ovec3 r0 = m3_row0(m), r1 = m3_row1(m), r2 = m3_row2(m);
    ovec3 sr0 = s3_mul_vec3(s, r0), sr1 = s3_mul_vec3(s, r1), sr2 = s3_mul_vec3(s, r2);
    osym3 r = {v3_dot(r0, sr0), v3_dot(r0, sr1), v3_dot(r0, sr2), v3_dot(r1, sr1), v3_dot(r1, sr2), v3_dot(r2, sr2)};
    return r;
}
/* lib.rs:517-529 : m^T * self * m */
static inline osym3 s3_transpose_congruence(osym3 s, omat3 m) {
    ovec3 sc0 = s3_mul_vec3(s, m.c0), sc1 = s3_mul_vec3(s, m.c1), sc2 = s3_mul_vec3(s, m.c2);
    osym3 r = {v3_dot(m.c0, sc0), v3_dot(m.c0, sc1), v3_dot(m.c0, sc2), v3_dot(m.c1, sc1), v3_dot(m.c1, sc2), v3_dot(m.c2, sc2)};
    return r;
}

/* lib.rs:571-576 */
static inline float orc_calc_sigma(float px, float py, osym2 conic, float mx, float my) {
    float dx = px - mx;
    float dy = py - my;
    return 0.5f * (conic.c00 * dx * dx + conic.c11 * dy * dy) + conic.c01 * dx * dy;
}

#endif
