/*
 * oracle/orc_loss_optim.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restatement of
 *   image_loss_forward_kernel  /root/reference/crates/brush-loss/src/lib.rs:180-359
 *   image_loss_backward_kernel /root/reference/crates/brush-loss/src/lib.rs:370-661
 *   AdamScaled::step/transform /root/reference/crates/brush-train/src/adam_scaled.rs:75-165
 *
 * "parity unpinned": the reference holds no stored numeric vector for SSIM
 * values, SSIM gradients or Adam outputs.  These
 * functions follow the source line by line; tests additionally check them
 * against float64 numpy restatements and finite differences.
 */
#include "orc_api.h"
#include "orc_math.h"

#include <stdlib.h>

#define C1 (0.01f * 0.01f)
#define C2 (0.03f * 0.03f)
#define INV_255 (1.0f / 255.0f)

/* lib.rs:55-68 gauss_taps(): f32 arithmetic, sigma = 1.5, normalised. */
static void gauss_taps(float w[11]) {
    float sigma = 1.5f;
    float sum = 0.0f;
    for (int i = 0; i < 11; i++) {
        float x = (float)i - 5.0f;
        w[i] = expf(-x * x / (2.0f * sigma * sigma)); /* host-side Rust f32::exp: libm expf */
        sum += w[i];
    }
    for (int i = 0; i < 11; i++) w[i] /= sum;
}

/* read_pred (lib.rs:107-118): zero outside the image */
static inline float rd_pred(const float *pred, uint32_t c, int64_t y, int64_t x, uint32_t h, uint32_t w) {
    if (y < 0 || x < 0 || y >= (int64_t)h || x >= (int64_t)w) return 0.0f;
    return pred[(size_t)c * h * w + (size_t)y * w + (size_t)x];
}
/* read_gt (lib.rs:126-142) + composite (lib.rs:252-256) */
static inline float rd_gt_eff(const uint32_t *gt, uint32_t c, int64_t y, int64_t x, uint32_t h, uint32_t w,
                              int composite, float bg_c) {
    if (y < 0 || x < 0 || y >= (int64_t)h || x >= (int64_t)w) {
        /* oob: gt_c = 0 and gt_a = 0 -> gt_eff = 0 + (1 - 0) * bg_c when compositing */
        return composite ? 0.0f + (1.0f - 0.0f) * bg_c : 0.0f;
    }
    uint32_t val = gt[(size_t)y * w + (size_t)x];
    float gt_c = (float)((val >> (c * 8u)) & 0xffu) * INV_255;
    float gt_a = (float)((val >> 24u) & 0xffu) * INV_255;
    return composite ? gt_c + (1.0f - gt_a) * bg_c : gt_c;
}
static inline float rd_gt_a(const uint32_t *gt, uint32_t y, uint32_t x, uint32_t w) {
    return (float)((gt[(size_t)y * w + x] >> 24u) & 0xffu) * INV_255;
}

/* The five separable sums at one pixel, in the reference's accumulation order:
 * horizontal pairs d=1..5 then centre (lib.rs:268-305), vertical pairs d=1..5 then centre (lib.rs:307-329). */
typedef struct { float s[5]; } Sums5;

static void hblur_row(const float *pred, const uint32_t *gt, uint32_t c, int64_t y, uint32_t h, uint32_t w,
                      int composite, float bg_c, const float *gw, Sums5 *out /* [w] */) {
    for (uint32_t x = 0; x < w; x++) {
        float sx = 0, sx2 = 0, sy = 0, sy2 = 0, sxy = 0;
        for (int d = 1; d <= 5; d++) {
            float wd = gw[5 - d];
            float xl = rd_pred(pred, c, y, (int64_t)x - d, h, w), xr = rd_pred(pred, c, y, (int64_t)x + d, h, w);
            float yl = rd_gt_eff(gt, c, y, (int64_t)x - d, h, w, composite, bg_c);
            float yr = rd_gt_eff(gt, c, y, (int64_t)x + d, h, w, composite, bg_c);
            sx += (xl + xr) * wd;
            sx2 += (xl * xl + xr * xr) * wd;
            sy += (yl + yr) * wd;
            sy2 += (yl * yl + yr * yr) * wd;
            sxy += (xl * yl + xr * yr) * wd;
        }
        float xc = rd_pred(pred, c, y, x, h, w), yc = rd_gt_eff(gt, c, y, x, h, w, composite, bg_c);
        float wc = gw[5];
        sx += xc * wc; sx2 += xc * xc * wc; sy += yc * wc; sy2 += yc * yc * wc; sxy += xc * yc * wc;
        out[x].s[0] = sx; out[x].s[1] = sx2; out[x].s[2] = sy; out[x].s[3] = sy2; out[x].s[4] = sxy;
    }
}

/* Blurred moments for every pixel of channel c: out[y*w+x].  Rows outside the image contribute the
 * h-blur of an all-zero row (exactly what the kernel's zero-filled halo gives); with compositing the
 * oob gt value is bg_c, as in rd_gt_eff. */
static Sums5 *blur_moments(const float *pred, const uint32_t *gt, uint32_t c, uint32_t h, uint32_t w, int composite,
                           float bg_c, const float *gw) {
    /* rows -5 .. h+4 */
    Sums5 *hb = (Sums5 *)malloc(sizeof(Sums5) * (size_t)w * (h + 10));
#pragma omp parallel for schedule(static)
    for (int64_t yy = 0; yy < (int64_t)h + 10; yy++) hblur_row(pred, gt, c, yy - 5, h, w, composite, bg_c, gw, hb + (size_t)yy * w);
    Sums5 *out = (Sums5 *)malloc(sizeof(Sums5) * (size_t)w * h);
#pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < (int64_t)h; y++) {
        for (uint32_t x = 0; x < w; x++) {
            float o[5] = {0, 0, 0, 0, 0};
            for (int d = 1; d <= 5; d++) {
                float wd = gw[5 - d];
                const Sums5 *bt = hb + (size_t)(y + 5 - d) * w + x, *bb = hb + (size_t)(y + 5 + d) * w + x;
                for (int k = 0; k < 5; k++) o[k] += (bt->s[k] + bb->s[k]) * wd;
            }
            const Sums5 *bc = hb + (size_t)(y + 5) * w + x;
            for (int k = 0; k < 5; k++) o[k] += bc->s[k] * gw[5];
            for (int k = 0; k < 5; k++) out[(size_t)y * w + x].s[k] = o[k];
        }
    }
    free(hb);
    return out;
}

void orc_image_loss_forward(const float *pred, const uint32_t *gt, uint32_t c_n, uint32_t h, uint32_t w, float l1_w,
                            float ssim_w, const float *bg3, int mask, float *loss_map) {
    float gw[11];
    gauss_taps(gw);
    const int composite = bg3 != NULL;
    for (uint32_t c = 0; c < c_n; c++) {
        if (c == 3) { /* alpha-match channel (lib.rs:215-227) */
            for (uint32_t y = 0; y < h; y++)
                for (uint32_t x = 0; x < w; x++) {
                    size_t idx = (size_t)3 * h * w + (size_t)y * w + x;
                    float ga = rd_gt_a(gt, y, x, w);
                    float v = fabsf(pred[idx] - ga);
                    if (mask) v = v * ga;
                    loss_map[idx] = v;
                }
            continue;
        }
        float bg_c = composite ? bg3[c] : 0.0f;
        Sums5 *mo = blur_moments(pred, gt, c, h, w, composite, bg_c, gw);
#pragma omp parallel for schedule(static)
        for (int64_t y = 0; y < (int64_t)h; y++)
            for (uint32_t x = 0; x < w; x++) {
                const float *o = mo[(size_t)y * w + x].s;
                float mu1 = o[0], mu2 = o[2];
                float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2;
                float sigma1_sq = orc_max(0.0f, o[1] - mu1_sq);
                float sigma2_sq = orc_max(0.0f, o[3] - mu2_sq);
                float sigma12 = o[4] - mu1 * mu2;
                float a = mu1_sq + mu2_sq + C1;
                float b = sigma1_sq + sigma2_sq + C2;
                float c_top = 2.0f * mu1 * mu2 + C1;
                float d_top = 2.0f * sigma12 + C2;
                float raw = (c_top * d_top) / (a * b);
                float val = orc_clamp(raw, -1.0f, 1.0f);
                float p1 = rd_pred(pred, c, y, x, h, w);
                float p2 = rd_gt_eff(gt, c, y, x, h, w, composite, bg_c);
                float l1 = fabsf(p1 - p2);
                float loss_v = l1_w * l1 + ssim_w * val;
                if (mask) loss_v = loss_v * rd_gt_a(gt, (uint32_t)y, x, w);
                loss_map[(size_t)c * h * w + (size_t)y * w + x] = loss_v;
            }
        free(mo);
    }
}

void orc_image_loss_backward(const float *pred, const uint32_t *gt, const float *dl_dmap, uint32_t c_n, uint32_t h,
                             uint32_t w, float l1_w, float ssim_w, const float *bg3, int mask, float *dl_dpred) {
    float gw[11];
    gauss_taps(gw);
    const int composite = bg3 != NULL;
    for (uint32_t c = 0; c < c_n; c++) {
        if (c == 3) { /* lib.rs:393-414 */
            for (uint32_t y = 0; y < h; y++)
                for (uint32_t x = 0; x < w; x++) {
                    size_t idx = (size_t)3 * h * w + (size_t)y * w + x;
                    float ga = rd_gt_a(gt, y, x, w);
                    float diff = pred[idx] - ga;
                    float sign = diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f);
                    float chain = dl_dmap[idx];
                    if (mask) chain = chain * ga;
                    dl_dpred[idx] = sign * chain;
                }
            continue;
        }
        float bg_c = composite ? bg3[c] : 0.0f;
        Sums5 *mo = blur_moments(pred, gt, c, h, w, composite, bg_c, gw);
        /* chain * partials at every in-image pixel; zero outside (coords() oob -> chain = 0), lib.rs:505-580 */
        float *part = (float *)calloc((size_t)w * h * 3, sizeof(float));
#pragma omp parallel for schedule(static)
        for (int64_t y = 0; y < (int64_t)h; y++)
            for (uint32_t x = 0; x < w; x++) {
                const float *o = mo[(size_t)y * w + x].s;
                float mu1 = o[0], mu2 = o[2];
                float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2;
                float sigma1_sq = orc_max(0.0f, o[1] - mu1_sq);
                float sigma2_sq = orc_max(0.0f, o[3] - mu2_sq);
                float sigma12 = o[4] - mu1 * mu2;
                float a = mu1_sq + mu2_sq + C1;
                float b = sigma1_sq + sigma2_sq + C2;
                float c_top = 2.0f * mu1 * mu2 + C1;
                float d_top = 2.0f * sigma12 + C2;
                float inv_ab = 1.0f / (a * b);
                float cd = c_top * d_top * inv_ab;
                int clamped = cd < -1.0f || cd > 1.0f;
                float dmu1 = clamped ? 0.0f : 2.0f * mu2 * inv_ab * (d_top - c_top) - 2.0f * mu1 * cd * (1.0f / a - 1.0f / b);
                float dsigma1 = clamped ? 0.0f : -cd / b;
                float dsigma12 = clamped ? 0.0f : 2.0f * c_top * inv_ab;
                float chain = dl_dmap[(size_t)c * h * w + (size_t)y * w + x];
                if (mask) chain = chain * rd_gt_a(gt, (uint32_t)y, x, w);
                float *p = part + ((size_t)y * w + x) * 3;
                p[0] = dmu1 * chain; p[1] = dsigma1 * chain; p[2] = dsigma12 * chain;
            }
        free(mo);
        /* second separable blur (lib.rs:582-631), zero padded */
        float *hb = (float *)calloc((size_t)w * (h + 10) * 3, sizeof(float));
#pragma omp parallel for schedule(static)
        for (int64_t y = 0; y < (int64_t)h; y++)
            for (uint32_t x = 0; x < w; x++) {
                float a3[3] = {0, 0, 0};
                for (int d = 1; d <= 5; d++) {
                    float wd = gw[5 - d];
                    for (int k = 0; k < 3; k++) {
                        float l = ((int64_t)x - d >= 0) ? part[((size_t)y * w + x - d) * 3 + k] : 0.0f;
                        float rr = (x + d < w) ? part[((size_t)y * w + x + d) * 3 + k] : 0.0f;
                        a3[k] += (l + rr) * wd;
                    }
                }
                for (int k = 0; k < 3; k++) a3[k] += part[((size_t)y * w + x) * 3 + k] * gw[5];
                for (int k = 0; k < 3; k++) hb[((size_t)(y + 5) * w + x) * 3 + k] = a3[k];
            }
#pragma omp parallel for schedule(static)
        for (int64_t y = 0; y < (int64_t)h; y++)
            for (uint32_t x = 0; x < w; x++) {
                float s3[3] = {0, 0, 0};
                for (int d = 1; d <= 5; d++) {
                    float wd = gw[5 - d];
                    for (int k = 0; k < 3; k++)
                        s3[k] += (hb[((size_t)(y + 5 - d) * w + x) * 3 + k] + hb[((size_t)(y + 5 + d) * w + x) * 3 + k]) * wd;
                }
                for (int k = 0; k < 3; k++) s3[k] += hb[((size_t)(y + 5) * w + x) * 3 + k] * gw[5];
                size_t idx = (size_t)c * h * w + (size_t)y * w + x;
                float p1 = pred[idx];
                float gt_eff = rd_gt_eff(gt, c, y, x, h, w, composite, bg_c);
                float ssim_grad = s3[0] + (2.0f * p1) * s3[1] + gt_eff * s3[2];
                float diff = p1 - gt_eff;
                float l1_sign = diff > 0.0f ? 1.0f : (diff < 0.0f ? -1.0f : 0.0f);
                float chain_centre = dl_dmap[idx];
                if (mask) chain_centre = chain_centre * rd_gt_a(gt, (uint32_t)y, x, w);
                dl_dpred[idx] = ssim_w * ssim_grad + l1_w * l1_sign * chain_centre;
            }
        free(hb);
        free(part);
    }
}

/* compiler-rt __powisf2, which Rust's f32::powi lowers to (adam_scaled.rs:135-142). */
static float powi_f32(float a, int b) {
    const int recip = b < 0;
    float r = 1.0f;
    while (1) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0f / r : r;
}

void orc_adam_step(float *p, const float *g, float *m, float *v, uint64_t rows, uint32_t cols,
                   const float *lr_scale_per_col, float lr, float beta1, float beta2, float eps, int t, int reduce_v) {
    const float f1 = 1.0f - beta1, f2 = 1.0f - beta2;
    const float bc1 = 1.0f - powi_f32(beta1, t), bc2 = 1.0f - powi_f32(beta2, t);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < (int64_t)rows; r++) {
        const float *gr = g + (size_t)r * cols;
        float *mr = m + (size_t)r * cols, *pr = p + (size_t)r * cols;
        float row_mean_sq = 0.0f;
        if (reduce_v) { /* mean_trailing_dims: sum_dim(1) / count (adam_scaled.rs:152-165) */
            float s = 0.0f;
            for (uint32_t c = 0; c < cols; c++) s += gr[c] * gr[c];
            row_mean_sq = s / (float)cols;
            float vv = (t == 1) ? row_mean_sq * f2 : v[r] * beta2 + row_mean_sq * f2;
            v[r] = vv;
        }
        for (uint32_t c = 0; c < cols; c++) {
            float gg = gr[c];
            float mm = (t == 1) ? gg * f1 : mr[c] * beta1 + gg * f1;
            mr[c] = mm;
            float vv;
            if (reduce_v) {
                vv = v[r];
            } else {
                float gsq = gg * gg;
                vv = (t == 1) ? gsq * f2 : v[(size_t)r * cols + c] * beta2 + gsq * f2;
                v[(size_t)r * cols + c] = vv;
            }
            float m_hat = mm / bc1;
            float v_hat = vv / bc2;
            float upd = m_hat / (sqrtf(v_hat) + eps);
            float step = lr_scale_per_col ? lr_scale_per_col[c] * lr : lr;
            pr[c] = pr[c] - upd * step;
        }
    }
}

/* ---- Mip-Splatting 3D filter floor ------------------------------------------------------------
 * orc_compute_min_scale follows compute_min_scale (brush-train/src/train.rs:102-125);
 * orc_fold_min_scale_fwd follows fold_min_scale (brush-render/src/gaussian_splats.rs:86-111) op by op;
 * orc_fold_min_scale_bwd is the reverse-mode chain of those ops, written out un-simplified (one
 * statement per burn op), i.e. what burn's autodiff evaluates.  view_cams: [views,4] = x,y,z,focal. */
void orc_compute_min_scale(const float *transforms, uint32_t n, const float *view_cams, uint32_t views, float factor,
                           float *f_out) {
    const float sf = sqrtf(factor);
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n; i++) {
        const float *t = transforms + (size_t)i * 10;
        float best = 0.0f;
        for (uint32_t v = 0; v < views; v++) {
            const float *c = view_cams + (size_t)v * 4;
            float dx = t[0] - c[0], dy = t[1] - c[1], dz = t[2] - c[2];
            float dist = sqrtf((dx * dx + dy * dy) + dz * dz);
            float ratio = dist / fmaxf(c[3], 1e-6f);
            best = (v == 0) ? ratio : fminf(best, ratio);
        }
        f_out[i] = best * sf;
    }
}

void orc_fold_min_scale_fwd(const float *transforms, const float *raw_opac, const float *f, uint32_t n,
                            float *transforms_out, float *raw_opac_out) {
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n; i++) {
        const float *t = transforms + (size_t)i * 10;
        float *o = transforms_out + (size_t)i * 10;
        float f2 = f[i] * f[i];
        float s2[3], s2f[3];
        for (int a = 0; a < 3; a++) { s2[a] = expf(t[7 + a] * 2.0f); s2f[a] = s2[a] + f2; }
        for (int c = 0; c < 7; c++) o[c] = t[c];
        for (int a = 0; a < 3; a++) o[7 + a] = logf(s2f[a]) * 0.5f;
        float det1 = s2[0] * s2[1] * s2[2], det2 = s2f[0] * s2f[1] * s2f[2];
        float coef = sqrtf(det1 / det2);
        float sig = 1.0f / (1.0f + expf(-raw_opac[i]));
        float opac = fminf(fmaxf(sig * coef, 1e-6f), 1.0f - 1e-6f);
        raw_opac_out[i] = logf(opac / (1.0f - opac));
    }
}

void orc_fold_min_scale_bwd(const float *transforms, const float *raw_opac, const float *f, uint32_t n,
                            float *v_transforms, float *v_raw_opac) {
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n; i++) {
        const float *t = transforms + (size_t)i * 10;
        float *vt = v_transforms + (size_t)i * 10;
        float f2 = f[i] * f[i];
        float s2[3], s2f[3];
        for (int a = 0; a < 3; a++) { s2[a] = expf(t[7 + a] * 2.0f); s2f[a] = s2[a] + f2; }
        float det1 = s2[0] * s2[1] * s2[2], det2 = s2f[0] * s2f[1] * s2f[2];
        float r = det1 / det2;
        float coef = sqrtf(r);
        float sig = 1.0f / (1.0f + expf(-raw_opac[i]));
        float pre = sig * coef;
        float opac = fminf(fmaxf(pre, 1e-6f), 1.0f - 1e-6f);
        /* raw' = log(q), q = opac / (1 - opac) */
        float one_m = 1.0f - opac;
        float q = opac / one_m;
        float v_q = v_raw_opac[i] / q;
        float v_opac = v_q / one_m + v_q * opac / (one_m * one_m);
        float v_pre = (pre >= 1e-6f && pre <= 1.0f - 1e-6f) ? v_opac : 0.0f; /* clamp passes gradient inside the range */
        float v_sig = v_pre * coef, v_coef = v_pre * sig;
        v_raw_opac[i] = v_sig * (sig * (1.0f - sig));
        float v_r = v_coef * 0.5f / coef;
        float v_det1 = v_r / det2, v_det2 = -v_r * det1 / (det2 * det2);
        for (int a = 0; a < 3; a++) {
            int b = (a + 1) % 3, c = (a + 2) % 3;
            float v_s2f = v_det2 * (s2f[b] * s2f[c]) + vt[7 + a] * 0.5f / s2f[a]; /* det2 product; log*0.5 */
            float v_s2 = v_det1 * (s2[b] * s2[c]) + v_s2f;                          /* det1 product; s2 + f2 */
            vt[7 + a] = v_s2 * s2[a] * 2.0f;                                        /* exp(2 ls) */
        }
    }
}
