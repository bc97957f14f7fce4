/*
 * oracle/orc_forward.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restatement of the forward render pipeline of the reference,
 *   /root/reference/crates/brush-render/src/render.rs:37-315
 * stage by stage.  Each function cites the kernel it follows.  f32 everywhere,
 * no FMA contraction (build flag -ffp-contract=off), same operation order as
 * the reference source.
 *
 * One deliberate choice inside the reference's behaviour set: visible
 * Gaussians are compacted in index order (the reference takes an atomic slot,
 * project_forward.rs:122-124, so ties in depth fall back to a
 * nondeterministic arrival order; index order is one legal arrival order and
 * makes results reproducible).  SURVEY.md H2.
 */
#include "orc_api.h"
#include "orc_math.h"
#include "orc_camera.h"

#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE_WIDTH 16u
#define TILE_SIZE 256u

float orc_expf_det(float x) { return orc_expf(x); }
float orc_logf_det(float x) { return orc_logf(x); }
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- uniforms helpers (kernels/types.rs:82-114) ---- */
static inline omat3 view_rotation(const OrcCamera *u) {
    omat3 m;
    m.c0 = v3(u->viewmat[0], u->viewmat[1], u->viewmat[2]);
    m.c1 = v3(u->viewmat[3], u->viewmat[4], u->viewmat[5]);
    m.c2 = v3(u->viewmat[6], u->viewmat[7], u->viewmat[8]);
    return m;
}
static inline ovec3 view_translation(const OrcCamera *u) { return v3(u->viewmat[9], u->viewmat[10], u->viewmat[11]); }

/* kernels/helpers.rs:311-314 */
static inline ovec3 world_to_cam(ovec3 mean, const OrcCamera *u) {
    return v3_add(m3_mul_vec3(view_rotation(u), mean), view_translation(u));
}

/* kernels/helpers.rs:145-174 */
osym2 orc_calc_cov2d(ovec3 scale, oquat quat, ovec3 mean_c, const OrcCamera *u) {
    omat3 ns = m3_mul_diag(m3_mul_mat3(view_rotation(u), q_to_mat3(quat)), scale);
    omat2x3 jac = orc_cam_jacobian(mean_c, u);
    omat2x3 v = m23_mul_mat3(jac, ns);
    osym2 raw = m23_gram(v);
    float lim = 1.0e18f;
    float max_abs = s2_max_abs(raw);
    float scale_down = (max_abs > lim) ? lim / max_abs : 1.0f;
    return s2_scale(raw, scale_down);
}

/* kernels/helpers.rs:179-195 */
osym2 orc_compensate_cov2d(osym2 c, int mip, float *filter_comp) {
    float cov_blur = mip ? 0.1f : 0.3f;
    osym2 blurred = {c.c00 + cov_blur, c.c01, c.c11 + cov_blur};
    *filter_comp = 1.0f;
    if (mip) {
        float det_raw = orc_max(s2_det2_strict(c), 0.0f);
        float det_blurred = s2_det2_strict(blurred);
        *filter_comp = sqrtf(det_raw / det_blurred);
    }
    return blurred;
}

/* kernels/helpers.rs:84-96 */
static inline void compute_bbox_extent(osym2 conic, float power_threshold, float *ex, float *ey) {
    float det = conic.c00 * conic.c11 - conic.c01 * conic.c01;
    int degenerate = det <= 0.0f;
    float inv_det = degenerate ? 0.0f : 1.0f / det;
    float x = sqrtf(2.0f * power_threshold * conic.c11 * inv_det);
    float y = sqrtf(2.0f * power_threshold * conic.c00 * inv_det);
    *ex = degenerate ? -1.0f : x;
    *ey = degenerate ? -1.0f : y;
}

typedef struct { uint32_t min_x, min_y, max_x, max_y; } TileBbox;

/* kernels/helpers.rs:112-140.  `as u32` of a clamped non-negative finite float truncates. */
static inline TileBbox get_tile_bbox(float px, float py, float ex, float ey, uint32_t bw, uint32_t bh) {
    float tw = (float)TILE_WIDTH;
    float cx = px / tw, cy = py / tw, dx = ex / tw, dy = ey / tw;
    float bwf = (float)bw, bhf = (float)bh;
    TileBbox b;
    b.min_x = (uint32_t)orc_clamp(cx - dx, 0.0f, bwf);
    b.min_y = (uint32_t)orc_clamp(cy - dy, 0.0f, bhf);
    b.max_x = (uint32_t)orc_clamp(cx + dx + 1.0f, 0.0f, bwf);
    b.max_y = (uint32_t)orc_clamp(cy + dy + 1.0f, 0.0f, bhf);
    return b;
}

/* kernels/helpers.rs:225-264 (StopThePop tile test); rect from helpers.rs:98-108 */
static inline int will_primitive_contribute(uint32_t tx, uint32_t ty, float mx, float my, osym2 conic,
                                            float power_threshold) {
    float rmin_x = (float)(tx * TILE_WIDTH);
    float rmin_y = (float)(ty * TILE_WIDTH);
    float rmax_x = rmin_x + (float)TILE_WIDTH;
    float rmax_y = rmin_y + (float)TILE_WIDTH;
    int x_left = mx < rmin_x;
    int x_right = mx > rmax_x;
    int in_x_range = !(x_left || x_right);
    int y_above = my < rmin_y;
    int y_below = my > rmax_y;
    int in_y_range = !(y_above || y_below);
    int hit = in_x_range && in_y_range;
    if (!hit) {
        float corner_x = x_left ? rmin_x : rmax_x;
        float corner_y = y_above ? rmin_y : rmax_y;
        float width = rmax_x - rmin_x;
        float height = rmax_y - rmin_y;
        float dxf = x_left ? width : -width;
        float dyf = y_above ? height : -height;
        float diff_x = mx - corner_x;
        float diff_y = my - corner_y;
        float tx_raw = (dxf * conic.c00 * diff_x + dxf * conic.c01 * diff_y) / (dxf * conic.c00 * dxf);
        float ty_raw = (dyf * conic.c01 * diff_x + dyf * conic.c11 * diff_y) / (dyf * conic.c11 * dyf);
        float tx_ = in_y_range ? 0.0f : orc_clamp(tx_raw, 0.0f, 1.0f);
        float ty_ = in_x_range ? 0.0f : orc_clamp(ty_raw, 0.0f, 1.0f);
        float max_x = corner_x + tx_ * dxf;
        float max_y = corner_y + ty_ * dyf;
        hit = orc_calc_sigma(max_x, max_y, conic, mx, my) <= power_threshold;
    }
    (void)y_below;
    return hit;
}

/* kernels/helpers.rs:203-222 */
static inline uint32_t count_contributing_tiles(TileBbox bb, float mx, float my, osym2 conic, float pt) {
    uint32_t bb_w = bb.max_x - bb.min_x;
    uint32_t num = (bb.max_y - bb.min_y) * bb_w;
    uint32_t hits = 0;
    for (uint32_t i = 0; i < num; i++) {
        uint32_t tx = (i % bb_w) + bb.min_x;
        uint32_t ty = (i / bb_w) + bb.min_y;
        if (will_primitive_contribute(tx, ty, mx, my, conic, pt)) hits++;
    }
    return hits;
}

/* ---- SH (kernels/sh.rs:41-136) ---- */
static inline ovec3 rd3(const float *c, uint32_t base) { return v3(c[base], c[base + 1], c[base + 2]); }

ovec3 orc_sh_to_color(const float *c, uint32_t degree, ovec3 v) {
    ovec3 color = v3_scale(rd3(c, 0), 0.2820948f);
    if (degree >= 1) {
        float f0a = 0.4886025f;
        color = v3_add(color, v3_scale(rd3(c, 3), -f0a * v.y));
        color = v3_add(color, v3_scale(rd3(c, 6), f0a * v.z));
        color = v3_add(color, v3_scale(rd3(c, 9), -f0a * v.x));
        if (degree >= 2) {
            float z2 = v.z * v.z;
            float f0b = -1.0925485f * v.z;
            float f1a = 0.54627424f;
            float fc1 = v.x * v.x - v.y * v.y;
            float fs1 = 2.0f * v.x * v.y;
            float p4 = f1a * fs1, p5 = f0b * v.y, p6 = 0.9461747f * z2 - 0.31539157f, p7 = f0b * v.x, p8 = f1a * fc1;
            color = v3_add(color, v3_scale(rd3(c, 12), p4));
            color = v3_add(color, v3_scale(rd3(c, 15), p5));
            color = v3_add(color, v3_scale(rd3(c, 18), p6));
            color = v3_add(color, v3_scale(rd3(c, 21), p7));
            color = v3_add(color, v3_scale(rd3(c, 24), p8));
            if (degree >= 3) {
                float f0c = -2.285229f * z2 + 0.4570458f;
                float f1b = 1.4453057f * v.z;
                float f2a = -0.5900436f;
                float fc2 = v.x * fc1 - v.y * fs1;
                float fs2 = v.x * fs1 + v.y * fc1;
                float p12 = v.z * (1.8658817f * z2 - 1.119529f);
                float p9 = f2a * fs2, p10 = f1b * fs1, p11 = f0c * v.y, p13 = f0c * v.x, p14 = f1b * fc1,
                      p15 = f2a * fc2;
                color = v3_add(color, v3_scale(rd3(c, 27), p9));
                color = v3_add(color, v3_scale(rd3(c, 30), p10));
                color = v3_add(color, v3_scale(rd3(c, 33), p11));
                color = v3_add(color, v3_scale(rd3(c, 36), p12));
                color = v3_add(color, v3_scale(rd3(c, 39), p13));
                color = v3_add(color, v3_scale(rd3(c, 42), p14));
                color = v3_add(color, v3_scale(rd3(c, 45), p15));
                if (degree >= 4) {
                    float f0d = v.z * (-4.683326f * z2 + 2.0071396f);
                    float f1c = 3.3116114f * z2 - 0.47308735f;
                    float f2b = -1.7701308f * v.z;
                    float f3a = 0.62583575f;
                    float fc3 = v.x * fc2 - v.y * fs2;
                    float fs3 = v.x * fs2 + v.y * fc2;
                    float p20 = 1.9843135f * v.z * p12 - 1.0062306f * p6;
                    float p16 = f3a * fs3, p17 = f2b * fs2, p18 = f1c * fs1, p19 = f0d * v.y, p21 = f0d * v.x,
                          p22 = f1c * fc1, p23 = f2b * fc2, p24 = f3a * fc3;
                    color = v3_add(color, v3_scale(rd3(c, 48), p16));
                    color = v3_add(color, v3_scale(rd3(c, 51), p17));
                    color = v3_add(color, v3_scale(rd3(c, 54), p18));
                    color = v3_add(color, v3_scale(rd3(c, 57), p19));
                    color = v3_add(color, v3_scale(rd3(c, 60), p20));
                    color = v3_add(color, v3_scale(rd3(c, 63), p21));
                    color = v3_add(color, v3_scale(rd3(c, 66), p22));
                    color = v3_add(color, v3_scale(rd3(c, 69), p23));
                    color = v3_add(color, v3_scale(rd3(c, 72), p24));
                }
            }
        }
    }
    return color;
}

/* brush-render/src/sh.rs:10-19 */
static uint32_t sh_degree_from_coeffs(uint32_t k) {
    switch (k) { case 1: return 0; case 4: return 1; case 9: return 2; case 16: return 3; case 25: return 4; default: return 0xFFFFFFFFu; }
}

/* ---- sort / scan specs ---- */
void orc_radix_argsort_u32(const uint32_t *keys, const uint32_t *vals, uint32_t n, uint32_t bits,
                           uint32_t *keys_out, uint32_t *vals_out) {
    /* Spec (brush-sort/src/lib.rs:147-151): equal to a stable host argsort.  LSD radix over the
     * low `bits` bits in 8-bit digits; higher bits are ignored exactly as the reference ignores them. */
    uint32_t *ka = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint32_t *va = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint32_t *kb = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint32_t *vb = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    memcpy(ka, keys, sizeof(uint32_t) * n);
    memcpy(va, vals, sizeof(uint32_t) * n);
    for (uint32_t shift = 0; shift < bits; shift += 8) {
        uint32_t width = bits - shift < 8 ? bits - shift : 8;
        uint32_t mask = (1u << width) - 1u;
        uint32_t hist[257];
        memset(hist, 0, sizeof(hist));
        for (uint32_t i = 0; i < n; i++) hist[((ka[i] >> shift) & mask) + 1]++;
        for (uint32_t d = 0; d < 256; d++) hist[d + 1] += hist[d];
        for (uint32_t i = 0; i < n; i++) {
            uint32_t d = (ka[i] >> shift) & mask;
            uint32_t o = hist[d]++;
            kb[o] = ka[i];
            vb[o] = va[i];
        }
        uint32_t *t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    memcpy(keys_out, ka, sizeof(uint32_t) * n);
    memcpy(vals_out, va, sizeof(uint32_t) * n);
    free(ka); free(va); free(kb); free(vb);
}

void orc_inclusive_scan_u32(const uint32_t *in, uint32_t n, uint32_t *out) {
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n; i++) { acc += in[i]; out[i] = acc; }
}

/* ---- smooth cutoff (kernels/helpers.rs:23-47) ---- */
#define ALPHA_CUTOFF_MID (1.0f / 255.0f)
#define ALPHA_CUTOFF_BAND 1.0e-3f
float orc_alpha_cutoff_weight(float alpha) {
    float t = orc_clamp((alpha - (ALPHA_CUTOFF_MID - 0.5f * ALPHA_CUTOFF_BAND)) / ALPHA_CUTOFF_BAND, 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
float orc_alpha_cutoff_weight_deriv(float alpha) {
    float low = ALPHA_CUTOFF_MID - 0.5f * ALPHA_CUTOFF_BAND;
    float high = ALPHA_CUTOFF_MID + 0.5f * ALPHA_CUTOFF_BAND;
    int inside = alpha > low && alpha < high;
    float t = (alpha - low) / ALPHA_CUTOFF_BAND;
    return inside ? (6.0f * t - 6.0f * t * t) / ALPHA_CUTOFF_BAND : 0.0f;
}

static void *xcalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz); }

OrcRender *orc_render_forward(const OrcCamera *cam, uint32_t w, uint32_t h, uint32_t n, uint32_t k,
                              const float *transforms, const float *sh, const float *raw_opac, int mip,
                              const float *bg3, int pass) {
    if (!cam || w == 0 || h == 0) return NULL;
    uint32_t sh_degree = sh_degree_from_coeffs(k);
    if (sh_degree == 0xFFFFFFFFu) return NULL;
    OrcRender *r = (OrcRender *)calloc(1, sizeof(OrcRender));
    r->n = n; r->k = k; r->w = w; r->h = h;
    r->tiles_x = (w + TILE_WIDTH - 1) / TILE_WIDTH;   /* render.rs:30-35 */
    r->tiles_y = (h + TILE_WIDTH - 1) / TILE_WIDTH;
    r->pass = pass; r->mip = mip;
    const uint32_t tiles_x = r->tiles_x, tiles_y = r->tiles_y, num_tiles = tiles_x * tiles_y;
    const int bwd_info = pass != ORC_PASS_FORWARD;
    const int smooth = pass == ORC_PASS_BACKWARD_SMOOTH;

    /* ---- K1 project_forward (kernels/project_forward.rs:20-125) ---- */
    r->intersect_counts = (uint32_t *)xcalloc(n, 4);
    r->max_radius = (float *)xcalloc(n, 4);
    float *depth_of = (float *)xcalloc(n, 4);
    uint8_t *vis_flag = (uint8_t *)xcalloc(n, 1);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t gi = 0; gi < (int64_t)n; gi++) {
        const float *t = transforms + (size_t)gi * 10;
        ovec3 mean_c = world_to_cam(v3(t[0], t[1], t[2]), cam);
        if (!(v3_is_finite(mean_c) && mean_c.z <= 1.0e10f)) continue;
        if (!orc_cam_in_front(mean_c, cam)) continue; /* project_forward.rs:47-61 */
        ovec3 scale = v3(orc_expf(t[7]), orc_expf(t[8]), orc_expf(t[9]));
        if (!v3_is_finite(scale)) continue;
        oquat qu = {t[3], t[4], t[5], t[6]};
        float qn = q_dot(qu, qu);
        if (!(qn >= 1.0e-6f && orc_is_finite(qn))) continue;
        float ro = raw_opac[gi];
        if (!orc_is_finite(ro)) continue;
        oquat quat = q_normalize(qu);
        osym2 raw_cov = orc_calc_cov2d(scale, quat, mean_c, cam);
        float filter_comp;
        osym2 cov = orc_compensate_cov2d(raw_cov, mip, &filter_comp);
        float opac = orc_sigmoid(ro) * filter_comp;
        if (!s2_is_finite(cov)) continue;
        float mx, my;
        orc_cam_project(mean_c, cam, &mx, &my);
        if (!(opac >= 1.0f / 255.0f)) continue;
        float power_threshold = orc_logf(opac * 255.0f);
        osym2 conic = s2_inverse(cov);
        float ex, ey;
        compute_bbox_extent(conic, power_threshold, &ex, &ey);
        if (!(ex >= 0.0f && ey >= 0.0f)) continue;
        float wf = (float)w, hf = (float)h;
        int on_screen = mx + ex > 0.0f && mx - ex < wf && my + ey > 0.0f && my - ey < hf;
        if (!on_screen) continue;
        TileBbox bb = get_tile_bbox(mx, my, ex, ey, tiles_x, tiles_y);
        r->intersect_counts[gi] = count_contributing_tiles(bb, mx, my, conic, power_threshold);
        r->max_radius[gi] = orc_max(ex / wf, ey / hf);
        depth_of[gi] = mean_c.z;
        vis_flag[gi] = 1;
    }
    /* compaction in index order (see header note) */
    uint32_t V = 0;
    for (uint32_t gi = 0; gi < n; gi++) V += vis_flag[gi];
    r->num_visible = V;
    uint32_t *presort_gid = (uint32_t *)xcalloc(V, 4);
    uint32_t *presort_key = (uint32_t *)xcalloc(V, 4);
    {
        uint32_t o = 0;
        for (uint32_t gi = 0; gi < n; gi++)
            if (vis_flag[gi]) { presort_gid[o] = gi; presort_key[o] = orc_f2u(depth_of[gi]); o++; }
    }
    /* ---- depth sort, 32 bits (render.rs:177-184) ---- */
    uint32_t *sorted_key = (uint32_t *)xcalloc(V, 4);
    r->gid_from_cgid = (uint32_t *)xcalloc(V, 4);
    orc_radix_argsort_u32(presort_key, presort_gid, V, 32, sorted_key, r->gid_from_cgid);
    r->depths_sorted = (float *)xcalloc(V, 4);
    for (uint32_t i = 0; i < V; i++) r->depths_sorted[i] = orc_u2f(sorted_key[i]);
    /* ---- int_gather + prefix_sum (render.rs:185-187) ---- */
    uint32_t *compact_counts = (uint32_t *)xcalloc(V, 4);
    for (uint32_t i = 0; i < V; i++) compact_counts[i] = r->intersect_counts[r->gid_from_cgid[i]];
    r->cum_tiles_hit = (uint32_t *)xcalloc(V, 4);
    orc_inclusive_scan_u32(compact_counts, V, r->cum_tiles_hit);
    uint32_t I = V ? r->cum_tiles_hit[V - 1] : 0;
    r->num_intersections = I;

    /* ---- K2 project_visible (kernels/project_visible.rs:22-88) ---- */
    r->projected = (float *)xcalloc((size_t)V * 9, 4);
    ovec3 cam_pos = v3(cam->cam_pos[0], cam->cam_pos[1], cam->cam_pos[2]);
#pragma omp parallel for schedule(static)
    for (int64_t ci = 0; ci < (int64_t)V; ci++) {
        uint32_t gi = r->gid_from_cgid[ci];
        const float *t = transforms + (size_t)gi * 10;
        ovec3 mean = v3(t[0], t[1], t[2]);
        ovec3 scale = v3(orc_expf(t[7]), orc_expf(t[8]), orc_expf(t[9]));
        oquat qu = {t[3], t[4], t[5], t[6]};
        oquat quat = q_normalize(qu);
        ovec3 mean_c = world_to_cam(mean, cam);
        osym2 raw_cov = orc_calc_cov2d(scale, quat, mean_c, cam);
        float filter_comp;
        osym2 cov = orc_compensate_cov2d(raw_cov, mip, &filter_comp);
        float opac = orc_sigmoid(raw_opac[gi]) * filter_comp;
        osym2 conic = s2_inverse(cov);
        float mx, my;
        orc_cam_project(mean_c, cam, &mx, &my);
        ovec3 v = v3_normalize(v3_sub(mean, cam_pos));
        ovec3 raw = orc_sh_to_color(sh + (size_t)gi * k * 3, sh_degree, v);
        float cr = raw.x + 0.5f, cg = raw.y + 0.5f, cb = raw.z + 0.5f;
        float *p = r->projected + (size_t)ci * 9;
        p[0] = mx; p[1] = my; p[2] = conic.c00; p[3] = conic.c01; p[4] = conic.c11; p[5] = opac;
        p[6] = orc_clamp(orc_is_finite(cr) ? cr : 0.0f, -100.0f, 100.0f);
        p[7] = orc_clamp(orc_is_finite(cg) ? cg : 0.0f, -100.0f, 100.0f);
        p[8] = orc_clamp(orc_is_finite(cb) ? cb : 0.0f, -100.0f, 100.0f);
    }

    /* ---- K3 map_gaussians_to_intersect (kernels/map_gaussians.rs:14-80) ---- */
    uint32_t *tile_id_unsorted = (uint32_t *)xcalloc(I, 4);
    uint32_t *cgid_unsorted = (uint32_t *)xcalloc(I, 4);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t ci = 0; ci < (int64_t)V; ci++) {
        const float *p = r->projected + (size_t)ci * 9;
        osym2 conic = {p[2], p[3], p[4]};
        float pt = orc_logf(p[5] * 255.0f);
        float ex, ey;
        compute_bbox_extent(conic, pt, &ex, &ey);
        TileBbox bb = get_tile_bbox(p[0], p[1], ex, ey, tiles_x, tiles_y);
        uint32_t base = ci == 0 ? 0 : r->cum_tiles_hit[ci - 1];
        uint32_t pf_count = r->cum_tiles_hit[ci] - base;
        uint32_t sentinel = tiles_x * tiles_y;
        uint32_t bb_w = bb.max_x - bb.min_x;
        uint32_t num = (bb.max_y - bb.min_y) * bb_w;
        uint32_t hits = 0;
        for (uint32_t i = 0; i < num; i++) {
            uint32_t tx = (i % bb_w) + bb.min_x;
            uint32_t ty = (i / bb_w) + bb.min_y;
            if (will_primitive_contribute(tx, ty, p[0], p[1], conic, pt) && hits < pf_count) {
                tile_id_unsorted[base + hits] = tx + ty * tiles_x;
                cgid_unsorted[base + hits] = (uint32_t)ci;
                hits++;
            }
        }
        for (uint32_t pad = hits; pad < pf_count; pad++) {
            tile_id_unsorted[base + pad] = sentinel;
            cgid_unsorted[base + pad] = (uint32_t)ci;
        }
    }
    /* ---- tile sort (render.rs:228-230): bits = 32 - clz(num_tiles) ---- */
    uint32_t bits = 0;
    while (bits < 32 && (num_tiles >> bits) != 0) bits++;
    r->tile_id_from_isect = (uint32_t *)xcalloc(I, 4);
    r->cgid_from_isect = (uint32_t *)xcalloc(I, 4);
    orc_radix_argsort_u32(tile_id_unsorted, cgid_unsorted, I, bits, r->tile_id_from_isect, r->cgid_from_isect);

    /* ---- K4 get_tile_offsets (get_tile_offset.rs:10-58) ---- */
    r->tile_offsets = (uint32_t *)xcalloc((size_t)num_tiles * 2, 4);
    r->tile_offsets_untrimmed = (uint32_t *)xcalloc((size_t)num_tiles * 2, 4);
    for (uint32_t i = 0; i < I; i++) {
        uint32_t tid = r->tile_id_from_isect[i];
        if (tid < num_tiles) {
            if (i == I - 1) r->tile_offsets[tid * 2 + 1] = i + 1;
            if (i == 0) {
                r->tile_offsets[tid * 2] = 0;
            } else {
                uint32_t prev = r->tile_id_from_isect[i - 1];
                if (tid != prev) {
                    if (prev < num_tiles) r->tile_offsets[prev * 2 + 1] = i;
                    r->tile_offsets[tid * 2] = i;
                }
            }
        }
    }
    memcpy(r->tile_offsets_untrimmed, r->tile_offsets, (size_t)num_tiles * 2 * 4);

    /* ---- K5 rasterize (kernels/rasterize.rs:25-190) ---- */
    if (bwd_info) {
        r->out_img = (float *)xcalloc((size_t)w * h * 4, 4);
        r->visible = (float *)xcalloc(n, 4);
    } else {
        r->out_packed = (uint32_t *)xcalloc((size_t)w * h, 4);
        r->visible = (float *)xcalloc(1, 4);
    }
    float bg_r = bg3 ? bg3[0] : 0.0f, bg_g = bg3 ? bg3[1] : 0.0f, bg_b = bg3 ? bg3[2] : 0.0f;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t tile = 0; tile < (int64_t)num_tiles; tile++) {
        uint32_t range_lo = r->tile_offsets[tile * 2], range_hi = r->tile_offsets[tile * 2 + 1];
        uint32_t tile_x0 = ((uint32_t)tile % tiles_x) * TILE_WIDTH, tile_y0 = ((uint32_t)tile / tiles_x) * TILE_WIDTH;
        uint32_t max_useful = range_lo;
        for (uint32_t ly = 0; ly < TILE_WIDTH; ly++) {
            for (uint32_t lx = 0; lx < TILE_WIDTH; lx++) {
                uint32_t pix_x = tile_x0 + lx, pix_y = tile_y0 + ly;
                if (!(pix_x < w && pix_y < h)) continue;
                float pcx = (float)pix_x + 0.5f, pcy = (float)pix_y + 0.5f;
                float t_acc = 1.0f, pr = 0.0f, pg = 0.0f, pb = 0.0f;
                uint32_t last_useful = range_lo;
                for (uint32_t is = range_lo; is < range_hi; is++) {
                    uint32_t cg = r->cgid_from_isect[is];
                    const float *p = r->projected + (size_t)cg * 9;
                    osym2 conic = {p[2], p[3], p[4]};
                    float sigma = orc_calc_sigma(pcx, pcy, conic, p[0], p[1]);
                    float alpha = orc_min(0.999f, p[5] * orc_expf(-sigma));
                    float w_cut = smooth ? orc_alpha_cutoff_weight(alpha) : (alpha >= ALPHA_CUTOFF_MID ? 1.0f : 0.0f);
                    if (sigma >= 0.0f && w_cut > 0.0f) {
                        float alpha_eff = alpha * w_cut;
                        float next_t = t_acc * (1.0f - alpha_eff);
                        if (next_t <= 1.0e-4f) break; /* done = true: this splat is not blended */
                        if (bwd_info) {
                            /* benign race: every writer stores 1.0f */
                            r->visible[r->gid_from_cgid[cg]] = 1.0f;
                        }
                        float vis = alpha_eff * t_acc;
                        pr += orc_max(p[6], 0.0f) * vis;
                        pg += orc_max(p[7], 0.0f) * vis;
                        pb += orc_max(p[8], 0.0f) * vis;
                        t_acc = next_t;
                        last_useful = is + 1;
                    }
                }
                float fr = pr + t_acc * bg_r, fg = pg + t_acc * bg_g, fb = pb + t_acc * bg_b, fa = 1.0f - t_acc;
                size_t pix_id = (size_t)pix_x + (size_t)pix_y * w;
                if (bwd_info) {
                    float *o = r->out_img + pix_id * 4;
                    o[0] = fr; o[1] = fg; o[2] = fb; o[3] = fa;
                } else {
                    uint32_t R = (uint32_t)orc_clamp(fr * 255.0f, 0.0f, 255.0f);
                    uint32_t G = (uint32_t)orc_clamp(fg * 255.0f, 0.0f, 255.0f);
                    uint32_t B = (uint32_t)orc_clamp(fb * 255.0f, 0.0f, 255.0f);
                    uint32_t A = (uint32_t)orc_clamp(fa * 255.0f, 0.0f, 255.0f);
                    r->out_packed[pix_id] = R | (G << 8) | (B << 16) | (A << 24);
                }
                if (last_useful > max_useful) max_useful = last_useful;
            }
        }
        if (bwd_info) r->tile_offsets[tile * 2 + 1] = max_useful; /* rasterize.rs:183-189 */
    }

    free(depth_of); free(vis_flag); free(presort_gid); free(presort_key); free(sorted_key);
    free(compact_counts); free(tile_id_unsorted); free(cgid_unsorted);
    return r;
}

void orc_render_free(OrcRender *r) {
    if (!r) return;
    free(r->out_img); free(r->out_packed); free(r->visible); free(r->max_radius); free(r->intersect_counts);
    free(r->depths_sorted); free(r->gid_from_cgid); free(r->cum_tiles_hit); free(r->projected);
    free(r->tile_id_from_isect); free(r->cgid_from_isect); free(r->tile_offsets); free(r->tile_offsets_untrimmed);
    free(r);
}

float orc_atan2f_det(float y, float x) { return orc_atan2f(y, x); }
