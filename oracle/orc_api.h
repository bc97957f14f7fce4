/*
 * oracle/orc_api.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C entry points of the CPU oracle: a restatement of the reference's hot path
 * (brush-render forward/backward, brush-sort, brush-prefix-sum, brush-loss,
 * brush-train AdamScaled) following the reference files cited at each
 * function.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
 * load this library.  It is the checker, never the product.
 *
 * Parity pins: oracle/README.md (golden vectors of
 * crates/brush-bench-test/test_cases, finite-difference suite restated from
 * crates/brush-bench-test/tests/finite_diff.rs).
 */
#ifndef ORC_API_H
#define ORC_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Host mirror of ProjectUniforms (brush-render/src/kernels/types.rs:51-80,
 * brush-render/src/render.rs:70-99).  viewmat: 3x4 column major, column i at
 * viewmat[3*i..3*i+3], column 3 = translation. */
typedef struct {
    float viewmat[12];
    float fx, fy, cx, cy;
    float cam_pos[3];
    float lim_pos_x, lim_pos_y, lim_neg_x, lim_neg_y;
    float half_max_render_fov;
    uint32_t camera_model;  /* 0 pinhole, 1 Kannala-Brandt 4, 2 radial-tangential 8, 3 thin-prism fisheye */
    float model_params[8];  /* distortion coefficients, see orc_camera.h */
} OrcCamera;

enum { ORC_PASS_FORWARD = 0, ORC_PASS_BACKWARD = 1, ORC_PASS_BACKWARD_SMOOTH = 2 };

/* Mirror of RenderOutput (brush-render/src/render_aux.rs:16-28). All arrays
 * are owned by the struct; free with orc_render_free. */
typedef struct {
    uint32_t n, k, w, h, tiles_x, tiles_y;
    uint32_t num_visible, num_intersections;
    int pass, mip;
    float *out_img;            /* [h,w,4] f32 (pass != forward) */
    uint32_t *out_packed;      /* [h,w] rgba8 (pass == forward) */
    float *visible;            /* [n] */
    float *max_radius;         /* [n] */
    uint32_t *intersect_counts;/* [n] by global gid */
    float *depths_sorted;      /* [V] */
    uint32_t *gid_from_cgid;   /* [V] global_from_compact_gid */
    uint32_t *cum_tiles_hit;   /* [V] inclusive */
    float *projected;          /* [V,9] */
    uint32_t *tile_id_from_isect; /* [I] sorted */
    uint32_t *cgid_from_isect; /* [I] sorted */
    uint32_t *tile_offsets;    /* [tiles_y,tiles_x,2], end trimmed when pass != forward */
    uint32_t *tile_offsets_untrimmed; /* [tiles_y,tiles_x,2] as written by get_tile_offsets */
} OrcRender;

OrcRender *orc_render_forward(const OrcCamera *cam, uint32_t w, uint32_t h, uint32_t n, uint32_t k,
                              const float *transforms, const float *sh, const float *raw_opac,
                              int mip, const float *bg3, int pass);
void orc_render_free(OrcRender *r);

/* rasterize_bwd (brush-render/src/bwd/render_bwd.rs:22-99). v_combined [V,10], zeroed here. */
void orc_rasterize_backward(const OrcRender *r, const float *bg3, const float *v_output /*[h,w,4]*/,
                            int smooth, float *v_combined);
/* project_bwd (brush-render/src/bwd/render_bwd.rs:102-171). Dense outputs, zeroed here. */
void orc_project_backward(const OrcCamera *cam, const OrcRender *r, const float *transforms, const float *sh,
                          const float *raw_opac, const float *v_combined, float *v_transforms /*[n,10]*/,
                          float *v_sh /*[n,k,3]*/, float *v_raw_opac /*[n]*/, float *v_refine /*[n]*/);

/* brush-sort/src/lib.rs:16-125 spec: stable ascending on the low `bits` bits. */
void orc_radix_argsort_u32(const uint32_t *keys, const uint32_t *vals, uint32_t n, uint32_t bits,
                           uint32_t *keys_out, uint32_t *vals_out);
/* brush-prefix-sum/src/lib.rs:11-89 : inclusive scan. */
void orc_inclusive_scan_u32(const uint32_t *in, uint32_t n, uint32_t *out);

/* brush-loss/src/lib.rs:180-359 / 370-661. pred, out: [c,h,w]; gt_packed [h,w]. bg3 may be NULL (no composite). */
void orc_image_loss_forward(const float *pred_chw, const uint32_t *gt_packed, uint32_t c, uint32_t h, uint32_t w,
                            float l1_w, float ssim_w, const float *bg3, int mask, float *loss_map);
void orc_image_loss_backward(const float *pred_chw, const uint32_t *gt_packed, const float *dl_dmap, uint32_t c,
                             uint32_t h, uint32_t w, float l1_w, float ssim_w, const float *bg3, int mask,
                             float *dl_dpred);

/* brush-train/src/adam_scaled.rs:75-165. p,g,m: [rows,cols]; v: [rows,cols] or [rows] when reduce_v.
 * lr_scale_per_col may be NULL. t = step index after increment (1 on the first step, which
 * initialises the moments). */
void orc_adam_step(float *p, const float *g, float *m, float *v, uint64_t rows, uint32_t cols,
                   const float *lr_scale_per_col, float lr, float beta1, float beta2, float eps, int t,
                   int reduce_v);

/* Mip-Splatting 3D filter floor: compute_min_scale (brush-train/src/train.rs:102-125), fold_min_scale
 * (brush-render/src/gaussian_splats.rs:86-111) and the reverse-mode chain of the fold (in place on the
 * gradients w.r.t. the folded values).  view_cams: [views,4] = x, y, z, focal_px. */
void orc_compute_min_scale(const float *transforms, uint32_t n, const float *view_cams, uint32_t views, float factor,
                           float *f_out);
void orc_fold_min_scale_fwd(const float *transforms, const float *raw_opac, const float *f, uint32_t n,
                            float *transforms_out, float *raw_opac_out);
void orc_fold_min_scale_bwd(const float *transforms, const float *raw_opac, const float *f, uint32_t n,
                            float *v_transforms, float *v_raw_opac);

float orc_expf_det(float x);
float orc_logf_det(float x);
float orc_atan2f_det(float y, float x); /* y >= 0 */
int orc_num_threads(void);
void orc_set_num_threads(int n);   /* launchers such as torchrun export OMP_NUM_THREADS=1 */

#ifdef __cplusplus
}
#endif
#endif
