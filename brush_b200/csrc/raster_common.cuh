// raster_common.cuh -- pieces shared by the forward and backward blend kernels.
//
// Tile = 16x16 pixels = one CTA of 256 threads.  Warp k owns the 8x4 pixel block
//   origin (8*(k&1), 4*(k>>1)), lane l -> pixel (l&7, l>>3) inside the block.
// The tile's depth-ordered splat list is consumed in batches of 256: every thread fetches one
// 48-byte projected row with three 16-byte cp.async (LDGSTS) into a double-buffered shared
// staging area, then tests "its" splat against the eight 8x4 blocks and publishes one hit bit
// per block.  Each warp afterwards walks only the set bits of its own hit words, in order, so
// the blend order -- and therefore the result -- is exactly the reference's, while pairs that
// cannot reach alpha >= 1/255 anywhere in the block are never evaluated.
#pragma once
#include "bg_common.cuh"

namespace bg {

constexpr int RB = 256;                   // splats per staged batch
constexpr int ROW = BG_PROJECTED_STRIDE;  // 12 floats

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Exact minimum of sigma(p) = 0.5 (p-m)^T C (p-m) over the rectangle of pixel centres
// [x0,x1] x [y0,y1] (C = conic, symmetric).  The minimiser of a convex quadratic over a box whose
// centre lies outside is on a face that is visible from the centre, so at most two clamped 1-D
// minimisations are needed.  Also returns a bound on the rounding error of the evaluation.
__device__ __forceinline__ float min_sigma_rect(float mx, float my, float a, float b, float c, float x0, float x1,
                                                float y0, float y1, float &err) {
    float xc = fminf(fmaxf(mx, x0), x1);
    float yc = fminf(fmaxf(my, y0), y1);
    bool out_x = xc != mx, out_y = yc != my;
    float best = 0.0f;
    err = 0.0f;
    if (out_x || out_y) {
        best = 3.0e38f;
        if (out_x) {  // face x = xc, free y
            float dx = xc - mx;
            float ys = (c > 0.0f) ? my - (b / c) * dx : yc;
            ys = fminf(fmaxf(ys, y0), y1);
            float dy = ys - my;
            float t0 = a * dx * dx, t1 = c * dy * dy, t2 = b * dx * dy;
            float s = 0.5f * (t0 + t1) + t2;
            if (s < best) { best = s; err = fabsf(t0) + fabsf(t1) + 2.0f * fabsf(t2); }
        }
        if (out_y) {  // face y = yc, free x
            float dy = yc - my;
            float xs = (a > 0.0f) ? mx - (b / a) * dy : xc;
            xs = fminf(fmaxf(xs, x0), x1);
            float dx = xs - mx;
            float t0 = a * dx * dx, t1 = c * dy * dy, t2 = b * dx * dy;
            float s = 0.5f * (t0 + t1) + t2;
            if (s < best) { best = s; err = fabsf(t0) + fabsf(t1) + 2.0f * fabsf(t2); }
        }
    }
    return best;
}

// 8-bit mask: bit k set <=> the splat may contribute to some pixel of block k of the tile.
// `thr` = ln(opac / alpha_min): a pixel can only pass the alpha test when sigma <= thr.
__device__ __forceinline__ uint32_t block_hit_mask(float mx, float my, float a, float b, float c, float thr,
                                                   float tile_x0, float tile_y0) {
    // Only a positive definite conic makes the box minimisation valid; anything else (rounding
    // at extreme scales, NaN) skips the culling and falls back to the per-pixel test alone.
    if (!(a > 0.0f && c > 0.0f && a * c > b * b)) return 0xFFu;
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float x0 = tile_x0 + (float)(8 * (k & 1)) + 0.5f;
        float y0 = tile_y0 + (float)(4 * (k >> 1)) + 0.5f;
        float err;
        float s = min_sigma_rect(mx, my, a, b, c, x0, x0 + 7.0f, y0, y0 + 3.0f, err);
        // keep unless clearly above the threshold (NaN compares false -> kept)
        bool cull = s > thr + 0.05f + 4.0e-6f * err;
        m |= cull ? 0u : (1u << k);
    }
    return m;
}

// kernels/helpers.rs:26-47 (test-only smooth cutoff)
__device__ __forceinline__ float cutoff_weight(float alpha) {
    float t = fminf(fmaxf((alpha - (ALPHA_CUTOFF_MID - 0.5f * ALPHA_CUTOFF_BAND)) / ALPHA_CUTOFF_BAND, 0.0f), 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
__device__ __forceinline__ float cutoff_weight_deriv(float alpha) {
    const float low = ALPHA_CUTOFF_MID - 0.5f * ALPHA_CUTOFF_BAND, high = ALPHA_CUTOFF_MID + 0.5f * ALPHA_CUTOFF_BAND;
    bool inside = alpha > low && alpha < high;
    float t = (alpha - low) / ALPHA_CUTOFF_BAND;
    return inside ? (6.0f * t - 6.0f * t * t) / ALPHA_CUTOFF_BAND : 0.0f;
}
// ln(alpha_cutoff_mid / smallest alpha with non-zero weight) for the smooth variant
constexpr float SMOOTH_THR_EXTRA = 0.1365f;

}  // namespace bg
