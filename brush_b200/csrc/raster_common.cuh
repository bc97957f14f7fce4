// raster_common.cuh -- pieces shared by the forward and backward blend kernels.
//
// Tile = 16x16 pixels = one CTA of 128 threads.  Warp k owns the 8x8 pixel block with origin
//   (8*(k&1), 8*(k>>1)); lane l owns the two pixels (l&7, l>>3) and (l&7, (l>>3)+4) of the block.
// Every warp walks the tile's depth-ordered splat list on its own, 32 splats at a time: lane l
// fetches the 48-byte projected row of splat l with three 16-byte cp.async (LDGSTS) into the
// warp's double-buffered shared staging area, tests that splat against the warp's 8x8 block
// (exact minimum of the conic over the rectangle of pixel centres), and the ballot of the
// survivors is then consumed bit by bit, in order.  Blend order -- and therefore the result -- is
// exactly the reference's, pairs that cannot reach alpha >= 1/255 anywhere in the block are never
// evaluated, and no block-wide barrier sits inside the loop: a warp whose 32 pixels are saturated
// retires at once (a warp's 64 pixels), the next batch's loads overlap the current batch's math.
#pragma once
#include "bg_common.cuh"

namespace bg {

constexpr int RASTER_WARPS = 4;
constexpr int RASTER_THREADS = RASTER_WARPS * 32;
constexpr int WB = 32;                    // splats per warp batch
constexpr int ROW = BG_PROJECTED_STRIDE;  // 16 floats
constexpr int ROW_PT = 12;                // lane of ln(255 opacity), the block-cull threshold

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// True when the splat may contribute to some pixel centre of the rectangle [x0,x1] x [y0,y1].
// sigma(p) = 0.5 (p-m)^T C (p-m) is convex for a positive definite conic; with the centre outside
// the box its minimum over the box lies on a face visible from the centre, so at most two clamped
// 1-D minimisations give the exact minimum.  A pixel passes the alpha test only if sigma <= thr
// (thr = ln(opacity / alpha_min)); the comparison carries a margin for the rounding of both sides.
__device__ __forceinline__ bool block_may_hit(float mx, float my, float a, float b, float c, float thr, float x0,
                                              float x1, float y0, float y1) {
    // anything but a positive definite conic (rounding at extreme scales, NaN): no culling
    if (!(a > 0.0f && c > 0.0f && a * c > b * b)) return true;
    const float xc = fminf(fmaxf(mx, x0), x1);
    const float yc = fminf(fmaxf(my, y0), y1);
    const bool out_x = xc != mx, out_y = yc != my;
    if (!(out_x || out_y)) return true;
    float best = 3.0e38f, err = 0.0f;
    if (out_x) {  // face x = xc, free y
        float dx = xc - mx;
        float ys = fminf(fmaxf(my - (b / c) * dx, y0), y1);
        float dy = ys - my;
        float t0 = a * dx * dx, t1 = c * dy * dy, t2 = b * dx * dy;
        float s = 0.5f * (t0 + t1) + t2;
        if (s < best) { best = s; err = fabsf(t0) + fabsf(t1) + 2.0f * fabsf(t2); }
    }
    if (out_y) {  // face y = yc, free x
        float dy = yc - my;
        float xs = fminf(fmaxf(mx - (b / a) * dy, x0), x1);
        float dx = xs - mx;
        float t0 = a * dx * dx, t1 = c * dy * dy, t2 = b * dx * dy;
        float s = 0.5f * (t0 + t1) + t2;
        if (s < best) { best = s; err = fabsf(t0) + fabsf(t1) + 2.0f * fabsf(t2); }
    }
    return !(best > thr + 0.05f + 4.0e-6f * err);  // NaN compares false -> kept
}

__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

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
