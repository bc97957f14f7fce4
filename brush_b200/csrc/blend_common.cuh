// blend_common.cuh -- the production (hard alpha cutoff) blend kernels: shared layout and the pair test.
//
// Same tile walk as raster_common.cuh (tile = CTA of 4 warps, warp = 8x8 pixel block, lane = the two
// pixels (x, y) and (x, y+4) of the block, every warp walks the tile's depth-ordered list on its own in
// batches of 32 splats).  What is new here:
//
//  * packed FP32.  The two pixels of a lane run ONE instruction stream, so all per-pixel arithmetic is
//    written on float2 with the sm_100 packed instructions (FFMA2 / FMUL2 / FADD2; per-splat scalars ride
//    in the broadcast operand form), which halves the FMA-pipe instruction count of both blend loops.
//  * forward -> backward hand-off.  The forward kernel records, per (tile, batch of 32, warp), the 32-bit
//    set of splats that changed any pixel of the warp's block (blended OR stopped a pixel), and per
//    (tile, warp) the number of batches it walked before all its pixels saturated.  The backward kernel
//    stages and evaluates exactly those splats: no block test, no vote, no dead iteration, no re-staging of
//    rows the forward proved irrelevant.  Both kernels evaluate the pair test below with the same
//    explicitly rounded operations, so the replayed transmittance is bit-identical to the forward's.
//
// Pair test (rasterize.rs:116-155; the backward's replay rasterize_backwards.rs:279-330), per pixel:
//   d = mean - pixel centre;  s2 = log2(e) * sigma = hx + (cy*dy)*dy + bdx*dy  with hx = (cz*dx)*dx, bdx = cw*dx
//   (cy, cz, cw = log2(e)/2*c, log2(e)/2*a, log2(e)*b are lanes 9..11 of the projected row)
//   g = ex2.approx(-s2);  oa = opac*g;  alpha = min(0.999, oa);  T' = T*(1 - alpha)
//   acts      = pixel not done  &&  s2 >= 0  &&  oa >= 1/255
//   blends    = acts && T' > 1e-4      (T <- T')
//   stops     = acts && T' <= 1e-4     (pixel done; this splat is NOT blended)
#pragma once
#include "raster_common.cuh"

namespace bg {

struct BlendUniforms {
    uint32_t tiles_x, img_w, img_h;
    float bg_r, bg_g, bg_b;
};

__device__ __forceinline__ float2 bcast2(float a) { return make_float2(a, a); }

// index of the first hand-off word of a tile: one uint4-sized group (4 warps) per batch.  Tiles own disjoint
// slot ranges: floor(lo/32) + tile is strictly increasing by at least ceil(len/32) from tile to tile.
__device__ __forceinline__ size_t blend_mask_base(uint32_t range_lo, uint32_t tile) {
    return ((size_t)(range_lo >> 5) + tile) * RASTER_WARPS;
}

// log2(e)-scaled exponent of the pair of pixels of this lane.  npy2 = (-py0, -(py0+4)), dx = mx - px.
__device__ __forceinline__ float2 pair_sigma(float dx, float my, float cy, float cz, float cw, float2 npy2, float2 &dy2) {
    const float hx = __fmul_rn(__fmul_rn(cz, dx), dx);
    const float bdx = __fmul_rn(cw, dx);
    dy2 = __fadd2_rn(bcast2(my), npy2);
    float2 t = __fmul2_rn(bcast2(cy), dy2);
    t = __ffma2_rn(t, dy2, bcast2(hx));
    return __ffma2_rn(bcast2(bdx), dy2, t);
}

}  // namespace bg
