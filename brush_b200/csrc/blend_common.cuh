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
#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, nothing links libcuda)

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

// ---- TMA staging of a batch of projected rows (north_star: "TMA staging of each tile's sorted Gaussian slice").
// `projected` is described to the TMA unit as a 2-D tensor [rows][16 f32] (one 64-byte row per visible Gaussian, box
// 16 x 1).  A tile's slice is an index list, not a contiguous range, so the copy is the sm_100 gather form:
// cp.async.bulk.tensor.2d ... tile::gather4 takes FOUR row coordinates and lands the four rows as 256 contiguous bytes
// of shared memory (SASS UTMALDG.2D.GATHER4).  One elected lane issues ceil(count/4) of them per batch against one
// mbarrier (complete_tx), the other lanes only park the row ids in shared memory first; nothing occupies the LSU
// pipe or the register file for the copy.
__device__ __forceinline__ void tma_gather4(void *dst_smem, const CUtensorMap *tm, uint4 rows, unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_u32(dst_smem)), "l"(tm), "r"(0), "r"((int)rows.x), "r"((int)rows.y), "r"((int)rows.z), "r"((int)rows.w), "r"(smem_u32(bar))
        : "memory");
}

// Per-warp staging state: two buffers of 32 dense 64-byte rows, the ids of the rows in flight, one mbarrier each.
struct __align__(128) BlendStage {
    float rows[2][WB * ROW];
    uint32_t ids[2][WB];
    unsigned long long bar[2];
};

// Issues the copy of `count` rows (ids already compacted in st.ids[buf][0..count), padded to a multiple of four with
// a valid id) into st.rows[buf].  Call with the whole warp converged; returns after the elected lane has issued.
__device__ __forceinline__ void stage_rows_tma(BlendStage &st, uint32_t buf, uint32_t count, const CUtensorMap *tm, uint32_t lane) {
    __syncwarp();   // ids visible to the elected lane
    if (lane == 0 && count > 0) {
        const uint32_t groups = (count + 3u) >> 2;
        mbar_expect_tx(&st.bar[buf], groups * 256u);
        for (uint32_t g = 0; g < groups; g++)
            tma_gather4(&st.rows[buf][g * 4 * ROW], tm, *reinterpret_cast<const uint4 *>(&st.ids[buf][g * 4]), &st.bar[buf]);
    }
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
