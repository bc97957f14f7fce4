// raster_fwd.cu -- per-tile front-to-back alpha blend.
// Replaces rasterize_kernel (kernels/rasterize.rs:25-190).  Blend semantics (rasterize.rs:116-181):
//   sigma = 0.5 (a dx^2 + c dy^2) + b dx dy at the pixel centre; alpha = min(0.999, o e^-sigma);
//   skip unless sigma >= 0 and alpha >= 1/255 (smoothstep weight in the test-only variant);
//   T' = T (1 - alpha); if T' <= 1e-4 the pixel is done and this splat is NOT blended;
//   rgb += max(c,0) alpha T; output rgb + T bg, a = 1 - T.
// With bwd_info: rgba f32 output, visible[gid] = 1 for every blended splat, and the tile's range
// end is trimmed to one past the last blended splat (rasterize.rs:183-189).
//
// Bound: FP32 issue + MUFU (ex2), not HBM.  See raster_common.cuh for the layout.
#include "raster_common.cuh"

namespace bg {

struct RasterUniforms {
    uint32_t tiles_x, img_w, img_h;
    float bg_r, bg_g, bg_b;
};

template <bool BWD_INFO, bool SMOOTH>
__global__ void __launch_bounds__(256)
rasterize_fwd_kernel(const uint32_t *__restrict__ cgid_from_isect, uint32_t *__restrict__ tile_offsets,
                     const float *__restrict__ projected, const uint32_t *__restrict__ gid_from_cgid,
                     float4 *__restrict__ out_f32, uint32_t *__restrict__ out_packed, float *__restrict__ visible,
                     RasterUniforms u) {
    __shared__ __align__(16) float s_rows[2][RB * ROW];
    __shared__ uint32_t s_hits[2][8][8];  // [buffer][target warp][word of 32 splats]
    __shared__ uint8_t s_used[RB];
    __shared__ uint32_t s_max_useful;

    const uint32_t tile = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t tile_x0 = (tile % u.tiles_x) * TILE_W, tile_y0 = (tile / u.tiles_x) * TILE_W;
    const uint32_t pix_x = tile_x0 + 8u * (wid & 1u) + (lane & 7u);
    const uint32_t pix_y = tile_y0 + 4u * (wid >> 1) + (lane >> 3);
    const bool inside = pix_x < u.img_w && pix_y < u.img_h;
    const float px = (float)pix_x + 0.5f, py = (float)pix_y + 0.5f;

    const uint32_t range_lo = tile_offsets[tile * 2], range_hi = tile_offsets[tile * 2 + 1];
    if (BWD_INFO && tid == 0) s_max_useful = range_lo;

    float T = 1.0f, acc_r = 0.0f, acc_g = 0.0f, acc_b = 0.0f;
    bool done = !inside;
    uint32_t last_useful = range_lo;

    const uint32_t num_batches = (range_hi - range_lo + RB - 1) / RB;
    uint32_t my_id = 0, next_id = 0;
    auto prefetch = [&](uint32_t b, uint32_t &id_out) {
        uint32_t idx = range_lo + b * RB + tid;
        if (idx < range_hi) {
            uint32_t id = __ldg(cgid_from_isect + idx);
            id_out = id;
            const float *src = projected + (size_t)id * ROW;
            float *dst = &s_rows[b & 1][tid * ROW];
            cp_async16(dst, src);
            cp_async16(dst + 4, src + 4);
            cp_async16(dst + 8, src + 8);
        }
        cp_async_commit();
    };
    if (num_batches > 0) prefetch(0, next_id);

    for (uint32_t b = 0; b < num_batches; b++) {
        const uint32_t buf = b & 1u;
        const uint32_t batch_start = range_lo + b * RB;
        const uint32_t count = min((uint32_t)RB, range_hi - batch_start);
        my_id = next_id;
        if (b + 1 < num_batches) {
            prefetch(b + 1, next_id);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        // ---- per-splat block culling
        {
            uint32_t mask = 0;
            if (tid < count) {
                const float4 A = *reinterpret_cast<const float4 *>(&s_rows[buf][tid * ROW]);
                const float4 B = *reinterpret_cast<const float4 *>(&s_rows[buf][tid * ROW + 4]);
                const float4 Cc = *reinterpret_cast<const float4 *>(&s_rows[buf][tid * ROW + 8]);
                float thr = Cc.y + (SMOOTH ? SMOOTH_THR_EXTRA : 0.0f);
                mask = block_hit_mask(A.x, A.y, A.z, A.w, B.x, thr, (float)tile_x0, (float)tile_y0);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                uint32_t w = __ballot_sync(0xffffffffu, (mask >> k) & 1u);
                if (lane == 0) s_hits[buf][k][wid] = w;
            }
            if (BWD_INFO) s_used[tid] = 0;
        }
        __syncthreads();
        // ---- blend: this warp walks its own hit bits in depth order
        bool warp_done = __all_sync(0xffffffffu, done);
        if (!warp_done) {
#pragma unroll 1
            for (int j = 0; j < 8; j++) {
                uint32_t bits = s_hits[buf][wid][j];
                while (bits) {
                    const uint32_t s = (uint32_t)(j * 32 + __ffs(bits) - 1);
                    bits &= bits - 1;
                    const float *row = &s_rows[buf][s * ROW];
                    const float4 A = *reinterpret_cast<const float4 *>(row);       // mx my a b
                    const float2 B = *reinterpret_cast<const float2 *>(row + 4);   // c opac
                    const float dx = px - A.x, dy = py - A.y;
                    const float sigma = 0.5f * (A.z * dx * dx + B.x * dy * dy) + A.w * dx * dy;
                    const float alpha = fminf(0.999f, B.y * __expf(-sigma));
                    float alpha_eff;
                    bool contrib;
                    if (SMOOTH) {
                        float wc = cutoff_weight(alpha);
                        alpha_eff = alpha * wc;
                        contrib = !done && sigma >= 0.0f && wc > 0.0f;
                    } else {
                        alpha_eff = alpha;
                        contrib = !done && sigma >= 0.0f && alpha >= ALPHA_CUTOFF_MID;
                    }
                    const float next_T = T * (1.0f - alpha_eff);
                    if (contrib && next_T <= 1.0e-4f) { done = true; contrib = false; }
                    if (__any_sync(0xffffffffu, contrib)) {
                        const float2 Cl = *reinterpret_cast<const float2 *>(row + 6);  // r g
                        const float cb = row[8];
                        if (contrib) {
                            const float vis = alpha_eff * T;
                            acc_r += fmaxf(Cl.x, 0.0f) * vis;
                            acc_g += fmaxf(Cl.y, 0.0f) * vis;
                            acc_b += fmaxf(cb, 0.0f) * vis;
                            T = next_T;
                        }
                        if (BWD_INFO && lane == 0) s_used[s] = 1;
                    } else if (__all_sync(0xffffffffu, done)) {
                        bits = 0;
                        j = 8;
                    }
                }
            }
            warp_done = __all_sync(0xffffffffu, done);
        }
        const int any_active = __syncthreads_or(warp_done ? 0 : 1);
        if (BWD_INFO) {
            if (tid < count && s_used[tid]) {
                visible[__ldg(gid_from_cgid + my_id)] = 1.0f;
                last_useful = batch_start + tid + 1;
            }
        }
        if (!any_active) break;
    }
    cp_async_wait<0>();

    if (inside) {
        const float fr = acc_r + T * u.bg_r, fg = acc_g + T * u.bg_g, fb = acc_b + T * u.bg_b, fa = 1.0f - T;
        const size_t pix_id = (size_t)pix_x + (size_t)pix_y * u.img_w;
        if (BWD_INFO) {
            out_f32[pix_id] = make_float4(fr, fg, fb, fa);
        } else {
            uint32_t r = (uint32_t)fminf(fmaxf(fr * 255.0f, 0.0f), 255.0f);
            uint32_t g = (uint32_t)fminf(fmaxf(fg * 255.0f, 0.0f), 255.0f);
            uint32_t bl = (uint32_t)fminf(fmaxf(fb * 255.0f, 0.0f), 255.0f);
            uint32_t a = (uint32_t)fminf(fmaxf(fa * 255.0f, 0.0f), 255.0f);
            out_packed[pix_id] = r | (g << 8) | (bl << 16) | (a << 24);
        }
    }
    if (BWD_INFO) {
        if (last_useful > range_lo) atomicMax(&s_max_useful, last_useful);
        __syncthreads();
        if (tid == 0) tile_offsets[tile * 2 + 1] = s_max_useful;
    }
}

cudaError_t launch_rasterize_fwd(cudaStream_t s, bool bwd_info, bool smooth, uint32_t num_tiles,
                                 const uint32_t *cgid_from_isect, uint32_t *tile_offsets, const float *projected,
                                 const uint32_t *gid_from_cgid, void *out_img, float *visible, uint32_t tiles_x,
                                 uint32_t w, uint32_t h, const float *bg) {
    RasterUniforms u;
    u.tiles_x = tiles_x; u.img_w = w; u.img_h = h; u.bg_r = bg[0]; u.bg_g = bg[1]; u.bg_b = bg[2];
    if (!bwd_info)
        rasterize_fwd_kernel<false, false><<<num_tiles, 256, 0, s>>>(cgid_from_isect, tile_offsets, projected,
                                                                      gid_from_cgid, nullptr, (uint32_t *)out_img,
                                                                      visible, u);
    else if (!smooth)
        rasterize_fwd_kernel<true, false><<<num_tiles, 256, 0, s>>>(cgid_from_isect, tile_offsets, projected,
                                                                     gid_from_cgid, (float4 *)out_img, nullptr, visible, u);
    else
        rasterize_fwd_kernel<true, true><<<num_tiles, 256, 0, s>>>(cgid_from_isect, tile_offsets, projected,
                                                                    gid_from_cgid, (float4 *)out_img, nullptr, visible, u);
    return cudaGetLastError();
}

}  // namespace bg
