// raster_fwd.cu -- per-tile front-to-back alpha blend.
// Replaces rasterize_kernel (kernels/rasterize.rs:25-190).  Blend semantics (rasterize.rs:116-181):
//   sigma = 0.5 (a dx^2 + c dy^2) + b dx dy at the pixel centre; alpha = min(0.999, o e^-sigma);
//   skip unless sigma >= 0 and alpha >= 1/255 (smoothstep weight in the test-only variant);
//   T' = T (1 - alpha); if T' <= 1e-4 the pixel is done and this splat is NOT blended;
//   rgb += max(c,0) alpha T; output rgb + T bg, a = 1 - T.
// With bwd_info: rgba f32 output, visible[gid] = 1 for every blended splat, and the tile's range
// end is trimmed to one past the last blended splat (rasterize.rs:183-189).
//
// Bound: FP32 issue + MUFU (ex2), not HBM (ncu: issue slots ~80% busy, DRAM ~1%).
// The lever is instructions per pixel-splat pair: every lane owns TWO pixels (same column, rows y
// and y+4 of the warp's 8x8 block), so the row loads, loop control, votes and the x-dependent half
// of sigma are paid once per two pairs.  See raster_common.cuh for the tile walk.
#include "raster_common.cuh"

namespace bg {

#ifdef BG_STATS
// development-only counters (build with BG_STATS=1): [0] splats tested, [1] warp-splat iterations,
// [2] iterations with a contribution, [3] contributing pixel-splat pairs
__device__ unsigned long long g_fwd_stats[4];
#endif

struct RasterUniforms {
    uint32_t tiles_x, img_w, img_h;
    float bg_r, bg_g, bg_b;
};

struct FwdPixel {
    float T, r, g, b;
    bool done;
};

template <bool SMOOTH>
__device__ __forceinline__ void fwd_pair(FwdPixel &p, float sigma, float opac, float &vis_out, bool &contrib, bool &stop) {
    const float alpha = fminf(0.999f, opac * ex2_approx(-sigma));   // sigma arrives scaled by log2(e)
    float alpha_eff;
    if (SMOOTH) {
        const float wc = cutoff_weight(alpha);
        alpha_eff = alpha * wc;
        contrib = !p.done && sigma >= 0.0f && wc > 0.0f;
    } else {
        alpha_eff = alpha;
        contrib = !p.done && sigma >= 0.0f && alpha >= ALPHA_CUTOFF_MID;
    }
    const float next_T = p.T * (1.0f - alpha_eff);
    stop = contrib && next_T <= 1.0e-4f;
    if (stop) { p.done = true; contrib = false; }
    vis_out = alpha_eff * p.T;
    if (contrib) p.T = next_T;
}

template <bool BWD_INFO, bool SMOOTH>
__global__ void __launch_bounds__(RASTER_THREADS)
rasterize_fwd_kernel(const uint32_t *__restrict__ cgid_from_isect, uint32_t *__restrict__ tile_offsets,
                     const float *__restrict__ projected, const uint32_t *__restrict__ gid_from_cgid,
                     float4 *__restrict__ out_f32, uint32_t *__restrict__ out_packed, float *__restrict__ visible,
                     RasterUniforms u) {
    __shared__ __align__(16) float s_rows[RASTER_WARPS][2][WB * ROW];  // per warp, double buffered
    __shared__ uint32_t s_max_useful;

    const uint32_t tile = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t tile_x0 = (tile % u.tiles_x) * TILE_W, tile_y0 = (tile / u.tiles_x) * TILE_W;
    const uint32_t blk_x0 = tile_x0 + 8u * (wid & 1u), blk_y0 = tile_y0 + 8u * (wid >> 1);
    const uint32_t pix_x = blk_x0 + (lane & 7u), pix_y0 = blk_y0 + (lane >> 3), pix_y1 = pix_y0 + 4u;
    const bool inside0 = pix_x < u.img_w && pix_y0 < u.img_h;
    const bool inside1 = pix_x < u.img_w && pix_y1 < u.img_h;
    const float px = (float)pix_x + 0.5f, py0 = (float)pix_y0 + 0.5f;
    // rectangle of this warp's pixel centres
    const float rx0 = (float)blk_x0 + 0.5f, rx1 = rx0 + 7.0f, ry0 = (float)blk_y0 + 0.5f, ry1 = ry0 + 7.0f;

    const uint32_t range_lo = tile_offsets[tile * 2], range_hi = tile_offsets[tile * 2 + 1];
    if (BWD_INFO && tid == 0) s_max_useful = range_lo;

#ifdef BG_STATS
    unsigned long long st_tested = 0, st_iters = 0, st_useful = 0, st_pairs = 0;
#endif
    FwdPixel p0, p1;
    p0.T = 1.0f; p0.r = p0.g = p0.b = 0.0f; p0.done = !inside0;
    p1.T = 1.0f; p1.r = p1.g = p1.b = 0.0f; p1.done = !inside1;
    uint32_t last_useful = range_lo;

    const uint32_t num_batches = (range_hi - range_lo + WB - 1) / WB;
    uint32_t next_id = 0;
    auto prefetch = [&](uint32_t b) {
        uint32_t idx = range_lo + b * WB + lane;
        if (idx < range_hi) {
            uint32_t id = __ldg(cgid_from_isect + idx);
            next_id = id;
            const float *src = projected + (size_t)id * ROW;
            float *dst = &s_rows[wid][b & 1u][lane * ROW];
            cp_async16(dst, src);
            cp_async16(dst + 4, src + 4);
            cp_async16(dst + 8, src + 8);
            cp_async16(dst + 12, src + 12);
        }
        cp_async_commit();
    };
    // a warp whose pixels are all outside the image has nothing to blend
    if (num_batches > 0 && !__all_sync(0xffffffffu, p0.done && p1.done)) {
        prefetch(0);
        for (uint32_t b = 0; b < num_batches; b++) {
            const uint32_t batch_start = range_lo + b * WB;
            const uint32_t count = min((uint32_t)WB, range_hi - batch_start);
            const uint32_t my_id = next_id;
            if (b + 1 < num_batches) {
                prefetch(b + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncwarp();
            const float *rows = s_rows[wid][b & 1u];
            bool hit = false;
            if (lane < count) {
                const float4 A = *reinterpret_cast<const float4 *>(rows + lane * ROW);
                const float2 B = *reinterpret_cast<const float2 *>(rows + lane * ROW + 4);
                const float pt = rows[lane * ROW + ROW_PT];
                hit = block_may_hit(A.x, A.y, A.z, A.w, B.x, pt + (SMOOTH ? SMOOTH_THR_EXTRA : 0.0f), rx0, rx1, ry0, ry1);
            }
            uint32_t bits = __ballot_sync(0xffffffffu, hit);
            uint32_t used_l = 0;
#ifdef BG_STATS
            st_tested += count; st_iters += __popc(bits);
#endif
            while (bits) {
                const uint32_t s = (uint32_t)__ffs(bits) - 1u;
                bits &= bits - 1u;
                const float *row = rows + s * ROW;
                const float4 A = *reinterpret_cast<const float4 *>(row);       // mx my a b
                const float4 B = *reinterpret_cast<const float4 *>(row + 4);   // c opac r g
                const float4 C = *reinterpret_cast<const float4 *>(row + 8);   // b_col, then log2(e)-scaled c/2, a/2, b
                const float dx = px - A.x, dy0 = py0 - A.y, dy1 = dy0 + 4.0f;
                const float hx = (C.z * dx) * dx, bdx = C.w * dx;
                // sigma * log2(e): alpha = opacity * 2^-sigma2
                const float sigma0 = fmaf(bdx, dy0, fmaf(C.y * dy0, dy0, hx));
                const float sigma1 = fmaf(bdx, dy1, fmaf(C.y * dy1, dy1, hx));
                float vis0, vis1;
                bool c0, c1, st0, st1;
                fwd_pair<SMOOTH>(p0, sigma0, B.y, vis0, c0, st0);
                fwd_pair<SMOOTH>(p1, sigma1, B.y, vis1, c1, st1);
#ifdef BG_STATS
                { uint32_t m0 = __ballot_sync(0xffffffffu, c0), m1 = __ballot_sync(0xffffffffu, c1);
                  st_pairs += __popc(m0) + __popc(m1); st_useful += (m0 | m1) ? 1 : 0; }
#endif
                // No votes inside the splat loop: after the exact block cull nearly every surviving splat has a
                // contributing lane, so the blend runs unconditionally (weights are zero where it does not apply),
                // each lane remembers which splats its pixels used, and "all pixels saturated" is checked once per
                // batch (saturated pixels ignore the remaining splats of the batch; nothing they do is observable).
                const float cr = fmaxf(B.z, 0.0f), cg = fmaxf(B.w, 0.0f), cb = fmaxf(C.x, 0.0f);
                const float v0 = c0 ? vis0 : 0.0f, v1 = c1 ? vis1 : 0.0f;
                p0.r = fmaf(cr, v0, p0.r); p0.g = fmaf(cg, v0, p0.g); p0.b = fmaf(cb, v0, p0.b);
                p1.r = fmaf(cr, v1, p1.r); p1.g = fmaf(cg, v1, p1.g); p1.b = fmaf(cb, v1, p1.b);
                if (c0 || c1) used_l |= 1u << s;
            }
            const uint32_t used = __reduce_or_sync(0xffffffffu, used_l);
            if (BWD_INFO && ((used >> lane) & 1u)) {
                visible[__ldg(gid_from_cgid + my_id)] = 1.0f;
                last_useful = batch_start + lane + 1;
            }
            if (__all_sync(0xffffffffu, p0.done && p1.done)) break;
            __syncwarp();  // all lanes are done with this buffer before the next prefetch overwrites its twin
        }
        cp_async_wait<0>();
    }

#ifdef BG_STATS
    if (lane == 0) {
        atomicAdd(&g_fwd_stats[0], st_tested); atomicAdd(&g_fwd_stats[1], st_iters);
        atomicAdd(&g_fwd_stats[2], st_useful); atomicAdd(&g_fwd_stats[3], st_pairs);
    }
#endif
    auto write_pixel = [&](const FwdPixel &p, uint32_t pix_y) {
        const float fr = p.r + p.T * u.bg_r, fg = p.g + p.T * u.bg_g, fb = p.b + p.T * u.bg_b, fa = 1.0f - p.T;
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
    };
    if (inside0) write_pixel(p0, pix_y0);
    if (inside1) write_pixel(p1, pix_y1);
    if (BWD_INFO) {
        // one block barrier, after all blending: publish the trimmed range end
        uint32_t m = last_useful;
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        __syncthreads();
        if (lane == 0 && m > range_lo) atomicMax(&s_max_useful, m);
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
        rasterize_fwd_kernel<false, false><<<num_tiles, RASTER_THREADS, 0, s>>>(
            cgid_from_isect, tile_offsets, projected, gid_from_cgid, nullptr, (uint32_t *)out_img, visible, u);
    else if (!smooth)
        rasterize_fwd_kernel<true, false><<<num_tiles, RASTER_THREADS, 0, s>>>(
            cgid_from_isect, tile_offsets, projected, gid_from_cgid, (float4 *)out_img, nullptr, visible, u);
    else
        rasterize_fwd_kernel<true, true><<<num_tiles, RASTER_THREADS, 0, s>>>(
            cgid_from_isect, tile_offsets, projected, gid_from_cgid, (float4 *)out_img, nullptr, visible, u);
    return cudaGetLastError();
}

#ifdef BG_STATS
extern "C" int bg_debug_fwd_stats(unsigned long long *out4, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out4, g_fwd_stats, sizeof(g_fwd_stats));
    if (e == cudaSuccess && reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        e = cudaMemcpyToSymbol(g_fwd_stats, z, sizeof(z));
    }
    return (int)e;
}
#endif

}  // namespace bg
