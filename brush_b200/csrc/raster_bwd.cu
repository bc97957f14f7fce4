// raster_bwd.cu -- adjoint of the per-tile blend.
// Replaces rasterize_backwards_kernel (bwd/kernels/rasterize_backwards.rs:100-391).
//
// The reference assigns one thread per splat and walks pixels through shared memory on a diagonal
// schedule with a barrier per step (a design for 32-wide Apple SIMD groups, rasterize_backwards.rs:25-27).
// On B200 the same forward-replay recurrence runs with one thread per pixel (state in registers, same
// warp-autonomous tile walk and block culling as the forward kernel); the per-splat sums over the 32
// pixels of a warp are formed with a 12-shuffle reduce-scatter (10 values -> 10 lanes), then one f32
// atomic (RED) per value.  As in the forward kernel every lane owns two pixels of the warp's 8x8 block,
// so one reduction and one set of atomics serves 64 pixel-splat pairs.
//
// Replay semantics (rasterize_backwards.rs:186-228, 279-383): pixel state starts at
// (final_rgb - T_final*bg, T = 1); per splat, with the forward's skip/stop rules:
//   vis = alpha*T; v_rgb += vis*v_out_rgb (gated on c >= 0); ra = 1/(1-alpha);
//   v_alpha = (sum_k (T*c_k - rem_k) v_out_k) ra + (v_out_a - bg.v_out_rgb) T_final ra;
//   v_sigma = -alpha v_alpha; if o*e^-sigma <= 0.999: v_conic += (0.5 v_sigma dx^2, v_sigma dx dy,
//   0.5 v_sigma dy^2), v_xy += v_sigma (a dx + b dy, b dx + c dy) [dx = mean - pixel],
//   v_opac += v_alpha e^-sigma, refine += |(v_x W, v_y H)| / max(alpha_final, 1e-5);
//   rem -= vis*c; T <- T'.
#include "raster_common.cuh"

namespace bg {

struct RasterBwdUniforms {
    uint32_t tiles_x, img_w, img_h;
    float bg_r, bg_g, bg_b;
};

struct BwdPixel {
    float rem_r, rem_g, rem_b, T;            // running state (forward replay)
    float vo_r, vo_g, vo_b, vo_w, inv_fa;    // per-pixel constants
};

__device__ __forceinline__ BwdPixel load_bwd_pixel(bool inside, size_t pix_id, const float4 *__restrict__ out_img,
                                                   const float4 *__restrict__ v_output, const RasterBwdUniforms &u) {
    BwdPixel p;
    p.rem_r = p.rem_g = p.rem_b = p.T = 0.0f;
    p.vo_r = p.vo_g = p.vo_b = p.vo_w = p.inv_fa = 0.0f;
    if (inside) {  // load_pixel_state, rasterize_backwards.rs:186-228
        const float4 o = __ldg(out_img + pix_id);
        const float4 vo = __ldg(v_output + pix_id);
        const float t_final = 1.0f - o.w;
        p.rem_r = o.x - t_final * u.bg_r;
        p.rem_g = o.y - t_final * u.bg_g;
        p.rem_b = o.z - t_final * u.bg_b;
        p.T = 1.0f;
        p.vo_r = vo.x; p.vo_g = vo.y; p.vo_b = vo.z;
        p.vo_w = (vo.w - (u.bg_r * vo.x + u.bg_g * vo.y + u.bg_b * vo.z)) * t_final;
        p.inv_fa = 1.0f / fmaxf(o.w, 1.0e-5f);
    }
    return p;
}

// Forward-replay test of one pixel against one splat.  Returns whether the pair contributes; on the
// stop rule the pixel's T drops to 0 (rasterize_backwards.rs:318-320).
template <bool SMOOTH>
__device__ __forceinline__ bool bwd_test(BwdPixel &p, float sigma, float opac, float &gaussian, float &oa, float &alpha,
                                         float &w_cut, float &next_T, bool &stop) {
    gaussian = ex2_approx(-sigma);   // sigma arrives scaled by log2(e)
    oa = opac * gaussian;
    alpha = fminf(0.999f, oa);
    w_cut = 1.0f;
    bool contrib;
    if (SMOOTH) {
        w_cut = cutoff_weight(alpha);
        contrib = p.T > 1.0e-4f && sigma >= 0.0f && w_cut > 0.0f;
    } else {
        contrib = p.T > 1.0e-4f && sigma >= 0.0f && alpha >= ALPHA_CUTOFF_MID;
    }
    next_T = p.T * (1.0f - alpha * w_cut);
    stop = contrib && next_T <= 1.0e-4f;
    if (stop) { p.T = 0.0f; contrib = false; }
    return contrib;
}

__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Adds one pair's terms to the lane's partial sums g[0..9] and advances the pixel state.  Branch-free:
// a pair that does not contribute (c == false) is multiplied out, so both pixels of a lane run the same
// instruction stream.  cr/cg/cb = max(colour, 0).
template <bool SMOOTH>
__device__ __forceinline__ void bwd_accumulate(BwdPixel &p, float *g, bool c, float dx, float dy, float ca, float cb,
                                               float cc, float cr, float cg, float cbl,
                                               float gaussian, float oa, float alpha, float w_cut, float next_T,
                                               float img_wf, float img_hf) {
    const float wf = c ? 1.0f : 0.0f;
    const float alpha_eff = alpha * w_cut;
    const float vis = alpha_eff * p.T * wf;
    // colour sums are accumulated ungated and the 0.5 of the conic diagonal is left out: both are per-splat
    // constants, applied once to the lane's partial sums before the reduction
    g[5] = fmaf(vis, p.vo_r, g[5]);
    g[6] = fmaf(vis, p.vo_g, g[6]);
    g[7] = fmaf(vis, p.vo_b, g[7]);
    const float ra = rcp_approx(1.0f - alpha_eff);
    float dot = fmaf(p.T, cr, -p.rem_r) * p.vo_r;
    dot = fmaf(fmaf(p.T, cg, -p.rem_g), p.vo_g, dot);
    dot = fmaf(fmaf(p.T, cbl, -p.rem_b), p.vo_b, dot);
    float v_alpha = (dot + p.vo_w) * ra * wf;
    if (SMOOTH) v_alpha *= (w_cut + alpha * cutoff_weight_deriv(alpha));
    // alpha-saturated pairs keep only the colour gradient (rasterize_backwards.rs:357-372)
    const float v_alpha_g = (oa <= 0.999f) ? v_alpha : 0.0f;
    const float v_sigma = -alpha * v_alpha_g;
    const float vsx = v_sigma * dx, vsy = v_sigma * dy;
    const float vxy_x = fmaf(ca, vsx, cb * vsy);
    const float vxy_y = fmaf(cb, vsx, cc * vsy);
    g[0] += vxy_x;
    g[1] += vxy_y;
    g[2] = fmaf(vsx, dx, g[2]);
    g[3] = fmaf(vsx, dy, g[3]);
    g[4] = fmaf(vsy, dy, g[4]);
    g[8] = fmaf(v_alpha_g, gaussian, g[8]);
    const float sx = vxy_x * img_wf, sy = vxy_y * img_hf;
    g[9] = fmaf(sqrt_approx(fmaf(sx, sx, sy * sy)), p.inv_fa, g[9]);
    p.rem_r = fmaf(-vis, cr, p.rem_r);
    p.rem_g = fmaf(-vis, cg, p.rem_g);
    p.rem_b = fmaf(-vis, cbl, p.rem_b);
    p.T = c ? next_T : p.T;
}

template <bool SMOOTH>
__global__ void __launch_bounds__(RASTER_THREADS)
rasterize_bwd_kernel(const uint32_t *__restrict__ cgid_from_isect, const uint32_t *__restrict__ tile_offsets,
                     const float *__restrict__ projected, const float4 *__restrict__ out_img,
                     const float4 *__restrict__ v_output, float *__restrict__ v_combined, RasterBwdUniforms u) {
    __shared__ __align__(16) float s_rows[RASTER_WARPS][2][WB * ROW];  // per warp, double buffered

    const uint32_t tile = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t range_lo = tile_offsets[tile * 2], range_hi = tile_offsets[tile * 2 + 1];
    if (range_hi <= range_lo) return;
    const uint32_t tile_x0 = (tile % u.tiles_x) * TILE_W, tile_y0 = (tile / u.tiles_x) * TILE_W;
    const uint32_t blk_x0 = tile_x0 + 8u * (wid & 1u), blk_y0 = tile_y0 + 8u * (wid >> 1);
    const uint32_t pix_x = blk_x0 + (lane & 7u), pix_y0 = blk_y0 + (lane >> 3), pix_y1 = pix_y0 + 4u;
    const bool inside0 = pix_x < u.img_w && pix_y0 < u.img_h;
    const bool inside1 = pix_x < u.img_w && pix_y1 < u.img_h;
    const float px = (float)pix_x + 0.5f, py0 = (float)pix_y0 + 0.5f;
    const float rx0 = (float)blk_x0 + 0.5f, rx1 = rx0 + 7.0f, ry0 = (float)blk_y0 + 0.5f, ry1 = ry0 + 7.0f;

    BwdPixel p0 = load_bwd_pixel(inside0, (size_t)pix_x + (size_t)pix_y0 * u.img_w, out_img, v_output, u);
    BwdPixel p1 = load_bwd_pixel(inside1, (size_t)pix_x + (size_t)pix_y1 * u.img_w, out_img, v_output, u);
    const float img_wf = (float)u.img_w, img_hf = (float)u.img_h;

    // reduce-scatter bookkeeping: which of the 10 sums this lane ends up owning
    const bool b4 = lane & 16u, b3 = lane & 8u, b2 = lane & 4u, b1 = lane & 2u;
    const uint32_t idx5 = (b3 ? 3u : 0u) + (b2 ? 2u : 0u) + (b1 ? 1u : 0u);
    const bool owner = !(lane & 1u) && !(b2 && b1) && !(b3 && b2);
    const uint32_t slot = (b4 ? 5u : 0u) + idx5;

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
    auto all_done = [&]() { return __all_sync(0xffffffffu, !(p0.T > 1.0e-4f) && !(p1.T > 1.0e-4f)); };
    if (all_done()) return;  // nothing inside the image in this block
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
        while (bits) {
            const uint32_t s = (uint32_t)__ffs(bits) - 1u;
            bits &= bits - 1u;
            const float *row = rows + s * ROW;
            const float4 A = *reinterpret_cast<const float4 *>(row);      // mx my a b
            const float4 B = *reinterpret_cast<const float4 *>(row + 4);  // c opac r g
            const float4 C = *reinterpret_cast<const float4 *>(row + 8);  // b_col, then log2(e)-scaled c/2, a/2, b
            const float dx = A.x - px, dy0 = A.y - py0, dy1 = dy0 - 4.0f;
            const float hx = (C.z * dx) * dx, bdx = C.w * dx;
            // sigma * log2(e): the exponent of 2^-x; its sign test is the test on sigma
            const float sigma0 = fmaf(bdx, dy0, fmaf(C.y * dy0, dy0, hx));
            const float sigma1 = fmaf(bdx, dy1, fmaf(C.y * dy1, dy1, hx));
            float ga0, oa0, al0, wc0, nt0, ga1, oa1, al1, wc1, nt1;
            bool st0, st1;
            const bool c0 = bwd_test<SMOOTH>(p0, sigma0, B.y, ga0, oa0, al0, wc0, nt0, st0);
            const bool c1 = bwd_test<SMOOTH>(p1, sigma1, B.y, ga1, oa1, al1, wc1, nt1, st1);
            // one vote per splat: saturation of the whole block is only checked between batches (a saturated
            // pixel ignores the remaining splats of its batch, nothing it does is observable)
            if (!__any_sync(0xffffffffu, c0 || c1)) continue;
            float g[10];
#pragma unroll
            for (int i = 0; i < 10; i++) g[i] = 0.0f;
            const float col_b = C.x;
            const float cr = fmaxf(B.z, 0.0f), cg = fmaxf(B.w, 0.0f), cbl = fmaxf(col_b, 0.0f);
            bwd_accumulate<SMOOTH>(p0, g, c0, dx, dy0, A.z, A.w, B.x, cr, cg, cbl, ga0, oa0, al0, wc0, nt0, img_wf, img_hf);
            bwd_accumulate<SMOOTH>(p1, g, c1, dx, dy1, A.z, A.w, B.x, cr, cg, cbl, ga1, oa1, al1, wc1, nt1, img_wf, img_hf);
            g[2] *= 0.5f; g[4] *= 0.5f;
            g[5] = (B.z >= 0.0f) ? g[5] : 0.0f;   // colour VJP gate (rasterize_backwards.rs:335-337)
            g[6] = (B.w >= 0.0f) ? g[6] : 0.0f;
            g[7] = (col_b >= 0.0f) ? g[7] : 0.0f;
            // ---- reduce-scatter 10 values over 32 lanes: 5+3+2+1+1 shuffles
            float a5[6];
#pragma unroll
            for (int i = 0; i < 5; i++) {
                float send = b4 ? g[i] : g[i + 5];
                float keep = b4 ? g[i + 5] : g[i];
                a5[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
            a5[5] = 0.0f;
            float b3v[4];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                float send = b3 ? a5[i] : a5[i + 3];
                float keep = b3 ? a5[i + 3] : a5[i];
                b3v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            b3v[3] = 0.0f;
            float c2[2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                float send = b2 ? b3v[i] : b3v[i + 2];
                float keep = b2 ? b3v[i + 2] : b3v[i];
                c2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            float d1;
            {
                float send = b1 ? c2[0] : c2[1];
                float keep = b1 ? c2[1] : c2[0];
                d1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
            const uint32_t id = __shfl_sync(0xffffffffu, my_id, s);
            if (owner && d1 != 0.0f) atomicAdd(v_combined + (size_t)id * BG_VCOMBINED_STRIDE + slot, d1);
        }
        if (all_done()) break;
        __syncwarp();
    }
    cp_async_wait<0>();
}

cudaError_t launch_rasterize_bwd(cudaStream_t s, bool smooth, uint32_t num_tiles, const uint32_t *cgid_from_isect,
                                 const uint32_t *tile_offsets, const float *projected, const float *out_img,
                                 const float *v_output, float *v_combined, uint32_t tiles_x, uint32_t w, uint32_t h,
                                 const float *bg) {
    RasterBwdUniforms u;
    u.tiles_x = tiles_x; u.img_w = w; u.img_h = h; u.bg_r = bg[0]; u.bg_g = bg[1]; u.bg_b = bg[2];
    if (smooth)
        rasterize_bwd_kernel<true><<<num_tiles, RASTER_THREADS, 0, s>>>(cgid_from_isect, tile_offsets, projected,
                                                             (const float4 *)out_img, (const float4 *)v_output,
                                                             v_combined, u);
    else
        rasterize_bwd_kernel<false><<<num_tiles, RASTER_THREADS, 0, s>>>(cgid_from_isect, tile_offsets, projected,
                                                              (const float4 *)out_img, (const float4 *)v_output,
                                                              v_combined, u);
    return cudaGetLastError();
}

}  // namespace bg
