// blend_bwd.cu -- adjoint of the per-tile blend, hard alpha cutoff (the production path).
// Replaces rasterize_backwards_kernel (bwd/kernels/rasterize_backwards.rs:100-391); the smooth-cutoff test
// variant stays in raster_bwd.cu.
//
// Replay semantics (rasterize_backwards.rs:186-228, 279-383): pixel state starts at (final_rgb - T_final*bg, T = 1);
// per splat, with the forward's skip/stop rules:
//   vis = alpha*T; v_rgb += vis*v_out_rgb (gated on c >= 0); ra = 1/(1-alpha);
//   v_alpha = (sum_k (T*c_k - rem_k) v_out_k) ra + (v_out_a - bg.v_out_rgb) T_final ra;
//   v_sigma = -alpha v_alpha; if o*e^-sigma <= 0.999: v_conic += (0.5 v_sigma dx^2, v_sigma dx dy, 0.5 v_sigma dy^2),
//   v_xy += v_sigma (a dx + b dy, b dx + c dy) [dx = mean - pixel], v_opac += v_alpha e^-sigma,
//   refine += |(v_x W, v_y H)| / max(alpha_final, 1e-5);  rem -= vis*c; T <- T'.
//
// Structure.  One thread per pixel pair as in the forward (state in registers), but the walk is driven by the
// forward's hand-off (blend_common.cuh): per batch the warp reads the 32-bit set of splats that changed its block,
// stages ONLY those rows (compacted, by TMA tile::gather4: blend_common.cuh), and runs a plain counted loop over them -- no block test, no vote,
// no dead iteration.  The staging lane also forms the per-splat constants once (clamped colour, -opacity, the ten
// post-reduction factors), so the loop body is per-pixel work only, written on float2 with packed FP32 (FFMA2 /
// FMUL2 / FADD2; splat scalars in the broadcast operand form).  Signs are chosen so that no negation is ever an
// instruction: the loop accumulates -vis, -v_alpha, -v_sigma and the factors put the signs back.
// v_opac uses e^-sigma = alpha/opacity on the unsaturated pairs (the only ones that count), i.e. it is
// -(sum v_sigma)/opacity: one factor per splat instead of one FMA per pair.
// The 10 per-splat sums over the warp's 64 pixels are formed with a 12-shuffle reduce-scatter, one RED.F32 each.
#include "blend_common.cuh"

namespace bg {

constexpr int BROW_ID = 12;       // lane of the staged row that receives the compact Gaussian id (bits)
constexpr int BFACT = 12;         // floats per row of the side table: ten post-reduction factors (+2 pad)

__device__ __forceinline__ float rcp_approx_f(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx_f(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// stats[0] warp-splat iterations, [1] pixel-splat pairs that blended, [2] pairs that stopped a pixel
template <bool STATS>
__global__ void __launch_bounds__(RASTER_THREADS)
blend_bwd_kernel(const __grid_constant__ CUtensorMap tm_projected, const uint32_t *__restrict__ cgid_from_isect,
                 const uint32_t *__restrict__ tile_offsets, const float4 *__restrict__ out_img,
                 const float4 *__restrict__ v_output, const uint32_t *__restrict__ live_masks,
                 const uint32_t *__restrict__ warp_batches, float *__restrict__ v_combined,
                 unsigned long long *__restrict__ stats, BlendUniforms u) {
    __shared__ BlendStage s_stage[RASTER_WARPS];                         // per warp, double buffered rows (TMA destination)
    __shared__ __align__(16) float s_fact[RASTER_WARPS][2][WB * BFACT];  // post-reduction factors of the staged rows

    const uint32_t tile = blockIdx.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    const uint32_t num_batches = __ldg(warp_batches + tile * RASTER_WARPS + wid);
    BlendStage &st = s_stage[wid];
    if (lane == 0) { mbar_init(&st.bar[0], 1); mbar_init(&st.bar[1], 1); }
    __syncthreads();   // barriers initialised before any copy is issued
    if (num_batches == 0) return;
    uint32_t phase_bits = 0u;   // bit b: parity the next wait on buffer b expects
    const uint32_t range_lo = tile_offsets[tile * 2];
    const uint32_t tile_x0 = (tile % u.tiles_x) * TILE_W, tile_y0 = (tile / u.tiles_x) * TILE_W;
    const uint32_t blk_x0 = tile_x0 + 8u * (wid & 1u), blk_y0 = tile_y0 + 8u * (wid >> 1);
    const uint32_t pix_x = blk_x0 + (lane & 7u), pix_y0 = blk_y0 + (lane >> 3), pix_y1 = pix_y0 + 4u;
    const bool inside0 = pix_x < u.img_w && pix_y0 < u.img_h;
    const bool inside1 = pix_x < u.img_w && pix_y1 < u.img_h;
    const float px = (float)pix_x + 0.5f, py0 = (float)pix_y0 + 0.5f;
    const float2 npy2 = make_float2(-py0, -(py0 + 4.0f));
    const float img_wf = (float)u.img_w, img_hf = (float)u.img_h;
    const float hw = img_hf / img_wf, hw2 = hw * hw;   // |(vx W, vy H)| = W sqrt(vx^2 + (H/W)^2 vy^2)

    // ---- pixel state (load_pixel_state, rasterize_backwards.rs:186-228), as pairs
    float2 T2 = make_float2(0.0f, 0.0f);                       // 0 outside the image: nothing ever blends
    float2 rem_r = T2, rem_g = T2, rem_b = T2;                 // colour still to come (positive)
    float2 vo_r = T2, vo_g = T2, vo_b = T2, nvo_w = T2, ifa = T2;
    {
        auto load = [&](bool inside, uint32_t pix_y, float &T, float &rr, float &rg, float &rb, float &vr, float &vg,
                        float &vb, float &nvw, float &inv) {
            if (!inside) return;
            const size_t pix_id = (size_t)pix_x + (size_t)pix_y * u.img_w;
            const float4 o = __ldg(out_img + pix_id);
            const float4 vo = __ldg(v_output + pix_id);
            const float t_final = 1.0f - o.w;
            T = 1.0f;
            rr = o.x - t_final * u.bg_r; rg = o.y - t_final * u.bg_g; rb = o.z - t_final * u.bg_b;
            vr = vo.x; vg = vo.y; vb = vo.z;
            nvw = -((vo.w - (u.bg_r * vo.x + u.bg_g * vo.y + u.bg_b * vo.z)) * t_final);
            inv = img_wf / fmaxf(o.w, 1.0e-5f);
        };
        load(inside0, pix_y0, T2.x, rem_r.x, rem_g.x, rem_b.x, vo_r.x, vo_g.x, vo_b.x, nvo_w.x, ifa.x);
        load(inside1, pix_y1, T2.y, rem_r.y, rem_g.y, rem_b.y, vo_r.y, vo_g.y, vo_b.y, nvo_w.y, ifa.y);
    }

    // reduce-scatter bookkeeping: which of the 10 sums this lane ends up owning
    const bool b4 = lane & 16u, b3 = lane & 8u, b2 = lane & 4u, b1 = lane & 2u;
    const uint32_t idx5 = (b3 ? 3u : 0u) + (b2 ? 2u : 0u) + (b1 ? 1u : 0u);
    const bool owner = !(lane & 1u) && !(b2 && b1) && !(b3 && b2);
    const uint32_t slot = (b4 ? 5u : 0u) + idx5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    const size_t mbase = blend_mask_base(range_lo, tile) + wid;
    auto load_mask = [&](uint32_t b) -> uint32_t {
        return b < num_batches ? __ldg(live_masks + mbase + (size_t)b * RASTER_WARPS) : 0u;
    };
    // stage the rows of batch b selected by mask m, compacted in list order, into buffer b&1: the lanes park the row
    // ids (padded to a multiple of four with the last one), one elected lane issues the TMA gathers
    auto stage = [&](uint32_t b, uint32_t m) -> uint32_t {
        uint32_t id = 0;
        const uint32_t n = (uint32_t)__popc(m);
        if (n == 0) return 0u;
        if ((m >> lane) & 1u) {
            id = __ldg(cgid_from_isect + range_lo + b * WB + lane);
            st.ids[b & 1u][__popc(m & lt_mask)] = id;
        }
        const uint32_t last = __shfl_sync(0xffffffffu, id, 31u - (uint32_t)__clz(m));
        if (lane < 3u && n + lane < ((n + 3u) & ~3u)) st.ids[b & 1u][n + lane] = last;
        stage_rows_tma(st, b & 1u, n, &tm_projected, lane);
        return id;
    };
    unsigned long long st_iter = 0, st_blend = 0, st_stop = 0;
    uint32_t m_cur = load_mask(0), m_next = load_mask(1);
    uint32_t id_next = stage(0, m_cur);
    for (uint32_t b = 0; b < num_batches; b++) {
        const uint32_t m = m_cur, my_id = id_next;
        m_cur = m_next;
        m_next = load_mask(b + 2);
        id_next = stage(b + 1, m_cur);   // (an all-zero mask stages nothing)
        const uint32_t n = (uint32_t)__popc(m);
        if (n == 0) continue;
        mbar_wait(&st.bar[b & 1u], (phase_bits >> (b & 1u)) & 1u);
        phase_bits ^= 1u << (b & 1u);
        float *rows = st.rows[b & 1u];
        float *fact = s_fact[wid][b & 1u];
        if ((m >> lane) & 1u) {
            // per-splat constants, formed once by the lane that parked the row's id
            const uint32_t slot_r = (uint32_t)__popc(m & lt_mask);
            float *mine = rows + slot_r * ROW;
            const float4 B = *reinterpret_cast<const float4 *>(mine + 4);   // c opac r g
            const float bcol = mine[8];
            *reinterpret_cast<float4 *>(mine + 4) = make_float4(B.x, -B.y, fmaxf(B.z, 0.0f), fmaxf(B.w, 0.0f));
            mine[8] = fmaxf(bcol, 0.0f);
            mine[BROW_ID] = __uint_as_float(my_id);
            // the loop accumulates -v_xy, -v_conic (without the 1/2 of the diagonal), -v_rgb (ungated), -sum v_sigma
            float *f = fact + slot_r * BFACT;
            *reinterpret_cast<float4 *>(f) = make_float4(-1.0f, -1.0f, -0.5f, -1.0f);
            *reinterpret_cast<float4 *>(f + 4) =
                make_float4(-0.5f, B.z >= 0.0f ? -1.0f : 0.0f, B.w >= 0.0f ? -1.0f : 0.0f, bcol >= 0.0f ? -1.0f : 0.0f);
            *reinterpret_cast<float2 *>(f + 8) = make_float2(1.0f / B.y, 1.0f);
        }
        __syncwarp();
        if (STATS) st_iter += n;
        for (uint32_t j = 0; j < n; j++) {
            const float *row = rows + j * ROW;
            const float4 A = *reinterpret_cast<const float4 *>(row);       // mx my a b
            const float4 B = *reinterpret_cast<const float4 *>(row + 4);   // c -opac r+ g+
            const float4 C = *reinterpret_cast<const float4 *>(row + 8);   // b+, then log2(e)-scaled c/2, a/2, b
            const float dx = A.x - px;
            float2 dy2;
            const float2 sg = pair_sigma(dx, A.y, C.y, C.z, C.w, npy2, dy2);
            const float2 gs = make_float2(ex2_approx(-sg.x), ex2_approx(-sg.y));
            const float2 noa = __fmul2_rn(gs, bcast2(B.y));                                 // -opac*g
            const float2 nal = make_float2(fmaxf(-0.999f, noa.x), fmaxf(-0.999f, noa.y));   // -alpha
            const float2 oma = __fadd2_rn(nal, bcast2(1.0f));
            const float2 nT = __fmul2_rn(T2, oma);
            // the forward's tests (blend_common.cuh); T of a stopped pixel stays <= 1e-4, so it never blends again
            const bool act0 = sg.x >= 0.0f && noa.x <= -ALPHA_CUTOFF_MID, act1 = sg.y >= 0.0f && noa.y <= -ALPHA_CUTOFF_MID;
            const bool c0 = act0 && nT.x > 1.0e-4f, c1 = act1 && nT.y > 1.0e-4f;
            if (STATS) {
                st_blend += __popc(__ballot_sync(0xffffffffu, c0)) + __popc(__ballot_sync(0xffffffffu, c1));
                st_stop += __popc(__ballot_sync(0xffffffffu, act0 && !c0 && T2.x > 1.0e-4f)) +
                           __popc(__ballot_sync(0xffffffffu, act1 && !c1 && T2.y > 1.0e-4f));
            }
            // -alpha where the pair blends; additionally gated on "not alpha-saturated" for everything but the
            // colour terms (rasterize_backwards.rs:357-372)
            const float2 nalc = make_float2(c0 ? nal.x : 0.0f, c1 ? nal.y : 0.0f);
            const float2 nals = make_float2((c0 && noa.x >= -0.999f) ? nal.x : 0.0f, (c1 && noa.y >= -0.999f) ? nal.y : 0.0f);
            const float2 ra = make_float2(rcp_approx_f(oma.x), rcp_approx_f(oma.y));
            const float2 nvis = __fmul2_rn(nalc, T2);                        // -vis
            const float2 G5 = __fmul2_rn(nvis, vo_r), G6 = __fmul2_rn(nvis, vo_g), G7 = __fmul2_rn(nvis, vo_b);
            // u_k = rem_k - T c_k
            const float2 u_r = __ffma2_rn(T2, bcast2(-B.z), rem_r);
            const float2 u_g = __ffma2_rn(T2, bcast2(-B.w), rem_g);
            const float2 u_b = __ffma2_rn(T2, bcast2(-C.x), rem_b);
            float2 nd = __fmul2_rn(u_r, vo_r);
            nd = __ffma2_rn(u_g, vo_g, nd);
            nd = __ffma2_rn(u_b, vo_b, nd);                                   // -dot
            const float2 nva = __fmul2_rn(__fadd2_rn(nd, nvo_w), ra);         // -v_alpha
            const float2 nvs = __fmul2_rn(nals, nva);                         // -v_sigma  (= alpha v_alpha)
            const float2 vsx = __fmul2_rn(nvs, bcast2(dx)), vsy = __fmul2_rn(nvs, dy2);
            const float2 G0 = __ffma2_rn(bcast2(A.z), vsx, __fmul2_rn(bcast2(A.w), vsy));   // -v_xy.x
            const float2 G1 = __ffma2_rn(bcast2(A.w), vsx, __fmul2_rn(bcast2(B.x), vsy));   // -v_xy.y
            const float2 G2 = __fmul2_rn(vsx, bcast2(dx)), G3 = __fmul2_rn(vsx, dy2), G4 = __fmul2_rn(vsy, dy2);
            const float2 nn = __ffma2_rn(G0, G0, __fmul2_rn(__fmul2_rn(G1, bcast2(hw2)), G1));
            const float2 G9 = __fmul2_rn(make_float2(sqrt_approx_f(nn.x), sqrt_approx_f(nn.y)), ifa);
            // advance the pixel state
            rem_r = __ffma2_rn(nvis, bcast2(B.z), rem_r);
            rem_g = __ffma2_rn(nvis, bcast2(B.w), rem_g);
            rem_b = __ffma2_rn(nvis, bcast2(C.x), rem_b);
            T2.x = act0 ? nT.x : T2.x;
            T2.y = act1 ? nT.y : T2.y;
            // ---- the lane's two pixels, then reduce-scatter 10 values over 32 lanes: 5+3+2+1+1 shuffles
            float g[10];
            g[0] = G0.x + G0.y; g[1] = G1.x + G1.y; g[2] = G2.x + G2.y; g[3] = G3.x + G3.y; g[4] = G4.x + G4.y;
            g[5] = G5.x + G5.y; g[6] = G6.x + G6.y; g[7] = G7.x + G7.y; g[8] = nvs.x + nvs.y; g[9] = G9.x + G9.y;
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
            if (owner && d1 != 0.0f) {
                const uint32_t id = __float_as_uint(row[BROW_ID]);
                atomicAdd(v_combined + (size_t)id * BG_VCOMBINED_STRIDE + slot, d1 * fact[j * BFACT + slot]);
            }
        }
        __syncwarp();  // all lanes are done with this buffer before the next stage() overwrites it
    }
    if (STATS && lane == 0) {
        atomicAdd(stats + 0, st_iter); atomicAdd(stats + 1, st_blend); atomicAdd(stats + 2, st_stop);
    }
}

cudaError_t launch_blend_bwd(cudaStream_t s, uint32_t num_tiles, const CUtensorMap &tm_projected, const uint32_t *cgid_from_isect,
                             const uint32_t *tile_offsets, const float *out_img, const float *v_output, const uint32_t *live_masks,
                             const uint32_t *warp_batches, float *v_combined, unsigned long long *stats, uint32_t tiles_x,
                             uint32_t w, uint32_t h, const float *bg) {
    BlendUniforms u;
    u.tiles_x = tiles_x; u.img_w = w; u.img_h = h; u.bg_r = bg[0]; u.bg_g = bg[1]; u.bg_b = bg[2];
    if (stats)
        blend_bwd_kernel<true><<<num_tiles, RASTER_THREADS, 0, s>>>(tm_projected, cgid_from_isect, tile_offsets, (const float4 *)out_img,
                                                                   (const float4 *)v_output, live_masks, warp_batches, v_combined, stats, u);
    else
        blend_bwd_kernel<false><<<num_tiles, RASTER_THREADS, 0, s>>>(tm_projected, cgid_from_isect, tile_offsets, (const float4 *)out_img,
                                                                    (const float4 *)v_output, live_masks, warp_batches, v_combined,
                                                                    nullptr, u);
    return cudaGetLastError();
}

}  // namespace bg
