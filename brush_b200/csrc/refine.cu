// refine.cu -- SplatTrainer::refine (brush-train/src/train.rs:431-893) as device kernels: prune / resample /
// force-split / grow / split / opacity decay, and the percentile bounds (splat_init.rs:130-160).
//
// The reference builds refine from generic tensor ops with a host round trip at every decision (argwhere, a full
// weight readback into rand's weighted sampler, a HashSet of indices).  Here every decision stays on the device:
//   * counts live in a control block (RefineCtl) that later kernels read; the host reads it once, at the end;
//   * "argwhere + select" is a flag scan + row compaction (bg_inclusive_scan_u32's kernel);
//   * weighted sampling WITHOUT replacement (multinomial.rs:1-26, rand's sample_weighted = Efraimidis-Spirakis) is
//     "the k largest keys log(u_i) / w_i": keys from a counter-based uniform stream, the context's one-sweep radix sort,
//     and "rank < k" with k read on the device;
//   * the HashSet union of the three selections is a flag array; children are appended in index order.
// Arithmetic uses the deterministic exp/log of bg_math.cuh and this file is compiled with -fmad=false, so the
// selections are a pure function of (inputs, seed): tests/refine_ref.py restates them in numpy and must agree exactly.
#include <algorithm>

#include "bg_common.cuh"
#include "bg_math.cuh"
#include "bg_refine.cuh"
#include "bg_rng.cuh"

namespace bg {

constexpr float R_MIN_OPACITY = 1.0f / 255.0f;          // train.rs:33
constexpr float R_FRAC_1_SQRT_2 = 0.70710678118654752440f;

__device__ __forceinline__ float sigmoid_det(float x) { return 1.0f / (1.0f + det_expf(-x)); }
__device__ __forceinline__ float inv_sigmoid_det(float x) { return det_logf(x / (1.0f - x)); }
// ascending radix order of the returned word == DESCENDING order of the float (NaN never occurs: keys are -inf or finite)
__device__ __forceinline__ uint32_t key_desc(float f) {
    const uint32_t u = __float_as_uint(f);
    return ~(u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u));
}
__device__ __forceinline__ uint32_t key_asc(float f) {
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key_asc_inv(uint32_t k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long stream, uint32_t i) {
    const unsigned long long ctr = (stream << 40) + (i >> 2);
    const uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0x52464e45u, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t w = (i & 3u) == 0 ? r.x : ((i & 3u) == 1 ? r.y : ((i & 3u) == 2 ? r.z : r.w));
    return u01(w);
}

// ---- 1. prune mask (train.rs:487-535): keep[i] = !(alpha < 1/255 | scale > max | out of bounds | non-finite)
__global__ void __launch_bounds__(256)
refine_classify_kernel(uint32_t n0, uint32_t kf, const float *__restrict__ transforms, const float *__restrict__ sh,
                       const float *__restrict__ raw_opac, float cx, float cy, float cz, float max_allowed,
                       uint32_t *__restrict__ keep, uint32_t *__restrict__ ctl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (i < n0) {
        float t[10];
#pragma unroll
        for (int c = 0; c < 10; c++) t[c] = __ldg(transforms + (size_t)i * 10 + c);
        const float raw = __ldg(raw_opac + i);
        bool nf = !is_finite(raw);
#pragma unroll
        for (int c = 0; c < 10; c++) nf = nf || !is_finite(t[c]);
        const float *row = sh + (size_t)i * kf;
        for (uint32_t c = 0; c < kf; c++) nf = nf || !is_finite(__ldg(row + c));
        const bool alpha = sigmoid_det(raw) < R_MIN_OPACITY;
        const bool big = det_expf(t[7]) > max_allowed || det_expf(t[8]) > max_allowed || det_expf(t[9]) > max_allowed;
        const bool far = fabsf(t[0] - cx) > max_allowed || fabsf(t[1] - cy) > max_allowed || fabsf(t[2] - cz) > max_allowed;
        keep[i] = (alpha || big || far || nf) ? 0u : 1u;
        bad = nf;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(ctl + RC_NON_FINITE, (uint32_t)__popc(m));
}

// ---- 2. prune_points (train.rs:848-893): nothing is pruned when no splat or every splat would go
__global__ void refine_plan_prune_kernel(uint32_t n0, const uint32_t *__restrict__ keep_incl, uint32_t *__restrict__ ctl) {
    const uint32_t kept = n0 ? keep_incl[n0 - 1] : 0u;
    const bool identity = kept == 0u || kept == n0;
    ctl[RC_IDENTITY] = identity ? 1u : 0u;
    ctl[RC_PRUNED] = identity ? 0u : n0 - kept;
    ctl[RC_N] = identity ? n0 : kept;
}

struct RefineArrays {   // the twelve per-Gaussian arrays that follow a splat through prune and split
    const float *transforms, *sh, *raw_opac, *m_t, *v_t, *m_sh, *v_sh, *m_o, *v_o, *refine_norm, *vis_weight, *max_screen;
};
struct RefineOut {
    float *transforms, *sh, *raw_opac, *m_t, *v_t, *m_sh, *v_sh, *m_o, *v_o;
    float *refine_norm, *vis_weight, *max_screen;   // compacted statistics (scratch: the record restarts after refine)
};

// one warp per source row: lanes stride over the row's elements
__global__ void __launch_bounds__(256)
refine_compact_kernel(uint32_t n0, uint32_t kf, RefineArrays a, RefineOut o, const uint32_t *__restrict__ keep,
                      const uint32_t *__restrict__ keep_incl, const uint32_t *__restrict__ ctl) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n0) return;
    const bool identity = ctl[RC_IDENTITY] != 0u;
    if (!identity && !keep[i]) return;
    const size_t d = identity ? i : keep_incl[i] - 1u;
    for (uint32_t c = lane; c < 10; c += 32) {
        o.transforms[d * 10 + c] = __ldg(a.transforms + (size_t)i * 10 + c);
        o.m_t[d * 10 + c] = __ldg(a.m_t + (size_t)i * 10 + c);
        o.v_t[d * 10 + c] = __ldg(a.v_t + (size_t)i * 10 + c);
    }
    for (uint32_t c = lane; c < kf; c += 32) {
        o.sh[d * kf + c] = __ldg(a.sh + (size_t)i * kf + c);
        o.m_sh[d * kf + c] = __ldg(a.m_sh + (size_t)i * kf + c);
    }
    if (lane == 0) {
        o.raw_opac[d] = __ldg(a.raw_opac + i); o.m_o[d] = __ldg(a.m_o + i); o.v_o[d] = __ldg(a.v_o + i);
        o.v_sh[d] = __ldg(a.v_sh + i);
        o.refine_norm[d] = __ldg(a.refine_norm + i); o.vis_weight[d] = __ldg(a.vis_weight + i); o.max_screen[d] = __ldg(a.max_screen + i);
    }
}

// ---- 3. sampling keys.  mode 0: replacement weights = opacity x visible (train.rs:544-556);
//         mode 1: growth weights = refine weight where it is above the threshold and visible (train.rs:590-632)
__global__ void __launch_bounds__(256)
refine_keys_kernel(uint32_t n_max, int mode, const float *__restrict__ raw_opac, const float *__restrict__ refine_norm,
                   const float *__restrict__ vis_weight, float grad_threshold, unsigned long long seed, unsigned long long stream,
                   uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ ctl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = ctl[RC_N];
    bool pos = false, above = false;
    if (i < n) {
        const bool vis = __ldg(vis_weight + i) > 0.0f;
        float w;
        if (mode == 0) w = vis ? sigmoid_det(__ldg(raw_opac + i)) : 0.0f;
        else { const float r = __ldg(refine_norm + i); above = vis && r > grad_threshold; w = above ? r : 0.0f; }
        pos = is_finite(w) && w > 0.0f;   // non-finite or negative weights count as zero (multinomial.rs:8-14)
        const float key = pos ? det_logf(uniform01(seed, stream, i)) / w : __int_as_float(0xff800000);
        keys[i] = key_desc(key);
        vals[i] = i;
    } else if (i < n_max) {
        keys[i] = 0xFFFFFFFFu;   // never sorted (the sort reads its count on the device); defined for tidiness
        vals[i] = i;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, pos), ma = __ballot_sync(0xffffffffu, above);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(ctl + (mode == 0 ? RC_POS0 : RC_POS1), (uint32_t)__popc(m));
    if (mode == 1 && (threadIdx.x & 31u) == 0 && ma) atomicAdd(ctl + RC_THRESHOLD_COUNT, (uint32_t)__popc(ma));
}

// growth count (train.rs:604-617): round(threshold_count * fraction) - pruned, capped by the headroom
__global__ void refine_plan_growth_kernel(float fraction, uint32_t max_splats, int enabled, uint32_t *__restrict__ ctl) {
    const uint32_t grow_count = (uint32_t)roundf((float)ctl[RC_THRESHOLD_COUNT] * fraction);
    const uint32_t pruned = ctl[RC_PRUNED];
    const uint32_t sample = grow_count > pruned ? grow_count - pruned : 0u;
    const uint32_t cur = ctl[RC_N] + ctl[RC_SPLIT_REPLACE] + ctl[RC_SPLIT_OVERSIZED];
    const uint32_t headroom = max_splats > cur ? max_splats - cur : 0u;
    ctl[RC_GROW] = enabled ? min(sample, headroom) : 0u;
}

// the k best keys: split[sorted_vals[r]] = 1 for r < min(k, positives); counts the newly set flags
__global__ void __launch_bounds__(256)
refine_mark_topk_kernel(uint32_t n_max, const uint32_t *__restrict__ sorted_vals, uint32_t k_slot, uint32_t pos_slot,
                        uint32_t count_slot, uint32_t *__restrict__ split, uint32_t *__restrict__ ctl) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t k = min(ctl[k_slot], ctl[pos_slot]);
    bool fresh = false;
    if (r < k && r < n_max) {
        const uint32_t g = sorted_vals[r];
        fresh = atomicExch(split + g, 1u) == 0u;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, fresh);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(ctl + count_slot, (uint32_t)__popc(m));
}

// ---- 4. force-split of splats that are too big on screen (train.rs:562-586): candidates in index order
__global__ void __launch_bounds__(256)
refine_oversize_flags_kernel(uint32_t n_max, float screen_threshold, const float *__restrict__ max_screen,
                             const float *__restrict__ vis_weight, const uint32_t *__restrict__ split,
                             uint32_t *__restrict__ cand, const uint32_t *__restrict__ ctl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_max) return;
    const bool c = i < ctl[RC_N] && screen_threshold > 0.0f && __ldg(max_screen + i) > screen_threshold && __ldg(vis_weight + i) > 0.0f &&
                   split[i] == 0u;
    cand[i] = c ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
refine_oversize_mark_kernel(uint32_t n_max, uint32_t max_splats, const uint32_t *__restrict__ cand,
                            const uint32_t *__restrict__ cand_incl, uint32_t *__restrict__ split, uint32_t *__restrict__ ctl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cur = ctl[RC_N] + ctl[RC_SPLIT_REPLACE];
    const uint32_t budget = max_splats > cur ? max_splats - cur : 0u;
    bool take = false;
    if (i < n_max && cand[i]) take = cand_incl[i] - 1u < budget;
    if (take) split[i] = 1u;
    const uint32_t m = __ballot_sync(0xffffffffu, take);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(ctl + RC_SPLIT_OVERSIZED, (uint32_t)__popc(m));
}

__global__ void refine_plan_split_kernel(uint32_t n_max, const uint32_t *__restrict__ split_incl, uint32_t capacity, uint32_t *__restrict__ ctl) {
    const uint32_t n = ctl[RC_N];
    const uint32_t count = n ? split_incl[n - 1] : 0u;
    ctl[RC_REFINE_COUNT] = count;
    ctl[RC_N_NEW] = n + count;
    if (n + count > capacity) ctl[RC_OVERFLOW] = 1u;
    (void)n_max;
}

// ---- 5. refine_splats (train.rs:665-821): parent moves to mean - offset and shrinks, the child sits at mean + offset
// with the normalised rotation; both halves restart with zero Adam moments.
__global__ void __launch_bounds__(128)
refine_split_kernel(uint32_t n_max, uint32_t kf, uint32_t capacity, float screen_threshold, RefineOut o,
                    const uint32_t *__restrict__ split, const uint32_t *__restrict__ split_incl, const uint32_t *__restrict__ ctl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = ctl[RC_N];
    if (i >= n || i >= n_max || !split[i]) return;
    const size_t child = (size_t)n + split_incl[i] - 1u;
    if (child >= capacity) return;
    float t[10];
#pragma unroll
    for (int c = 0; c < 10; c++) t[c] = o.transforms[(size_t)i * 10 + c];
    const float mag = fmaxf(sqrtf(t[3] * t[3] + t[4] * t[4] + t[5] * t[5] + t[6] * t[6]), 1e-32f);
    const float qw = t[3] / mag, qx = t[4] / mag, qy = t[5] / mag, qz = t[6] / mag;
    const float raw = o.raw_opac[i];
    const float inv_opac = 1.0f - sigmoid_det(raw);
    // inv_opac^(1/sqrt2) through the deterministic exp/log
    const float powed = inv_opac > 0.0f ? det_expf(R_FRAC_1_SQRT_2 * det_logf(inv_opac)) : 0.0f;
    const float new_opac = fminf(fmaxf(1.0f - powed, R_MIN_OPACITY), 1.0f - R_MIN_OPACITY);
    const float new_raw = inv_sigmoid_det(new_opac);
    float sc[3], sq[3];
#pragma unroll
    for (int a = 0; a < 3; a++) { sc[a] = det_expf(t[7 + a]); sq[a] = sc[a] * sc[a]; }
    const float max_sq = fmaxf(fmaxf(fmaxf(sq[0], sq[1]), sq[2]), 1e-30f);
    float k_max = R_FRAC_1_SQRT_2;
    if (screen_threshold > 0.0f) k_max = fminf((1.0f / fmaxf(o.max_screen[i], 1e-6f)) * screen_threshold, R_FRAC_1_SQRT_2);
    float off[3], nls[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float ratio = sq[a] / max_sq;
        const float k_axis = -(ratio * (-k_max + 1.0f)) + 1.0f;
        off[a] = sqrtf(fmaxf(-(k_axis * k_axis) + 1.0f, 0.0f)) * sc[a];
        nls[a] = t[7 + a] + det_logf(k_axis);
    }
    // quat_vec.rs: rotate the local offset by the (normalised) rotation
    const float qw2 = qw * qw, qx2 = qx * qx, qy2 = qy * qy, qz2 = qz * qz;
    const float xy = qx * qy, xz = qx * qz, yz = qy * qz, wx = qw * qx, wy = qw * qy, wz = qw * qz;
    const float sx = (qw2 + qx2 - qy2 - qz2) * off[0] + (xy * off[1] + xz * off[2] + wy * off[2] - wz * off[1]) * 2.0f;
    const float sy = (qw2 - qx2 + qy2 - qz2) * off[1] + (xy * off[0] + yz * off[2] + wz * off[0] - wx * off[2]) * 2.0f;
    const float sz = (qw2 - qx2 - qy2 + qz2) * off[2] + (xz * off[0] + yz * off[1] + wx * off[1] - wy * off[0]) * 2.0f;
    float *pt = o.transforms + (size_t)i * 10, *ct = o.transforms + child * 10;
    ct[0] = t[0] + sx; ct[1] = t[1] + sy; ct[2] = t[2] + sz;
    ct[3] = qw; ct[4] = qx; ct[5] = qy; ct[6] = qz;
    pt[0] = t[0] - sx; pt[1] = t[1] - sy; pt[2] = t[2] - sz;
#pragma unroll
    for (int a = 0; a < 3; a++) { ct[7 + a] = nls[a]; pt[7 + a] = nls[a]; }
    o.raw_opac[i] = new_raw; o.raw_opac[child] = new_raw;
    for (uint32_t c = 0; c < kf; c++) {
        o.sh[child * kf + c] = o.sh[(size_t)i * kf + c];
        o.m_sh[(size_t)i * kf + c] = 0.0f; o.m_sh[child * kf + c] = 0.0f;
    }
#pragma unroll
    for (int c = 0; c < 10; c++) {
        o.m_t[(size_t)i * 10 + c] = 0.0f; o.v_t[(size_t)i * 10 + c] = 0.0f;
        o.m_t[child * 10 + c] = 0.0f; o.v_t[child * 10 + c] = 0.0f;
    }
    o.v_sh[i] = 0.0f; o.v_sh[child] = 0.0f; o.m_o[i] = 0.0f; o.m_o[child] = 0.0f; o.v_o[i] = 0.0f; o.v_o[child] = 0.0f;
}

// ---- 6. opacity decay on every splat (train.rs:808-816)
__global__ void __launch_bounds__(256)
refine_decay_kernel(uint32_t cap, float minus_opac, float *__restrict__ raw_opac, const uint32_t *__restrict__ ctl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap || i >= ctl[RC_N_NEW]) return;
    const float v = fminf(fmaxf(sigmoid_det(raw_opac[i]) - minus_opac, 1e-12f), 1.0f - 1e-12f);
    raw_opac[i] = inv_sigmoid_det(v);
}

// ---- bounds (splat_init.rs:130-160): per axis, the (1-p)/2 and (1+p)/2 order statistics of the finite means
__global__ void __launch_bounds__(256)
bounds_keys_kernel(uint32_t n, int axis, const float *__restrict__ transforms, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
                   uint32_t *__restrict__ count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool fin = false;
    if (i < n) {
        const float v = __ldg(transforms + (size_t)i * 10 + axis);
        fin = is_finite(v);
        keys[i] = fin ? key_asc(v) : 0xFFFFFFFFu;   // non-finite values sort last and are not counted
        vals[i] = i;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, fin);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(count, (uint32_t)__popc(m));
}
__global__ void bounds_pick_kernel(const uint32_t *__restrict__ sorted_keys, const uint32_t *__restrict__ count, float percentile,
                                   float *__restrict__ out2) {
    const uint32_t m = *count;
    if (m == 0) { out2[0] = __int_as_float(0x7fc00000); out2[1] = __int_as_float(0x7fc00000); return; }
    const uint32_t lo = (uint32_t)(((1.0f - percentile) / 2.0f) * (float)m);
    const uint32_t hi = min(m - 1u, (uint32_t)(((1.0f + percentile) / 2.0f) * (float)m));
    out2[0] = key_asc_inv(sorted_keys[min(lo, m - 1u)]);
    out2[1] = key_asc_inv(sorted_keys[hi]);
}

// ---------------------------------------------------------------------------------------------- launchers
static unsigned blocks(uint32_t n, unsigned per) { return (unsigned)std::max<uint64_t>(((uint64_t)n + per - 1) / per, 1); }

cudaError_t launch_refine_classify(cudaStream_t s, uint32_t n0, uint32_t kf, const float *transforms, const float *sh,
                                   const float *raw_opac, const float *center, float max_allowed, uint32_t *keep, uint32_t *ctl) {
    refine_classify_kernel<<<blocks(n0, 256), 256, 0, s>>>(n0, kf, transforms, sh, raw_opac, center[0], center[1], center[2], max_allowed, keep, ctl);
    return cudaGetLastError();
}
cudaError_t launch_refine_plan_prune(cudaStream_t s, uint32_t n0, const uint32_t *keep_incl, uint32_t *ctl) {
    refine_plan_prune_kernel<<<1, 1, 0, s>>>(n0, keep_incl, ctl);
    return cudaGetLastError();
}
cudaError_t launch_refine_compact(cudaStream_t s, uint32_t n0, uint32_t kf, const RefinePtrs &p, const uint32_t *keep,
                                  const uint32_t *keep_incl, const uint32_t *ctl) {
    RefineArrays a{p.transforms, p.sh, p.raw_opac, p.m_t, p.v_t, p.m_sh, p.v_sh, p.m_o, p.v_o, p.refine_norm, p.vis_weight, p.max_screen};
    RefineOut o{p.transforms_out, p.sh_out, p.raw_opac_out, p.m_t_out, p.v_t_out, p.m_sh_out, p.v_sh_out, p.m_o_out, p.v_o_out,
                p.refine_norm_tmp, p.vis_weight_tmp, p.max_screen_tmp};
    refine_compact_kernel<<<blocks(n0, 8), 256, 0, s>>>(n0, kf, a, o, keep, keep_incl, ctl);
    return cudaGetLastError();
}
cudaError_t launch_refine_keys(cudaStream_t s, uint32_t n_max, int mode, const RefinePtrs &p, float grad_threshold, uint64_t seed,
                               uint64_t stream_id, uint32_t *keys, uint32_t *vals, uint32_t *ctl) {
    refine_keys_kernel<<<blocks(n_max, 256), 256, 0, s>>>(n_max, mode, p.raw_opac_out, p.refine_norm_tmp, p.vis_weight_tmp, grad_threshold,
                                                        seed, stream_id, keys, vals, ctl);
    return cudaGetLastError();
}
cudaError_t launch_refine_plan_growth(cudaStream_t s, float fraction, uint32_t max_splats, bool enabled, uint32_t *ctl) {
    refine_plan_growth_kernel<<<1, 1, 0, s>>>(fraction, max_splats, enabled ? 1 : 0, ctl);
    return cudaGetLastError();
}
cudaError_t launch_refine_mark_topk(cudaStream_t s, uint32_t n_max, const uint32_t *sorted_vals, uint32_t k_slot, uint32_t pos_slot,
                                    uint32_t count_slot, uint32_t *split, uint32_t *ctl) {
    refine_mark_topk_kernel<<<blocks(n_max, 256), 256, 0, s>>>(n_max, sorted_vals, k_slot, pos_slot, count_slot, split, ctl);
    return cudaGetLastError();
}
cudaError_t launch_refine_oversize_flags(cudaStream_t s, uint32_t n_max, float thr, const RefinePtrs &p, const uint32_t *split,
                                         uint32_t *cand, const uint32_t *ctl) {
    refine_oversize_flags_kernel<<<blocks(n_max, 256), 256, 0, s>>>(n_max, thr, p.max_screen_tmp, p.vis_weight_tmp, split, cand, ctl);
    return cudaGetLastError();
}
cudaError_t launch_refine_oversize_mark(cudaStream_t s, uint32_t n_max, uint32_t max_splats, const uint32_t *cand, const uint32_t *cand_incl,
                                        uint32_t *split, uint32_t *ctl) {
    refine_oversize_mark_kernel<<<blocks(n_max, 256), 256, 0, s>>>(n_max, max_splats, cand, cand_incl, split, ctl);
    return cudaGetLastError();
}
cudaError_t launch_refine_plan_split(cudaStream_t s, uint32_t n_max, const uint32_t *split_incl, uint32_t capacity, uint32_t *ctl) {
    refine_plan_split_kernel<<<1, 1, 0, s>>>(n_max, split_incl, capacity, ctl);
    return cudaGetLastError();
}
cudaError_t launch_refine_split(cudaStream_t s, uint32_t n_max, uint32_t kf, uint32_t capacity, float thr, const RefinePtrs &p,
                                const uint32_t *split, const uint32_t *split_incl, const uint32_t *ctl) {
    RefineOut o{p.transforms_out, p.sh_out, p.raw_opac_out, p.m_t_out, p.v_t_out, p.m_sh_out, p.v_sh_out, p.m_o_out, p.v_o_out,
                p.refine_norm_tmp, p.vis_weight_tmp, p.max_screen_tmp};
    refine_split_kernel<<<blocks(n_max, 128), 128, 0, s>>>(n_max, kf, capacity, thr, o, split, split_incl, ctl);
    return cudaGetLastError();
}
cudaError_t launch_refine_decay(cudaStream_t s, uint32_t cap, float minus_opac, float *raw_opac, const uint32_t *ctl) {
    refine_decay_kernel<<<blocks(cap, 256), 256, 0, s>>>(cap, minus_opac, raw_opac, ctl);
    return cudaGetLastError();
}
cudaError_t launch_bounds_keys(cudaStream_t s, uint32_t n, int axis, const float *transforms, uint32_t *keys, uint32_t *vals, uint32_t *count) {
    bounds_keys_kernel<<<blocks(n, 256), 256, 0, s>>>(n, axis, transforms, keys, vals, count);
    return cudaGetLastError();
}
cudaError_t launch_bounds_pick(cudaStream_t s, const uint32_t *sorted_keys, const uint32_t *count, float percentile, float *out2) {
    bounds_pick_kernel<<<1, 1, 0, s>>>(sorted_keys, count, percentile, out2);
    return cudaGetLastError();
}

}  // namespace bg
