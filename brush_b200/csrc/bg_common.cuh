// bg_common.cuh -- shared device utilities: uniforms, control block layout, decoupled look-back.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/brush_b200.h"

namespace bg {

constexpr uint32_t TILE_W = 16;          // kernels/helpers.rs:15-16
constexpr uint32_t TILE_PIX = 256;
constexpr float ALPHA_CUTOFF_MID = 1.0f / 255.0f;   // kernels/helpers.rs:23-24
constexpr float ALPHA_CUTOFF_BAND = 1.0e-3f;

// Device-side control block, zeroed with one memset at the start of every forward.
// counters: [0]=num_visible [1]=num_intersections [2]=overflow flag [3]=reserved
enum CtlOffsets : uint32_t {
    CTL_COUNTERS = 0,        // 16 u32
    CTL_TICKETS = 16,        // 48 u32: one work-distribution ticket per persistent kernel launch
    CTL_HIST_DEPTH = 64,     // 4*256 u32: digit histograms of the depth keys
    CTL_HIST_TILE = 64 + 1024,   // 4*256 u32: digit histograms of the tile keys
    CTL_WORDS = 64 + 2048
};
enum TicketIds : uint32_t {
    TK_PROJECT = 0, TK_DEPTH_HIST = 1, TK_DEPTH_PASS0 = 2 /* ..5 */, TK_SCAN = 6, TK_VISIBLE = 7,
    TK_TILE_HIST = 8, TK_TILE_PASS0 = 9 /* ..12 */, TK_MISC = 13
};

// ---- decoupled look-back state: 64-bit word = [epoch:30 | status:2 | value:32].
// A word is meaningful only if its epoch equals the launch's epoch, so the state arrays never
// need zeroing between launches.
constexpr uint32_t LB_INVALID = 0, LB_AGGREGATE = 1, LB_INCLUSIVE = 2;

__device__ __forceinline__ void lb_store(unsigned long long *p, uint32_t epoch, uint32_t status, uint32_t value) {
    unsigned long long w = ((unsigned long long)((epoch << 2) | status) << 32) | (unsigned long long)value;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long lb_load(const unsigned long long *p) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ uint32_t lb_status(unsigned long long w, uint32_t epoch) {
    uint32_t hi = (uint32_t)(w >> 32);
    return ((hi >> 2) == epoch) ? (hi & 3u) : LB_INVALID;
}
__device__ __forceinline__ uint32_t lb_value(unsigned long long w) { return (uint32_t)w; }

// Warp-parallel look-back for one running sum per tile.  Called by one full warp; returns the
// exclusive prefix of `tile` (sum of the aggregates of tiles 0..tile-1) in every lane.
__device__ __forceinline__ uint32_t lb_lookback_warp(const unsigned long long *state, uint32_t tile, uint32_t epoch) {
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t prefix = 0;
    int64_t pos = (int64_t)tile - 1;  // nearest predecessor inspected by lane 0
    while (pos >= 0) {
        int64_t idx = pos - (int64_t)lane;
        uint32_t st, val;
        // spin until every in-range lane sees a published predecessor
        while (true) {
            st = LB_INCLUSIVE;  // out-of-range lanes act like a zero-valued inclusive sentinel
            val = 0;
            if (idx >= 0) {
                unsigned long long w = lb_load(state + idx);
                st = lb_status(w, epoch);
                val = lb_value(w);
            }
            if (__all_sync(0xffffffffu, st != LB_INVALID)) break;
        }
        uint32_t incl_mask = __ballot_sync(0xffffffffu, st == LB_INCLUSIVE);
        // nearest inclusive predecessor = lowest lane with INCLUSIVE
        uint32_t first = incl_mask ? (uint32_t)(__ffs(incl_mask) - 1) : 32u;
        uint32_t contrib = (lane <= first) ? val : 0u;
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        prefix += contrib;
        if (incl_mask) break;
        pos -= 32;
    }
    return prefix;
}

// ---- TMA 1-D bulk copy (cp.async.bulk, SASS: UBLKCP) + mbarrier completion, for contiguous tiles.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 256-bit global accesses (LDG.E.256 / STG.E.256 on sm_100): one whole 32-byte sector per lane, so a row-per-lane
// gather does not depend on L1 keeping half-used sectors between two 128-bit loads.  p must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const float *p, float *o) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]), "=f"(o[6]), "=f"(o[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(float *p, const float *v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Block-wide exclusive scan of one uint per thread (blockDim.x <= 1024, multiple of 32).
// `warp_sums` is shared scratch of >= 33 words.  Returns the exclusive prefix; *total gets the block sum.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = (lane < nw) ? warp_sums[lane] : 0u;
        uint32_t si = s;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= (uint32_t)o) si += t;
        }
        if (lane < nw) warp_sums[lane] = si - s;
        if (lane == 31) warp_sums[32] = si;
    }
    __syncthreads();
    uint32_t r = warp_sums[wid] + incl - v;
    *total = warp_sums[32];
    __syncthreads();
    return r;
}

}  // namespace bg
