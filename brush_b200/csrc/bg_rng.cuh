// bg_rng.cuh -- counter-based normal noise (Philox4x32-10 + Box-Muller), shared by optim.cu (bg_normal_noise) and
// update.cu (the mean noise drawn inside the fused optimiser pass).  Element e of the stream keyed by (seed, offset)
// is normal_quad(seed, offset + e/4)[e%4].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bg {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)

// the four standard-normal draws of counter `ctr`
__device__ __forceinline__ void normal_quad(uint64_t seed, uint64_t ctr, float *z) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float r0 = sqrtf(-2.0f * logf(u01(r.x))), r1 = sqrtf(-2.0f * logf(u01(r.z)));
    float s0, c0, s1, c1;
    sincospif(2.0f * u01(r.y), &s0, &c0);
    sincospif(2.0f * u01(r.w), &s1, &c1);
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}

}  // namespace bg
