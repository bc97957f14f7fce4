// bg_sh.cuh -- real spherical-harmonics basis shared by project_bwd.cu (coefficient VJP) and update.cu (the
// SH gradient rebuilt from per-view colour gradients inside the optimiser pass).
#pragma once
#include "bg_math.cuh"

namespace bg {

// SH basis values Y[0..K) for unit direction v (kernels/sh.rs:265-355 uses the same polynomials).
template <int DEG>
__device__ __forceinline__ void sh_basis(V3 v, float *Y) {
    Y[0] = 0.2820948f;
    if (DEG >= 1) {
        const float f0a = 0.4886025f;
        Y[1] = -f0a * v.y; Y[2] = f0a * v.z; Y[3] = -f0a * v.x;
    }
    float z2 = v.z * v.z;
    float fc1 = v.x * v.x - v.y * v.y, fs1 = 2.0f * v.x * v.y;
    float p6 = 0.9461747f * z2 - 0.31539157f;
    if (DEG >= 2) {
        float f0b = -1.0925485f * v.z;
        const float f1a = 0.54627424f;
        Y[4] = f1a * fs1; Y[5] = f0b * v.y; Y[6] = p6; Y[7] = f0b * v.x; Y[8] = f1a * fc1;
    }
    float fc2 = v.x * fc1 - v.y * fs1, fs2 = v.x * fs1 + v.y * fc1;
    float p12 = v.z * (1.8658817f * z2 - 1.119529f);
    if (DEG >= 3) {
        float f0c = -2.285229f * z2 + 0.4570458f;
        float f1b = 1.4453057f * v.z;
        const float f2a = -0.5900436f;
        Y[9] = f2a * fs2; Y[10] = f1b * fs1; Y[11] = f0c * v.y; Y[12] = p12; Y[13] = f0c * v.x; Y[14] = f1b * fc1;
        Y[15] = f2a * fc2;
    }
    if (DEG >= 4) {
        float f0d = v.z * (-4.683326f * z2 + 2.0071396f);
        float f1c = 3.3116114f * z2 - 0.47308735f;
        float f2b = -1.7701308f * v.z;
        const float f3a = 0.62583575f;
        float fc3 = v.x * fc2 - v.y * fs2, fs3 = v.x * fs2 + v.y * fc2;
        Y[16] = f3a * fs3; Y[17] = f2b * fs2; Y[18] = f1c * fs1; Y[19] = f0d * v.y;
        Y[20] = 1.9843135f * v.z * p12 + -1.0062306f * p6;
        Y[21] = f0d * v.x; Y[22] = f1c * fc1; Y[23] = f2b * fc2; Y[24] = f3a * fc3;
    }
}

}  // namespace bg
