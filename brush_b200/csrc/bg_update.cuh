// bg_update.cuh -- argument block of train_update_kernel (update.cu), filled by api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bg {

struct UpdateParams {
    uint32_t g_begin, count;         // this launch updates Gaussians [g_begin, g_begin + count)
    float *transforms, *sh, *raw_opac;
    float *m_t, *v_t, *m_sh, *v_sh, *m_o, *v_o;
    float *refine_norm, *vis_weight, *max_screen;
    const float *g_t, *g_o;          // [n,10], [n] (summed over the views when grad_scale != 1)
    const float *g_sh;               // dense [n,K,3] gradient; unused when factored
    float grad_scale;                // applied to g_t and g_o (1/views)
    // factored SH gradient + MAX statistics: the slice of the gathered records that holds [g_begin, g_begin+count)
    // (bg_dp.cuh DpLayout): colours [local][world][count][3], refine [world][count], radius [world][count]
    const float *colours, *refine_all, *radius_all;
    const float *cam_all;            // device [views][4]: camera positions in global view order
    uint32_t views, local, world;
    float sh_grad_scale;             // 1/views
    const float *v_refine, *max_radius;   // [n] statistics of the step when not factored
    const float *visible;            // [n] visibility count of the step (sum over the views)
    float lr_t[10];
    float lr_sh_dc, lr_sh_rest, lr_opac;
    float beta1, beta2, eps, f1, f2, inv_bc1, inv_bc2;   // 1 / (1 - beta^t), rounded once on the host
    int first;
    int noisy;
    float noise_scale, median_scale;
    unsigned long long seed, noise_offset;
};

cudaError_t launch_train_update(cudaStream_t s, int deg, const UpdateParams &P, bool factored);

}  // namespace bg
