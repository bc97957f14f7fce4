// bg_refine.cuh -- control block and pointer bundle shared by refine.cu (kernels) and api.cu (bg_refine).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bg {

// device control words of one refine (zeroed at its start, read back once at its end)
enum RefineCtlSlots : uint32_t {
    RC_NON_FINITE = 0,      // splats with a non-finite parameter
    RC_IDENTITY = 1,        // 1: nothing pruned (no splat, or every splat, matched the prune mask)
    RC_PRUNED = 2,
    RC_N = 3,               // splats after the prune
    RC_POS0 = 4,            // positive replacement weights
    RC_SPLIT_REPLACE = 5,   // splits that replace pruned splats
    RC_SPLIT_OVERSIZED = 6,
    RC_THRESHOLD_COUNT = 7,
    RC_POS1 = 8,            // positive growth weights
    RC_GROW = 9,            // growth samples requested
    RC_SPLIT_GROWTH = 10,   // new splits from the growth sample
    RC_REFINE_COUNT = 11,
    RC_N_NEW = 12,
    RC_OVERFLOW = 13,       // n_new exceeded the capacity of the destination arrays
    RC_WORDS = 16
};

struct RefinePtrs {
    const float *transforms, *sh, *raw_opac, *m_t, *v_t, *m_sh, *v_sh, *m_o, *v_o, *refine_norm, *vis_weight, *max_screen;
    float *transforms_out, *sh_out, *raw_opac_out, *m_t_out, *v_t_out, *m_sh_out, *v_sh_out, *m_o_out, *v_o_out;
    float *refine_norm_tmp, *vis_weight_tmp, *max_screen_tmp;
};

cudaError_t launch_refine_classify(cudaStream_t, uint32_t, uint32_t, const float *, const float *, const float *, const float *, float,
                                   uint32_t *, uint32_t *);
cudaError_t launch_refine_plan_prune(cudaStream_t, uint32_t, const uint32_t *, uint32_t *);
cudaError_t launch_refine_compact(cudaStream_t, uint32_t, uint32_t, const RefinePtrs &, const uint32_t *, const uint32_t *, const uint32_t *);
cudaError_t launch_refine_keys(cudaStream_t, uint32_t, int, const RefinePtrs &, float, uint64_t, uint64_t, uint32_t *, uint32_t *, uint32_t *);
cudaError_t launch_refine_plan_growth(cudaStream_t, float, uint32_t, bool, uint32_t *);
cudaError_t launch_refine_mark_topk(cudaStream_t, uint32_t, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t *, uint32_t *);
cudaError_t launch_refine_oversize_flags(cudaStream_t, uint32_t, float, const RefinePtrs &, const uint32_t *, uint32_t *, const uint32_t *);
cudaError_t launch_refine_oversize_mark(cudaStream_t, uint32_t, uint32_t, const uint32_t *, const uint32_t *, uint32_t *, uint32_t *);
cudaError_t launch_refine_plan_split(cudaStream_t, uint32_t, const uint32_t *, uint32_t, uint32_t *);
cudaError_t launch_refine_split(cudaStream_t, uint32_t, uint32_t, uint32_t, float, const RefinePtrs &, const uint32_t *, const uint32_t *,
                                const uint32_t *);
cudaError_t launch_refine_decay(cudaStream_t, uint32_t, float, float *, const uint32_t *);
cudaError_t launch_bounds_keys(cudaStream_t, uint32_t, int, const float *, uint32_t *, uint32_t *, uint32_t *);
cudaError_t launch_bounds_pick(cudaStream_t, const uint32_t *, const uint32_t *, float, float *);

}  // namespace bg
