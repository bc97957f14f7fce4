// bg_dp.cuh -- types shared by dp.cu (NCCL binding, exchange) and api.cu (the multi-view step).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bg {

constexpr int DP_MAX_CHUNKS = 16;
constexpr int DP_MAX_VIEWS = 16;

struct NcclUniqueId { char internal[128]; };

struct DpComm {
    void *comm = nullptr;          // ncclComm_t
    int device = 0, rank = 0, world = 1;
    cudaStream_t stream = nullptr;     // the exchange runs here, beside the caller's stream
    cudaEvent_t ev_ready = nullptr;    // caller's stream -> exchange stream: gradients of the step are complete
    cudaEvent_t ev_chunk[DP_MAX_CHUNKS] = {};   // exchange stream -> caller's stream: slice c has arrived
};

struct DpHeader { float pos[DP_MAX_VIEWS][3]; };

// Float offsets of the exchange buffers (see dp.cu).
struct DpLayout {
    uint32_t n, local, world;
    size_t rec_refine, rec_radius, rec_floats;   // record: colours [local][n][3] | refine [n] | radius [n]
    size_t small_floats;                          // v_transforms [n][10] | v_raw_opac [n] | visible [n]
    size_t recv_floats;                           // world * rec_floats, laid out per slice
    __host__ __device__ size_t chunk_base(uint32_t g0) const { return (size_t)(3 * local + 2) * world * g0; }
    __host__ __device__ size_t colour_off(uint32_t len, uint32_t li) const { return (size_t)li * world * len * 3; }
    __host__ __device__ size_t refine_off(uint32_t len) const { return (size_t)3 * local * world * len; }
    __host__ __device__ size_t radius_off(uint32_t len) const { return (size_t)(3 * local + 1) * world * len; }
};
inline DpLayout dp_layout(uint32_t n, uint32_t local, uint32_t world) {
    DpLayout L;
    L.n = n; L.local = local; L.world = world;
    L.rec_refine = (size_t)3 * local * n;
    L.rec_radius = L.rec_refine + n;
    L.rec_floats = L.rec_radius + n;
    L.small_floats = (size_t)12 * n;
    L.recv_floats = L.rec_floats * world;
    return L;
}

const char *dp_nccl_error(int code);
int dp_unique_id(NcclUniqueId *out);
DpComm *dp_comm_create(int device, const NcclUniqueId &id, int rank, int world, int *nccl_rc);
void dp_comm_destroy(DpComm *c);
void dp_chunk_range(uint32_t n, uint32_t chunks, uint32_t c, uint32_t *g0, uint32_t *len);
int dp_exchange_chunk(DpComm *cm, uint32_t n, uint32_t local, uint32_t chunks, uint32_t c, float *small, const float *record,
                      float *recv);
int dp_exchange_header(DpComm *cm, uint32_t local, const float *hdr, float *hdr_all);
cudaError_t launch_write_header(cudaStream_t s, float *hdr, const DpHeader &h, uint32_t local);
cudaError_t launch_accumulate_view(cudaStream_t s, uint32_t n, float *small, const float *tmp, const float *vis_view,
                                   float *refine, const float *refine_view, float *radius, const float *radius_view);

}  // namespace bg
