// micro test: TMA tile::gather4 of 64-byte rows (4 arbitrary rows -> 256 contiguous bytes of shared memory)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <stdint.h>
#include <vector>
typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, const uint32_t *ids, float *out) {
    __shared__ __align__(128) float rows[8 * 16];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), r = (uint32_t)__cvta_generic_to_shared(rows);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(b));
        for (int g = 0; g < 2; g++)
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                         ::"r"(r + g * 256), "l"(&tm), "r"(0), "r"((int)ids[4 * g]), "r"((int)ids[4 * g + 1]), "r"((int)ids[4 * g + 2]), "r"((int)ids[4 * g + 3]), "r"(b)
                         : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}" ::"r"(b));
    out[threadIdx.x] = rows[threadIdx.x];
}
int main(int argc, char **argv) {
    const int boxrows = argc > 1 ? atoi(argv[1]) : 1;
    const int N = 1000;
    std::vector<float> h(N * 16);
    for (int i = 0; i < N * 16; i++) h[i] = (float)i;
    float *d, *out; uint32_t *ids;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 128 * 4);
    uint32_t hid[8] = {5, 900, 17, 3, 999, 0, 500, 501};
    cudaMalloc(&ids, 32); cudaMemcpy(ids, hid, 32, cudaMemcpyHostToDevice);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tm;
    cuuint64_t dims[2] = {16, (cuuint64_t)N}; cuuint64_t strides[1] = {64};
    cuuint32_t box[2] = {16, (cuuint32_t)boxrows}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeTiled)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (box rows %d): %d\n", boxrows, (int)r);
    if (r != CUDA_SUCCESS) return 2;
    k<<<1, 128>>>(tm, ids, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    float ho[128]; cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
    int ok = 1;
    for (int g = 0; g < 8; g++) for (int c = 0; c < 16; c++) if (ho[g * 16 + c] != (float)(hid[g] * 16 + c)) ok = 0;
    printf("rows: %g %g %g %g | %g -> %s\n", ho[0], ho[16], ho[32], ho[48], ho[64], ok ? "GATHER4_OK" : "MISMATCH");
    return ok ? 0 : 3;
}
