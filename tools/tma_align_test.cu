// micro test: 3-D tensor-map load of a u8 image viewed as u16/u8/u32 elements; args: etype box0_bytes c0_bytes c1
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t sa(const void* p){return (uint32_t)__cvta_generic_to_shared(p);}
__global__ void k(const __grid_constant__ CUtensorMap m, int c0, int c1, int c2, int bytes, uint8_t* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(sa(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(sa(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            :: "r"(sa(smem)), "l"(&m), "r"(c0), "r"(c1), "r"(c2), "r"(sa(&bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra W;\n}\n" :: "r"(sa(&bar)) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}
int main(int argc, char** argv)
{
    int esz = atoi(argv[1]), box0b = atoi(argv[2]), c0b = atoi(argv[3]), c1 = atoi(argv[4]);
    const int W = 1920, H = 1080, n = 2, rows = 74;
    std::vector<uint8_t> img((size_t)n * H * W * 3);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *o; cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    int bytes = box0b * rows; cudaMalloc(&o, bytes);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
    alignas(64) CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)W * 3 / esz, (cuuint64_t)H, (cuuint64_t)n}, str[2] = {(cuuint64_t)W * 3, (cuuint64_t)H * W * 3};
    cuuint32_t box[3] = {(cuuint32_t)(box0b / esz), (cuuint32_t)rows, 1}, es[3] = {1, 1, 1};
    CUtensorMapDataType dt = esz == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    CUresult r = enc(&m, dt, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("esz %d box0 %dB c0 %dB c1 %d: encode rc %d; ", esz, box0b, c0b, c1, (int)r);
    if (r) { printf("\n"); return 0; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    k<<<1, 256, 100000>>>(m, c0b / esz, c1, 1, bytes, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s; ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint8_t> got(bytes); cudaMemcpy(got.data(), o, bytes, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int y = 0; y < rows; ++y) for (int b = 0; b < box0b; ++b) {
            long gy = c1 + y, gb = c0b + b;
            uint8_t want = (gy < 0 || gy >= H || gb < 0 || gb >= W * 3) ? 0 : img[(size_t)1 * H * W * 3 + gy * W * 3 + gb];
            bad += got[y * box0b + b] != want;
        }
        printf("mismatches %ld", bad);
    }
    printf("\n");
    return 0;
}
