// Standalone TMA probe: loads one 2-D box (box_cols x box_rows of `esz`-byte elements) at (c0, r0) into smem.
// usage: tma_probe box_cols box_rows esz c0 r0 [align_shift]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int r0, uint32_t bytes, float *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(smem)),
                     "l"(&map), "r"(c0), "r"(r0), "r"(s32(&bar))
                     : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
    float acc = 0;
    for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) acc += ((float *)smem)[i];
    atomicAdd(out, acc);
}
int main(int argc, char **argv) {
    int bc = atoi(argv[1]), br = atoi(argv[2]), esz = atoi(argv[3]), c0 = atoi(argv[4]), r0 = atoi(argv[5]);
    const uint64_t cols = 1024, rows = 1024;
    void *buf; cudaMalloc(&buf, cols * rows * esz); cudaMemset(buf, 0, cols * rows * esz);
    float *out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    CUtensorMap m; cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * (uint64_t)esz}; cuuint32_t box[2] = {(cuuint32_t)bc, (cuuint32_t)br}, es[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fp)(&m, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    uint32_t bytes = (uint32_t)bc * br * esz;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe<<<2, 256, bytes + 128>>>(m, c0, r0, bytes, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("box %dx%d esz %d at (%d,%d): encode=%d run=%s\n", bc, br, esz, c0, r0, (int)r, cudaGetErrorString(e));
    return 0;
}
