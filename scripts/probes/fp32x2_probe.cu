// Throughput probe (B200): scalar FADD / FMUL / FFMA vs packed add.rn.f32x2 / fma.rn.f32x2, per SM per clock.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32x2_probe fp32x2_probe.cu && ./fp32x2_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
constexpr int ITERS = 4096, ILP = 8;
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float addf(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float mulf(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmaf_(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
template <int MODE> __global__ void k(float *out, float seed) {
    float x[ILP]; uint64_t y[ILP];
    for (int i = 0; i < ILP; ++i) { x[i] = seed + i + threadIdx.x; y[i] = ((uint64_t)__float_as_uint(x[i]) << 32) | __float_as_uint(x[i] + 1.f); }
    const float c = seed * 0.5f; const uint64_t c2 = ((uint64_t)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) x[i] = addf(x[i], c);
            if (MODE == 1) x[i] = mulf(x[i], 0.25f);
            if (MODE == 2) x[i] = fmaf_(x[i], c, c);
            if (MODE == 3) y[i] = add2(y[i], c2);
            if (MODE == 4) y[i] = fma2(y[i], c2, c2);
            if (MODE == 5) x[i] = addf(x[i], x[(i + 1) % ILP]);   // two varying register operands
        }
    }
    float s = 0; for (int i = 0; i < ILP; ++i) s += x[i] + __uint_as_float((uint32_t)y[i]) + __uint_as_float((uint32_t)(y[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int warps, float *out, int sms, float ghz) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<sms, 32 * warps>>>(out, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(a); k<MODE><<<sms, 32 * warps>>>(out, 1.0f); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double insts = (double)ITERS * ILP * warps;   // warp instructions per SM
    printf("%-28s warps/SM=%2d  %.3f ms  %.2f warp-inst/clk/SM (at %.2f GHz)  lanes/clk/SM=%.0f\n", name, warps, ms,
           insts / (ms * 1e-3 * ghz * 1e9), ghz, insts / (ms * 1e-3 * ghz * 1e9) * 32 * (MODE >= 3 && MODE <= 4 ? 2 : 1));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float ghz = khz * 1e-6f; float *out; cudaMalloc(&out, 148 * 1024 * 4);
    printf("%s, %d SMs, clock attr %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    for (int w : {8, 16, 32}) {
        run<0>("FADD r,r", w, out, p.multiProcessorCount, ghz);
        run<5>("FADD r,r (2 varying)", w, out, p.multiProcessorCount, ghz);
        run<1>("FMUL r,imm", w, out, p.multiProcessorCount, ghz);
        run<2>("FFMA r,r,r", w, out, p.multiProcessorCount, ghz);
        run<3>("FADD2 (add.f32x2)", w, out, p.multiProcessorCount, ghz);
        run<4>("FFMA2 (fma.f32x2)", w, out, p.multiProcessorCount, ghz);
    }
    return 0;
}
