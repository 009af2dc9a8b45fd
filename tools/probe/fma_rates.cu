// Issue-rate probe for the fp32 / fp16 multiply-add flavours of sm_100a: FFMA, FFMA2 (fma.rn.f32x2), FHFMA (fma.rn.f32.f16),
// HFMA2 (fma.rn.f16x2), HADD2.F32 conversions. Each kernel runs independent dependency chains per thread, so the result is
// the pipe rate, not the latency.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rates fma_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int kIters = 4096, kChains = 8;

__global__ void k_ffma(float* out, float a, float b) {
    float acc[kChains];
    for (int i = 0; i < kChains; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < kIters; ++it)
#pragma unroll
        for (int i = 0; i < kChains; ++i) acc[i] = fmaf(acc[i], a, b);
    float s = 0; for (int i = 0; i < kChains; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long acc[kChains], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    for (int i = 0; i < kChains; ++i) { float v = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(v)); }
    for (int it = 0; it < kIters; ++it)
#pragma unroll
        for (int i = 0; i < kChains; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(aa), "l"(bb));
    float s = 0;
    for (int i = 0; i < kChains; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i])); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fhfma(float* out, float a, float b) {
    float acc[kChains];
    unsigned short ha = __half_as_ushort(__float2half(a)), hb = __half_as_ushort(__float2half(b));
    for (int i = 0; i < kChains; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < kIters; ++it)
#pragma unroll
        for (int i = 0; i < kChains; ++i) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(ha), "h"(hb));
    float s = 0; for (int i = 0; i < kChains; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_hfma2(float* out, float a, float b) {
    __half2 acc[kChains], aa = __float2half2_rn(a), bb = __float2half2_rn(b);
    for (int i = 0; i < kChains; ++i) acc[i] = __float2half2_rn(threadIdx.x * 0.001f + i);
    for (int it = 0; it < kIters; ++it)
#pragma unroll
        for (int i = 0; i < kChains; ++i) acc[i] = __hfma2(acc[i], aa, bb);
    float s = 0; for (int i = 0; i < kChains; ++i) s += __low2float(acc[i]) + __high2float(acc[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_cvt(float* out, float a, float b) {      // half2 -> float2 conversions feeding an FADD each (the FIR's input pattern)
    float acc[kChains];
    unsigned hv = 0x3c003c00u + threadIdx.x;
    for (int i = 0; i < kChains; ++i) acc[i] = i;
    for (int it = 0; it < kIters; ++it)
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            __half2 h = *reinterpret_cast<__half2*>(&hv);
            float2 f = __half22float2(h);
            acc[i] += f.x; acc[i] += f.y;
            hv += 0x00010001u * (unsigned)(acc[i] > 1e30f);
        }
    float s = 0; for (int i = 0; i < kChains; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static void run(const char* name, K kern, double flop_per_instr, float* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256;
    kern<<<blocks, threads>>>(out, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) kern<<<blocks, threads>>>(out, 0.999f, 0.001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double instr = (double)blocks * threads / 32 * kIters * kChains;     // warp instructions
    const double per_clk_sm = instr / (ms * 1e-3 * 1.965e9 * 148);
    printf("%-8s %8.3f ms  %6.3f warp-instr/clk/SM  %7.1f lane-ops/clk/SM (x%.0f ops per lane-instr)  err=%s\n", name, ms, per_clk_sm,
           per_clk_sm * 32 * flop_per_instr, flop_per_instr, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    run("FFMA", k_ffma, 1, out);
    run("FFMA2", k_ffma2, 2, out);
    run("FHFMA", k_fhfma, 1, out);
    run("HFMA2", k_hfma2, 2, out);
    run("CVT+ADD", k_cvt, 1, out);
    return 0;
}
