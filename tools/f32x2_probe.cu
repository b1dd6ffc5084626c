// Throughput probe: scalar FADD/FFMA vs packed add.f32x2 / fma.rn.f32x2 on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/f32x2_probe tools/f32x2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define NACC 8
__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
template <int MODE> __global__ void k(float* out, int iters, float s) {
    float ax[NACC], ay[NACC];
    unsigned long long p[NACC];
    for (int i = 0; i < NACC; ++i) { ax[i] = threadIdx.x * 0.001f + i; ay[i] = ax[i] * 0.5f; p[i] = pk(ax[i], ay[i]); }
    unsigned long long ps = pk(s, s * 0.5f), pm = pk(0.999f, 1.001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) { ax[i] += s; ay[i] += s * 0.5f; }                                   // 2 FADD
            if (MODE == 1) { asm volatile("add.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps)); }    // 1 FADD2
            if (MODE == 2) { ax[i] = fmaf(ax[i], 0.999f, s); ay[i] = fmaf(ay[i], 1.001f, s * 0.5f); }  // 2 FFMA
            if (MODE == 3) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pm), "l"(ps)); }  // 1 FFMA2
        }
    }
    float acc = 0;
    for (int i = 0; i < NACC; ++i) { float a, b; upk(p[i], a, b); acc += ax[i] + ay[i] + a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, float* out) {
    int iters = 20000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 100, 1e-3f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters, 1e-3f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * (double)iters * NACC * 2;   // scalar-equivalent FP ops (an FMA counted once)
    printf("%-28s %.3f ms  %.2f T scalar-op/s  (%.1f op/clk/SM at 1.965 GHz)\n", name, ms, ops / ms * 1e-9, ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("FADD  (2 per pair)", out); run<1>("add.f32x2 (1 per pair)", out);
    run<2>("FFMA  (2 per pair)", out); run<3>("fma.rn.f32x2 (1 per pair)", out);
    return 0;
}
