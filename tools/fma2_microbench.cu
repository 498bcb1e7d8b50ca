#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float v[8]; uint64_t w[8]; int q[8];
    for (int k = 0; k < 8; ++k) { v[k] = threadIdx.x + k; w[k] = pack(v[k], v[k] + 1); q[k] = threadIdx.x * k; }
    uint64_t a2 = pack(a, a), b2 = pack(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0 || MODE == 2) v[k] = fmaf(v[k], a, b);
            if (MODE == 1 || MODE == 3) w[k] = fma2(w[k], a2, b2);
            if (MODE == 2 || MODE == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[k]) : "r"(it), "r"(k));
        }
    }
    float s = 0; for (int k = 0; k < 8; ++k) s += v[k] + __uint_as_float((uint32_t)w[k]) + q[k];
    if (s == -1.f) out[0] = s;
}
template <int MODE> void run(const char* name, double flops_per_iter) {
    float* d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 16384, blocks = 148 * 8;
    float best = 1e9;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
    printf("%-28s %8.3f ms  %7.2f TFLOP/s   %6.2f G warp-instr-slots/s/SMSP-norm\n", name, best, flops_per_iter * iters * blocks * 256.0 / (best * 1e-3) / 1e12, 0.0);
}
int main() {
    run<0>("FFMA x8", 16); run<1>("FFMA2 x8", 32); run<2>("FFMA x8 + LOP3 x8", 16); run<3>("FFMA2 x8 + LOP3 x8", 32);
    return 0;
}
