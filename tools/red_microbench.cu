// What rate of no-return floating-point reductions (RED) into global memory does the B200 L2 sustain?
// Decides whether a Newton's-third-law force sweep (partner forces flushed with one vector RED per (tile, partner))
// can pay on this part.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_microbench red_microbench.cu
//   pattern 0: red.global.add.v4.f32, one warp -> 32 consecutive 16-byte slots of a random 512-byte block (record order)
//   pattern 1: red.global.add.v4.f32, every lane a random 16-byte slot (particle order, random particle numbering)
//   pattern 2: 3 x red.global.add.f32, every lane a random 16-byte slot
//   pattern 3: red.global.add.v2.f32 + red.global.add.f32 (12-byte AoS rows), random rows
//   pattern 4: like 0, but 8 FFMA-chains of work (64 FFMA per lane) between two REDs: does the RED traffic overlap?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_v2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int PATTERN>
__global__ void __launch_bounds__(128) k_red(float* buf, unsigned nslots, int iters, float v) {
    const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = v + k + lane;
    for (int it = 0; it < iters; ++it) {
        unsigned slot;
        if (PATTERN == 0 || PATTERN == 4) slot = ((hash(gw * 7919u + it) % (nslots / 32)) * 32 + lane);
        else slot = hash((gw * 32 + lane) * 2654435761u + it * 97u) % nslots;
        if (PATTERN == 4) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(acc[k], 0.999f, 0.001f);
        }
        if (PATTERN == 0 || PATTERN == 1 || PATTERN == 4) red_v4(buf + (size_t)slot * 4, acc[0], acc[1], acc[2], 0.f);
        else if (PATTERN == 2) { atomicAdd(buf + (size_t)slot * 4, acc[0]); atomicAdd(buf + (size_t)slot * 4 + 1, acc[1]); atomicAdd(buf + (size_t)slot * 4 + 2, acc[2]); }
        else if (PATTERN == 3) {
            // 12-byte rows: the v2 needs 8-byte alignment -> rows at even float offsets only when slot*3 is even
            float* p = buf + (size_t)slot * 3;
            if ((slot & 1u) == 0) { red_v2(p, acc[0], acc[1]); atomicAdd(p + 2, acc[2]); }
            else { atomicAdd(p, acc[0]); red_v2(p + 1, acc[1], acc[2]); }
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    if (s == -1.f) buf[0] = s;
}

template <int PATTERN> void run(const char* name, float* d, unsigned nslots, int blocks, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_red<PATTERN><<<blocks, 128>>>(d, nslots, iters, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double lane_ops = (double)blocks * 128 * iters;
    printf("%-58s %8.3f ms  %10.3e lane-REDs/s  (%6.1f GB/s of 16-byte payload)\n", name, best, lane_ops / (best * 1e-3), lane_ops * 16 / (best * 1e-3) / 1e9);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("no device\n"); return 1; }
    for (unsigned nslots : {1u << 20, 1u << 23}) {   // 16 MB (1M particles, L2 resident) and 128 MB
        float* d;
        cudaMalloc(&d, (size_t)nslots * 16);
        cudaMemset(d, 0, (size_t)nslots * 16);
        printf("%s, %d SMs, buffer %u slots x 16 B\n", prop.name, prop.multiProcessorCount, nslots);
        const int blocks = prop.multiProcessorCount * 8, iters = 2000;
        run<0>("v4 RED, warp-coalesced 512 B blocks", d, nslots, blocks, iters);
        run<1>("v4 RED, random 16 B slot per lane", d, nslots, blocks, iters);
        run<2>("3 scalar REDs, random 16 B slot per lane", d, nslots, blocks, iters);
        run<3>("v2 + scalar RED, random 12 B row per lane", d, nslots, blocks, iters);
        run<4>("v4 RED coalesced + 64 FFMA per lane between REDs", d, nslots, blocks, iters);
        cudaFree(d);
    }
    return 0;
}
