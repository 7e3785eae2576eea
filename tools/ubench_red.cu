// Shared-memory reduction (RED.shared.add.u32) throughput and what ncu's "bank conflict" counter makes of patterns
// that cannot conflict (50.8 M conflicts reported for k_cycle_stats' 126 M reductions).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_red tools/ubench_red.cu
//   tools/ubench_red                                     -> reductions per clock per SM for each pattern
//   ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum,smsp__inst_executed_op_shared_atom.sum,\
//       l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum tools/ubench_red
// Patterns (32 lanes of one instruction):
//   0  lane-private column: word = code * 32 + lane, code random per lane   (the 5-mer tables: 32 different banks)
//   1  one word per lane, same row: word = lane                              (trivially conflict-free)
//   2  the per-cycle counters: word = bin * 544 + row * 33 + lane, bin random per lane (bank = (row + lane) mod 32)
//   3  all lanes the same word                                               (a genuine 32-way collision)
//   4  random word per lane                                                  (random banks: real conflicts)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2048

__device__ __forceinline__ void red(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

template <int PAT>
__global__ void __launch_bounds__(1024) k_red(uint32_t* out, uint32_t seed) {
    extern __shared__ uint32_t tab[];                 // 32768 words
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    const int lane = threadIdx.x & 31;
    uint32_t x = seed ^ (threadIdx.x * 2654435761u) ^ blockIdx.x;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x = x * 1664525u + 1013904223u;
            uint32_t w;
            if (PAT == 0) w = ((x >> 12) & 1023u) * 32u + lane;
            else if (PAT == 1) w = lane + 32u * u;
            else if (PAT == 2) w = ((x >> 12) & 7u) * 544u + (uint32_t)u * 33u + lane;
            else if (PAT == 3) w = 7u;
            else w = (x >> 12) & 32767u;
            red(base + 4u * w, 1u);
        }
    }
    __syncthreads();
    if (tab[threadIdx.x] == 0xdeadbeefu) out[0] = 1;
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 4);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int mhz = 0;
    cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    const int blocks = p.multiProcessorCount, threads = 1024, smem = 32768 * 4;
    const char* names[] = {"lane-private [code][lane]", "one word per lane", "cycle counters (bin*544+row*33+lane)", "all lanes one word", "random words"};
#define RUN(P) { cudaFuncSetAttribute(k_red<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                       \
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);                                                      \
        k_red<P><<<blocks, threads, smem>>>(d, 1u); cudaDeviceSynchronize();                                             \
        cudaEventRecord(a); k_red<P><<<blocks, threads, smem>>>(d, 2u); cudaEventRecord(b); cudaEventSynchronize(b);      \
        float ms = 0; cudaEventElapsedTime(&ms, a, b);                                                                   \
        const double reds = (double)blocks * (threads / 32) * ITERS * 8.0;                                               \
        printf("%-40s %8.3f ms  %6.3f warp-RED/clk/SM\n", names[P], ms, reds / (ms * 1e-3) / (mhz * 1e3) / blocks); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
    return 0;
}
