// Which issue pipe do the integer instructions of the hot kernels use on B200, and do they overlap?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_pipes tools/ubench_pipes.cu && tools/ubench_pipes
// Each test runs N dependent-free chains of one instruction kind per thread (8 independent accumulators), 1024
// threads x 148*2 blocks; reported: warp-instructions per cycle per SM sub-partition.  "A+B" interleaves two kinds:
// if the pair runs in max(tA, tB) they sit on different pipes, if in tA + tB on the same one.
// Also: one 16x16 global Levenshtein as (a) the 8-logic-op Myers column used by k_trim and (b) a row-wise DP on the
// DPX min instructions (__viaddmin_s32 / __vimin3_s32), the north_star's suggestion — cells per second of each.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096

template <int KIND>
__global__ void k_pipe(uint32_t* out, uint32_t seed, uint32_t one) {
    uint32_t a[8], b = seed ^ threadIdx.x, c = seed * 2654435761u + blockIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i * 977u + threadIdx.x;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            b = a[(i + 1) & 7];          // operands keep changing: nothing folds into a closed form
            if (KIND == 0 || KIND == 10 || KIND == 11 || KIND == 12) a[i] = (a[i] & b) ^ c;                    // LOP3
            if (KIND == 1 || KIND == 10) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(one | 2u), "r"(c));   // IMAD
            if (KIND == 2 || KIND == 11) a[i] = __dp4a(a[i], 0x08040201u, b);                                // IDP.4A
            if (KIND == 3 || KIND == 12) a[i] = __funnelshift_r(a[i], b, 7);                                 // SHF
            if (KIND == 4) a[i] = __byte_perm(a[i], b, 0x2103);                                              // PRMT
            if (KIND == 5) a[i] = a[i] + b + c;                                                              // IADD3
            if (KIND == 6) a[i] = __popc(a[i]) + b;                                                          // POPC
            if (KIND == 7) a[i] = (uint32_t)__viaddmin_s32((int)a[i], (int)b, (int)c);                       // VIADDMNMX
            if (KIND == 8) a[i] = (uint32_t)__vimin3_s32((int)a[i], (int)b, (int)c);                         // VIMNMX3
            if (KIND == 9) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(one | 2u), "r"(c)); a[i] = __dp4a(a[i], 0x08040201u, b); }  // IMAD + IDP
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= a[i];
    if (s == 0x12345678u) out[0] = s;
}

// (a) Myers/Hyyro bit-parallel column, the form of fpl_trim.cu:myers16 (pattern and text from registers)
__global__ void k_myers(uint32_t* out, uint32_t seed) {
    uint32_t acc = 0;
    uint32_t eq[4] = {0x1111u * (seed | 1u), 0x2222u ^ seed, 0x4444u + threadIdx.x, 0x8888u ^ threadIdx.x};
    for (int it = 0; it < ITERS / 4; it++) {
        uint32_t VP = 0xFFFFu, VN = 0, aP = 0, aN = 0;
        uint32_t text = seed + it * 2654435761u + threadIdx.x;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t Eq = eq[(text >> (2 * i)) & 3u];
            const uint32_t Xv = Eq | VN;
            const uint32_t Xh = (((Eq & VP) + VP) ^ VP) | Eq;
            uint32_t HP = VN | ~(Xh | VP);
            uint32_t HN = VP & Xh;
            aP += HP & 0x8000u; aN += HN & 0x8000u;
            HP = HP * 2u + 1u; HN = HN * 2u;
            VP = HN | ~(Xv | HP);
            VN = HP & Xv;
        }
        acc += 16 + (aP >> 15) - (aN >> 15);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// (b) the same 16x16 global distance, one DP row at a time on the DPX min instructions:
// d[j] = min(d_prev[j] + 1, d[j-1] + 1, d_prev[j-1] + (a_i != b_j))
__global__ void k_dpx(uint32_t* out, uint32_t seed) {
    uint32_t acc = 0;
    for (int it = 0; it < ITERS / 4; it++) {
        const uint32_t pat = seed * 747796405u + it, text = seed + it * 2654435761u + threadIdx.x;
        int row[17];
#pragma unroll
        for (int j = 0; j <= 16; j++) row[j] = j;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t ai = (pat >> (2 * i)) & 3u;
            int diag = row[0];
            row[0] = i + 1;
#pragma unroll
            for (int j = 1; j <= 16; j++) {
                const int sub = diag + (int)(((text >> (2 * (j - 1))) & 3u) != ai);
                diag = row[j];
                row[j] = __viaddmin_s32(__vimin3_s32(row[j], row[j - 1], sub - 1), 1, 0x7fffffff);   // min3 + 1
            }
        }
        acc += row[16];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    uint32_t* d;
    cudaMalloc(&d, 4);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int mhz = 0;
    cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    const int blocks = p.multiProcessorCount * 2, threads = 1024;
    const double warps = (double)blocks * threads / 32, smsp = p.multiProcessorCount * 4.0;
    printf("%s, %d SMs, %.0f MHz nominal\n", p.name, p.multiProcessorCount, mhz / 1e3);
    const char* names[] = {"LOP3", "IMAD", "IDP.4A", "SHF", "PRMT", "IADD3", "POPC", "VIADDMNMX", "VIMNMX3", "IMAD+IDP", "LOP3+IMAD", "LOP3+IDP", "LOP3+SHF"};
    const int per[] = {1, 1, 1, 1, 1, 1, 2, 1, 1, 2, 2, 2, 2};
#define RUN(K) { float ms = time_ms([&] { k_pipe<K><<<blocks, threads>>>(d, 12345u, 1u); });                                   \
        const double insts = warps * ITERS * 8.0 * per[K];                                                                     \
        printf("%-10s %8.3f ms  %6.3f warp-inst/clk/SMSP (at %d MHz)\n", names[K], ms, insts / (ms * 1e-3) / (mhz * 1e3) / smsp, mhz / 1000); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
    {
        float ms = time_ms([&] { k_myers<<<blocks, threads>>>(d, 12345u); });
        const double probes = (double)blocks * threads * (ITERS / 4);
        printf("Myers16    %8.3f ms  %8.2f G 16x16 distances/s\n", ms, probes / (ms * 1e-3) / 1e9);
        ms = time_ms([&] { k_dpx<<<blocks, threads>>>(d, 12345u); });
        printf("DPX rows   %8.3f ms  %8.2f G 16x16 distances/s\n", ms, probes / (ms * 1e-3) / 1e9);
    }
    return 0;
}
