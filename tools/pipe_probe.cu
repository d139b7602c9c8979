// tools/pipe_probe.cu — integer-pipe throughput microbenchmarks for sm_100a (design input for the
// Poseidon2 / NTT kernels; results are recorded in profiles/).  Each kernel runs NCHAIN independent
// dependency chains per thread of one instruction kind; reports thread-instructions per clock per SM.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define NCHAIN 8
#define ITERS 2048

template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, uint32_t m) {
    uint32_t a[NCHAIN];
    uint64_t w[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) { a[i] = seed + threadIdx.x + i; w[i] = a[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NCHAIN; i++) {
            if (KIND == 0) a[i] = a[i] * m + seed;                                        // IMAD
            if (KIND == 1) w[i] = (uint64_t)(uint32_t)w[i] * m + w[i];                     // IMAD.WIDE
            if (KIND == 2) a[i] = __umulhi(a[i], m) + seed;                                // IMAD.HI
            if (KIND == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));      // IADD3
            if (KIND == 4) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));      // LOP3
            if (KIND == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(m)); // SHF
            if (KIND == 6) { uint32_t s = a[i] + m; uint32_t t = s - 0x7f000001u; a[i] = s < t ? s : t; }  // add + VIADDMNMX
            if (KIND == 7) {  // Montgomery multiply (3 IMAD-class + sub + add/min)
                uint64_t t = (uint64_t)a[i] * m;
                uint32_t q = (uint32_t)t * 0x81000001u;
                uint32_t u = __umulhi(q, 0x7f000001u);
                uint32_t r = (uint32_t)(t >> 32) - u;
                uint32_t r2 = r + 0x7f000001u;
                a[i] = r < r2 ? r : r2;
            }
            if (KIND == 8) {  // mixed: 1 IMAD + 1 IADD + 1 LOP
                a[i] = a[i] * m + seed;
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(m));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(seed));
            }
            if (KIND == 9) {  // mod add a = (a + m) mod p, a = (a + seed) mod p : pure add+min stream
                uint32_t s = a[i] + m; uint32_t t = s - 0x7f000001u; a[i] = s < t ? s : t;
                s = a[i] + seed; t = s - 0x7f000001u; a[i] = s < t ? s : t;
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) r ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    if (r == 0x12345678u) out[0] = r;
}

template <int KIND>
void run(const char* name, double instr_per_iter, int sm_count, double clk_ghz_hint) {
    uint32_t* d;
    cudaMalloc(&d, 4);
    int blocks = sm_count * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<KIND><<<blocks, 256>>>(d, 12345u, 0x9e3779b1u);
    cudaEventRecord(a);
    k<KIND><<<blocks, 256>>>(d, 12345u, 0x9e3779b1u);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double n = (double)blocks * 256 * ITERS * NCHAIN * instr_per_iter;
    double gips = n / (ms * 1e-3) / 1e9;
    printf("%-28s %8.3f ms  %9.1f Ginstr/s  %6.1f instr/clk/SM @%.3f GHz\n", name, ms, gips, gips / sm_count / clk_ghz_hint, clk_ghz_hint);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double ghz = clk_khz / 1e6;
    printf("%s  SMs=%d  clockRate=%.3f GHz (nominal max; actual clock may differ)\n", p.name, p.multiProcessorCount, ghz);
    int sm = p.multiProcessorCount;
    run<0>("IMAD", 1, sm, ghz);
    run<1>("IMAD.WIDE", 1, sm, ghz);
    run<2>("IMAD.HI (+add)", 1, sm, ghz);
    run<3>("IADD3", 1, sm, ghz);
    run<4>("LOP3", 1, sm, ghz);
    run<5>("SHF", 1, sm, ghz);
    run<6>("modadd (IADD+VIADDMNMX)", 2, sm, ghz);
    run<7>("montmul (5 instr)", 5, sm, ghz);
    run<8>("mix IMAD+IADD+LOP", 3, sm, ghz);
    run<9>("2x modadd", 4, sm, ghz);
    return 0;
}
