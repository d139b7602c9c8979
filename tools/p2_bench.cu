// tools/p2_bench.cu — Poseidon2 kernel-shape experiments (instruction-cache behaviour, CTA size,
// round barriers).  Each thread runs NPERM chained permutations (like the leaf sponge: 32 per leaf).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include "../zkvm-brainfuck_b200/csrc/poseidon2.cuh"
#include "../zkvm-brainfuck_b200/csrc/rc_16_30.h"

#define NPERM 32

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) k(uint32_t* out, uint32_t seed) {
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = (seed + threadIdx.x * 16 + blockIdx.x + i) % kb::P;
#pragma unroll 1
    for (int it = 0; it < NPERM; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = (s[i] ^ it) & 0x3fffffff;  // stand-in for the absorbed words
        if (MODE == 0) p2::permute<false>(s);
        if (MODE == 1) p2::permute<true>(s);
        if (MODE == 2) p2::permute_unrolled(s);
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) r ^= s[i];
    out[blockIdx.x * THREADS + threadIdx.x] = r;
}

template <int MODE, int THREADS>
void run(const char* name, uint32_t* d, size_t nthreads) {
    int blocks = (int)(nthreads / THREADS);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, THREADS><<<blocks, THREADS>>>(d, 1);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a);
        k<MODE, THREADS><<<blocks, THREADS>>>(d, 1);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE, THREADS>, THREADS, 0);
    double perms = (double)nthreads * NPERM;
    printf("%-34s threads=%4d occ=%2d blk/SM  %8.3f ms  %7.3f Gperm/s  %6.1f clk/perm/SM@1.965GHz\n", name, THREADS, occ, best,
           perms / best / 1e6, 148 * 1.965e9 / (perms / (best * 1e-3)));
}

int main() {
    p2::Consts h; memset(&h, 0, sizeof h);
    for (int r = 0; r < 4; r++) for (int i = 0; i < 16; i++) { h.ext[r][i] = kb::to_mont(BFGPU_RC_16_30[r][i]); h.ext[4 + r][i] = kb::to_mont(BFGPU_RC_16_30[17 + r][i]); }
    for (int r = 0; r < 13; r++) h.internal[r] = kb::to_mont(BFGPU_RC_16_30[4 + r][0]);
    for (int i = 0; i < 16; i++) h.diag[i] = kb::to_mont(i + 2);
    cudaMemcpyToSymbol(p2::c_p2, &h, sizeof h);
    size_t n = 1 << 21;
    uint32_t* d; cudaMalloc(&d, n * 4);
    run<0, 128>("rolled", d, n);
    run<0, 256>("rolled", d, n);
    run<0, 512>("rolled", d, n);
    run<1, 128>("rolled+barrier/round", d, n);
    run<1, 256>("rolled+barrier/round", d, n);
    run<1, 512>("rolled+barrier/round", d, n);
    run<1, 1024>("rolled+barrier/round", d, n);
    run<2, 128>("unrolled", d, n);
    run<2, 512>("unrolled", d, n);
    return 0;
}
