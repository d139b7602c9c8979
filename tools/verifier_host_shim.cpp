// tools/verifier_host_shim.cpp — host-only build of the native verifier (csrc/verifier.h) for sanitizer / fuzz runs:
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -shared -fPIC \
//       -o tools/bin/libbfverify_asan.so tools/verifier_host_shim.cpp
// The CUDA qualifiers are defined away so that the very same headers the library is built from compile with g++
// (tools/fuzz_verifier.py loads the result instead of libbfgpu.so; see there).
#define __host__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __global__
#define __constant__
#define __shared__
#define __restrict__
#include <algorithm>
#include <array>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
#include "../include/bfgpu.h"
#include "../zkvm-brainfuck_b200/csrc/kb31.cuh"
#include "../zkvm-brainfuck_b200/csrc/rc_16_30.h"
#include "../zkvm-brainfuck_b200/csrc/challenger.h"
namespace air {
struct Selectors {
    uint32_t is_first, is_last, is_trans;
};
struct Challenges {
    kb::Ext alpha;
    kb::Ext beta_pow[8];
    kb::Ext cumulative_sum;
};
constexpr int NUM_CHIPS = 8;
struct ChipInfo { const char* name; int main_w, prep_w, perm_w, n_constraints, local_only, log_quotient_degree; };
extern const ChipInfo CHIPS[NUM_CHIPS];
}  // namespace air
static int chip_index(const char* name);
#include "../zkvm-brainfuck_b200/csrc/verifier.h"
namespace air {
const ChipInfo CHIPS[NUM_CHIPS] = {  // copy of the table at the top of gen_air.cuh (tools/fuzz_verifier.py checks it against the library)
    {"Cpu", 31, 0, 9, 31, 0, 1},    {"Program", 1, 6, 2, 4, 0, 1}, {"AddSub", 7, 0, 4, 14, 1, 1},        {"Jump", 45, 0, 2, 48, 1, 1},
    {"Memory", 12, 0, 3, 5, 0, 1}, {"Byte", 2, 2, 2, 4, 0, 1},    {"MemoryInstrs", 41, 0, 2, 44, 0, 1}, {"IO", 5, 0, 2, 7, 1, 1},
};
}  // namespace air
static int chip_index(const char* name) {
    for (int i = 0; i < air::NUM_CHIPS; i++)
        if (!strcmp(name, air::CHIPS[i].name)) return i;
    return -1;
}
