// tracegen.cuh — native executor and device-side trace generation.
//
// The reference runs the program on the host (`Executor::run`, crates/core/executor/src/executor.rs:71-79,106-325;
// `Program::from`, program.rs:22-44), then every chip's `MachineAir::generate_trace` builds its RowMajorMatrix on the
// host (Cpu cpu/trace.rs:28-150 + memory/consistency/trace.rs:9-77, Program program/mod.rs:64-137, AddSub
// alu/mod.rs:62-157, Jump jump/trace.rs:31-96, Memory memory/memory.rs:84-126, Byte bytes/trace.rs:39-60,
// MemoryInstrs memory/instructions/trace.rs:36-96, IO io/mod.rs:72-124; padding utils/mod.rs:25-53) and
// `generate_dependencies` (machine.rs:228-248) counts the byte-lookup multiplicities.  ≈1.1 GB of traces then cross
// PCIe for a 2^22-cycle run.
//
// Here the executor (C++, one pass) emits ONE 16-byte record per cycle {pc, mp, previous timestamp of the cell,
// value | previous value << 8}; everything else in the eight traces is a function of that record, its successor and
// the program, so the GPU rebuilds the traces directly in the prover's device layout (column-major, Montgomery,
// bit-reversed rows — what `ingest` would have produced), 67 MB over PCIe instead of 1.1 GB:
//   k_classify + prefix scan + k_index : event index lists of the AddSub / Jump / MemoryInstrs / IO chips
//   k_byte_hist                        : u8 / u16 lookup multiplicities (shared-memory privatised histograms)
//   k_*_trace                          : one thread per stored row; reads its cycle record(s), writes its row coalesced
// Parity: tests/test_gpu_tracegen_parity.py compares every generated trace with the numpy restatement of the reference's
// generators and the resulting proofs word for word.
#pragma once
#include <unordered_map>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "gen_layout.cuh"

// ---- host executor -------------------------------------------------------------------------------------------------------
struct bfgpu_record {
    std::vector<uint8_t> ops;       // opcode per instruction
    std::vector<uint32_t> args;     // jump targets (program.rs:22-44)
    uint4* cycles = nullptr;        // n_cycles + 1 records (the last one is a sentinel holding the final pc / mp)
    bool pinned = false;
    bfgpu_ctx* owner = nullptr;     // context whose page-locked buffer cache the cycle buffer returns to
    uint64_t n_cycles = 0, cap = 0;
    std::vector<uint32_t> mem_events;  // 5 words per touched cell, first-access order: addr, initial ts, initial value, final ts, final value
    std::vector<uint32_t> prog_counts;  // executions per instruction (Program chip multiplicities)
    uint64_t n_alu = 0, n_jump = 0, n_mem = 0, n_io = 0;
    std::vector<uint8_t> output;
    std::string err;
};

static bool record_grow(bfgpu_record* r, uint64_t want, uint64_t used) {
    if (want <= r->cap) return true;
    uint64_t cap = std::max<uint64_t>(r->cap * 2, 1u << 16);
    while (cap < want) cap *= 2;
    // page-locked buffers are allocated once at the shard limit (2^23 cycles + sentinel, 128 MB) and recycled through the
    // context: cudaHostAlloc / cudaFreeHost cost tens of milliseconds and serialise against the GPU work of other threads
    if (r->pinned) cap = std::max<uint64_t>(cap, (1ull << 23) + 2);
    uint4* p = nullptr;
    if (r->pinned) {
        if (cudaHostAlloc((void**)&p, cap * sizeof(uint4), cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
    } else {
        p = (uint4*)realloc(r->cycles, cap * sizeof(uint4));  // large blocks are remapped, not copied
        if (!p) return false;
        r->cycles = p;
        r->cap = cap;
        return true;
    }
    if (r->cycles) {
        memcpy(p, r->cycles, used * sizeof(uint4));
        if (r->pinned) cudaFreeHost(r->cycles);
        else free(r->cycles);
    }
    r->cycles = p;
    r->cap = cap;
    return true;
}

extern "C" void bfgpu_record_free(bfgpu_record* r) {
    if (!r) return;
    if (r->cycles) {
        if (!r->pinned) free(r->cycles);
        else {
            // page-locking 128 MB costs tens of milliseconds: hand the buffer to the owning context for the next execution
            std::lock_guard<std::mutex> g(g_ctx_mutex);
            if (g_live_ctx.count(r->owner) && r->owner->pinned_pool.size() < 4) r->owner->pinned_pool.emplace_back(r->cycles, r->cap);
            else cudaFreeHost(r->cycles);
        }
    }
    delete r;
}
extern "C" const char* bfgpu_record_error(const bfgpu_record* r) { return r ? r->err.c_str() : "null record"; }

// Compile and run.  ctx may be null (plain host memory; used by the CPU tests); with a context the cycle records are
// written into page-locked memory so the later copy runs at PCIe speed.
extern "C" int32_t bfgpu_execute(bfgpu_ctx* ctx, const char* code, const uint8_t* stdin_bytes, uint64_t n_stdin, uint64_t max_cycles, bfgpu_record** out) {
    if (!code || !out) return BFGPU_ERR_INVALID;
    *out = nullptr;
    bfgpu_record* r = new bfgpu_record();
    *out = r;  // returned even on failure so the caller can read the message
    r->pinned = ctx != nullptr;
    r->owner = ctx;
    if (ctx) {
        cudaSetDevice(ctx->device);  // may run on a helper thread (CudaProver.prove_many): page-lock against the right device
        std::lock_guard<std::mutex> g(g_ctx_mutex);
        if (!ctx->pinned_pool.empty()) {
            r->cycles = (uint4*)ctx->pinned_pool.back().first;
            r->cap = ctx->pinned_pool.back().second;
            ctx->pinned_pool.pop_back();
        }
    }
    using namespace lay;
    {  // Program::from
        std::vector<uint32_t> stack;
        for (const char* c = code; *c; c++) {
            int op = -1;
            switch (*c) {
                case '>': op = OP_MEM_FWD; break;
                case '<': op = OP_MEM_BWD; break;
                case '+': op = OP_ADD; break;
                case '-': op = OP_SUB; break;
                case '.': op = OP_OUTPUT; break;
                case ',': op = OP_INPUT; break;
                case '[': op = OP_LOOP_START; break;
                case ']': op = OP_LOOP_END; break;
                case ' ': case '\n': case '\r': continue;
                default:
                    r->err = std::string("unexpected character '") + *c + "' in program";
                    return BFGPU_ERR_INVALID;
            }
            uint32_t arg = 0;
            if (op == OP_LOOP_START) stack.push_back((uint32_t)r->ops.size());
            if (op == OP_LOOP_END) {
                if (stack.empty()) {
                    r->err = "unmatched ']'";
                    return BFGPU_ERR_INVALID;
                }
                uint32_t start = stack.back();
                stack.pop_back();
                r->args[start] = (uint32_t)r->ops.size();  // the LOOP_END itself (executor.rs quirk kept: dst = index of ']')
                arg = start + 1;
            }
            r->ops.push_back((uint8_t)op);
            r->args.push_back(arg);
        }
        if (!stack.empty()) {
            r->err = "unmatched '['";
            return BFGPU_ERR_INVALID;
        }
    }
    const uint32_t n = (uint32_t)r->ops.size();
    r->prog_counts.assign(n, 0);
    if (max_cycles == 0 || max_cycles > (1ull << 23)) max_cycles = 1ull << 23;  // clk = 2 * cycle must stay below 2^24 (24-bit range checks)
    struct Cell { uint32_t ts; uint8_t val; uint8_t seen; };
    std::vector<Cell> flat(1u << 16, Cell{0, 0, 0});
    std::unordered_map<uint32_t, Cell> far;
    auto slow_cell = [&flat, &far](uint32_t a) -> Cell* {  // beyond the flat table: grow it (addresses below 2^26) or use the map (wrapped pointers)
        if (a < (1u << 26)) {
            size_t s = flat.size();
            while (s <= a) s *= 2;
            flat.resize(s, Cell{0, 0, 0});
            return flat.data() + a;
        }
        return &far[a];
    };
    std::vector<uint32_t> first_order;  // addresses in first-access order
    // Everything the loop touches every cycle lives in locals whose address is never taken (the lambda above captures only the two
    // containers): with `flat_p / flat_n / i` captured by reference, every store through `counts`, `cyc` or a cell pointer could
    // alias them and the compiler reloaded them from the stack each cycle.
    const uint8_t* __restrict__ ops = r->ops.data();
    const uint32_t* __restrict__ args = r->args.data();
    uint32_t* __restrict__ counts = r->prog_counts.data();
    uint4* __restrict__ cyc = r->cycles;
    Cell* flat_p = flat.data();
    uint32_t flat_n = (uint32_t)flat.size();
    // the limit applies to recycled (pooled, already large) record buffers as well: clamp before the loop, not only after a growth
    uint64_t cap = std::min<uint64_t>(r->cap, max_cycles + 1);
    uint32_t pc = 0, mp = 0;
    uint64_t i = 0;
    // 16 bytes per cycle into a buffer that is only read back by the DMA engine: non-temporal stores skip the read-for-ownership of
    // every cache line (half of the interpreter's memory traffic).  $BFGPU_EXEC_NT=0/1 overrides (default: on for page-locked buffers,
    // whose pages exist already; on fresh pageable memory the page faults dominate either way).
    bool nt_store = r->pinned;
    if (const char* e = getenv("BFGPU_EXEC_NT")) nt_store = atoi(e) != 0;
    auto put = [nt_store](uint4* dst, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
#if defined(__x86_64__)
        if (nt_store) {
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst), _mm_set_epi32((int)w, (int)z, (int)y, (int)x));
            return;
        }
#endif
        *dst = make_uint4(x, y, z, w);
    };
    while (pc != n) {
        if (i + 2 > cap) {
            if (i >= max_cycles) {
                r->err = "cycle limit exceeded (one shard holds at most 2^23 cycles)";
                return BFGPU_ERR_INVALID;
            }
            if (!record_grow(r, i + 2, i)) {
                r->err = "out of host memory";
                return BFGPU_ERR_OOM;
            }
            cyc = r->cycles;
            cap = std::min<uint64_t>(r->cap, max_cycles + 1);  // re-enter this branch exactly when the limit is reached
        }
        const uint32_t op = ops[pc], clk = (uint32_t)(2 * i);
        counts[pc]++;
        if (op == OP_MEM_FWD || op == OP_MEM_BWD) {
            put(cyc + i, pc, mp, 0, 0);
            mp = op == OP_MEM_FWD ? mp + 1 : mp - 1;
            pc++;
            i++;
            continue;
        }
        Cell* c;
        if (mp < flat_n) c = flat_p + mp;
        else {
            c = slow_cell(mp);
            flat_p = flat.data();
            flat_n = (uint32_t)flat.size();
        }
        if (!c->seen) {
            c->seen = 1;
            first_order.push_back(mp);
            r->mem_events.insert(r->mem_events.end(), {mp, c->ts, (uint32_t)c->val, 0u, 0u});
        }
        const uint32_t pv = c->val, pt = c->ts;
        uint32_t mv = pv, next_pc = pc + 1;
        if (op == OP_ADD || op == OP_SUB) {
            c->val = (uint8_t)(op == OP_ADD ? mv + 1 : mv - 1);
            c->ts = clk + 2;
        } else {
            c->ts = clk + 1;
            if (op == OP_LOOP_START || op == OP_LOOP_END) {
                if ((op == OP_LOOP_START) == (mv == 0)) next_pc = args[pc];  // '[' jumps on zero, ']' on non-zero
            } else if (op == OP_INPUT) {
                if (n_stdin == 0) {
                    r->err = "',' executed with empty stdin";
                    return BFGPU_ERR_INVALID;
                }
                mv = stdin_bytes[0];  // the reference never advances the input pointer (executor.rs:183)
                c->val = (uint8_t)mv;
            } else {  // OUTPUT
                r->output.push_back((uint8_t)mv);
            }
        }
        put(cyc + i, pc, mp, pt, mv | (pv << 8));
        pc = next_pc;
        i++;
    }
    // event counts per chip: every execution of an instruction is one event of its class
    uint64_t n_alu = 0, n_jump = 0, n_mem = 0, n_io = 0;
    for (uint32_t k = 0; k < n; k++) {
        const uint32_t op = ops[k];
        const uint64_t cnt = counts[k];
        if (op == OP_MEM_FWD || op == OP_MEM_BWD) n_mem += cnt;
        else if (op == OP_ADD || op == OP_SUB) n_alu += cnt;
        else if (op == OP_LOOP_START || op == OP_LOOP_END) n_jump += cnt;
        else n_io += cnt;
    }
    r->n_alu = n_alu;
    r->n_jump = n_jump;
    r->n_mem = n_mem;
    r->n_io = n_io;
    auto cell = [&flat, &far](uint32_t a) -> Cell& { return a < flat.size() ? flat[a] : far[a]; };
    if (!record_grow(r, i + 1, i)) {
        r->err = "out of host memory";
        return BFGPU_ERR_OOM;
    }
#if defined(__x86_64__)
    _mm_sfence();  // non-temporal record stores are globally visible before anything (DMA included) reads the buffer
#endif
    r->cycles[i] = make_uint4(pc, mp, 0, 0);  // sentinel: next_pc / next_mp of the last cycle
    r->n_cycles = i;
    for (size_t k = 0; k < first_order.size(); k++) {
        const Cell& c = cell(first_order[k]);
        r->mem_events[5 * k + 3] = c.ts;
        r->mem_events[5 * k + 4] = c.val;
    }
    return BFGPU_OK;
}

// counts: cycles, instructions, alu / jump / memory-instruction / io events, touched cells, output bytes
extern "C" int32_t bfgpu_record_info(const bfgpu_record* r, uint64_t counts[8]) {
    if (!r || !counts) return BFGPU_ERR_INVALID;
    counts[0] = r->n_cycles;
    counts[1] = r->ops.size();
    counts[2] = r->n_alu;
    counts[3] = r->n_jump;
    counts[4] = r->n_mem;
    counts[5] = r->n_io;
    counts[6] = r->mem_events.size() / 5;
    counts[7] = r->output.size();
    return BFGPU_OK;
}
extern "C" int32_t bfgpu_record_output(const bfgpu_record* r, uint8_t* out) {
    if (!r || (!out && !r->output.empty())) return BFGPU_ERR_INVALID;
    if (!r->output.empty()) memcpy(out, r->output.data(), r->output.size());
    return BFGPU_OK;
}
// raw views for the parity tests: cycle records (n_cycles + 1 x 4 words), memory events (cells x 5), program (ops, args)
extern "C" const uint32_t* bfgpu_record_cycles(const bfgpu_record* r) { return r ? (const uint32_t*)r->cycles : nullptr; }
extern "C" const uint32_t* bfgpu_record_mem_events(const bfgpu_record* r) { return r ? r->mem_events.data() : nullptr; }
extern "C" int32_t bfgpu_record_program(const bfgpu_record* r, uint32_t* ops, uint32_t* args) {
    if (!r || !ops || !args) return BFGPU_ERR_INVALID;
    for (size_t k = 0; k < r->ops.size(); k++) {
        ops[k] = r->ops[k];
        args[k] = r->args[k];
    }
    return BFGPU_OK;
}

// ---- device kernels ----------------------------------------------------------------------------------------------------------
namespace tg {
using namespace lay;

__device__ __forceinline__ uint32_t M(uint32_t v) { return kb::mul(v, kb::R2); }  // any 32-bit value -> Montgomery form of v mod p

struct Dec {
    uint32_t op, pc, next_pc, mp, next_mp, mv, pv, next_mv, pt, clk;
    bool acc, nacc;
};
__device__ __forceinline__ Dec decode(const uint4* __restrict__ cyc, const uint8_t* __restrict__ ops, uint32_t i) {
    const uint4 c = cyc[i], nx = cyc[i + 1];
    Dec d;
    d.pc = c.x;
    d.mp = c.y;
    d.pt = c.z;
    d.op = ops[c.x];
    d.next_pc = nx.x;
    d.next_mp = nx.y;
    d.clk = 2 * i;
    const bool mem = d.op == OP_MEM_FWD || d.op == OP_MEM_BWD;
    d.acc = !mem;
    d.mv = mem ? 0u : (c.w & 0xFFu);
    d.pv = mem ? 0u : ((c.w >> 8) & 0xFFu);
    d.nacc = d.op == OP_ADD || d.op == OP_SUB;
    d.next_mv = d.op == OP_ADD ? ((d.mv + 1) & 0xFFu) : d.op == OP_SUB ? ((d.mv - 1) & 0xFFu) : 0u;
    return d;
}

// flags[i] = (is alu, is jump, is memory instruction, is io) of cycle i
__global__ void k_classify(const uint4* __restrict__ cyc, const uint8_t* __restrict__ ops, uint32_t n, uint4* __restrict__ flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t op = ops[cyc[i].x];
    flags[i] = make_uint4(op == OP_ADD || op == OP_SUB, op == OP_LOOP_START || op == OP_LOOP_END, op == OP_MEM_FWD || op == OP_MEM_BWD,
                          op == OP_INPUT || op == OP_OUTPUT);
}
// scan = inclusive prefix sums of flags: the k-th event of a class happened at cycle idx_class[k]
__global__ void k_index(const uint4* __restrict__ cyc, const uint8_t* __restrict__ ops, const uint4* __restrict__ scan, uint32_t n,
                        uint32_t* __restrict__ idx_alu, uint32_t* __restrict__ idx_jump, uint32_t* __restrict__ idx_mem, uint32_t* __restrict__ idx_io) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t op = ops[cyc[i].x];
    uint4 s = scan[i];
    if (op == OP_ADD || op == OP_SUB) idx_alu[s.x - 1] = i;
    else if (op == OP_LOOP_START || op == OP_LOOP_END) idx_jump[s.y - 1] = i;
    else if (op == OP_MEM_FWD || op == OP_MEM_BWD) idx_mem[s.z - 1] = i;
    else idx_io[s.w - 1] = i;
}

// byte-lookup multiplicities of the whole shard (what generate_dependencies collects): u8[256], u16[65536]
constexpr int HIST_THREADS = 256, HIST_LOW = 2048;
__global__ void __launch_bounds__(HIST_THREADS) k_byte_hist(const uint4* __restrict__ cyc, const uint8_t* __restrict__ ops, uint32_t n,
                                                            unsigned int* __restrict__ u8, unsigned int* __restrict__ u16) {
    __shared__ unsigned int s8[256], s16[HIST_LOW];
    for (int k = threadIdx.x; k < 256; k += HIST_THREADS) s8[k] = 0;
    for (int k = threadIdx.x; k < HIST_LOW; k += HIST_THREADS) s16[k] = 0;
    __syncthreads();
    auto add8 = [&](uint32_t v) { atomicAdd(&s8[v & 0xFF], 1u); };
    auto add16 = [&](uint32_t v) {
        v &= 0xFFFF;
        if (v < HIST_LOW) atomicAdd(&s16[v], 1u);
        else atomicAdd(&u16[v], 1u);
    };
    for (uint32_t i = blockIdx.x * HIST_THREADS + threadIdx.x; i < n; i += gridDim.x * HIST_THREADS) {
        Dec d = decode(cyc, ops, i);
        add16(d.clk);  // eval_clk range check (cpu/trace.rs)
        add8(d.clk >> 16);
        if (d.acc) {  // memory access timestamp differences (memory/consistency/trace.rs:9-77)
            uint32_t diff = d.clk - d.pt;  // (clk + 1) - prev_ts - 1
            add16(diff);
            add8(diff >> 16);
        }
        if (d.nacc) {  // second access of an ALU cycle: (clk + 2) - (clk + 1) - 1 = 0
            add16(0);
            add8(0);
        }
        add8(d.mv);
        if (d.nacc) {  // AddSub chip (alu/mod.rs:62-157): operand_1, operand_2 = 1, value
            uint32_t op1 = d.op == OP_ADD ? d.mv : d.next_mv;
            add8(op1);
            add8(1);
            add8(op1 + 1);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 256; k += HIST_THREADS)
        if (s8[k]) atomicAdd(&u8[k], s8[k]);
    for (int k = threadIdx.x; k < HIST_LOW; k += HIST_THREADS)
        if (s16[k]) atomicAdd(&u16[k], s16[k]);
}

// All trace kernels: thread t owns STORED row t = natural row bitrev(t); the matrix was zeroed, padding rows return.
#define TG_ROW_PROLOGUE(n_real)                                   \
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;     \
    const uint64_t rows = 1ull << log_rows;                       \
    if (t >= rows) return;                                        \
    const uint32_t r = kb::bitrev(t, log_rows);                   \
    if (r >= (n_real)) return;                                    \
    auto put = [&](int col, uint32_t v) { out[(uint64_t)col * rows + t] = M(v); }

__device__ __forceinline__ void put_word(uint32_t* out, uint64_t rows, uint32_t t, int col, uint32_t v) {
#pragma unroll
    for (int k = 0; k < 4; k++) out[(uint64_t)(col + k) * rows + t] = M((v >> (8 * k)) & 0xFFu);
}
// KoalaBearWordRangeChecker::populate (operations/koala_bear_word.rs:37-50): 8 bits of the top byte + running ANDs
__device__ __forceinline__ void put_range_checker(uint32_t* out, uint64_t rows, uint32_t t, int col, uint32_t v) {
    uint32_t bits[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        bits[k] = (v >> (24 + k)) & 1u;
        out[(uint64_t)(col + k) * rows + t] = M(bits[k]);
    }
    uint32_t a = bits[0] & bits[1];
    out[(uint64_t)(col + 8) * rows + t] = M(a);
#pragma unroll
    for (int k = 2; k < 7; k++) {
        a &= bits[k];
        out[(uint64_t)(col + 7 + k) * rows + t] = M(a);
    }
}

__global__ void __launch_bounds__(256) k_cpu_trace(const uint4* __restrict__ cyc, const uint8_t* __restrict__ ops, const uint32_t* __restrict__ args,
                                                   uint32_t n, unsigned log_rows, uint32_t* __restrict__ out) {
    TG_ROW_PROLOGUE(n);
    const Dec d = decode(cyc, ops, r);
    namespace L = lay::cpu;
    put(L::clk_16bit_limb, d.clk & 0xFFFF);
    put(L::clk_8bit_limb, (d.clk >> 16) & 0xFF);
    put(L::pc, d.pc);
    put(L::next_pc, d.next_pc);
    put(L::mp, d.mp);
    put(L::next_mp, d.next_mp);
    put(L::mv, d.mv);
    put(L::next_mv, d.next_mv);
    put(L::opcode, d.op);
    put_word(out, rows, t, L::op_a, args[d.pc]);
    const uint32_t diff = d.clk - d.pt;
    put(L::mv_access_prev_value, d.acc ? d.pv : 0);
    put(L::mv_access_value, d.mv);
    put(L::mv_access_prev_clk, d.acc ? d.pt : 0);
    put(L::mv_access_diff_16bit_limb, d.acc ? (diff & 0xFFFF) : 0);
    put(L::mv_access_diff_8bit_limb, d.acc ? ((diff >> 16) & 0xFF) : 0);
    put(L::next_mv_access_prev_value, d.nacc ? d.mv : 0);
    put(L::next_mv_access_value, d.next_mv);
    put(L::next_mv_access_prev_clk, d.nacc ? d.clk + 1 : 0);
    put(L::next_mv_access_diff_16bit_limb, 0);
    put(L::next_mv_access_diff_8bit_limb, 0);
    put(L::mv_accessed, d.acc);
    put(L::next_mv_accessed, d.nacc);
    const bool is_jump = d.op == OP_LOOP_START || d.op == OP_LOOP_END, is_mem = !d.acc, is_io = d.op == OP_INPUT || d.op == OP_OUTPUT;
    put(L::is_mv_immutable, d.nacc || is_jump || d.op == OP_OUTPUT);
    put(L::is_alu, d.nacc);
    put(L::is_jump, is_jump);
    put(L::is_io, is_io);
    put(L::is_memory_instr, is_mem);
    put(L::is_real, 1);
}

__global__ void __launch_bounds__(256) k_addsub_trace(const uint4* __restrict__ cyc, const uint8_t* __restrict__ ops, const uint32_t* __restrict__ idx,
                                                      uint32_t n_ev, unsigned log_rows, uint32_t* __restrict__ out) {
    TG_ROW_PROLOGUE(n_ev);
    const Dec d = decode(cyc, ops, idx[r]);
    namespace L = lay::addsub;
    const bool is_add = d.op == OP_ADD;
    const uint32_t op1 = is_add ? d.mv : d.next_mv;
    put(L::pc, d.pc);
    put(L::value, (op1 + 1) & 0xFF);
    put(L::carry, op1 + 1 > 255);
    put(L::operand_1, op1);
    put(L::operand_2, 1);
    put(L::is_add, is_add);
    put(L::is_sub, !is_add);
}

__global__ void __launch_bounds__(256) k_jump_trace(const uint4* __restrict__ cyc, const uint8_t* __restrict__ ops, const uint32_t* __restrict__ idx,
                                                    uint32_t n_ev, unsigned log_rows, const uint32_t* __restrict__ inv_table, uint32_t* __restrict__ out) {
    TG_ROW_PROLOGUE(n_ev);
    const Dec d = decode(cyc, ops, idx[r]);
    namespace L = lay::jump;
    put_word(out, rows, t, L::pc, d.pc);
    put_range_checker(out, rows, t, L::pc_msb_decomp, d.pc);
    put_word(out, rows, t, L::next_pc, d.next_pc);
    put_range_checker(out, rows, t, L::next_pc_msb_decomp, d.next_pc);
    put_word(out, rows, t, L::dst, d.next_pc);  // the event's dst is the taken next_pc (executor.rs:121-124)
    put(L::mv, d.mv);
    out[(uint64_t)L::is_mv_zero_inverse * rows + t] = inv_table[d.mv];  // already Montgomery
    put(L::is_mv_zero_result, d.mv == 0);
    put(L::is_loop_start, d.op == OP_LOOP_START);
    put(L::is_loop_end, d.op == OP_LOOP_END);
}

__global__ void __launch_bounds__(256) k_meminstr_trace(const uint4* __restrict__ cyc, const uint8_t* __restrict__ ops, const uint32_t* __restrict__ idx,
                                                        uint32_t n_ev, unsigned log_rows, uint32_t* __restrict__ out) {
    TG_ROW_PROLOGUE(n_ev);
    const Dec d = decode(cyc, ops, idx[r]);
    namespace L = lay::meminstr;
    put(L::pc, d.pc);
    put(L::clk, d.clk);
    put_word(out, rows, t, L::mp, d.mp);
    put_range_checker(out, rows, t, L::mp_msb_decomp, d.mp);
    put_word(out, rows, t, L::next_mp, d.next_mp);
    put_range_checker(out, rows, t, L::next_mp_msb_decomp, d.next_mp);
    put(L::is_step_forward, d.op == OP_MEM_FWD);
    put(L::is_step_backward, d.op == OP_MEM_BWD);
    put(L::is_real, 1);
}

__global__ void __launch_bounds__(256) k_io_trace(const uint4* __restrict__ cyc, const uint8_t* __restrict__ ops, const uint32_t* __restrict__ idx,
                                                  uint32_t n_ev, unsigned log_rows, uint32_t* __restrict__ out) {
    TG_ROW_PROLOGUE(n_ev);
    const Dec d = decode(cyc, ops, idx[r]);
    namespace L = lay::io;
    put(L::pc, d.pc);
    put(L::mp, d.mp);
    put(L::mv, d.mv);
    put(L::is_input, d.op == OP_INPUT);
    put(L::is_output, d.op == OP_OUTPUT);
}

// Memory chip: two touched cells per row (memory/memory.rs:84-126); ev = 5 words per cell
__global__ void __launch_bounds__(256) k_memory_trace(const uint32_t* __restrict__ ev, uint32_t n_cells, unsigned log_rows, uint32_t* __restrict__ out) {
    TG_ROW_PROLOGUE((n_cells + 1) / 2);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        // entry k of row r is cell 2r + k ... as laid out by the reference: even cells fill entry 0 of rows 0.., odd cells entry 1
        uint32_t c = 2 * r + k;
        if (c >= n_cells) break;
        const uint32_t* e = ev + 5 * (uint64_t)c;
        put(6 * k + 0, e[0]);  // addr
        put(6 * k + 1, e[1]);  // initial_clk
        put(6 * k + 2, e[3]);  // final_clk
        put(6 * k + 3, e[2]);  // initial_value
        put(6 * k + 4, e[4]);  // final_value
        put(6 * k + 5, 1);     // is_real
    }
}

__global__ void __launch_bounds__(256) k_program_trace(const uint32_t* __restrict__ counts, uint32_t n_instr, unsigned log_rows, uint32_t* __restrict__ out) {
    TG_ROW_PROLOGUE(n_instr);
    put(lay::program::multiplicity, counts[r]);
}

// Byte chip main trace: multiplicities (bytes/trace.rs:39-60): column U8_RANGE holds u8[row] for row < 256, U16_RANGE u16[row]
__global__ void __launch_bounds__(256) k_byte_trace(const unsigned int* __restrict__ u8, const unsigned int* __restrict__ u16, uint32_t* __restrict__ out) {
    const unsigned log_rows = 16;
    TG_ROW_PROLOGUE(1u << 16);
    put(lay::byte::multiplicities + U8_RANGE, r < 256 ? u8[r] : 0);
    put(lay::byte::multiplicities + U16_RANGE, u16[r]);
}

}  // namespace tg

// ---- host orchestration ----------------------------------------------------------------------------------------------------------
static uint64_t tg_pow2(uint64_t n, uint64_t minimum) {  // utils/mod.rs:25-53
    uint64_t p = 1;
    while (p < n) p <<= 1;
    return std::max(p, minimum);
}

// MachineProver::commit (prover.rs:209-236) fed by the execution record instead of host traces
static int32_t machine_commit_record_impl(bfgpu_ctx* ctx, const bfgpu_record* rec, uint32_t root[8], bfgpu_shard** out);
extern "C" int32_t bfgpu_machine_commit_record(bfgpu_ctx* ctx, const bfgpu_record* rec, uint32_t root[8], bfgpu_shard** out) {
    AllocScope scope(ctx);  // blocks taken by a failing call go back to the cache (see AllocScope)
    return scope.ok(machine_commit_record_impl(ctx, rec, root, out));
}
// Every included chip's main trace generated on the device from the execution record, in the prover's layout (column-major, Montgomery,
// bit-reversed rows), sorted by (height desc, name) as `MachineProver::commit` orders them (prover.rs:214).  The caller owns the blocks.
// d_cyc_in: the (cycles + 1) 16-byte cycle records already on the device (sharded prover: every rank uploads a slice and the ranks
// exchange them over NVLink), or null to copy them from the record's page-locked host buffer here.
static int32_t record_traces(bfgpu_ctx* ctx, const bfgpu_record* rec, std::vector<std::string>* out_names, std::vector<int>* out_chip,
                             std::vector<DMat>* out_traces, const uint4* d_cyc_in = nullptr) {
    const uint32_t n = (uint32_t)rec->n_cycles, n_instr = (uint32_t)rec->ops.size(), n_cells = (uint32_t)(rec->mem_events.size() / 5);
    if (n == 0) return fail(ctx, BFGPU_ERR_INVALID, "empty execution");
    // The Cpu trace is padded to a power of two with no minimum (utils/mod.rs:25-53): one cycle gives a ONE-row trace whose
    // LDE is as short as the FRI blow-up, and p3-fri's verifier never consumes a reduced opening of that height — the
    // reference would emit an unverifiable proof.  Refuse instead.
    if (n == 1) return fail(ctx, BFGPU_ERR_INVALID, "a one-cycle execution cannot be proven: its one-row Cpu trace is below the FRI verifier's minimum height");
    // ---- inputs to the device ----
    Scratch scratch(ctx);  // inputs and index lists: back to the block cache on every exit path
    uint4 *d_cyc = nullptr, *d_flags = nullptr;
    uint8_t* d_ops = nullptr;
    uint32_t *d_args = nullptr, *d_counts = nullptr, *d_mem = nullptr, *d_idx = nullptr, *d_inv = nullptr;
    unsigned int* d_hist = nullptr;
    {
        Phase ph(ctx, BFGPU_PHASE_H2D);
        if (d_cyc_in) d_cyc = const_cast<uint4*>(d_cyc_in);
        else TRY(scratch.alloc((void**)&d_cyc, (size_t)(n + 1) * 16));
        TRY(scratch.alloc((void**)&d_ops, n_instr));
        TRY(scratch.alloc((void**)&d_args, (size_t)n_instr * 4));
        TRY(scratch.alloc((void**)&d_counts, (size_t)n_instr * 4));
        TRY(scratch.alloc((void**)&d_mem, std::max<size_t>(rec->mem_events.size(), 1) * 4));
        if (!d_cyc_in) CU(cudaMemcpyAsync(d_cyc, rec->cycles, (size_t)(n + 1) * 16, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_ops, rec->ops.data(), n_instr, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_args, rec->args.data(), (size_t)n_instr * 4, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_counts, rec->prog_counts.data(), (size_t)n_instr * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (n_cells) CU(cudaMemcpyAsync(d_mem, rec->mem_events.data(), rec->mem_events.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    std::unique_ptr<Phase> ph_gen(new Phase(ctx, BFGPU_PHASE_TRACEGEN));
    // ---- event index lists ----
    const uint64_t ne[4] = {rec->n_alu, rec->n_jump, rec->n_mem, rec->n_io};
    uint64_t off[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < 4; k++) off[k + 1] = off[k] + ne[k];
    TRY(scratch.alloc((void**)&d_flags, (size_t)n * 16));
    TRY(scratch.alloc((void**)&d_idx, std::max<uint64_t>(off[4], 1) * 4));
    const unsigned gb = (n + 255) / 256;
    tg::k_classify<<<gb, 256, 0, ctx->stream>>>(d_cyc, d_ops, n, d_flags);
    LAUNCHED(ctx);
    TRY(scan_ext(ctx, (uint32_t*)d_flags, n));  // counts stay far below p, so the field addition is the integer one
    tg::k_index<<<gb, 256, 0, ctx->stream>>>(d_cyc, d_ops, d_flags, n, d_idx + off[0], d_idx + off[1], d_idx + off[2], d_idx + off[3]);
    LAUNCHED(ctx);
    // ---- byte multiplicities ----
    TRY(scratch.alloc((void**)&d_hist, (256 + 65536) * 4));
    CU(cudaMemsetAsync(d_hist, 0, (256 + 65536) * 4, ctx->stream));
    tg::k_byte_hist<<<std::min(gb, 148u * 8u), tg::HIST_THREADS, 0, ctx->stream>>>(d_cyc, d_ops, n, d_hist, d_hist + 256);
    LAUNCHED(ctx);
    // inverses of 1..255 for IsZeroOperation (operations/is_zero.rs:29-40), Montgomery form
    if (!ctx->d_inv256) {  // built once per context
        std::vector<uint32_t> inv(256, 0);
        for (uint32_t v = 1; v < 256; v++) inv[v] = kb::inv(kb::to_mont(v));
        CU(cudaMalloc(&ctx->d_inv256, 256 * 4));
        CU(cudaMemcpy(ctx->d_inv256, inv.data(), 256 * 4, cudaMemcpyHostToDevice));
    }
    d_inv = ctx->d_inv256;
    // ---- traces, straight into the prover's layout ----
    std::vector<std::string> names;
    std::vector<DMat> traces;
    auto new_trace = [&](const char* name, uint64_t rows, DMat* m) -> int32_t {
        int ci = chip_index(name);
        m->rows = rows;
        m->cols = (uint32_t)air::CHIPS[ci].main_w;
        m->rs = 1;
        TRY(dalloc(ctx, (void**)&m->d, rows * m->cols * 4));
        names.push_back(name);
        traces.push_back(*m);  // owned by `traces` from here on (released on the error path below)
        CU(cudaMemsetAsync(m->d, 0, rows * m->cols * 4, ctx->stream));
        return BFGPU_OK;
    };
    auto blocks = [](uint64_t rows) { return (unsigned)((rows + 255) / 256); };
    DMat m;
    int32_t rc = BFGPU_OK;
    do {
        if ((rc = new_trace("Cpu", tg_pow2(n, 1), &m)) != BFGPU_OK) break;
        tg::k_cpu_trace<<<blocks(m.rows), 256, 0, ctx->stream>>>(d_cyc, d_ops, d_args, n, ilog2(m.rows), m.d);
        LAUNCHED(ctx);
        if ((rc = new_trace("Program", tg_pow2(n_instr, 16), &m)) != BFGPU_OK) break;
        tg::k_program_trace<<<blocks(m.rows), 256, 0, ctx->stream>>>(d_counts, n_instr, ilog2(m.rows), m.d);
        LAUNCHED(ctx);
        if (ne[0]) {
            if ((rc = new_trace("AddSub", tg_pow2(ne[0], 16), &m)) != BFGPU_OK) break;
            tg::k_addsub_trace<<<blocks(m.rows), 256, 0, ctx->stream>>>(d_cyc, d_ops, d_idx + off[0], (uint32_t)ne[0], ilog2(m.rows), m.d);
            LAUNCHED(ctx);
        }
        if (ne[1]) {
            if ((rc = new_trace("Jump", tg_pow2(ne[1], 16), &m)) != BFGPU_OK) break;
            tg::k_jump_trace<<<blocks(m.rows), 256, 0, ctx->stream>>>(d_cyc, d_ops, d_idx + off[1], (uint32_t)ne[1], ilog2(m.rows), d_inv, m.d);
            LAUNCHED(ctx);
        }
        if (n_cells) {
            if ((rc = new_trace("Memory", tg_pow2((n_cells + 1) / 2, 16), &m)) != BFGPU_OK) break;
            tg::k_memory_trace<<<blocks(m.rows), 256, 0, ctx->stream>>>(d_mem, n_cells, ilog2(m.rows), m.d);
            LAUNCHED(ctx);
        }
        if (ne[2]) {
            if ((rc = new_trace("MemoryInstrs", tg_pow2(ne[2], 16), &m)) != BFGPU_OK) break;
            tg::k_meminstr_trace<<<blocks(m.rows), 256, 0, ctx->stream>>>(d_cyc, d_ops, d_idx + off[2], (uint32_t)ne[2], ilog2(m.rows), m.d);
            LAUNCHED(ctx);
        }
        if (ne[3]) {
            if ((rc = new_trace("IO", tg_pow2(ne[3], 16), &m)) != BFGPU_OK) break;
            tg::k_io_trace<<<blocks(m.rows), 256, 0, ctx->stream>>>(d_cyc, d_ops, d_idx + off[3], (uint32_t)ne[3], ilog2(m.rows), m.d);
            LAUNCHED(ctx);
        }
        if ((rc = new_trace("Byte", 1u << 16, &m)) != BFGPU_OK) break;
        tg::k_byte_trace<<<blocks(m.rows), 256, 0, ctx->stream>>>(d_hist, d_hist + 256, m.d);
        LAUNCHED(ctx);
        if (cudaGetLastError() != cudaSuccess) rc = fail(ctx, BFGPU_ERR_CUDA, "trace generation kernels failed to launch");
    } while (0);
    ph_gen.reset();  // the commit below is accounted under its own phases
    if (rc != BFGPU_OK) {
        for (DMat& t : traces) dfree(ctx, t.d);
        return rc;
    }
    const size_t nt = traces.size();
    std::vector<size_t> order(nt);
    for (size_t k = 0; k < nt; k++) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
        if (traces[a].rows != traces[b].rows) return traces[a].rows > traces[b].rows;
        return names[a] < names[b];
    });
    for (size_t k = 0; k < nt; k++) {
        out_names->push_back(names[order[k]]);
        out_chip->push_back(chip_index(names[order[k]].c_str()));
        out_traces->push_back(traces[order[k]]);
    }
    return BFGPU_OK;
}

static int32_t machine_commit_record_impl(bfgpu_ctx* ctx, const bfgpu_record* rec, uint32_t root[8], bfgpu_shard** out) {
    if (!ctx || !rec || !root || !out) return fail(ctx, BFGPU_ERR_INVALID, "null argument");
    *out = nullptr;
    // ---- commit: traces sorted by (height desc, name) (prover.rs:214), LDE a scratch copy of every trace, Merkle tree ----
    auto* sd = new bfgpu_shard();
    sd->ctx = ctx;
    int32_t rc = record_traces(ctx, rec, &sd->names, &sd->chip, &sd->traces);
    if (rc != BFGPU_OK) {
        delete sd;
        return rc;
    }
    std::vector<DMat>& traces = sd->traces;
    const size_t nt = traces.size();
    std::vector<DMat> coefs(nt);
    std::vector<uint32_t> shifts(nt, kb::to_mont(kb::GEN));
    for (size_t k = 0; k < nt && rc == BFGPU_OK; k++) {
        const DMat& t = traces[k];
        coefs[k] = t;
        coefs[k].d = nullptr;
        rc = dalloc(ctx, (void**)&coefs[k].d, t.rows * t.cols * 4);
        if (rc == BFGPU_OK && cudaMemcpyAsync(coefs[k].d, t.d, t.rows * t.cols * 4, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
            rc = fail(ctx, BFGPU_ERR_CUDA, "trace copy failed");
    }
    if (rc == BFGPU_OK) rc = commit_bitrev_device(ctx, coefs, shifts, &sd->data, sd->commit);
    if (rc != BFGPU_OK) {
        for (DMat& c : coefs) dfree(ctx, c.d);
        for (DMat& t : traces) dfree(ctx, t.d);
        delete sd;
        return rc;
    }
    for (int i = 0; i < 8; i++) root[i] = out_word(ctx, sd->commit[i]);
    *out = sd;
    return BFGPU_OK;
}

// StarkMachine::setup (machine.rs:154-224) from the compiled program: the preprocessed Program and Byte traces
// (program/mod.rs:64-100, bytes/mod.rs:31-62) are tiny and built on the host.
extern "C" int32_t bfgpu_machine_setup_record(bfgpu_ctx* ctx, const bfgpu_record* rec, uint32_t commit[8], bfgpu_pk** out) {
    if (!ctx || !rec || !commit || !out) return fail(ctx, BFGPU_ERR_INVALID, "null argument");
    const uint32_t n_instr = (uint32_t)rec->ops.size();
    const uint64_t prows = tg_pow2(n_instr, 16);
    std::vector<uint32_t> prog(prows * 6, 0), byte((size_t)65536 * 2);
    for (uint32_t k = 0; k < n_instr; k++) {
        prog[6 * k + 0] = k;
        prog[6 * k + 1] = rec->ops[k];
        for (int b = 0; b < 4; b++) prog[6 * k + 2 + b] = (rec->args[k] >> (8 * b)) & 0xFF;
    }
    for (uint32_t k = 0; k < 65536; k++) {
        byte[2 * k + 0] = k & 0xFF;
        byte[2 * k + 1] = k;
    }
    const char* names[2] = {"Program", "Byte"};
    bfgpu_mat mats[2] = {{prog.data(), prows, 6}, {byte.data(), 65536, 2}};
    const int repr = ctx->repr, space = ctx->input_space;
    ctx->repr = BFGPU_REPR_CANONICAL;  // the matrices above are plain residues in host memory
    ctx->input_space = BFGPU_MEM_HOST;
    int32_t rc = bfgpu_machine_setup(ctx, names, mats, 2, commit, out);
    ctx->repr = repr;
    ctx->input_space = space;
    if (rc == BFGPU_OK && repr != BFGPU_REPR_CANONICAL)
        for (int i = 0; i < 8; i++) commit[i] = (*out)->commit[i];
    return rc;
}

// test / debug hooks: the traces held by a shard (commit order), copied out row-major in natural row order
extern "C" int32_t bfgpu_shard_num_traces(const bfgpu_shard* sd) { return sd ? (int32_t)sd->traces.size() : 0; }
extern "C" int32_t bfgpu_shard_trace_info(const bfgpu_shard* sd, int32_t i, const char** name, uint64_t* rows, uint64_t* cols) {
    if (!sd || i < 0 || (size_t)i >= sd->traces.size()) return BFGPU_ERR_INVALID;
    if (name) *name = sd->names[i].c_str();
    if (rows) *rows = sd->traces[i].rows;
    if (cols) *cols = sd->traces[i].cols;
    return BFGPU_OK;
}
static int32_t shard_get_trace_impl(const bfgpu_shard* sd, int32_t i, uint32_t* out);
extern "C" int32_t bfgpu_shard_get_trace(const bfgpu_shard* sd, int32_t i, uint32_t* out) {
    AllocScope scope(sd ? sd->ctx : nullptr);  // blocks taken by a failing call go back to the cache (see AllocScope)
    return scope.ok(shard_get_trace_impl(sd, i, out));
}
static int32_t shard_get_trace_impl(const bfgpu_shard* sd, int32_t i, uint32_t* out) {
    if (!sd || i < 0 || (size_t)i >= sd->traces.size() || !out) return BFGPU_ERR_INVALID;
    return egress(sd->ctx, sd->traces[i], /*bitrev=*/true, out);  // stored bit-reversed: undo for natural order
}
