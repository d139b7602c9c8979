"""TEST INFRASTRUCTURE (part of the CPU oracle; PARITY UNPINNED like the rest of it): trace generation for the eight
chips (numpy, vectorised over events).  The product generates the traces on the device (csrc/tracegen.cuh); this is the
restatement it is tested against.  Restates the reference's `MachineAir::generate_trace` /
`generate_preprocessed_trace` / `generate_dependencies` of every chip:
  Cpu cpu/trace.rs:28-150 (+ memory/consistency/trace.rs:9-77), Program program/mod.rs:64-137,
  AddSub alu/mod.rs:62-157 (+ operations/add.rs:21-40), Jump jump/trace.rs:31-96
  (+ operations/is_zero.rs:29-40, koala_bear_word.rs:37-50), Memory memory/memory.rs:84-126,
  Byte bytes/{mod.rs:31-62,trace.rs:39-60}, MemoryInstrs memory/instructions/trace.rs:36-96, IO io/mod.rs:72-124;
padding rules crates/core/machine/src/utils/mod.rs:25-53 (Cpu pads to a power of two with NO minimum,
everything else to >= 16 rows).  All values are canonical residues (uint32).
"""
import numpy as np

import importlib

C = importlib.import_module("zkvm-brainfuck_b200.air.chips")  # the declarative AIR (layouts, opcodes) is shared with the code generator

P = 2130706433


def _pow2(n, minimum=16):
    p = 1 if n <= 1 else 1 << (n - 1).bit_length()
    return max(p, minimum)


def _word(v):
    v = np.asarray(v, np.int64)
    return np.stack([(v >> (8 * i)) & 0xFF for i in range(4)], axis=1)


def _range_checker(v):
    """KoalaBearWordRangeChecker::populate: 8 bits of the top byte + the running products."""
    v = np.asarray(v, np.int64)
    bits = np.stack([(v >> (24 + i)) & 1 for i in range(8)], axis=1)
    ands = [bits[:, 0] * bits[:, 1]]
    for i in range(2, 7):
        ands.append(ands[-1] * bits[:, i])
    return np.concatenate([bits, np.stack(ands, axis=1)], axis=1)


def _inv_mod(v):
    v = np.asarray(v, np.int64)
    out = np.zeros_like(v)
    for x in np.unique(v):
        if x:
            out[v == x] = pow(int(x), P - 2, P)
    return out


def generate_traces(rec):
    """-> (main traces {name: (rows, width) uint32}, byte lookup multiplicities) for the included chips."""
    prog = rec.program
    u8 = np.zeros(256, np.int64)
    u16 = np.zeros(65536, np.int64)

    def add_u8(v):
        np.add.at(u8, np.asarray(v, np.int64) & 0xFF, 1)

    def add_u16(v):
        np.add.at(u16, np.asarray(v, np.int64) & 0xFFFF, 1)

    traces = {}
    # ---- Cpu ------------------------------------------------------------------------------------------
    ev = rec.cpu
    n = ev.shape[0]
    L = C.CPU_LAYOUT
    t = np.zeros((_pow2(n, 1), C.CPU_WIDTH), np.int64)
    clk, pc = ev[:, 0], ev[:, 1]
    t[:n, L["clk_16bit_limb"]] = clk & 0xFFFF
    t[:n, L["clk_8bit_limb"]] = (clk >> 16) & 0xFF
    add_u16(clk & 0xFFFF); add_u8((clk >> 16) & 0xFF)
    for name, col in (("pc", 1), ("next_pc", 2), ("mp", 3), ("next_mp", 4), ("mv", 5), ("next_mv", 6)):
        t[:n, L[name]] = ev[:, col]
    opc = prog.opcodes[pc].astype(np.int64)
    t[:n, L["opcode"]] = opc
    t[:n, L["op_a"][0]:L["op_a"][0] + 4] = _word(prog.op_a[pc])
    t[:n, L["mv_access_value"]] = ev[:, 5]        # *mv_access.value_mut() = mv (cpu/trace.rs:108-109)
    t[:n, L["next_mv_access_value"]] = ev[:, 6]
    for pref, base in (("mv_access", 7), ("next_mv_access", 12)):
        acc = ev[:, base] == 1
        prev_value, value, prev_ts, ts = ev[:, base + 1], ev[:, base + 2], ev[:, base + 3], ev[:, base + 4]
        diff = ts - prev_ts - 1
        t[:n, L[pref + "_prev_value"]] = np.where(acc, prev_value, 0)
        t[:n, L[pref + "_value"]] = np.where(acc, value, t[:n, L[pref + "_value"]])
        t[:n, L[pref + "_prev_clk"]] = np.where(acc, prev_ts, 0)
        t[:n, L[pref + "_diff_16bit_limb"]] = np.where(acc, diff & 0xFFFF, 0)
        t[:n, L[pref + "_diff_8bit_limb"]] = np.where(acc, (diff >> 16) & 0xFF, 0)
        add_u16((diff & 0xFFFF)[acc]); add_u8(((diff >> 16) & 0xFF)[acc])
        t[:n, L["mv_accessed" if pref == "mv_access" else "next_mv_accessed"]] = acc
    add_u8(ev[:, 5])
    is_alu = (opc == C.ADD) | (opc == C.SUB)
    is_jump = (opc == C.LOOP_START) | (opc == C.LOOP_END)
    is_mem = (opc == C.MEM_FWD) | (opc == C.MEM_BWD)
    is_io = (opc == C.INPUT) | (opc == C.OUTPUT)
    t[:n, L["is_mv_immutable"]] = is_alu | is_jump | (opc == C.OUTPUT)
    t[:n, L["is_alu"]], t[:n, L["is_jump"]], t[:n, L["is_memory_instr"]], t[:n, L["is_io"]] = is_alu, is_jump, is_mem, is_io
    t[:n, L["is_real"]] = 1
    traces["Cpu"] = t
    # ---- Program (main = multiplicities; always included) -----------------------------------------------
    t = np.zeros((_pow2(len(prog)), 1), np.int64)
    np.add.at(t[:, 0], pc, 1)
    traces["Program"] = t
    # ---- AddSub -------------------------------------------------------------------------------------------
    ev = rec.alu
    if ev.shape[0]:
        n = ev.shape[0]
        L = C.ADDSUB_LAYOUT
        t = np.zeros((_pow2(n), C.ADDSUB_WIDTH), np.int64)
        is_add = ev[:, 1] == C.ADD
        op1 = np.where(is_add, ev[:, 3], ev[:, 2])  # mv for add, next_mv for sub
        value = (op1 + 1) & 0xFF
        t[:n, L["pc"]], t[:n, L["value"]], t[:n, L["carry"]] = ev[:, 0], value, (op1 + 1) > 255
        t[:n, L["operand_1"]], t[:n, L["operand_2"]] = op1, 1
        t[:n, L["is_add"]], t[:n, L["is_sub"]] = is_add, ~is_add
        add_u8(op1); add_u8(np.ones(n, np.int64)); add_u8(value)
        traces["AddSub"] = t
    # ---- Jump -------------------------------------------------------------------------------------------------
    ev = rec.jump
    if ev.shape[0]:
        n = ev.shape[0]
        L = C.JUMP_LAYOUT
        t = np.zeros((_pow2(n), C.JUMP_WIDTH), np.int64)
        t[:n, L["pc"][0]:L["pc"][0] + 4] = _word(ev[:, 0])
        t[:n, L["pc_msb_decomp"][0]:L["pc_and_0_to_7"] + 1] = _range_checker(ev[:, 0])
        t[:n, L["next_pc"][0]:L["next_pc"][0] + 4] = _word(ev[:, 1])
        t[:n, L["next_pc_msb_decomp"][0]:L["next_pc_and_0_to_7"] + 1] = _range_checker(ev[:, 1])
        t[:n, L["dst"][0]:L["dst"][0] + 4] = _word(ev[:, 3])
        t[:n, L["mv"]] = ev[:, 4]
        t[:n, L["is_mv_zero_inverse"]] = _inv_mod(ev[:, 4])
        t[:n, L["is_mv_zero_result"]] = ev[:, 4] == 0
        t[:n, L["is_loop_start"]], t[:n, L["is_loop_end"]] = ev[:, 2] == C.LOOP_START, ev[:, 2] == C.LOOP_END
        traces["Jump"] = t
    # ---- Memory (two entries per row) ---------------------------------------------------------------------------
    ev = rec.memory
    if ev.shape[0]:
        n = ev.shape[0]
        rows = -(-n // 2)
        t = np.zeros((_pow2(rows), C.MEMORY_WIDTH), np.int64)
        for k in range(2):
            e = ev[k::2]
            m = e.shape[0]
            t[:m, 6 * k + 0], t[:m, 6 * k + 1], t[:m, 6 * k + 2] = e[:, 0], e[:, 1], e[:, 3]  # addr, initial_clk, final_clk
            t[:m, 6 * k + 3], t[:m, 6 * k + 4], t[:m, 6 * k + 5] = e[:, 2], e[:, 4], 1         # initial_value, final_value, is_real
        traces["Memory"] = t
    # ---- MemoryInstrs ----------------------------------------------------------------------------------------------
    ev = rec.mem_instr
    if ev.shape[0]:
        n = ev.shape[0]
        L = C.MEMINSTR_LAYOUT
        t = np.zeros((_pow2(n), C.MEMINSTR_WIDTH), np.int64)
        t[:n, L["pc"]], t[:n, L["clk"]] = ev[:, 1], ev[:, 0]
        t[:n, L["mp"][0]:L["mp"][0] + 4] = _word(ev[:, 3])
        t[:n, L["mp_msb_decomp"][0]:L["mp_and_0_to_7"] + 1] = _range_checker(ev[:, 3])
        t[:n, L["next_mp"][0]:L["next_mp"][0] + 4] = _word(ev[:, 4])
        t[:n, L["next_mp_msb_decomp"][0]:L["next_mp_and_0_to_7"] + 1] = _range_checker(ev[:, 4])
        t[:n, L["is_step_forward"]], t[:n, L["is_step_backward"]] = ev[:, 2] == C.MEM_FWD, ev[:, 2] == C.MEM_BWD
        t[:n, L["is_real"]] = 1
        traces["MemoryInstrs"] = t
    # ---- IO ------------------------------------------------------------------------------------------------------------
    ev = rec.io
    if ev.shape[0]:
        n = ev.shape[0]
        t = np.zeros((_pow2(n), C.IO_WIDTH), np.int64)
        t[:n, 0], t[:n, 1], t[:n, 2] = ev[:, 0], ev[:, 2], ev[:, 3]
        t[:n, 3], t[:n, 4] = ev[:, 1] == C.INPUT, ev[:, 1] == C.OUTPUT
        traces["IO"] = t
    # ---- Byte (multiplicities; always included) ---------------------------------------------------------------------------
    t = np.zeros((1 << 16, 2), np.int64)
    t[:256, C.U8_RANGE] = u8
    t[:, C.U16_RANGE] = u16
    traces["Byte"] = t
    return {k: np.ascontiguousarray(v % P, np.uint32) for k, v in traces.items()}


def preprocessed_traces(program):
    """StarkMachine::setup inputs (machine.rs:154-196): Program (pc, opcode, op_a word) and the Byte table."""
    n = len(program)
    t = np.zeros((_pow2(n), C.PROGRAM_PREP_WIDTH), np.int64)
    t[:n, 0] = np.arange(n)
    t[:n, 1] = program.opcodes
    t[:n, 2:6] = _word(program.op_a)
    b = np.zeros((1 << 16, 2), np.int64)
    b[:, 0] = np.arange(1 << 16) & 0xFF   # value_u8 = c for row b*256 + c
    b[:, 1] = np.arange(1 << 16)          # value_u16
    return {"Program": t.astype(np.uint32), "Byte": b.astype(np.uint32)}
