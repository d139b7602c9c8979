"""TEST INFRASTRUCTURE (part of the CPU oracle; PARITY UNPINNED like the rest of it): Brainfuck compiler + interpreter
emitting the event streams the chips' traces are built from.  The product's executor is the native one behind
`bfgpu_execute` (csrc/tracegen.cuh); this Python restatement is what it is tested against and what feeds the numpy prover.
  Restates the reference's
`Program::from` (crates/core/executor/src/program.rs:22-44) and `Executor::{run, execute_instruction,
emit_events, rr_traced, rw_traced}` (crates/core/executor/src/executor.rs:71-79,106-325), including its quirks:
`+` and `-` both land in `add_events` (:219-226), `,` never advances the input pointer (:183), a jump event's
`dst` is the taken `next_pc` (:121-124).  The reference drains `memory_events` from a randomly seeded hash map
(:74-76), so its Memory-chip row order differs between processes; here rows are in first-access order.
"""
import numpy as np

LOOP_START, LOOP_END, ADD, SUB, MEM_FWD, MEM_BWD, INPUT, OUTPUT = range(8)
_DECODE = {">": MEM_FWD, "<": MEM_BWD, "+": ADD, "-": SUB, ".": OUTPUT, ",": INPUT, "[": LOOP_START, "]": LOOP_END}


class Program:
    def __init__(self, code):
        ops, args, stack = [], [], []
        for c in code:
            if c == "[":
                ops.append(LOOP_START); args.append(0); stack.append(len(ops) - 1)
            elif c == "]":
                start = stack.pop()
                args[start] = len(ops)
                ops.append(LOOP_END); args.append(start + 1)
            elif c not in " \n\r":
                ops.append(_DECODE[c]); args.append(0)
        self.opcodes = np.array(ops, np.uint32)
        self.op_a = np.array(args, np.uint32)

    def __len__(self):
        return len(self.opcodes)


class ExecutionRecord:
    """Event streams as numpy arrays (one row per event)."""

    def __init__(self, program):
        self.program = program
        self.output = []


def execute(program, stdin=()):
    """Run to completion. Returns an ExecutionRecord with
       cpu:  [clk, pc, next_pc, mp, next_mp, mv, next_mv, mv_acc(0/1), mv_prev_value, mv_value, mv_prev_ts, mv_ts,
              nmv_acc(0/1), nmv_prev_value, nmv_value, nmv_prev_ts, nmv_ts]
       alu:  [pc, opcode, next_mv, mv]        jump: [pc, next_pc, opcode, dst, mv]
       mem_instr: [clk, pc, opcode, mp, next_mp]   io: [pc, opcode, mp, mv]
       memory: [addr, initial_ts, initial_value, final_ts, final_value]"""
    ops, args = program.opcodes.tolist(), program.op_a.tolist()
    n = len(ops)
    mem_val, mem_ts = {}, {}
    first = {}  # addr -> (initial value, initial ts), insertion ordered
    cpu, alu, jump, meminstr, io = [], [], [], [], []
    out = []
    pc = mp = clk = 0
    stdin = list(stdin)
    M32 = 0xFFFFFFFF
    while pc != n:
        op = ops[pc]
        next_pc = (pc + 1) & M32
        mv = next_mv = 0
        mv_acc = nmv_acc = 0
        a0 = a1 = a2 = a3 = b0 = b1 = b2 = b3 = 0
        cur_mp = mp
        if op == MEM_FWD or op == MEM_BWD:
            mp = (mp + 1) & M32 if op == MEM_FWD else (mp - 1) & M32
            meminstr.append((clk, pc, op, cur_mp, mp))
        else:
            pv, pt = mem_val.get(mp, 0), mem_ts.get(mp, 0)
            if mp not in first:
                first[mp] = (pv, pt)
            if op == INPUT:  # write at clk+1 recorded in mv_access
                mv = stdin[0]
                mem_val[mp], mem_ts[mp] = mv, clk + 1
                mv_acc, a0, a1, a2, a3 = 1, pv, mv, pt, clk + 1
                io.append((pc, op, cur_mp, mv))
            else:  # read at clk+1
                mv = pv
                mem_val[mp], mem_ts[mp] = pv, clk + 1
                mv_acc, a0, a1, a2, a3 = 1, pv, pv, pt, clk + 1
                if op == ADD or op == SUB:
                    next_mv = (mv + 1) & 0xFF if op == ADD else (mv - 1) & 0xFF
                    mem_val[mp], mem_ts[mp] = next_mv, clk + 2
                    nmv_acc, b0, b1, b2, b3 = 1, mv, next_mv, clk + 1, clk + 2
                    alu.append((pc, op, next_mv, mv))
                elif op == LOOP_START or op == LOOP_END:
                    if (op == LOOP_START and mv == 0) or (op == LOOP_END and mv != 0):
                        next_pc = args[pc]
                    jump.append((pc, next_pc, op, next_pc, mv))
                else:  # OUTPUT
                    out.append(mv)
                    io.append((pc, op, cur_mp, mv))
        cpu.append((clk, pc, next_pc, cur_mp, mp, mv, next_mv, mv_acc, a0, a1, a2, a3, nmv_acc, b0, b1, b2, b3))
        pc = next_pc
        clk += 2
    rec = ExecutionRecord(program)
    arr = lambda rows, w: np.array(rows, np.int64).reshape(-1, w)
    rec.cpu, rec.alu, rec.jump, rec.mem_instr, rec.io = arr(cpu, 17), arr(alu, 4), arr(jump, 5), arr(meminstr, 5), arr(io, 4)
    rec.memory = arr([(a, its, iv, mem_ts[a], mem_val[a]) for a, (iv, its) in first.items()], 5)
    rec.output = out
    rec.cycles = len(cpu)
    return rec
