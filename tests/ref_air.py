"""Hand-written per-row evaluators of all eight chips' AIRs, transcribed DIRECTLY from the reference's Rust
(no import of air/dsl.py or air/chips.py): test infrastructure that breaks the "one declarative description feeds the
CUDA codegen, the oracle prover and both verifiers" blind spot.  tests/test_air_independent.py
cross-checks air/chips.py against these on valid, corrupted and random rows.

Sources (reference, /root/reference/crates):
  core/machine/src/cpu/air.rs:28-185, cpu/cols.rs:29-78            CpuChip::eval and the CpuCols layout
  core/machine/src/air/memory.rs:17-125                             eval_memory_access / _timestamp / eval_range_check_24bits
  core/machine/src/air/program.rs:17-30, air/u8_air.rs:8-17          send_program (opcode appears twice), range_check_u8
  core/machine/src/jump/air.rs:22-82, jump/cols.rs:11-31            JumpChip::eval and JumpCols
  core/machine/src/operations/is_zero.rs:45-62, koala_bear_word.rs:50-108   IsZeroOperation::eval, KoalaBearWordRangeChecker
  core/machine/src/memory/consistency/cols.rs:5-45                  MemoryReadWriteCols / MemoryWriteCols / MemoryAccessCols
  core/machine/src/alu/mod.rs:23-38,157-192, operations/add.rs:41-75   AddSubChip::eval, AddSubCols, AddOperation::eval
  core/machine/src/memory/instructions/air.rs:26-76, cols.rs:13-35  MemoryInstructionsChip::eval and its columns
  core/machine/src/memory/memory.rs:22-44,131-145                   MemoryChip::eval (two entries per row)
  core/machine/src/io/mod.rs:19-31,127-141                          IoChip::eval
  core/machine/src/program/mod.rs:31-41,152-163, cpu/cols.rs:18-24  ProgramChip::eval (preprocessed pc + InstructionCols, main multiplicity)
  core/machine/src/bytes/air.rs:22-44, bytes/cols.rs:13-30, executor/src/events/byte.rs:109-113   ByteChip::eval
  stark/src/air/builder.rs:31-33,52-229                             when_not = when_ne(c, 1); lookup tuples
  stark/src/word.rs:67-70                                           Word::reduce
  stark/src/lookup/lookup.rs:19-40, core/executor/src/opcode.rs:12-42   LookupKind, Opcode, ByteOpcode numbering
Semantics of the filtered builder (p3-air FilteredAirBuilder): `when(c).assert_zero(x)` records c * x, `assert_eq(x, y)`
records x - y, `assert_bool(x)` records x * (x - 1), `assert_one(x)` records x - 1; conditions do not apply to lookups.
Every function works on Python ints mod p; rows are sequences of canonical residues."""
P = 2130706433
MEMORY, PROGRAM, ALU, JUMP, MEMINSTR, IO, BYTE = 1, 2, 3, 4, 5, 6, 7   # LookupKind
U8RANGE, U16RANGE = 0, 1                                                # ByteOpcode
LOOP_START, LOOP_END, ADD, SUB, MEM_STEP_FORWARD, MEM_STEP_BACKWARD, INPUT, OUTPUT = range(8)   # Opcode


class _Rec:
    """Collects constraints (call order = alpha-fold order) and lookups like the reference's builders."""

    def __init__(self):
        self.constraints, self.sends, self.receives = [], [], []

    def zero(self, x, *conds):
        v = x % P
        for c in conds:
            v = v * (c % P) % P
        self.constraints.append(v)

    def eq(self, x, y, *conds):
        self.zero(x - y, *conds)

    def boolean(self, x, *conds):
        self.zero(x * (x - 1), *conds)

    def send(self, kind, values, mult):
        self.sends.append((kind, tuple(v % P for v in values), mult % P))

    def receive(self, kind, values, mult):
        self.receives.append((kind, tuple(v % P for v in values), mult % P))


def _not(c):
    return c - 1  # when_not(c) == when_ne(c, ONE): the recorded factor is (c - 1)


# ---- CpuCols (cpu/cols.rs:29-78): #[repr(C)] field order ------------------------------------------------------------
class CpuRow:
    def __init__(self, r):
        (self.clk16, self.clk8, self.pc, self.next_pc, self.mp, self.next_mp, self.mv, self.next_mv, self.opcode) = [int(x) for x in r[0:9]]
        self.op_a = [int(x) for x in r[9:13]]
        # MemoryReadWriteCols { prev_value, access { value, prev_clk, diff_16bit_limb, diff_8bit_limb } }
        self.mva_prev_value, self.mva_value, self.mva_prev_clk, self.mva_diff16, self.mva_diff8 = [int(x) for x in r[13:18]]
        # MemoryWriteCols { prev_value, access { ... } }
        self.nva_prev_value, self.nva_value, self.nva_prev_clk, self.nva_diff16, self.nva_diff8 = [int(x) for x in r[18:23]]
        (self.mv_accessed, self.next_mv_accessed, self.is_mv_immutable, self.is_alu, self.is_jump, self.is_io, self.is_memory_instr,
         self.is_real) = [int(x) for x in r[23:31]]
        assert len(r) == 31


def _range_check_24bits(b, value, limb16, limb8, do_check):          # air/memory.rs:97-124
    b.eq(value, limb16 + limb8 * (1 << 16), do_check)
    b.send(BYTE, [U16RANGE, 0, limb16], do_check)
    b.send(BYTE, [U8RANGE, limb8, 0], do_check)


def _memory_access(b, clk, addr, prev_value, value, prev_clk, diff16, diff8, do_check):   # air/memory.rs:17-56
    b.boolean(do_check)
    _range_check_24bits(b, clk - prev_clk - 1, diff16, diff8, do_check)                 # eval_memory_access_timestamp :63-90
    b.send(MEMORY, [prev_clk, addr, prev_value], do_check)
    b.receive(MEMORY, [clk, addr, value], do_check)


def cpu_eval(local, nxt, is_first_row, is_last_row, is_transition):
    """CpuChip::eval (cpu/air.rs:28-64) on one (local, next) row pair."""
    l, n, b = CpuRow(local), CpuRow(nxt), _Rec()
    clk = (1 << 16) * l.clk8 + l.clk16
    b.send(PROGRAM, [l.pc, l.opcode, l.opcode] + l.op_a, l.is_real)                       # air/program.rs:17-30: opcode, then (opcode, op_a)
    # eval_instruction (cpu/air.rs:67-98)
    b.send(ALU, [l.pc, l.opcode, l.next_mv, l.mv], l.is_alu)
    b.send(JUMP, [l.pc, l.next_pc, l.opcode, l.mv], l.is_jump)
    b.send(MEMINSTR, [clk, l.pc, l.opcode, l.mp, l.next_mp], l.is_memory_instr)
    b.send(IO, [l.pc, l.opcode, l.mp, l.mv], l.is_io)
    # eval_registers (cpu/air.rs:160-185)
    _memory_access(b, clk + 1, l.mp, l.mva_prev_value, l.mva_value, l.mva_prev_clk, l.mva_diff16, l.mva_diff8, l.mv_accessed)
    _memory_access(b, clk + 2, l.mp, l.nva_prev_value, l.nva_value, l.nva_prev_clk, l.nva_diff16, l.nva_diff8, l.next_mv_accessed)
    b.send(BYTE, [U8RANGE, l.mv, 0], l.is_real)                                          # range_check_u8
    b.eq(l.mva_value, l.mva_prev_value, l.is_mv_immutable)
    # eval_clk (cpu/air.rs:106-130)
    b.zero(clk, is_first_row)
    next_clk = (1 << 16) * n.clk8 + n.clk16
    b.eq(clk + 2, next_clk, is_transition, n.is_real)
    _range_check_24bits(b, clk, l.clk16, l.clk8, l.is_real)
    # eval_pc (cpu/air.rs:133-146)
    b.eq(l.next_pc, n.pc, is_transition, n.is_real)
    b.eq(l.next_pc, l.pc + 1, is_transition, l.is_real, _not(l.is_jump))
    # eval_is_real (cpu/air.rs:152-157)
    b.boolean(l.is_real)
    b.eq(l.is_real, 1, is_first_row)
    b.zero(n.is_real, is_transition, _not(l.is_real))
    for flag in (l.is_alu, l.is_jump, l.is_memory_instr, l.is_io, l.is_mv_immutable, l.mv_accessed, l.next_mv_accessed):   # cpu/air.rs:57-63
        b.boolean(flag)
    return b


# ---- JumpCols (jump/cols.rs:11-31) ------------------------------------------------------------------------------------------
class JumpRow:
    def __init__(self, r):
        r = [int(x) for x in r]
        assert len(r) == 45
        self.pc, self.pc_rc = r[0:4], r[4:18]                # Word, KoalaBearWordRangeChecker (8 bits + 6 running products)
        self.next_pc, self.next_pc_rc = r[18:22], r[22:36]
        self.dst, self.mv = r[36:40], r[40]
        self.inverse, self.result = r[41], r[42]              # IsZeroOperation { inverse, result }
        self.is_loop_start, self.is_loop_end = r[43], r[44]


def _reduce(word):                                           # Word::reduce (stark/src/word.rs:67-70)
    return sum(x << (8 * i) for i, x in enumerate(word))


def _word_range_check(b, value, cols, is_real):              # operations/koala_bear_word.rs:50-108
    bits, ands = cols[0:8], cols[8:14]
    recomposed = 0
    for i, bit in enumerate(bits):
        b.boolean(bit, is_real)
        recomposed += (1 << i) * bit
    b.eq(recomposed, value[3], is_real)
    b.zero(bits[7], is_real)
    b.eq(ands[0], bits[0] * bits[1], is_real)
    for k in range(1, 6):
        b.eq(ands[k], ands[k - 1] * bits[k + 1], is_real)
    b.zero(value[0] + value[1] + value[2], is_real, ands[5])


def jump_eval(local, nxt=None, is_first_row=0, is_last_row=0, is_transition=0):
    """JumpChip::eval (jump/air.rs:22-82); the chip reads the local row only."""
    l, b = JumpRow(local), _Rec()
    is_real = l.is_loop_start + l.is_loop_end
    b.boolean(l.is_loop_start)
    b.boolean(l.is_loop_end)
    b.boolean(is_real)
    # IsZeroOperation::eval (operations/is_zero.rs:45-62)
    b.eq(1 - l.inverse * l.mv, l.result, is_real)
    b.boolean(l.result, is_real)
    b.zero(l.mv, is_real, l.result)
    npc, pc, dst = _reduce(l.next_pc), _reduce(l.pc), _reduce(l.dst)
    b.eq(npc, dst, l.is_loop_start, l.result)                       # '[' jumps when mv == 0
    b.eq(npc, pc + 1, l.is_loop_start, _not(l.result))              # '[' falls through otherwise
    b.eq(npc, dst, l.is_loop_end, _not(l.result))                   # ']' jumps when mv != 0
    b.eq(npc, pc + 1, l.is_loop_end, l.result)
    _word_range_check(b, l.pc, l.pc_rc, is_real)
    _word_range_check(b, l.next_pc, l.next_pc_rc, is_real)
    opcode = l.is_loop_start * LOOP_START + l.is_loop_end * LOOP_END
    b.receive(JUMP, [pc, npc, opcode, l.mv], is_real)
    return b


# ---- AddSubCols (alu/mod.rs:23-38): pc, AddOperation { value, carry }, operand_1, operand_2, is_add, is_sub ---------------------
def addsub_eval(local, nxt=None, is_first_row=0, is_last_row=0, is_transition=0, prep=None):
    """AddSubChip::eval (alu/mod.rs:157-192) with AddOperation::eval (operations/add.rs:41-75)."""
    pc, value, carry, operand_1, operand_2, is_add, is_sub = [int(x) for x in local]
    b = _Rec()
    is_real = is_add + is_sub
    b.boolean(is_add)
    b.boolean(is_sub)
    b.boolean(is_real)
    overflow = operand_1 + operand_2 - value
    b.zero(overflow * (overflow - 256), is_real)          # the carried and the plain result differ by zero or the base
    b.zero(carry * (overflow - 256), is_real)
    b.zero((carry - 1) * overflow, is_real)
    b.boolean(carry, is_real)
    b.boolean(is_real, is_real)
    for v in (operand_1, operand_2, value):               # range_check_u8 (air/u8_air.rs:8-17)
        b.send(BYTE, [U8RANGE, v, 0], is_real)
    b.receive(ALU, [pc, ADD, value, operand_1], is_add)   # '+': next_mv = value, mv = operand_1
    b.receive(ALU, [pc, SUB, operand_1, value], is_sub)   # '-': the addition runs backwards (next_mv + 1 = mv)
    return b


# ---- MemoryInstructionsCols (memory/instructions/cols.rs:13-35) -----------------------------------------------------------------
class MemInstrRow:
    def __init__(self, r):
        r = [int(x) for x in r]
        assert len(r) == 41
        self.pc, self.clk = r[0], r[1]
        self.mp, self.mp_rc = r[2:6], r[6:20]
        self.next_mp, self.next_mp_rc = r[20:24], r[24:38]
        self.is_step_forward, self.is_step_backward, self.is_real = r[38], r[39], r[40]


def meminstr_eval(local, nxt, is_first_row=0, is_last_row=0, is_transition=0, prep=None):
    """MemoryInstructionsChip::eval (memory/instructions/air.rs:26-76)."""
    l, n, b = MemInstrRow(local), MemInstrRow(nxt), _Rec()
    is_real = l.is_step_forward + l.is_step_backward
    b.boolean(l.is_step_forward)
    b.boolean(l.is_step_backward)
    b.boolean(is_real)
    mp, next_mp = _reduce(l.mp), _reduce(l.next_mp)
    b.eq(next_mp, mp + 1, l.is_step_forward)
    b.eq(next_mp, mp - 1, l.is_step_backward)
    b.eq(next_mp, _reduce(n.mp), is_transition, n.is_real)
    _word_range_check(b, l.mp, l.mp_rc, l.is_real)           # the range checks hang on the is_real COLUMN,
    _word_range_check(b, l.next_mp, l.next_mp_rc, l.is_real)
    opcode = l.is_step_forward * MEM_STEP_FORWARD + l.is_step_backward * MEM_STEP_BACKWARD
    b.receive(MEMINSTR, [l.clk, l.pc, opcode, mp, next_mp], is_real)   # the lookup on the SUM of the two selectors
    return b


# ---- MemCols (memory/memory.rs:22-44): two SingleMemoryLocal { addr, initial_clk, final_clk, initial_value, final_value, is_real } --
def memory_eval(local, nxt=None, is_first_row=0, is_last_row=0, is_transition=0, prep=None):
    """MemoryChip::eval (memory/memory.rs:131-145): no constraints, one receive (initial access) and one send (final access) per entry."""
    r, b = [int(x) for x in local], _Rec()
    assert len(r) == 12
    for k in range(2):
        addr, initial_clk, final_clk, initial_value, final_value, is_real = r[6 * k:6 * k + 6]
        b.receive(MEMORY, [initial_clk, addr, initial_value], is_real)
        b.send(MEMORY, [final_clk, addr, final_value], is_real)
    return b


# ---- IoCols (io/mod.rs:19-31): pc, mp, mv, is_input, is_output -------------------------------------------------------------------
def io_eval(local, nxt=None, is_first_row=0, is_last_row=0, is_transition=0, prep=None):
    """IoChip::eval (io/mod.rs:127-141)."""
    pc, mp, mv, is_input, is_output = [int(x) for x in local]
    b = _Rec()
    is_real = is_input + is_output
    b.boolean(is_input)
    b.boolean(is_output)
    b.boolean(is_real)
    b.receive(IO, [pc, is_input * INPUT + is_output * OUTPUT, mp, mv], is_real)
    return b


# ---- ProgramPreprocessedCols { pc, InstructionCols { opcode, op_a: Word } } + ProgramMultiplicityCols (program/mod.rs:31-41) --------
def program_eval(local, nxt=None, is_first_row=0, is_last_row=0, is_transition=0, prep=None):
    """ProgramChip::eval (program/mod.rs:152-163): receive_program (air/program.rs:31-42) lists the opcode twice."""
    (multiplicity,), q, b = [int(x) for x in local], [int(x) for x in prep], _Rec()
    assert len(q) == 6
    pc, opcode, op_a = q[0], q[1], q[2:6]
    b.receive(PROGRAM, [pc, opcode, opcode] + op_a, multiplicity)
    return b


# ---- BytePreprocessedCols { value_u8, value_u16 } + ByteMultCols { multiplicities[2] } (bytes/cols.rs:13-30) ---------------------
def byte_eval(local, nxt=None, is_first_row=0, is_last_row=0, is_transition=0, prep=None):
    """ByteChip::eval (bytes/air.rs:22-44), opcodes in ByteOpcode::all() order (executor/src/events/byte.rs:109-113)."""
    mult, (value_u8, value_u16), b = [int(x) for x in local], [int(x) for x in prep], _Rec()
    assert len(mult) == 2
    b.receive(BYTE, [U8RANGE, value_u8, 0], mult[U8RANGE])
    b.receive(BYTE, [U16RANGE, 0, value_u16], mult[U16RANGE])
    return b
