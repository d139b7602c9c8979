"""The eight chips of the Brainfuck machine as constraint programs + lookup tables.

Each `eval_*` restates, call for call, the reference chip's `Air::eval` (the order of the calls is the
alpha-folding order of the constraints and the chunking order of the lookups):

  Cpu           crates/core/machine/src/cpu/air.rs:28-62,159-185,100-123,126-139,145-156 (cols cpu/cols.rs:29-71)
  Program       crates/core/machine/src/program/mod.rs:152-163 (cols :31-41)
  AddSub        crates/core/machine/src/alu/mod.rs:159-193 + operations/add.rs:42-75
  Jump          crates/core/machine/src/jump/air.rs:22-82 + operations/is_zero.rs:42-66 + operations/koala_bear_word.rs:52-106
  Memory        crates/core/machine/src/memory/memory.rs:132-145
  Byte          crates/core/machine/src/bytes/air.rs:22-44
  MemoryInstrs  crates/core/machine/src/memory/instructions/air.rs:25-76
  IO            crates/core/machine/src/io/mod.rs:127-141
Helper traits: crates/core/machine/src/air/{memory,program,u8_air}.rs; chip order and names:
crates/core/machine/src/brainfuck/mod.rs:53-81.
"""
from .dsl import Chip, Expr, layout, MEMORY, PROGRAM

# Opcode / ByteOpcode (crates/core/executor/src/opcode.rs:13-42)
LOOP_START, LOOP_END, ADD, SUB, MEM_FWD, MEM_BWD, INPUT, OUTPUT = range(8)
U8_RANGE, U16_RANGE = 0, 1

ACCESS = [("value", 1), ("prev_clk", 1), ("diff_16bit_limb", 1), ("diff_8bit_limb", 1)]  # MemoryAccessCols
RANGE_CHECKER = [("msb_decomp", 8), ("and_0_to_2", 1), ("and_0_to_3", 1), ("and_0_to_4", 1), ("and_0_to_5", 1), ("and_0_to_6", 1),
                 ("and_0_to_7", 1)]  # KoalaBearWordRangeChecker


def _prefixed(prefix, fields):
    return [(prefix + "_" + n, w) for n, w in fields]


def reduce_word(w):
    """Word::reduce (crates/stark/src/word.rs:67-73)."""
    return w[0] + w[1] * (1 << 8) + w[2] * (1 << 16) + w[3] * (1 << 24)


# ---- shared gadgets ----------------------------------------------------------------------------------
def eval_range_check_24bits(b, value, limb_16, limb_8, do_check):
    """MemoryAirBuilder::eval_range_check_24bits (air/memory.rs:97-125)."""
    b.when(do_check).assert_eq(value, limb_16 + limb_8 * (1 << 16))
    b.send_byte(U16_RANGE, 0, limb_16, do_check)
    b.send_byte(U8_RANGE, limb_8, 0, do_check)


def eval_memory_access(b, clk, addr, prev_value, value, prev_clk, diff16, diff8, do_check):
    """MemoryAirBuilder::eval_memory_access (+ _timestamp) (air/memory.rs:17-95)."""
    b.assert_bool(do_check)
    diff_minus_one = clk - prev_clk - 1
    eval_range_check_24bits(b, diff_minus_one, diff16, diff8, do_check)
    b.send(MEMORY, [prev_clk, addr, prev_value], do_check)
    b.receive(MEMORY, [clk, addr, value], do_check)


def range_check_u8(b, value, mult):
    """U8AirBuilder::range_check_u8 (air/u8_air.rs:8-15)."""
    b.send_byte(U8_RANGE, value, 0, mult)


def koala_bear_word_range_check(b, value, cols, prefix, is_real):
    """KoalaBearWordRangeChecker::range_check (operations/koala_bear_word.rs:52-106)."""
    bits = getattr(cols, prefix + "_msb_decomp")
    a = lambda k: getattr(cols, f"{prefix}_and_0_to_{k}")
    recomposed = Expr.const(0)
    for i, bit in enumerate(bits):
        b.when(is_real).assert_bool(bit)
        recomposed = recomposed + Expr.const(1 << i) * bit
    b.when(is_real).assert_eq(recomposed, value[3])
    b.when(is_real).assert_zero(bits[7])
    b.when(is_real).assert_eq(a(2), bits[0] * bits[1])
    b.when(is_real).assert_eq(a(3), a(2) * bits[2])
    b.when(is_real).assert_eq(a(4), a(3) * bits[3])
    b.when(is_real).assert_eq(a(5), a(4) * bits[4])
    b.when(is_real).assert_eq(a(6), a(5) * bits[5])
    b.when(is_real).assert_eq(a(7), a(6) * bits[6])
    b.when(is_real).when(a(7)).assert_zero(value[0] + value[1] + value[2])


# ---- Cpu ---------------------------------------------------------------------------------------------------
CPU_LAYOUT, CPU_WIDTH = layout(
    [("clk_16bit_limb", 1), ("clk_8bit_limb", 1), ("pc", 1), ("next_pc", 1), ("mp", 1), ("next_mp", 1), ("mv", 1), ("next_mv", 1),
     ("opcode", 1), ("op_a", 4), ("mv_access_prev_value", 1)] + _prefixed("mv_access", ACCESS) + [("next_mv_access_prev_value", 1)]
    + _prefixed("next_mv_access", ACCESS)
    + [("mv_accessed", 1), ("next_mv_accessed", 1), ("is_mv_immutable", 1), ("is_alu", 1), ("is_jump", 1), ("is_io", 1),
       ("is_memory_instr", 1), ("is_real", 1)])


def send_program(b, pc, opcode, op_a, mult, receive=False):
    """ProgramAirBuilder::{send,receive}_program (air/program.rs:13-42): [pc, opcode, opcode, op_a[0..4]]."""
    values = [pc, opcode, opcode] + list(op_a)
    (b.receive if receive else b.send)(PROGRAM, values, mult)


def eval_cpu(b):
    local, nxt = b.main(0), b.main(1)
    clk = Expr.const(1 << 16) * local.clk_8bit_limb + local.clk_16bit_limb
    send_program(b, local.pc, local.opcode, local.op_a, local.is_real)
    # eval_instruction
    b.send_alu(local.pc, local.opcode, local.next_mv, local.mv, local.is_alu)
    b.send_jump(local.pc, local.next_pc, local.opcode, local.mv, local.is_jump)
    b.send_memory_instr(clk, local.pc, local.opcode, local.mp, local.next_mp, local.is_memory_instr)
    b.send_io(local.pc, local.opcode, local.mp, local.mv, local.is_io)
    # eval_registers
    eval_memory_access(b, clk + 1, local.mp, local.mv_access_prev_value, local.mv_access_value, local.mv_access_prev_clk,
                       local.mv_access_diff_16bit_limb, local.mv_access_diff_8bit_limb, local.mv_accessed)
    eval_memory_access(b, clk + 2, local.mp, local.next_mv_access_prev_value, local.next_mv_access_value, local.next_mv_access_prev_clk,
                       local.next_mv_access_diff_16bit_limb, local.next_mv_access_diff_8bit_limb, local.next_mv_accessed)
    range_check_u8(b, local.mv, local.is_real)
    b.when(local.is_mv_immutable).assert_eq(local.mv_access_value, local.mv_access_prev_value)
    # eval_clk
    b.when_first_row().assert_zero(clk)
    next_clk = Expr.const(1 << 16) * nxt.clk_8bit_limb + nxt.clk_16bit_limb
    b.when_transition().when(nxt.is_real).assert_eq(clk + 2, next_clk)
    eval_range_check_24bits(b, clk, local.clk_16bit_limb, local.clk_8bit_limb, local.is_real)
    # eval_pc
    b.when_transition().when(nxt.is_real).assert_eq(local.next_pc, nxt.pc)
    b.when_transition().when(local.is_real).when_not(local.is_jump).assert_eq(local.next_pc, local.pc + 1)
    # eval_is_real
    b.assert_bool(local.is_real)
    b.when_first_row().assert_one(local.is_real)
    b.when_transition().when_not(local.is_real).assert_zero(nxt.is_real)
    for col in (local.is_alu, local.is_jump, local.is_memory_instr, local.is_io, local.is_mv_immutable, local.mv_accessed, local.next_mv_accessed):
        b.assert_bool(col)


# ---- Program --------------------------------------------------------------------------------------------------
PROGRAM_PREP_LAYOUT, PROGRAM_PREP_WIDTH = layout([("pc", 1), ("opcode", 1), ("op_a", 4)])
PROGRAM_LAYOUT, PROGRAM_WIDTH = layout([("multiplicity", 1)])


def eval_program(b):
    prep, mult = b.preprocessed(0), b.main(0)
    send_program(b, prep.pc, prep.opcode, prep.op_a, mult.multiplicity, receive=True)


# ---- AddSub ------------------------------------------------------------------------------------------------------
ADDSUB_LAYOUT, ADDSUB_WIDTH = layout([("pc", 1), ("value", 1), ("carry", 1), ("operand_1", 1), ("operand_2", 1), ("is_add", 1), ("is_sub", 1)])


def eval_addsub(b):
    local = b.main(0)
    is_real = local.is_add + local.is_sub
    b.assert_bool(local.is_add)
    b.assert_bool(local.is_sub)
    b.assert_bool(is_real)
    # AddOperation::eval (operations/add.rs:42-75)
    br = b.when(is_real)
    overflow = local.operand_1 + local.operand_2 - local.value
    br.assert_zero(overflow * (overflow - 256))
    br.assert_zero(local.carry * (overflow - 256))
    br.assert_zero((local.carry - 1) * overflow)
    br.assert_bool(local.carry)
    br.assert_bool(is_real)
    range_check_u8(b, local.operand_1, is_real)
    range_check_u8(b, local.operand_2, is_real)
    range_check_u8(b, local.value, is_real)
    b.receive_alu(local.pc, ADD, local.value, local.operand_1, local.is_add)
    b.receive_alu(local.pc, SUB, local.operand_1, local.value, local.is_sub)


# ---- Jump -------------------------------------------------------------------------------------------------------------
JUMP_LAYOUT, JUMP_WIDTH = layout([("pc", 4)] + _prefixed("pc", RANGE_CHECKER) + [("next_pc", 4)] + _prefixed("next_pc", RANGE_CHECKER)
                                 + [("dst", 4), ("mv", 1), ("is_mv_zero_inverse", 1), ("is_mv_zero_result", 1), ("is_loop_start", 1),
                                    ("is_loop_end", 1)])


def eval_jump(b):
    local = b.main(0)
    is_real = local.is_loop_start + local.is_loop_end
    b.assert_bool(local.is_loop_start)
    b.assert_bool(local.is_loop_end)
    b.assert_bool(is_real)
    # IsZeroOperation::eval (operations/is_zero.rs:42-66)
    is_zero = 1 - local.is_mv_zero_inverse * local.mv
    b.when(is_real).assert_eq(is_zero, local.is_mv_zero_result)
    b.when(is_real).assert_bool(local.is_mv_zero_result)
    b.when(is_real).when(local.is_mv_zero_result).assert_zero(local.mv)
    pc, next_pc, dst = reduce_word(local.pc), reduce_word(local.next_pc), reduce_word(local.dst)
    b.when(local.is_loop_start).when(local.is_mv_zero_result).assert_eq(next_pc, dst)
    b.when(local.is_loop_start).when_not(local.is_mv_zero_result).assert_eq(next_pc, pc + 1)
    b.when(local.is_loop_end).when_not(local.is_mv_zero_result).assert_eq(next_pc, dst)
    b.when(local.is_loop_end).when(local.is_mv_zero_result).assert_eq(next_pc, pc + 1)
    koala_bear_word_range_check(b, local.pc, local, "pc", is_real)
    koala_bear_word_range_check(b, local.next_pc, local, "next_pc", is_real)
    opcode = local.is_loop_start * LOOP_START + local.is_loop_end * LOOP_END
    b.receive_jump(pc, next_pc, opcode, local.mv, is_real)


# ---- Memory (initialize / finalize) ----------------------------------------------------------------------------------------
_MEM_ENTRY = [("addr", 1), ("initial_clk", 1), ("final_clk", 1), ("initial_value", 1), ("final_value", 1), ("is_real", 1)]
MEMORY_LAYOUT, MEMORY_WIDTH = layout(_prefixed("e0", _MEM_ENTRY) + _prefixed("e1", _MEM_ENTRY))


def eval_memory(b):
    local = b.main(0)
    for e in ("e0", "e1"):
        g = lambda n: getattr(local, f"{e}_{n}")
        b.receive(MEMORY, [g("initial_clk"), g("addr"), g("initial_value")], g("is_real"))
        b.send(MEMORY, [g("final_clk"), g("addr"), g("final_value")], g("is_real"))


# ---- Byte ------------------------------------------------------------------------------------------------------------------
BYTE_PREP_LAYOUT, BYTE_PREP_WIDTH = layout([("value_u8", 1), ("value_u16", 1)])
BYTE_LAYOUT, BYTE_WIDTH = layout([("multiplicities", 2)])


def eval_byte(b):
    mult, prep = b.main(0), b.preprocessed(0)
    b.receive_byte(U8_RANGE, prep.value_u8, 0, mult.multiplicities[U8_RANGE])
    b.receive_byte(U16_RANGE, 0, prep.value_u16, mult.multiplicities[U16_RANGE])


# ---- MemoryInstrs --------------------------------------------------------------------------------------------------------------
MEMINSTR_LAYOUT, MEMINSTR_WIDTH = layout([("pc", 1), ("clk", 1), ("mp", 4)] + _prefixed("mp", RANGE_CHECKER) + [("next_mp", 4)]
                                         + _prefixed("next_mp", RANGE_CHECKER) + [("is_step_forward", 1), ("is_step_backward", 1), ("is_real", 1)])


def eval_memory_instrs(b):
    local, nxt = b.main(0), b.main(1)
    is_real = local.is_step_forward + local.is_step_backward
    b.assert_bool(local.is_step_forward)
    b.assert_bool(local.is_step_backward)
    b.assert_bool(is_real)
    mp, next_mp = reduce_word(local.mp), reduce_word(local.next_mp)
    b.when(local.is_step_forward).assert_eq(next_mp, mp + 1)
    b.when(local.is_step_backward).assert_eq(next_mp, mp - 1)
    b.when_transition().when(nxt.is_real).assert_eq(next_mp, reduce_word(nxt.mp))
    koala_bear_word_range_check(b, local.mp, local, "mp", local.is_real)
    koala_bear_word_range_check(b, local.next_mp, local, "next_mp", local.is_real)
    opcode = local.is_step_forward * MEM_FWD + local.is_step_backward * MEM_BWD
    b.receive_memory_instr(local.clk, local.pc, opcode, mp, next_mp, is_real)


# ---- IO ---------------------------------------------------------------------------------------------------------------------------
IO_LAYOUT, IO_WIDTH = layout([("pc", 1), ("mp", 1), ("mv", 1), ("is_input", 1), ("is_output", 1)])


def eval_io(b):
    local = b.main(0)
    is_real = local.is_input + local.is_output
    b.assert_bool(local.is_input)
    b.assert_bool(local.is_output)
    b.assert_bool(is_real)
    opcode = local.is_input * INPUT + local.is_output * OUTPUT
    b.receive_io(local.pc, opcode, local.mp, local.mv, is_real)


def machine_chips():
    """BfAir::chips() order (brainfuck/mod.rs:53-81)."""
    return [
        Chip("Cpu", CPU_WIDTH, 0, False, eval_cpu, CPU_LAYOUT),
        Chip("Program", PROGRAM_WIDTH, PROGRAM_PREP_WIDTH, False, eval_program, PROGRAM_LAYOUT, PROGRAM_PREP_LAYOUT),
        Chip("AddSub", ADDSUB_WIDTH, 0, True, eval_addsub, ADDSUB_LAYOUT),
        Chip("Jump", JUMP_WIDTH, 0, True, eval_jump, JUMP_LAYOUT),
        Chip("Memory", MEMORY_WIDTH, 0, False, eval_memory, MEMORY_LAYOUT),
        Chip("Byte", BYTE_WIDTH, BYTE_PREP_WIDTH, False, eval_byte, BYTE_LAYOUT, BYTE_PREP_LAYOUT),
        Chip("MemoryInstrs", MEMINSTR_WIDTH, 0, False, eval_memory_instrs, MEMINSTR_LAYOUT),
        Chip("IO", IO_WIDTH, 0, True, eval_io, IO_LAYOUT),
    ]
