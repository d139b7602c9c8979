"""Native executor (`bfgpu_execute`, csrc/tracegen.cuh) against the Python restatement of the reference's
`Program::from` / `Executor::run` (zkvm-brainfuck_b200/machine/executor.py; reference crates/core/executor/src/
program.rs:22-44, executor.rs:71-79,106-325) on the reference's own test programs.  Runs without a GPU: with no
context the cycle records live in ordinary host memory."""
import importlib
import os

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf

ex = importlib.import_module("zkvm-brainfuck_b200.machine.executor")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PROGRAMS = [("++-.", []), (">><", []), ("[----]", []), (",.", [7]), ("++[>+<-]>.", []), ("<+>+", []), ("loop.bf", []), ("move.bf", []),
            ("printa.bf", []), ("hello.bf", []), ("fibo.bf", [17]), ("-[>-[>+>+>+<<<-]<-]", [])]


def _code(c):
    return open(os.path.join(GOLD, c)).read() if c.endswith(".bf") else c


@pytest.mark.parametrize("code,stdin", PROGRAMS)
def test_native_executor_matches_python_executor(code, stdin):
    code = _code(code)
    prog = ex.Program(code)
    ref = ex.execute(prog, stdin)
    rec = bf.Record(code, stdin)
    ops, args = rec.program()
    assert (ops == prog.opcodes).all() and (args == prog.op_a).all()
    assert rec.cycles == ref.cycles and rec.output == ref.output
    assert (rec.n_alu, rec.n_jump, rec.n_mem_instr, rec.n_io, rec.n_cells) == (
        ref.alu.shape[0], ref.jump.shape[0], ref.mem_instr.shape[0], ref.io.shape[0], ref.memory.shape[0])
    cyc, cpu = rec.cycle_records().astype(np.int64), ref.cpu
    assert (cyc[:-1, 0] == cpu[:, 1]).all() and (cyc[1:, 0] == cpu[:, 2]).all()          # pc, next_pc
    assert (cyc[:-1, 1] == cpu[:, 3]).all() and (cyc[1:, 1] == cpu[:, 4]).all()          # mp, next_mp
    assert ((cyc[:-1, 3] & 0xFF) == cpu[:, 5]).all()                                     # mv
    assert (((cyc[:-1, 3] >> 8) & 0xFF) == cpu[:, 8]).all() and (cyc[:-1, 2] == cpu[:, 10]).all()  # previous value / timestamp
    assert (rec.memory_events().astype(np.int64) == ref.memory).all()
    rec.free()


def test_fibo_output_is_pinned_by_the_reference():
    """fibo(17) = 85: the interpreter result the reference pins (crates/core/executor/src/executor.rs:335-416)."""
    assert bf.Record(_code("fibo.bf"), [17]).output == [85]
    assert bytes(bf.Record(_code("hello.bf")).output) == b"Hello"


@pytest.mark.parametrize("code,stdin,msg", [("[", [], "unmatched"), ("]", [], "unmatched"), (",", [], "stdin"), ("+x", [], "unexpected"),
                                            ("+[]", [], "cycle limit")])
def test_executor_errors(code, stdin, msg):
    with pytest.raises(bf.BfGpuError, match=msg):
        bf.Record(code, stdin, max_cycles=1 << 16)
