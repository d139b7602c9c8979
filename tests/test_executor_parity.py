"""Native executor (`bfgpu_execute`, csrc/tracegen.cuh) against the Python restatement of the reference's
`Program::from` / `Executor::run` (oracle/machine/executor.py; reference crates/core/executor/src/
program.rs:22-44, executor.rs:71-79,106-325) on the reference's own test programs.  Runs without a GPU: with no
context the cycle records live in ordinary host memory."""
import importlib
import os

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf

ex = importlib.import_module("oracle.machine.executor")
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


def _random_program(rng, n_ops, depth=0):
    """Random Brainfuck text with balanced brackets; loops are of the form [-...] so that they terminate quickly."""
    out = []
    while n_ops > 0:
        r = rng.random()
        if r < 0.12 and depth < 3 and n_ops > 4:
            body_len = int(rng.integers(1, min(n_ops - 2, 6) + 1))
            inner = _random_program(rng, body_len, depth + 1)
            # move away and back inside the body would break termination: keep the loop cell fixed by balancing moves
            bal = inner.count(">") - inner.count("<")
            inner += "<" * bal if bal > 0 else ">" * (-bal)
            out.append("[-" + inner + "]")
            n_ops -= body_len + 2
        else:
            out.append(str(rng.choice(list("+-><.,"), p=[0.3, 0.15, 0.2, 0.15, 0.1, 0.1])))
            n_ops -= 1
    return "".join(out)


def test_random_programs_match_the_python_executor():
    """200 random terminating programs (nested loops, pointer wrap-around below cell 0, input and output)."""
    rng = np.random.default_rng(2024)
    done = 0
    for _ in range(400):
        code = _random_program(rng, int(rng.integers(1, 40)))
        stdin = [int(rng.integers(0, 256))]
        try:
            rec = bf.Record(code, stdin, max_cycles=20000)
        except bf.BfGpuError as e:
            assert "cycle limit" in str(e)  # nested decrement loops over 255 can be long: skip those
            continue
        ref = ex.execute(ex.Program(code), stdin)
        assert rec.cycles == ref.cycles and rec.output == ref.output, code
        cyc = rec.cycle_records().astype(np.int64)
        if rec.cycles:
            assert (cyc[:-1, 0] == ref.cpu[:, 1]).all() and (cyc[:-1, 1] == ref.cpu[:, 3]).all(), code
            assert ((cyc[:-1, 3] & 0xFF) == ref.cpu[:, 5]).all() and (cyc[:-1, 2] == ref.cpu[:, 10]).all(), code
        assert (rec.memory_events().astype(np.int64).reshape(-1, 5) == ref.memory.reshape(-1, 5)).all(), code
        done += 1
    assert done >= 200
