"""air/chips.py (the declarative AIR every product and oracle component is generated from or evaluates) against
tests/ref_air.py, evaluators of all eight chips written by hand from the reference's Rust: identical constraint
values (in order) and identical lookup tuples on valid trace rows, on rows with one column corrupted, and on uniformly
random rows — so a wrong, missing or reordered constraint in chips.py can no longer hide behind the fact that the
oracle prover, the CUDA codegen and both verifiers share it."""
import importlib

import numpy as np
import pytest

import ref_air as R

P = R.P
chips_mod = importlib.import_module("zkvm-brainfuck_b200.air.chips")
ex = importlib.import_module("oracle.machine.executor")
tg = importlib.import_module("oracle.machine.tracegen")
CHIPS = {c.name: c for c in chips_mod.machine_chips()}
REF = {"Cpu": R.cpu_eval, "Jump": R.jump_eval, "AddSub": R.addsub_eval, "MemoryInstrs": R.meminstr_eval, "Memory": R.memory_eval, "IO": R.io_eval,
       "Program": R.program_eval, "Byte": R.byte_eval}
ALL = list(REF)
# constraints / sends / receives per chip, counted in the reference's eval functions
COUNTS = {"Cpu": (20, 14, 2), "Jump": (44, 0, 1), "AddSub": (8, 3, 2), "MemoryInstrs": (40, 0, 1), "Memory": (0, 2, 2), "IO": (3, 0, 1),
          "Program": (0, 0, 1), "Byte": (0, 0, 2)}
# columns whose sum is the chip's "this row is real" flag
REAL = {"Cpu": (30,), "Jump": (43, 44), "AddSub": (5, 6), "MemoryInstrs": (40,), "Memory": (5, 11), "IO": (3, 4), "Program": (0,), "Byte": (0, 1)}


def dsl_eval(chip, local, nxt, sel, prep=None):
    """Evaluate the chip's recorded constraints / lookups (air/dsl.py objects) on one row pair with plain ints."""
    from oracle import prover as PR

    class IntAlg:
        const = staticmethod(lambda v: int(v) % P)
        add = staticmethod(lambda a, b: (a + b) % P)
        sub = staticmethod(lambda a, b: (a - b) % P)
        mul = staticmethod(lambda a, b: a * b % P)

    def leaf(n):
        if n.op == "sel":
            return sel[n.args[0]] % P
        kind, off, idx = n.args
        if kind == "prep":
            assert not off
            return int(prep[idx]) % P
        assert kind == "main"
        return int((nxt if off else local)[idx]) % P

    cons = PR.eval_exprs(chip.constraints, leaf, IntAlg)

    def aff(a):
        v = a.const
        for (kind, idx), w in a.terms:
            assert kind in ("main", "prep")
            v = (v + w * int((local if kind == "main" else prep)[idx])) % P
        return v

    lk = lambda lst: [(l.kind, tuple(aff(v) for v in l.values), aff(l.multiplicity)) for l in lst]
    return [int(c) for c in cons], lk(chip.sends), lk(chip.receives)


def compare(name, local, nxt, sel, prep=None):
    kw = {} if prep is None else {"prep": prep}
    ref = REF[name](local, nxt, sel["is_first_row"], sel["is_last_row"], sel["is_transition"], **kw)
    cons, sends, recvs = dsl_eval(CHIPS[name], local, nxt, sel, prep)
    assert len(cons) == len(ref.constraints), (name, len(cons), len(ref.constraints))
    for k, (a, b) in enumerate(zip(cons, ref.constraints)):
        assert a == b, f"{name}: constraint #{k} differs: chips.py {a} vs reference transcription {b}"
    assert sends == ref.sends, f"{name}: sends differ"
    assert recvs == ref.receives, f"{name}: receives differ"
    return ref


@pytest.fixture(scope="module")
def traces(oracle):
    out = {}
    for code, stdin in (("++[>+<-]>,.", [9]), ("+++[>++[>+<-]<-]>>,.[-]", [200])):
        prog = ex.Program(code)
        t, q = tg.generate_traces(ex.execute(prog, stdin)), tg.preprocessed_traces(prog)
        for name in REF:
            out.setdefault(name, []).append(np.asarray(t[name], np.uint64))
            if name in q:
                out.setdefault("prep:" + name, []).append(np.asarray(q[name], np.uint64))
    return out


def prep_rows(traces, name, k):
    return traces["prep:" + name][k] if CHIPS[name].prep_width else None


def selectors(i, n, rng=None):
    if rng is not None:
        return {k: int(rng.integers(0, P)) for k in ("is_first_row", "is_last_row", "is_transition")}
    return {"is_first_row": int(i == 0), "is_last_row": int(i == n - 1), "is_transition": int(i != n - 1)}


@pytest.mark.parametrize("name", ALL)
def test_counts_match_the_reference_structure(name):
    chip = CHIPS[name]
    z = [0] * chip.main_width
    ref = REF[name](z, z, 0, 0, 0, **({"prep": [0] * chip.prep_width} if chip.prep_width else {}))
    assert (len(chip.constraints), len(chip.sends), len(chip.receives)) == (len(ref.constraints), len(ref.sends), len(ref.receives)) == COUNTS[name]
    assert (chip.main_width, chip.prep_width) == {"Cpu": (31, 0), "Jump": (45, 0), "AddSub": (7, 0), "MemoryInstrs": (41, 0), "Memory": (12, 0),
                                                  "IO": (5, 0), "Program": (1, 6), "Byte": (2, 2)}[name]
    # MachineAir::local_only (alu/mod.rs:122, jump/trace.rs:71, io/mod.rs:107 return true; default false, stark/src/air/machine.rs:51)
    assert chip.local_only == (name in ("AddSub", "Jump", "IO"))


@pytest.mark.parametrize("name", ALL)
def test_valid_rows_satisfy_both_and_agree(name, traces):
    for k, t in enumerate(traces[name]):
        n = t.shape[0]
        q = prep_rows(traces, name, k)
        rows = list(range(min(n, 40))) + list(range(max(0, n - 8), n))
        if name == "Byte":  # rows with a non-zero multiplicity
            rows += [int(i) for i in np.nonzero(t.sum(axis=1))[0][:60]]
        for i in rows:
            ref = compare(name, t[i], t[(i + 1) % n], selectors(i, n), None if q is None else q[i])
            assert not any(ref.constraints), f"{name} row {i}: a generated trace row violates the reference constraints"


@pytest.mark.parametrize("name", ALL)
def test_corrupted_and_random_rows_agree_and_every_constraint_fires(name, traces):
    rng = np.random.default_rng(20251018)
    width = CHIPS[name].main_width
    pw = CHIPS[name].prep_width
    fired = set()
    t = traces[name][1]
    q = prep_rows(traces, name, 1)
    n = t.shape[0]
    real = [i for i in range(n - 1) if sum(int(t[i, c]) for c in REAL[name])]
    assert real
    # (a) one column of a valid row (local or next) replaced by a random or an off-by-one value: real selectors
    for trial in range(600):
        i = int(rng.choice(real))
        local, nxt = t[i].copy(), t[(i + 1) % n].copy()
        tgt = local if rng.random() < 0.8 else nxt
        c = int(rng.integers(0, width))
        tgt[c] = (int(tgt[c]) + 1) % P if rng.random() < 0.5 else int(rng.integers(0, P))
        ref = compare(name, local, nxt, selectors(i, n), None if q is None else q[i])
        fired.update(k for k, v in enumerate(ref.constraints) if v)
    # first / last row variants of the same
    for i in (0, n - 1):
        for c in range(width):
            local, nxt = t[i].copy(), t[(i + 1) % n].copy()
            local[c] = (int(local[c]) + 1) % P
            ref = compare(name, local, nxt, selectors(i, n), None if q is None else q[i])
            fired.update(k for k, v in enumerate(ref.constraints) if v)
    # (b) uniformly random rows and random selector values: every product of conditions is exercised as a polynomial identity
    for trial in range(200):
        local = rng.integers(0, P, width, dtype=np.uint64)
        nxt = rng.integers(0, P, width, dtype=np.uint64)
        ref = compare(name, local, nxt, selectors(0, 2, rng), rng.integers(0, P, pw, dtype=np.uint64) if pw else None)
        fired.update(k for k, v in enumerate(ref.constraints) if v)
    # the Cpu chip's `clk == limb16 + limb8 * 2^16` check (cpu/air.rs:124-129) is an identity in its own columns: it can never fire
    never = {7} if name == "Cpu" else set()
    assert fired | never == set(range(len(CHIPS[name].constraints))), f"{name}: constraints never violated by the test rows: " \
        f"{sorted(set(range(len(CHIPS[name].constraints))) - fired - never)}"
    assert not (fired & never)
