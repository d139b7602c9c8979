"""air/chips.py (the declarative AIR every product and oracle component is generated from or evaluates) against
tests/ref_air.py, evaluators of the Cpu and Jump chips written by hand from the reference's Rust: identical constraint
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
REF = {"Cpu": R.cpu_eval, "Jump": R.jump_eval}


def dsl_eval(chip, local, nxt, sel):
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
        assert kind == "main"
        return int((nxt if off else local)[idx]) % P

    cons = PR.eval_exprs(chip.constraints, leaf, IntAlg)

    def aff(a):
        v = a.const
        for (kind, idx), w in a.terms:
            assert kind == "main"
            v = (v + w * int(local[idx])) % P
        return v

    lk = lambda lst: [(l.kind, tuple(aff(v) for v in l.values), aff(l.multiplicity)) for l in lst]
    return [int(c) for c in cons], lk(chip.sends), lk(chip.receives)


def compare(name, local, nxt, sel):
    ref = REF[name](local, nxt, sel["is_first_row"], sel["is_last_row"], sel["is_transition"])
    cons, sends, recvs = dsl_eval(CHIPS[name], local, nxt, sel)
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
        t = tg.generate_traces(ex.execute(ex.Program(code), stdin))
        for name in REF:
            out.setdefault(name, []).append(np.asarray(t[name], np.uint64))
    return out


def selectors(i, n, rng=None):
    if rng is not None:
        return {k: int(rng.integers(0, P)) for k in ("is_first_row", "is_last_row", "is_transition")}
    return {"is_first_row": int(i == 0), "is_last_row": int(i == n - 1), "is_transition": int(i != n - 1)}


@pytest.mark.parametrize("name", ["Cpu", "Jump"])
def test_counts_match_the_reference_structure(name):
    chip = CHIPS[name]
    z = [0] * chip.main_width
    ref = REF[name](z, z, 0, 0, 0)
    assert len(chip.constraints) == len(ref.constraints) == {"Cpu": 20, "Jump": 44}[name]
    assert (len(chip.sends), len(chip.receives)) == (len(ref.sends), len(ref.receives)) == {"Cpu": (14, 2), "Jump": (0, 1)}[name]


@pytest.mark.parametrize("name", ["Cpu", "Jump"])
def test_valid_rows_satisfy_both_and_agree(name, traces):
    for t in traces[name]:
        n = t.shape[0]
        for i in list(range(min(n, 40))) + list(range(max(0, n - 8), n)):
            ref = compare(name, t[i], t[(i + 1) % n], selectors(i, n))
            assert not any(ref.constraints), f"{name} row {i}: a generated trace row violates the reference constraints"


@pytest.mark.parametrize("name", ["Cpu", "Jump"])
def test_corrupted_and_random_rows_agree_and_every_constraint_fires(name, traces):
    rng = np.random.default_rng(20251018)
    width = CHIPS[name].main_width
    fired = set()
    t = traces[name][1]
    n = t.shape[0]
    real = [i for i in range(n - 1) if (t[i, 30] if name == "Cpu" else t[i, 43] + t[i, 44])]
    # (a) one column of a valid row (local or next) replaced by a random or an off-by-one value: real selectors
    for trial in range(600):
        i = int(rng.choice(real))
        local, nxt = t[i].copy(), t[(i + 1) % n].copy()
        tgt = local if rng.random() < 0.8 else nxt
        c = int(rng.integers(0, width))
        tgt[c] = (int(tgt[c]) + 1) % P if rng.random() < 0.5 else int(rng.integers(0, P))
        ref = compare(name, local, nxt, selectors(i, n))
        fired.update(k for k, v in enumerate(ref.constraints) if v)
    # first / last row variants of the same
    for i in (0, n - 1):
        for c in range(width):
            local, nxt = t[i].copy(), t[(i + 1) % n].copy()
            local[c] = (int(local[c]) + 1) % P
            ref = compare(name, local, nxt, selectors(i, n))
            fired.update(k for k, v in enumerate(ref.constraints) if v)
    # (b) uniformly random rows and random selector values: every product of conditions is exercised as a polynomial identity
    for trial in range(200):
        local = rng.integers(0, P, width, dtype=np.uint64)
        nxt = rng.integers(0, P, width, dtype=np.uint64)
        ref = compare(name, local, nxt, selectors(0, 2, rng))
        fired.update(k for k, v in enumerate(ref.constraints) if v)
    # the Cpu chip's `clk == limb16 + limb8 * 2^16` check (cpu/air.rs:124-129) is an identity in its own columns: it can never fire
    never = {7} if name == "Cpu" else set()
    assert fired | never == set(range(len(CHIPS[name].constraints))), f"{name}: constraints never violated by the test rows: " \
        f"{sorted(set(range(len(CHIPS[name].constraints))) - fired - never)}"
    assert not (fired & never)
