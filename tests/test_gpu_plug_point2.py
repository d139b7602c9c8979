"""Plug point #2 of SURVEY.md §8b: the steps of `CpuProver::open` exposed one by one for an integration at the Plonky3
trait level — `Chip::generate_permutation_trace` (chip.rs:117-136), `quotient_values` (quotient.rs:18-165) reading the
committed LDEs in place, and the device view of `get_evaluations_on_domain` — each against the CPU oracle."""
import ctypes as C
import importlib

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf

pytestmark = pytest.mark.gpu
ex = importlib.import_module("oracle.machine.executor")
tg = importlib.import_module("oracle.machine.tracegen")
chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()


@pytest.fixture(scope="module")
def setup(oracle):
    from oracle import prover as PR, stark as S
    prog = ex.Program("++[>+<-]>,.")
    traces, preps = tg.generate_traces(ex.execute(prog, [3])), tg.preprocessed_traces(prog)
    ctx = bf.Context()
    yield ctx, PR, S, traces, preps
    ctx.close()


@pytest.mark.parametrize("name", [c.name for c in chips])
def test_perm_trace_and_quotient_match_oracle(setup, name, oracle):
    ctx, PR, S, traces, preps = setup
    chip = {c.name: c for c in chips}[name]
    main = traces[name]
    prep = preps.get(name)
    rng = np.random.default_rng(hash(name) % 1000)
    alpha, beta, fold_alpha = (rng.integers(0, bf.P, 4, dtype=np.uint64) for _ in range(3))
    prep_or_empty = prep if prep is not None else np.zeros((main.shape[0], 0), np.uint32)
    operm, ocs = PR.generate_permutation_trace(chip, prep_or_empty, main, [alpha, beta])
    gperm, gcs = bf.generate_permutation_trace(ctx, name, main, prep, alpha, beta)
    assert (gperm == PR.flatten_to_base(operm)).all() and (gcs == ocs).all()
    # commit the three matrices, evaluate the quotient from the committed LDEs on both sides
    pcs = bf.TwoAdicFriPcs(ctx)
    _, dm = pcs.commit([main])
    _, dq = pcs.commit([gperm])
    dp = pcs.commit([prep])[1] if prep is not None else None
    ld = main.shape[0].bit_length() - 1
    nat = lambda m: oracle.coset_lde_batch(m, 1, 3)
    prep_q = nat(prep) if prep is not None else np.zeros((2 << ld, 1), np.uint32)
    want = PR.quotient_values(chip, ocs, ld, prep_q, nat(main), nat(gperm), [alpha, beta], fold_alpha)
    got = bf.quotient_values(ctx, name, dp, 0 if dp is not None else -1, dm, 0, dq, 0, fold_alpha, [alpha, beta], gcs)
    assert (got == np.asarray(want, np.uint32)).all()
    # device view of the committed LDE: same words as the host copy-out, Montgomery form, column-major, rows bit-reversed
    dev, rows, cols, stride = C.c_void_p(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    ctx.check(bf.lib().bfgpu_pcs_lde_device(dm._h, 0, C.byref(dev), C.byref(rows), C.byref(cols), C.byref(stride)))
    assert (rows.value, cols.value, stride.value) == (2 * main.shape[0], main.shape[1], 2 * main.shape[0])
    host = np.zeros(rows.value * cols.value, np.uint32)
    ctx.synchronize()
    try:
        from cuda.bindings import runtime as cudart
    except ImportError:
        from cuda import cudart
    err = cudart.cudaMemcpy(host.ctypes.data, dev.value, host.nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
    assert int(err[0]) == 0, err
    rinv = pow((1 << 32) % bf.P, bf.P - 2, bf.P)
    canon = (host.astype(np.uint64) * np.uint64(rinv) % np.uint64(bf.P)).astype(np.uint32).reshape(cols.value, rows.value).T
    assert (canon == pcs.get_evaluations_on_domain(dm, 0, bit_reversed_rows=True)).all()
    for d in (dm, dq, dp):
        if d is not None:
            d.free()


def test_plug_point_argument_errors(setup):
    ctx, PR, S, traces, preps = setup
    with pytest.raises(bf.BfGpuError, match="unknown chip"):
        bf.generate_permutation_trace(ctx, "Nope", traces["Cpu"], None, np.zeros(4), np.zeros(4))
    with pytest.raises(bf.BfGpuError, match="width"):
        bf.generate_permutation_trace(ctx, "Cpu", traces["AddSub"], None, np.zeros(4), np.zeros(4))
    with pytest.raises(bf.BfGpuError, match="preprocessed"):
        bf.generate_permutation_trace(ctx, "Program", traces["Program"], None, np.zeros(4), np.zeros(4))
