"""Error paths of the C ABI on the device: a failing entry point must hand every device block it took back
to the context (`bfgpu_debug_live_blocks` returns to where it was) and the context must keep working afterwards; the
executor's cycle limit must hold for recycled (pooled) record buffers too."""
import importlib

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf

pytestmark = pytest.mark.gpu
ex = importlib.import_module("oracle.machine.executor")
tg = importlib.import_module("oracle.machine.tracegen")


def _sweep(ctx, call, max_allocs=400):
    """Run `call()` with the n-th device allocation failing, n = 0, 1, 2, ... until it succeeds; returns the number of
    failures injected.  After every failure the context owns exactly the blocks it owned before."""
    base = ctx.live_blocks
    for n in range(max_allocs):
        ctx.fail_alloc(n)
        try:
            res = call()
        except bf.BfGpuError as e:
            assert "injected allocation failure" in str(e), e
            assert ctx.live_blocks == base, f"allocation #{n} failing leaked {ctx.live_blocks - base} blocks"
            continue
        finally:
            ctx.fail_alloc(-1)
        return n, res
    pytest.fail("call never succeeded")


def test_pcs_commit_releases_everything_on_allocation_failure(oracle):
    rng = np.random.default_rng(5)
    evals = [rng.integers(0, bf.P, s, dtype=np.uint32) for s in [(1 << 13, 40), (1 << 13, 3), (1 << 10, 9), (64, 2)]]
    ref = oracle.PcsData(evals)
    ctx = bf.Context()
    pcs = bf.TwoAdicFriPcs(ctx)

    def commit():
        root, data = pcs.commit(evals)
        data.free()
        return root

    n, root = _sweep(ctx, commit)
    assert n >= 10 and (root == ref.root).all()
    # the pipelined host path (one wide matrix): staging blocks, sponge states and events included
    wide = [rng.integers(0, bf.P, (1 << 13, 160), dtype=np.uint32)]
    refw = oracle.PcsData(wide)
    n, root = _sweep(ctx, lambda: (lambda r, d: (d.free(), r)[1])(*pcs.commit(wide)))
    assert n >= 5 and (root == refw.root).all()
    ctx.close()


def test_machine_prove_releases_everything_on_allocation_failure(oracle):
    prog = ex.Program("++[>+<-]>,.")
    rec = ex.execute(prog, [7])
    traces, preps = tg.generate_traces(rec), tg.preprocessed_traces(prog)
    ctx = bf.Context()
    ctx.set_fri_params(1, 6, 3)
    prover = bf.CudaProver(ctx)
    pk = prover.setup(preps)
    good, _ = prover.prove(pk, traces, bf.Challenger(ctx), raw=True)
    shard = prover.commit(traces)
    ch = bf.Challenger(ctx)
    bf.lib().bfgpu_pk_observe_into(pk._h, ch._h)
    n, words = _sweep(ctx, lambda: prover.open_raw(pk, shard, ch.clone()), max_allocs=1500)
    assert n >= 50 and (words == good).all(), "the proof after the failed attempts differs"
    shard.free()
    # program -> proof (device-side trace generation)
    n, (res, nrec) = _sweep(ctx, lambda: prover.prove_program("++[>+<-]>,.", [7], pk=pk, raw=True), max_allocs=1500)
    assert (res[0] == good).all()
    pk.free()
    ctx.close()


def test_cycle_limit_applies_to_pooled_record_buffers():
    """The second execution on a context gets the first one's (large) page-locked buffer back: max_cycles must still bind."""
    ctx = bf.Context()
    loop = "+[]"  # never terminates
    for attempt in range(3):
        with pytest.raises(bf.BfGpuError, match="cycle limit"):
            bf.Record(loop, ctx=ctx, max_cycles=5000)
    ok = bf.Record("+++.", ctx=ctx, max_cycles=5000)
    assert ok.cycles == 4 and ok.output == [3]
    # exactly at the limit: a program of max_cycles cycles runs, one more does not
    assert bf.Record("+" * 64, ctx=ctx, max_cycles=64).cycles == 64
    with pytest.raises(bf.BfGpuError, match="cycle limit"):
        bf.Record("+" * 65, ctx=ctx, max_cycles=64)
    ctx.close()
