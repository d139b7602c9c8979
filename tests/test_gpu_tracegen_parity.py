"""Device-side trace generation (csrc/tracegen.cuh: `bfgpu_machine_commit_record`) against the numpy restatement of the
reference's `MachineAir::generate_trace` implementations and `generate_dependencies`
(oracle/machine/tracegen.py; reference files cited there): every chip's main trace word for word, the
main commitment, and the whole proof from `prove_program` equal to the proof obtained from host traces (which
tests/test_gpu_prove_parity.py pins against the oracle prover and verifier)."""
import importlib
import os

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ex = importlib.import_module("oracle.machine.executor")
tg = importlib.import_module("oracle.machine.tracegen")

PROGRAMS = [("+.", []), ("++-.", []), (">><", []), ("[----]", []), (",.", [7]), ("++[>+<-]>.", []), ("<+>+", []), ("loop.bf", []), ("move.bf", []),
            ("printa.bf", []), ("hello.bf", []), ("fibo.bf", [17]), ("-[>-[>+>+>+<<<-]<-]", []), ("+>" * 3000 + ",.", [255])]


@pytest.fixture(scope="module")
def ctx():
    c = bf.Context()
    yield c
    c.close()


def _code(c):
    return open(os.path.join(GOLD, c)).read() if c.endswith(".bf") else c


@pytest.mark.parametrize("code,stdin", PROGRAMS)
def test_device_traces_equal_host_traces(ctx, code, stdin):
    code = _code(code)
    prog = ex.Program(code)
    ref = tg.generate_traces(ex.execute(prog, stdin))
    prover = bf.CudaProver(ctx)
    rec = prover.execute(code, stdin)
    shard = prover.commit_record(rec)
    got = prover.shard_traces(shard)
    assert sorted(got) == sorted(ref)
    for name in ref:
        assert got[name].shape == ref[name].shape, name
        bad = np.argwhere(got[name] != ref[name])
        assert bad.size == 0, (name, bad[:5].tolist())
    host = prover.commit(ref)
    assert shard.names == host.names and shard.heights == host.heights
    assert (shard.commit == host.commit).all()
    shard.free()
    host.free()
    rec.free()


@pytest.mark.parametrize("code,stdin", [("hello.bf", []), ("fibo.bf", [17]), (",.", [7])])
def test_prove_program_equals_proof_from_host_traces(ctx, oracle, code, stdin):
    from oracle import prover as PR, stark as S
    code = _code(code)
    queries, pow_bits = 12, 6
    ctx.set_fri_params(1, queries, pow_bits)
    try:
        prover = bf.CudaProver(ctx)
        (buf, decode), rec = prover.prove_program(code, stdin, raw=True)
        prog = ex.Program(code)
        pyrec = ex.execute(prog, stdin)
        traces, preps = tg.generate_traces(pyrec), tg.preprocessed_traces(prog)
        pk = prover.setup(preps)
        buf2, _ = prover.prove(pk, traces, bf.Challenger(ctx), raw=True)
        assert buf.shape == buf2.shape and (buf == buf2).all()
        assert rec.output == pyrec.output
        # and the restated verifier accepts it (crates/stark/src/verifier.rs:27-216)
        chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()
        local_only = dict((c[0], c[4]) for c in prover.chips)
        vk = dict(commit=pk.commit, chip_information=[(n, h.bit_length() - 1, local_only[n]) for n, h in zip(pk.names, pk.heights)])
        och = S.Challenger()
        och.observe_digest(pk.commit)
        for _ in range(7):
            och.observe(0)
        assert PR.verify_shard(chips, vk, decode(), och, S.FriConfig(1, queries, pow_bits)) is None
        pk.free()
    finally:
        ctx.set_fri_params(1, 84, 16)


def test_prove_many_pipelines_executor_and_gpu(ctx):
    """A stream of jobs through `prove_many` (interpreter of job k+1 overlapped with the GPU proof of job k) gives the
    same proofs as one-at-a-time `prove_program`."""
    ctx.set_fri_params(1, 12, 6)
    try:
        prover = bf.CudaProver(ctx)
        jobs = [(_code("hello.bf"), []), (",.", [9]), (_code("fibo.bf"), [17]), (_code("hello.bf"), [])]
        got = [(buf.copy(), rec.output) for buf, rec in prover.prove_many(jobs)]
        assert len(got) == len(jobs)
        for (code, stdin), (buf, output) in zip(jobs, got):
            (ref, _), rec = prover.prove_program(code, stdin, raw=True)
            assert output == rec.output and buf.shape == ref.shape and (buf == ref).all()
        assert (got[0][0] == got[3][0]).all()
    finally:
        ctx.set_fri_params(1, 84, 16)


@pytest.mark.parametrize("code,cycles,cpu_rows", [("-[>-[>+>+>+<<<-]<-]", 716807, 1 << 20), ("++++++++[>-[>-[>+>+<<-]<-]<-]", 4173897, 1 << 22)])
def test_fullsize_program_proof_is_accepted_by_the_restated_verifier(ctx, oracle, code, cycles, cpu_rows):
    """BASELINE config 3 and the north-star size, program -> proof entirely on the backend (native executor, device
    traces, 84 queries, 16 PoW bits); too big for a word-for-word comparison with the numpy prover, so the check is the
    reference's own acceptance test: `Verifier::verify_shard` (restated, oracle/prover.py) must accept, which includes
    C(zeta) = q(zeta) Z_H(zeta) for every chip, the FRI low-degree test, all Merkle openings and the zero total of the
    LogUp cumulative sums."""
    from oracle import prover as PR, stark as S
    prover = bf.CudaProver(ctx)
    rec = prover.execute(code)
    assert rec.cycles == cycles
    pk = prover.setup_record(rec)
    ch = bf.Challenger(ctx)
    bf.lib().bfgpu_pk_observe_into(pk._h, ch._h)
    shard = prover.commit_record(rec)
    assert shard.names[0] == "Cpu" and shard.heights[0] == cpu_rows
    words = prover.open_raw(pk, shard, ch.clone())
    proof = prover._parse(words, pk, None)
    shard.free()
    # native verifier (csrc/verifier.h) on the serialised proof, and rejection of a flipped word in the middle
    assert bf.verify_shard(pk.commit, pk.names, pk.heights, words) is None
    bad = words.copy()
    bad[len(bad) // 2] ^= 1
    assert bf.verify_shard(pk.commit, pk.names, pk.heights, bad) is not None
    chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()
    local_only = dict((c[0], c[4]) for c in prover.chips)
    vk = dict(commit=pk.commit, chip_information=[(n, h.bit_length() - 1, local_only[n]) for n, h in zip(pk.names, pk.heights)])
    och = S.Challenger()
    och.observe_digest(pk.commit)
    for _ in range(7):
        och.observe(0)
    assert PR.verify_shard(chips, vk, proof, och, S.FriConfig()) is None
    total = S.E_ZERO
    for c in proof["opened_values"]:
        total = S.e_add(total, c["cumulative_sum"])
    assert S.e_eq(total, S.E_ZERO)
    pk.free()
    rec.free()


def test_e2e_core_through_the_sdk_facade():
    """The reference's own end-to-end test (`test_e2e_core`, crates/sdk/src/lib.rs:185-196) with this backend behind the
    same calls: setup, prove fibo(17), check the output 85, verify; a proof for another stdin must not verify under a
    tampered statement."""
    client = bf.ProverClient()
    code = _code("fibo.bf")
    assert client.execute(code, [17]).run() == [85]
    pk, vk = client.setup(code)
    proof = client.prove(pk, [17]).run()
    assert proof.output == [85] and proof.stdin == [17]
    assert client.verify(proof, vk) is None
    proof.words[40] ^= 1
    assert client.verify(proof, vk) is not None
    other_pk, other_vk = client.setup(_code("hello.bf"))
    proof = client.prove(pk, [17]).run()
    assert client.verify(proof, other_vk) is not None  # verifying key of a different program


def test_random_programs_device_traces(ctx):
    """Random terminating programs (nested loops, pointer wrap-around, input/output): every device-generated trace equals
    the numpy restatement and the device commitment equals the commitment of the host traces."""
    from tests.test_executor_parity import _random_program
    rng = np.random.default_rng(77)
    prover = bf.CudaProver(ctx)
    done = 0
    while done < 25:
        code = _random_program(rng, int(rng.integers(1, 60)))
        stdin = [int(rng.integers(0, 256))]
        try:
            rec = prover.execute(code, stdin)
        except bf.BfGpuError:
            continue
        if rec.cycles == 0 or rec.cycles > 50000:
            continue
        ref = tg.generate_traces(ex.execute(ex.Program(code), stdin))
        shard = prover.commit_record(rec)
        got = prover.shard_traces(shard)
        assert sorted(got) == sorted(ref), code
        for name in ref:
            assert got[name].shape == ref[name].shape and (got[name] == ref[name]).all(), (code, name)
        host = prover.commit(ref)
        assert (shard.commit == host.commit).all(), code
        shard.free()
        host.free()
        done += 1


def test_one_cycle_execution_is_refused(ctx):
    """One cycle -> one-row Cpu trace -> LDE as short as the FRI blow-up, which p3-fri's verifier cannot consume (the
    reference would produce an unverifiable proof): the program path refuses it with a clear message."""
    prover = bf.CudaProver(ctx)
    rec = prover.execute("+")
    with pytest.raises(bf.BfGpuError, match="one-cycle"):
        prover.commit_record(rec)
