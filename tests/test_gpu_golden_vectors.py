"""The CUDA backend against the committed known-answer vectors (tests/golden/vectors.json; generator
scripts/gen_golden.py, CPU twin tests/test_golden_vectors.py): primitives, LDE, commitments, openings and the
transcript-visible parts of two whole proofs produced from the PROGRAM (native executor + device-side traces)."""
import json
import os

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf
from proofio import proof_sha256

pytestmark = pytest.mark.gpu
P = 2130706433
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
V = json.load(open(os.path.join(GOLD, "vectors.json")))


@pytest.fixture(scope="module")
def ctx():
    c = bf.Context()
    yield c
    c.close()


def seeded(seed, rows, cols):
    return np.random.default_rng(seed).integers(0, P, (rows, cols), dtype=np.uint32)


def test_primitive_vectors(ctx):
    assert ctx.permute(np.arange(16, dtype=np.uint32).reshape(1, 16))[0].tolist() == V["poseidon2_permute_0_to_15"]
    assert ctx.hash_rows(np.arange(31, dtype=np.uint32).reshape(1, 31))[0].tolist() == V["sponge_hash_0_to_30"]
    assert ctx.compress(np.arange(8, dtype=np.uint32).reshape(1, 8), np.arange(8, 16, dtype=np.uint32).reshape(1, 8))[0].tolist() == \
        V["compress_0_to_7_and_8_to_15"]
    lde = bf.Radix2Dit(ctx).coset_lde_batch(seeded(1, 64, 3), 1, 3)
    assert [lde[0].tolist(), lde[127].tolist()] == V["coset_lde_seed1_64x3_row0_row127"]


def test_commit_vectors(ctx):
    mats = [seeded(2, 1 << 10, 31), seeded(3, 1 << 10, 2), seeded(4, 1 << 6, 7), seeded(5, 16, 5)]
    root, tree = bf.MerkleTreeMmcs(ctx).commit(mats)
    assert root.tolist() == V["mmcs_root_seeds2to5"]
    tree.free()
    pcs = bf.TwoAdicFriPcs(ctx)
    root, data = pcs.commit(mats)
    assert root.tolist() == V["pcs_commit_root_seeds2to5"]
    assert bf.MerkleTreeMmcs(ctx).open_batch(1234, data.tree)[1].tolist() == V["pcs_open_batch_1234_siblings"]
    data.free()
    root, data = pcs.commit([seeded(6, 4096, 100)])
    assert root.tolist() == V["pcs_commit_root_seed6_4096x100"]
    data.free()


@pytest.mark.parametrize("name", ["hello", "fibo"])
def test_program_proof_vectors(ctx, name):
    g = V["proofs"][name]
    ctx.set_fri_params(*g["fri"])
    try:
        prover = bf.CudaProver(ctx)
        rec = prover.execute(open(os.path.join(GOLD, name + ".bf")).read(), g["stdin"])
        assert rec.cycles == g["cycles"] and rec.output == g["output"]
        pk = prover.setup_record(rec)
        assert pk.commit.tolist() == g["preprocessed_commit"]
        ch = bf.Challenger(ctx)
        bf.lib().bfgpu_pk_observe_into(pk._h, ch._h)
        shard = prover.commit_record(rec)
        proof = prover.open(pk, shard, ch.clone())
        shard.free()
        assert {k: np.asarray(proof["commitment"][k]).tolist() for k in g["commitments"]} == g["commitments"]
        assert proof["chip_ordering"] == g["chip_ordering"]
        assert [np.asarray(c["cumulative_sum"]).tolist() for c in proof["opened_values"]] == g["cumulative_sums"]
        fri = proof["opening_proof"]
        assert [np.asarray(c).tolist() for c in fri["commit_phase_commits"]] == g["fri_commit_phase_commits"]
        assert np.asarray(fri["final_poly"]).tolist() == g["final_poly"] and int(fri["pow_witness"]) == g["pow_witness"]
        assert [int(q["index"]) for q in fri["query_proofs"]] == g["query_indices"]
        # the whole serialised proof, byte for byte, against the digest of the oracle's proof
        words = prover.open_raw(pk, (shard2 := prover.commit_record(rec)), ch.clone())
        shard2.free()
        assert int(words.size) == g["proof_words"] and proof_sha256(words) == g["proof_sha256"]
        pk.free()
    finally:
        ctx.set_fri_params(1, 84, 16)


def test_fibo_proof_at_full_parameters_is_the_oracle_proof(ctx):
    """BASELINE config 1 (`test_e2e_core`: fibo.bf, stdin [17]) at the reference's own FRI parameters (84 queries, 16 PoW bits,
    kb31_poseidon2.rs:54-64): the GPU proof, produced from the program, is byte-identical to the CPU oracle's proof
    (SHA-256 of the ~0.7 MB serialisation, committed by scripts/gen_golden.py), same PoW witness, same query indices."""
    g = V["proofs"]["fibo_full_parameters"]
    ctx.set_fri_params(*g["fri"])
    prover = bf.CudaProver(ctx)
    (words, decode), rec = prover.prove_program(open(os.path.join(GOLD, "fibo.bf")).read(), g["stdin"], raw=True)
    assert rec.output == [85]
    proof = decode()
    assert int(proof["opening_proof"]["pow_witness"]) == g["pow_witness"]
    assert [int(q["index"]) for q in proof["opening_proof"]["query_proofs"]] == g["query_indices"]
    assert {k: np.asarray(proof["commitment"][k]).tolist() for k in g["commitments"]} == g["commitments"]
    assert int(words.size) == g["proof_words"] and proof_sha256(words) == g["proof_sha256"]


@pytest.mark.parametrize("log_rows", [16, 22])
def test_bench_workload_root_equals_cpu_oracle_root(ctx, log_rows):
    """The exact bench.py input (bench_workload, seed 0xB200, 2^22 x 256 = BASELINE config 4) committed on the GPU gives the
    root the CPU oracle computed for it (tests/golden/bench_roots.json)."""
    import ctypes as C
    import torch
    import bench_workload as W
    gold = json.load(open(os.path.join(GOLD, "bench_roots.json")))["roots"]
    key = "log_rows=16,seed=0xB200" if log_rows == 16 else "log_rows=22,seed=0xB200+0"
    trace = W.trace_torch(1 << log_rows, 256, device="cuda")
    torch.cuda.synchronize()
    ctx.set_input_space(bf.MEM_DEVICE)
    try:
        mat = bf.Mat(trace.data_ptr(), 1 << log_rows, 256)
        root = np.zeros(8, np.uint32)
        h = C.c_void_p()
        ctx.check(bf.lib().bfgpu_pcs_commit(ctx._h, C.byref(mat), None, 1, root.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(h)))
        bf.lib().bfgpu_pcs_data_free(h)
        assert root.tolist() == gold[key]
    finally:
        ctx.set_input_space(bf.MEM_HOST)
        del trace
        torch.cuda.empty_cache()
