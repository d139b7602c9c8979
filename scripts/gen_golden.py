#!/usr/bin/env python3
"""Generate tests/golden/vectors.json: known-answer vectors of every stage of the path, computed by the CPU oracle
(oracle/, itself cross-checked against independent restatements in tests/test_oracle_*.py).

The reference holds no byte-level fixtures for this path and cannot be built offline, so these vectors do
NOT pin the oracle to Plonky3's bytes — "parity unpinned" stands.  They pin the oracle and the CUDA path to each other and
to this commit: `-m "not gpu"` tests recompute them with the oracle, `-m gpu` tests with the CUDA backend.

Run from the repo root:  python scripts/gen_golden.py
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from proofio import serialize, proof_sha256  # noqa: E402
from oracle import prover as PR, stark as S  # noqa: E402

P = 2130706433
ex = importlib.import_module("oracle.machine.executor")
tg = importlib.import_module("oracle.machine.tracegen")
chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()
GOLD = os.path.join(ROOT, "tests", "golden")


def seeded(seed, rows, cols):
    return np.random.default_rng(seed).integers(0, P, (rows, cols), dtype=np.uint32)


def main():
    if len(sys.argv) == 3 and sys.argv[1] == "--inputs-only":
        # the seeded input matrices of the LDE / commitment vectors, for integration/dump_vectors.rs (numpy's generator is not
        # reproducible from Rust): BFGPU_GOLDEN_INPUTS=<this file> cargo test ...
        json.dump({"seed1_64x3": seeded(1, 64, 3).tolist(), "seed2_1024x31": seeded(2, 1 << 10, 31).tolist(), "seed3_1024x2": seeded(3, 1 << 10, 2).tolist(),
                   "seed4_64x7": seeded(4, 1 << 6, 7).tolist(), "seed5_16x5": seeded(5, 16, 5).tolist()}, open(sys.argv[2], "w"))
        print("wrote", sys.argv[2])
        return
    v = {}
    v["poseidon2_permute_0_to_15"] = oracle.permute(np.arange(16, dtype=np.uint32)).tolist()
    v["sponge_hash_0_to_30"] = oracle.sponge_hash(np.arange(31, dtype=np.uint32)).tolist()
    v["compress_0_to_7_and_8_to_15"] = oracle.compress(np.arange(8, dtype=np.uint32), np.arange(8, 16, dtype=np.uint32)).tolist()
    a = seeded(1, 64, 3)
    lde = oracle.coset_lde_batch(a, 1, 3)
    v["coset_lde_seed1_64x3_row0_row127"] = [lde[0].tolist(), lde[127].tolist()]
    mats = [seeded(2, 1 << 10, 31), seeded(3, 1 << 10, 2), seeded(4, 1 << 6, 7), seeded(5, 16, 5)]
    v["mmcs_root_seeds2to5"] = oracle.Tree(mats).root.tolist()
    pd = oracle.PcsData(mats)
    v["pcs_commit_root_seeds2to5"] = pd.root.tolist()
    rows, sib = pd.tree.open_batch(1234)
    v["pcs_open_batch_1234_siblings"] = sib.tolist()
    v["pcs_commit_root_seed6_4096x100"] = oracle.PcsData([seeded(6, 4096, 100)]).root.tolist()
    proofs = {}
    for name, stdin in (("hello", []), ("fibo", [17])):
        prog = ex.Program(open(os.path.join(GOLD, name + ".bf")).read())
        rec = ex.execute(prog, stdin)
        traces, preps = tg.generate_traces(rec), tg.preprocessed_traces(prog)
        cfg = S.FriConfig(1, 12, 6)
        pk = PR.setup(chips, preps)
        ch = S.Challenger()
        PR.observe_pk(pk, ch)
        proof = PR.prove_shard(chips, pk, traces, ch.clone(), cfg)
        proofs[name] = dict(
            stdin=stdin, cycles=rec.cycles, output=rec.output, fri=[1, 12, 6],
            preprocessed_commit=np.asarray(pk.commit).tolist(),
            commitments={k: np.asarray(proof["commitment"][k]).tolist() for k in ("main", "permutation", "quotient")},
            chip_ordering=proof["chip_ordering"],
            cumulative_sums=[np.asarray(c["cumulative_sum"]).tolist() for c in proof["opened_values"]],
            fri_commit_phase_commits=[np.asarray(c).tolist() for c in proof["opening_proof"]["commit_phase_commits"]],
            final_poly=np.asarray(proof["opening_proof"]["final_poly"]).tolist(),
            pow_witness=int(proof["opening_proof"]["pow_witness"]),
            query_indices=[int(q["index"]) for q in proof["opening_proof"]["query_proofs"]])
        proof.pop("_debug", None)
        words = serialize(proof, pk.names)
        proofs[name]["proof_words"] = int(words.size)
        proofs[name]["proof_sha256"] = proof_sha256(words)  # the WHOLE serialised proof (layout: include/bfgpu.h, bfgpu_machine_open)
        if name == "fibo":
            # BASELINE config 1 (`test_e2e_core`) at the reference's full parameters: 84 queries, 16 proof-of-work bits
            # (kb31_poseidon2.rs:54-64), smallest witness.  Only digests are committed (the proof is ~1 MB).
            full = PR.prove_shard(chips, pk, traces, ch.clone(), S.FriConfig(1, 84, 16))
            full.pop("_debug", None)
            fw = serialize(full, pk.names)
            proofs["fibo_full_parameters"] = dict(
                stdin=stdin, fri=[1, 84, 16], proof_words=int(fw.size), proof_sha256=proof_sha256(fw),
                pow_witness=int(full["opening_proof"]["pow_witness"]),
                commitments={k: np.asarray(full["commitment"][k]).tolist() for k in ("main", "permutation", "quotient")},
                final_poly=np.asarray(full["opening_proof"]["final_poly"]).tolist(),
                query_indices=[int(q["index"]) for q in full["opening_proof"]["query_proofs"]])
    v["proofs"] = proofs
    with open(os.path.join(GOLD, "vectors.json"), "w") as f:
        json.dump(v, f, indent=1, sort_keys=True)
    print("wrote", os.path.join(GOLD, "vectors.json"))


if __name__ == "__main__":
    main()
