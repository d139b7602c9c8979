#!/usr/bin/env python3
"""PIN KIT, backend side: compare vectors dumped from the REAL Rust prover (integration/dump_vectors.rs, run with cargo inside
the reference) with the backend's committed golden vectors (tests/golden/vectors.json, computed by the CPU oracle and
reproduced by the CUDA path in tests/test_gpu_golden_vectors.py).

    python scripts/compare_golden.py rust_vectors.json

Prints one line per stage in pipeline order — the FIRST mismatch is the place to look — and, for the stages that depend on a
choice the restatement could not confirm offline (include/bfgpu.h, bfgpu_set_transcript_option), which setting of the
oracle's switches (oracle/stark.py: OBSERVE_OPENED_VALUES, FRI_ROLLIN) reproduces the Rust value.  Exit status 0 = every
compared field equal.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STAGES = [  # pipeline order
    ("poseidon2_permute_0_to_15", "Poseidon2 permutation (constants carve-out kb31_poseidon2.rs:35-50, internal diagonal, x^3)"),
    ("sponge_hash_0_to_30", "PaddingFreeSponge<16,8,8> (overwrite mode, partial last block)"),
    ("compress_0_to_7_and_8_to_15", "TruncatedPermutation<2,8,16>"),
    ("coset_lde_seed1_64x3_row0_row127", "coset_lde_batch (shift = generator 3, natural order)"),
    ("pcs_commit_root_seeds2to5", "TwoAdicFriPcs::commit (bit-reversed LDE rows, MerkleTreeMmcs with injection of shorter matrices)"),
]
PROOF_FIELDS = [
    ("preprocessed_commit", "StarkMachine::setup commit (Program + Byte traces)"),
    ("commitments.main", "main commitment: trace generation + commit order (height desc, name)"),
    ("commitments.permutation", "LogUp traces: challenger sampling order (alpha, beta), batching, running sum"),
    ("cumulative_sums", "LogUp cumulative sums"),
    ("commitments.quotient", "quotient: constraint folding order, selectors on the coset, chunk domains"),
    ("fri_commit_phase_commits", "pcs.open: opened values, alpha, reduced openings, FRI folds   [OBSERVE_OPENED_VALUES / FRI_ROLLIN]"),
    ("final_poly", "FRI final polynomial                                              [FRI_ROLLIN]"),
]


def get(d, path):
    for k in path.split("."):
        d = d[k]
    return d


def main():
    if len(sys.argv) != 2:
        raise SystemExit(__doc__)
    rust = json.load(open(sys.argv[1]))
    ours = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
    bad = 0
    for key, what in STAGES:
        if key not in rust:
            print(f"  skipped   {key}: not in the Rust dump")
            continue
        ok = rust[key] == ours[key]
        bad += not ok
        print(f"{'  equal  ' if ok else 'DIFFERENT'}  {key}: {what}")
    for name, rp in rust.get("proofs", {}).items():
        op = ours["proofs"].get(name)
        if op is None or rp.get("fri") != op.get("fri"):
            print(f"  skipped   proof {name}: golden vector has FRI parameters {op and op.get('fri')}, dump has {rp.get('fri')} (set FRI_QUERIES)")
            continue
        print(f"proof {name}: cycles {rp.get('cycles')} (golden {op.get('cycles')}), output {rp.get('output')} (golden {op.get('output')})")
        for path, what in PROOF_FIELDS:
            try:
                ok = get(rp, path) == get(op, path)
            except KeyError:
                continue
            bad += not ok
            print(f"{'  equal  ' if ok else 'DIFFERENT'}  {name}.{path}: {what}")
            if not ok and path == "commitments.main":
                print("            (the Memory chip's row order is process-random in the reference: re-run gen_golden.py with the dumped "
                      "memory_trace before reading on — tests/test_oracle_machine.py shows how a trace is injected)")
        if "proof_bincode_len" in rp:
            print(f"            proofSize (bincode bytes): Rust {rp['proof_bincode_len']}; backend: python bfprove.py size <proof> --vk <vk>")
        # everything after the proof of work depends on the witness the Rust prover happened to find
        if rp.get("pow_witness") != op.get("pow_witness"):
            print(f"            pow_witness: Rust {rp['pow_witness']}, golden {op['pow_witness']} — both valid (find_any); "
                  f"re-prove with fixed_pow_witness={rp['pow_witness']} to compare the query openings")
    print("RESULT:", "all compared fields equal — the oracle (and with it the CUDA path) is pinned to the reference's bytes" if bad == 0
          else f"{bad} field(s) differ: the first DIFFERENT line above is where the restatement departs from Plonky3 rev 93967fce")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
