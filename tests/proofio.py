"""Shared test helper: ShardProof dict (oracle/prover.py, CudaProver._parse) -> the flat u32 serialisation documented at
bfgpu_machine_open / bfgpu_pcs_open (include/bfgpu.h), and its SHA-256 (golden pin of whole proofs)."""
import hashlib
import importlib

import numpy as np

chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()


def serialize(proof, pk_names):
    """ShardProof dict -> the flat u32 layout documented at bfgpu_machine_open / bfgpu_pcs_open (include/bfgpu.h)."""
    idx = {c.name: i for i, c in enumerate(chips)}
    by_name = {c.name: c for c in chips}
    order = sorted(proof["chip_ordering"], key=lambda k: proof["chip_ordering"][k])
    ov = proof["opened_values"]
    w = []
    put = lambda a: w.extend(int(x) for x in np.asarray(a).ravel())
    for k in ("main", "permutation", "quotient"):
        put(proof["commitment"][k])
    w.append(len(order))
    for name, v in zip(order, ov):
        w += [idx[name], v["log_degree"]]
        put(v["cumulative_sum"])

    def lv(vals, both):
        put(vals["local"])
        if both:
            put(vals["next"])

    for name in pk_names:
        lv(ov[proof["chip_ordering"][name]]["preprocessed"], not by_name[name].local_only)
    for name, v in zip(order, ov):
        lv(v["main"], not by_name[name].local_only)
    for v in ov:
        lv(v["permutation"], True)
    for v in ov:
        for q in v["quotient"]:
            put(q)
    fri = proof["opening_proof"]
    w.append(len(fri["commit_phase_commits"]))
    for c in fri["commit_phase_commits"]:
        put(c)
    put(fri["final_poly"])
    w.append(int(fri["pow_witness"]))
    w.append(len(fri["query_proofs"]))
    for q in fri["query_proofs"]:
        w.append(int(q["index"]))
        for rnd in q["input_proof"]:
            for row in rnd["opened_values"]:
                put(row)
            put(rnd["opening_proof"])
        for st in q["commit_phase_openings"]:
            put(st["sibling_value"])
            put(st["opening_proof"])
    return np.array(w, np.uint32)


def proof_sha256(words):
    return hashlib.sha256(np.ascontiguousarray(words, dtype="<u4").tobytes()).hexdigest()
