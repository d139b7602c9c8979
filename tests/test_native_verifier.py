"""Native verifier (`bfgpu_verify_shard`, csrc/verifier.h — host code, no GPU) against the oracle's restatement of
`Verifier::verify_shard` (oracle/prover.py; reference crates/stark/src/verifier.rs:27-216): same verdict on oracle-made
proofs and on corrupted copies (every error class of the reference that a single-word corruption can reach)."""
import importlib
import os

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ex = importlib.import_module("oracle.machine.executor")
tg = importlib.import_module("oracle.machine.tracegen")
chips = importlib.import_module("zkvm-brainfuck_b200.air.chips").machine_chips()
FRI = (1, 10, 5)


from proofio import serialize  # noqa: E402


@pytest.fixture(scope="module")
def hello(oracle):
    from oracle import prover as PR, stark as S
    prog = ex.Program(open(os.path.join(GOLD, "hello.bf")).read())
    rec = ex.execute(prog, [])
    traces, preps = tg.generate_traces(rec), tg.preprocessed_traces(prog)
    cfg = S.FriConfig(*FRI)
    pk = PR.setup(chips, preps)
    ch = S.Challenger()
    PR.observe_pk(pk, ch)
    proof = PR.prove_shard(chips, pk, traces, ch.clone(), cfg)
    proof.pop("_debug", None)
    vk = dict(commit=pk.commit, chip_information=[(n, t.shape[0].bit_length() - 1, lo) for n, t, lo in zip(pk.names, pk.traces, pk.local_only)])
    return PR, S, pk, vk, proof, serialize(proof, pk.names)


def native(pk, words):
    return bf.verify_shard(pk.commit, pk.names, [t.shape[0] for t in pk.traces], words, *FRI)


def oracle_verdict(PR, S, pk, vk, proof):
    och = S.Challenger()
    PR.observe_pk(pk, och)
    return PR.verify_shard(chips, vk, proof, och, S.FriConfig(*FRI))


def test_accepts_oracle_proof(hello):
    PR, S, pk, vk, proof, words = hello
    assert oracle_verdict(PR, S, pk, vk, proof) is None
    assert native(pk, words) is None
    # Montgomery-form words (what a Rust caller holds) are accepted as well
    R = (1 << 32) % bf.P
    mont = words.copy()
    # every field word is converted; structural words (counts, indices, witness) stay: rebuild through the serializer
    class M(dict):
        pass
    def conv(x):
        return (np.asarray(x, np.uint64) * R % bf.P).astype(np.uint64)
    mp = dict(commitment={k: conv(v) for k, v in proof["commitment"].items()}, chip_ordering=proof["chip_ordering"],
              opened_values=[dict(preprocessed={k: conv(v) for k, v in c["preprocessed"].items()}, main={k: conv(v) for k, v in c["main"].items()},
                                  permutation={k: conv(v) for k, v in c["permutation"].items()}, quotient=[conv(q) for q in c["quotient"]],
                                  cumulative_sum=conv(c["cumulative_sum"]), log_degree=c["log_degree"]) for c in proof["opened_values"]],
              opening_proof=dict(commit_phase_commits=[conv(c) for c in proof["opening_proof"]["commit_phase_commits"]],
                                 final_poly=conv(proof["opening_proof"]["final_poly"]), pow_witness=proof["opening_proof"]["pow_witness"],
                                 query_proofs=[dict(index=q["index"], input_proof=[dict(opened_values=[conv(r) for r in rnd["opened_values"]],
                                                                                       opening_proof=conv(rnd["opening_proof"])) for rnd in q["input_proof"]],
                                                    commit_phase_openings=[dict(sibling_value=conv(st["sibling_value"]), opening_proof=conv(st["opening_proof"]))
                                                                           for st in q["commit_phase_openings"]]) for q in proof["opening_proof"]["query_proofs"]]))
    assert bf.verify_shard(conv(pk.commit), pk.names, [t.shape[0] for t in pk.traces], serialize(mp, pk.names), *FRI, repr=bf.REPR_MONTY) is None


def _tweak(words, pos):
    bad = words.copy()
    bad[pos] = (int(bad[pos]) + 1) % bf.P
    return bad


def test_rejects_corruptions_like_the_reference(hello):
    PR, S, pk, vk, proof, words = hello
    n = len(proof["opened_values"])
    head = 24 + 1 + 6 * n
    # a cumulative sum: changes the transcript AND breaks the constraint / sum checks
    assert native(pk, _tweak(words, 24 + 1 + 2)) is not None
    # an opened value (first preprocessed column at zeta): the FRI input no longer matches the committed polynomial
    e = native(pk, _tweak(words, head))
    assert e is not None and e.startswith("Invalid")
    # a commitment
    assert native(pk, _tweak(words, 3)) is not None
    # truncation and trailing data
    assert "InvalidProofShape" in native(pk, words[:-1])
    assert "InvalidProofShape" in native(pk, np.concatenate([words, np.zeros(1, np.uint32)]))
    # proof-of-work witness
    fri = proof["opening_proof"]
    bad = dict(proof, opening_proof=dict(fri, pow_witness=(fri["pow_witness"] + 1) % bf.P))
    ev, eo = native(pk, serialize(bad, pk.names)), oracle_verdict(PR, S, pk, vk, bad)
    assert ev is not None and eo is not None and ("PowWitness" in ev) == ("PowWitness" in eo)
    # a Merkle sibling of the first query's first round
    q0 = fri["query_proofs"][0]
    r0 = q0["input_proof"][0]
    sib = np.array(r0["opening_proof"], np.uint32).copy()
    sib[0, 0] = (int(sib[0, 0]) + 1) % bf.P
    badq = dict(q0, input_proof=[dict(r0, opening_proof=sib)] + list(q0["input_proof"][1:]))
    bad = dict(proof, opening_proof=dict(fri, query_proofs=[badq] + list(fri["query_proofs"][1:])))
    ev, eo = native(pk, serialize(bad, pk.names)), oracle_verdict(PR, S, pk, vk, bad)
    assert "InputMmcsError" in ev and "InputMmcsError" in eo
    # the final polynomial
    fp = np.array(fri["final_poly"], np.uint64).copy()
    fp[1] = (int(fp[1]) + 1) % bf.P
    bad = dict(proof, opening_proof=dict(fri, final_poly=fp))
    ev, eo = native(pk, serialize(bad, pk.names)), oracle_verdict(PR, S, pk, vk, bad)
    assert ev is not None and eo is not None


def test_rejects_consistent_but_wrong_statement(hello, oracle):
    """A proof whose quotient opening is inconsistent with the constraints: swap in the quotient values of another chip
    position -> the PCS check fails before or the OOD check fails; both verifiers must reject."""
    PR, S, pk, vk, proof, words = hello
    ov = [dict(c) for c in proof["opened_values"]]
    ov[0]["quotient"], ov[1]["quotient"] = ov[1]["quotient"], ov[0]["quotient"]
    bad = dict(proof, opened_values=ov)
    assert native(pk, serialize(bad, pk.names)) is not None and oracle_verdict(PR, S, pk, vk, bad) is not None


def test_out_of_domain_check_and_cumulative_sums(oracle):
    """Proofs of traces that violate the AIR: a wrong carry bit fails `constraints(zeta)/Z_H(zeta) == quotient(zeta)`
    for the AddSub chip (verifier.rs:220-247); a wrong lookup multiplicity leaves every row constraint intact... except
    the LogUp totals, which no longer cancel (verifier.rs:210-213).  Same verdict from both verifiers."""
    from oracle import prover as PR, stark as S
    chips_mod = importlib.import_module("zkvm-brainfuck_b200.air.chips")
    prog = ex.Program("++[>+<-]>.")
    traces, preps = tg.generate_traces(ex.execute(prog)), tg.preprocessed_traces(prog)
    pk = PR.setup(chips, preps)
    ch = S.Challenger()
    PR.observe_pk(pk, ch)
    cfg = S.FriConfig(*FRI)
    vk = dict(commit=pk.commit, chip_information=[(n, t.shape[0].bit_length() - 1, lo) for n, t, lo in zip(pk.names, pk.traces, pk.local_only)])
    heights = [t.shape[0] for t in pk.traces]
    good = PR.prove_shard(chips, pk, traces, ch.clone(), cfg)
    assert bf.verify_shard(pk.commit, pk.names, heights, serialize(good, pk.names), *FRI) is None
    bad = {k: v.copy() for k, v in traces.items()}
    bad["AddSub"][0, chips_mod.ADDSUB_LAYOUT["carry"]] = 1
    proof = PR.prove_shard(chips, pk, bad, ch.clone(), cfg)
    eo = PR.verify_shard(chips, vk, proof, ch.clone(), cfg)
    ev = bf.verify_shard(pk.commit, pk.names, heights, serialize(proof, pk.names), *FRI)
    assert eo.startswith("OodEvaluationMismatch") and ev == eo
    bad = {k: v.copy() for k, v in traces.items()}
    bad["Byte"][5, chips_mod.U8_RANGE] = (int(bad["Byte"][5, chips_mod.U8_RANGE]) + 1) % bf.P
    proof = PR.prove_shard(chips, pk, bad, ch.clone(), cfg)
    eo = PR.verify_shard(chips, vk, proof, ch.clone(), cfg)
    ev = bf.verify_shard(pk.commit, pk.names, heights, serialize(proof, pk.names), *FRI)
    assert eo == "CumulativeSumsError" and ev == eo


def test_rejects_hostile_shapes_and_noncanonical_words(hello):
    """Untrusted 32-bit shape words must not wrap the range checks (log_degree = 2^32 - 1 once passed `ld + blowup > 24`),
    a field word >= p is a second encoding of the same proof and is refused, and a verifying key without preprocessed
    matrices is a shape error, not an out-of-bounds read."""
    PR, S, pk, vk, proof, words = hello
    n = len(proof["opened_values"])
    for ld in (0xFFFFFFFF, 0xFFFFFFF0, 24, 200):
        bad = words.copy()
        bad[24 + 1 + 1] = ld  # log_degree of the first chip
        assert "InvalidProofShape" in native(pk, bad)
    head = 24 + 1 + 6 * n
    nc = len(proof["opening_proof"]["commit_phase_commits"])
    # number of FRI commit-phase commitments: the word right before nc * 8 digest words, final poly, witness, query count
    cand = [i for i in range(head, len(words) - 8 * nc - 6) if int(words[i]) == nc and int(words[i + 8 * nc + 6]) == FRI[1]]
    assert cand
    for v in (0xFFFFFFFF, 24, 0):
        bad = words.copy()
        bad[cand[0]] = v
        assert native(pk, bad) is not None
    # non-canonical encodings: same residue, different word
    bad = words.copy()
    assert int(bad[0]) + bf.P < 2 ** 32
    bad[0] = int(bad[0]) + bf.P  # first word of the main commitment
    assert native(pk, bad) is not None
    bad = words.copy()
    bad[head] = int(bad[head]) + bf.P if int(bad[head]) + bf.P < 2 ** 32 else bad[head]
    if int(bad[head]) != int(words[head]):
        assert native(pk, bad) is not None
    com = np.asarray(pk.commit, np.uint64).copy()
    com[0] += bf.P
    assert bf.verify_shard(com.astype(np.uint32), pk.names, [t.shape[0] for t in pk.traces], words, *FRI) is not None
    # no preprocessed matrices at all / an absurd preprocessed height
    e = bf.verify_shard(pk.commit, [], [], words, *FRI)
    assert e is not None and "InvalidProofShape" in e
    e = bf.verify_shard(pk.commit, pk.names, [1 << 40 for _ in pk.traces], words, *FRI)
    assert e is not None


@pytest.mark.parametrize("observe,rollin", [(0, 0), (1, 1), (0, 1)])
def test_transcript_options_native_verifier_matches_oracle(oracle, monkeypatch, observe, rollin):
    """bfgpu_verify_shard_ex with non-default transcript options accepts exactly the proofs the oracle produces (and its verifier
    accepts) under the same switches, and rejects them under the default options: the switches are wired through both verifiers."""
    from oracle import prover as PR, stark as S
    monkeypatch.setattr(S, "OBSERVE_OPENED_VALUES", bool(observe))
    monkeypatch.setattr(S, "FRI_ROLLIN", rollin)
    prog = ex.Program("++[>+<-]>.")
    traces, preps = tg.generate_traces(ex.execute(prog)), tg.preprocessed_traces(prog)
    pk = PR.setup(chips, preps)
    ch = S.Challenger()
    PR.observe_pk(pk, ch)
    cfg = S.FriConfig(*FRI)
    proof = PR.prove_shard(chips, pk, traces, ch.clone(), cfg)
    proof.pop("_debug", None)
    vk = dict(commit=pk.commit, chip_information=[(n, t.shape[0].bit_length() - 1, lo) for n, t, lo in zip(pk.names, pk.traces, pk.local_only)])
    assert PR.verify_shard(chips, vk, proof, ch.clone(), cfg) is None
    words = serialize(proof, pk.names)
    heights = [t.shape[0] for t in pk.traces]
    opts = dict(observe_opened_values=observe, fri_rollin=rollin)
    assert bf.verify_shard(pk.commit, pk.names, heights, words, *FRI, options=opts) is None
    assert bf.verify_shard(pk.commit, pk.names, heights, words, *FRI) is not None


def test_core_proof_checks_of_the_prover_crate(hello):
    """`BfProver::verify` (crates/prover/src/verify.rs:10-36): Cpu chip present, Cpu log degree <= MAX_CPU_LOG_DEGREE = 22
    (crates/core/machine/src/cpu/mod.rs:8), then `StarkMachine::verify`, whose shard errors surface as InvalidShardProof
    (machine.rs:281,391-416)."""
    PR, S, pk, vk, proof, words = hello
    heights = [t.shape[0] for t in pk.traces]
    core = lambda w: bf.verify_core_proof(pk.commit, pk.names, heights, w, *FRI)  # noqa: E731
    assert core(words) is None
    order = sorted(proof["chip_ordering"], key=proof["chip_ordering"].get)
    names = [c.name for c in chips]
    n = len(order)
    entries = [(int(words[25 + 6 * i]), int(words[25 + 6 * i + 1])) for i in range(n)]
    assert [names[c] for c, _ in entries] == order  # the header layout this test edits
    at = order.index("Cpu")
    assert entries[at][1] == [c["log_degree"] for c in proof["opened_values"]][at]
    # the Cpu entry renamed to another chip: the reference's `chip_ordering.contains_key("Cpu")` is false
    other = names.index("Byte")
    bad = words.copy()
    bad[25 + 6 * at] = other
    assert core(bad) == "MissingCpuInFirstShard"
    # a proof cut off before the chip list holds no Cpu chip either
    assert core(words[:20]) == "MissingCpuInFirstShard"
    # log degree above 22: refused before the shard verifier looks at anything else
    for ld in (23, 24, 0xFFFFFFFF):
        bad = words.copy()
        bad[25 + 6 * at + 1] = ld
        assert core(bad) == f"CpuLogDegreeTooLarge: {ld}"
    bad = words.copy()
    bad[25 + 6 * at + 1] = 22  # allowed by this check, wrong for the proof
    e = core(bad)
    assert e is not None and e.startswith("InvalidShardProof: ")
    # shard-level failures carry the reference's wrapper name and the same inner error as bfgpu_verify_shard
    bad = _tweak(words, 3)
    assert core(bad) == "InvalidShardProof: " + native(pk, bad)
    # ProverClient.verify is this entry point
    client = bf.ProverClient()
    pv = bf.ProofWithPublicValues(words, None, [], [])
    key = dict(commit=pk.commit, names=pk.names, heights=heights)
    assert client.verify(pv, key, *FRI) is None
    assert client.verify(bf.ProofWithPublicValues(bad, None, [], []), key, *FRI).startswith("InvalidShardProof: ")


def test_differential_error_names_on_random_field_corruptions(hello):
    """Random single-word corruptions anywhere in the proof (commitments, cumulative sums, opened values, FRI commitments, final
    polynomial, Merkle siblings, opened rows): the native verifier and the oracle's restatement of `Verifier::verify_shard` return the
    SAME error string, including the `InvalidOpeningArgument:` wrapper of every PCS error (verifier.rs maps `pcs.verify` errors) and
    the place a wrong transcript surfaces (the Merkle opening at the re-sampled query index, as in the reference, whose QueryProof
    carries no index).  400 such mutants were compared by hand; 30 run here."""
    import copy
    PR, S, pk, vk, proof, words = hello
    local_only = {c.name for c in chips if c.local_only}
    order = sorted(proof["chip_ordering"], key=proof["chip_ordering"].get)

    def leaves(o, path=()):
        if isinstance(o, dict):
            for k, v in o.items():
                if k != "chip_ordering":
                    yield from leaves(v, path + (k,))
        elif isinstance(o, (list, tuple)):
            for i, v in enumerate(o):
                yield from leaves(v, path + (i,))
        elif isinstance(o, np.ndarray) and o.size:
            yield path

    # the `next` row of a local-only chip is not part of the proof (not serialised, not opened)
    paths = [p for p in leaves(proof) if not (p[0] == "opened_values" and p[-1] == "next" and p[2] in ("main", "preprocessed") and order[p[1]] in local_only)]
    rng = np.random.default_rng(77)
    seen = set()
    for _ in range(30):
        bad = copy.deepcopy(proof)
        path = paths[int(rng.integers(0, len(paths)))]
        o = bad
        for k in path[:-1]:
            o = o[k]
        a = np.array(o[path[-1]]).copy()
        flat = a.reshape(-1)
        j = int(rng.integers(0, flat.size))
        flat[j] = (int(flat[j]) + int(rng.integers(1, bf.P))) % bf.P
        o[path[-1]] = a
        eo, ev = oracle_verdict(PR, S, pk, vk, bad), native(pk, serialize(bad, pk.names))
        assert eo is not None and ev == eo, (path, eo, ev)
        seen.add(eo)
    assert len(seen) >= 2
