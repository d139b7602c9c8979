"""oracle/prover.py — TEST INFRASTRUCTURE (CPU oracle), not product code.

numpy restatement of the reference's shard prover and verifier on top of oracle/stark.py:

  generate_permutation_trace / eval_permutation_constraints   crates/stark/src/permutation.rs:27-69,75-148,157-272
  quotient_values + ProverConstraintFolder                     crates/stark/src/quotient.rs:18-165, folder.rs:14-123
  StarkMachine::setup (preprocessed commit)                    crates/stark/src/machine.rs:154-224
  CpuProver::{commit, open, prove}                             crates/stark/src/prover.rs:209-236,242-553,560-601
  Verifier::{verify_shard, verify_constraints, recompute_quotient}   crates/stark/src/verifier.rs:27-330
  debug_constraints (row-by-row constraint + lookup balance check)   crates/stark/src/debug.rs:24-105

The chips' constraint programs come from the declarative AIR description (air/dsl.py + air/chips.py, each
`eval_*` citing the reference chip it restates); this module only EVALUATES them, independently of the
generated CUDA code.  PARITY UNPINNED at the Plonky3 boundary (see oracle/stark.py); what this pins is that
oracle proofs and GPU proofs are accepted by this restated verifier and that single-word corruptions are not.
"""
import importlib

import numpy as np

import oracle as O
from oracle import stark as S

P = O.P
U = np.uint64
_dsl = importlib.import_module("zkvm-brainfuck_b200.air.dsl")


# ---- generic DAG evaluation -------------------------------------------------------------------------------
class BaseAlg:
    """values are uint64 arrays (or scalars) mod p."""
    const = staticmethod(lambda v: U(v))
    add = staticmethod(S.f_add)
    sub = staticmethod(S.f_sub)
    mul = staticmethod(S.f_mul)


class ExtAlg:
    """values are (..., 4) ext arrays."""
    const = staticmethod(lambda v: S.e_from_base(v))
    add = staticmethod(S.e_add)
    sub = staticmethod(S.e_sub)
    mul = staticmethod(S.e_mul)


def eval_exprs(roots, leaf, alg):
    """Evaluate DAG roots; leaf(node) supplies 'var' and 'sel' nodes."""
    vals = {}
    for n in _dsl.topo_order(list(roots)):
        if n.op == "const":
            vals[n._id] = alg.const(n.args[0])
        elif n.op in ("var", "sel"):
            vals[n._id] = leaf(n)
        else:
            a, b = vals[n.args[0]._id], vals[n.args[1]._id]
            vals[n._id] = getattr(alg, n.op)(a, b)
    return [vals[r._id] for r in roots]


def affine_base(aff, prep, main):
    """VirtualPairCol::apply on whole matrices -> uint64 vector."""
    n = main.shape[0]
    acc = np.full(n, aff.const, U)
    for (kind, idx), w in aff.terms:
        col = (prep if kind == "prep" else main)[:, idx].astype(U)
        acc = S.f_add(acc, S.f_mul(col, w))
    return acc


# ---- LogUp permutation trace (permutation.rs:27-148) ----------------------------------------------------------
def lookup_denominators(chip, prep, main, alpha, beta):
    """[(denominator (n,4), signed multiplicity (n,))] per lookup, sends first."""
    n = main.shape[0]
    out = []
    for lk, is_send in chip.lookups:
        betas = S.e_powers(beta, len(lk.values) + 1)
        den = np.broadcast_to(S.e_add(alpha, S.e_scale(betas[0], lk.kind)), (n, 4)).copy()
        for k, v in enumerate(lk.values):
            den = S.e_add(den, S.e_scale(np.broadcast_to(betas[k + 1], (n, 4)), affine_base(v, prep, main)))
        mult = affine_base(lk.multiplicity, prep, main)
        if not is_send:
            mult = S.f_sub(0, mult)
        out.append((den, mult))
    return out


def generate_permutation_trace(chip, prep, main, challenges):
    """-> (perm trace (n, perm_width, 4) ext, cumulative sum (4,))."""
    alpha, beta = challenges
    n = main.shape[0]
    dens = lookup_denominators(chip, prep, main, alpha, beta)
    W = chip.perm_width
    perm = np.zeros((n, W, 4), U)
    for j in range(W - 1):
        acc = np.zeros((n, 4), U)
        for den, mult in dens[j * chip.batch_size:(j + 1) * chip.batch_size]:
            acc = S.e_add(acc, S.e_scale(S.e_inv(den), mult))
        perm[:, j] = acc
    row_sums = np.zeros((n, 4), U)
    for j in range(W - 1):
        row_sums = S.e_add(row_sums, perm[:, j])
    csum = np.cumsum(row_sums.astype(object), axis=0) % P  # inclusive prefix sums (exact big-int)
    perm[:, W - 1] = csum.astype(U)
    return perm, perm[n - 1, W - 1].copy()


def flatten_to_base(perm):
    return np.ascontiguousarray(perm.reshape(perm.shape[0], -1), np.uint32)


# ---- constraint evaluation shared by the quotient, the verifier and debug_constraints ------------------------
def eval_chip_constraints(chip, leaf, alg, perm_local, perm_next, perm_challenges, cumulative_sum, selectors, to_ext):
    """Runs `Chip::eval` (chip.rs:216-229): the AIR's constraints then eval_permutation_constraints.
    Returns the list of constraint values in folding order; base constraints are lifted with `to_ext` so the
    caller folds one homogeneous list.  perm_* are lists of ext values (one per perm column)."""
    base_vals = eval_exprs(chip.constraints, leaf, alg)
    out = [to_ext(v) for v in base_vals]
    alpha, beta = perm_challenges
    E = ExtAlg
    lookups = chip.lookups
    W = chip.perm_width
    # values of every affine form over the local row, in the caller's algebra, lifted to ext
    def aff(a):
        acc = alg.const(a.const)
        for (kind, idx), w in a.terms:
            acc = alg.add(acc, alg.mul(leaf(_dsl.Expr.var(kind, 0, idx)), alg.const(w)))
        return to_ext(acc)
    for j in range(W - 1):
        chunk = lookups[j * chip.batch_size:(j + 1) * chip.batch_size]
        rlcs, mults = [], []
        for lk, is_send in chunk:
            betas = S.e_powers(beta, len(lk.values) + 1)
            rlc = S.e_add(alpha, S.e_scale(betas[0], lk.kind))
            for k, v in enumerate(lk.values):
                rlc = E.add(rlc, E.mul(betas[k + 1], aff(v)))
            rlcs.append(rlc)
            m = aff(lk.multiplicity)
            mults.append(m if is_send else S.e_neg(m))
        product = S.E_ONE
        numerator = S.E_ZERO
        for i, (m, rlc) in enumerate(zip(mults, rlcs)):
            product = E.mul(product, rlc)
            others = S.E_ONE
            for k, o in enumerate(rlcs):
                if k != i:
                    others = E.mul(others, o)
            numerator = E.add(numerator, E.mul(m, others))
        out.append(E.sub(E.mul(product, perm_local[j]), numerator))
    sum_local, sum_next = S.E_ZERO, S.E_ZERO
    for j in range(W - 1):
        sum_local = E.add(sum_local, perm_local[j])
        sum_next = E.add(sum_next, perm_next[j])
    phi_local, phi_next = perm_local[W - 1], perm_next[W - 1]
    out.append(E.mul(selectors["is_first_row"], E.sub(phi_local, sum_local)))
    out.append(E.mul(selectors["is_transition"], E.sub(E.sub(phi_next, phi_local), sum_next)))
    out.append(E.mul(selectors["is_last_row"], E.sub(phi_local, cumulative_sum)))
    return out


def fold(values, alpha):
    acc = S.E_ZERO
    for v in values:
        acc = S.e_add(S.e_mul(acc, alpha), v)
    return acc


# ---- quotient_values (quotient.rs:18-165) --------------------------------------------------------------------------
def quotient_values(chip, cumulative_sum, log_degree, prep_on_q, main_on_q, perm_on_q, perm_challenges, alpha):
    """*_on_q: LDE matrices on the quotient domain in NATURAL order (perm flattened to base columns).
    -> (2^(log_degree+lqd), 4) ext quotient values in natural order."""
    trace_domain = S.Domain(log_degree)
    qdom = trace_domain.create_disjoint_domain(1 << (log_degree + chip.log_quotient_degree))
    sels = trace_domain.selectors_on_coset(qdom)
    n = qdom.size
    next_step = 1 << chip.log_quotient_degree
    nxt = (np.arange(n) + next_step) % n
    mats = {"prep": prep_on_q, "main": main_on_q}

    def leaf(node):
        if node.op == "sel":
            return sels[node.args[0]]
        kind, offset, idx = node.args
        col = mats[kind][:, idx].astype(U)
        return col[nxt] if offset else col

    perm = np.asarray(perm_on_q, U).reshape(n, chip.perm_width, 4)
    perm_local = [perm[:, j] for j in range(chip.perm_width)]
    perm_next = [perm[nxt, j] for j in range(chip.perm_width)]
    ext_sels = {k: S.e_from_base(v) for k, v in sels.items()}
    vals = eval_chip_constraints(chip, leaf, BaseAlg, perm_local, perm_next, perm_challenges, cumulative_sum, ext_sels, S.e_from_base)
    vals = [np.broadcast_to(v, (n, 4)) for v in vals]
    acc = fold(vals, alpha)
    return S.e_scale(acc, sels["inv_zeroifier"])


# ---- debug_constraints (debug.rs:24-105): the trace satisfies every constraint on every row -------------------------
def debug_constraints(chip, prep, main, perm, perm_challenges, cumulative_sum):
    n = main.shape[0]
    nxt = (np.arange(n) + 1) % n
    first = np.zeros(n, U); first[0] = 1
    last = np.zeros(n, U); last[n - 1] = 1
    sels = {"is_first_row": first, "is_last_row": last, "is_transition": U(1) - last}
    mats = {"prep": prep, "main": main}

    def leaf(node):
        if node.op == "sel":
            return sels[node.args[0]]
        kind, offset, idx = node.args
        col = mats[kind][:, idx].astype(U)
        return col[nxt] if offset else col

    perm_local = [perm[:, j] for j in range(chip.perm_width)]
    perm_next = [perm[nxt, j] for j in range(chip.perm_width)]
    ext_sels = {k: S.e_from_base(v) for k, v in sels.items()}
    vals = eval_chip_constraints(chip, leaf, BaseAlg, perm_local, perm_next, perm_challenges, cumulative_sum, ext_sels, S.e_from_base)
    bad = []
    for i, v in enumerate(vals):
        v = np.broadcast_to(v, (n, 4))
        rows = np.nonzero(v.any(axis=1))[0]
        if rows.size:
            bad.append((i, rows[:5].tolist()))
    return bad


# ---- machine plumbing -----------------------------------------------------------------------------------------------
class ProvingKey:
    """StarkProvingKey (machine.rs:30-60): preprocessed commit, traces, prover data, chip ordering, local_only."""


def setup(chips, prep_traces):
    """StarkMachine::setup (machine.rs:154-224): traces sorted by (height desc, name)."""
    named = sorted(prep_traces.items(), key=lambda kv: (-kv[1].shape[0], kv[0]))
    pk = ProvingKey()
    pk.names = [k for k, _ in named]
    pk.traces = [v for _, v in named]
    pk.data = O.PcsData(pk.traces)
    pk.commit = pk.data.root.copy()
    pk.chip_ordering = {k: i for i, k in enumerate(pk.names)}
    by_name = {c.name: c for c in chips}
    pk.local_only = [by_name[k].local_only for k in pk.names]
    return pk


def observe_pk(pk, ch):
    """StarkProvingKey::observe_into (prover.rs:595-601): commitment then 7 zero elements."""
    ch.observe_digest(pk.commit)
    for _ in range(7):
        ch.observe(0)


def prove_shard(chips, pk, traces, ch, cfg=None, pow_witness=None):
    """CpuProver::commit + CpuProver::open.  traces: {chip name: main trace}.  `ch` must already have observed the pk.
    Returns the ShardProof as a dict (plus the intermediate data under '_debug')."""
    cfg = cfg or S.FriConfig()
    by_name = {c.name: c for c in chips}
    named = sorted(traces.items(), key=lambda kv: (-kv[1].shape[0], kv[0]))  # prover.rs:214
    names = [k for k, _ in named]
    mains = [v for _, v in named]
    ordered = [by_name[k] for k in names]
    main_data = O.PcsData(mains)
    log_degrees = [m.shape[0].bit_length() - 1 for m in mains]
    ch.observe_digest(main_data.root)
    perm_challenges = [ch.sample_ext(), ch.sample_ext()]
    preps = [pk.traces[pk.chip_ordering[c.name]] if c.name in pk.chip_ordering else None for c in ordered]
    perms, csums = [], []
    for c, prep, main in zip(ordered, preps, mains):
        pt, cs = generate_permutation_trace(c, prep if prep is not None else np.zeros((main.shape[0], 0), np.uint32), main, perm_challenges)
        perms.append(pt)
        csums.append(cs)
    perm_flat = [flatten_to_base(p) for p in perms]
    perm_data = O.PcsData(perm_flat)
    ch.observe_digest(perm_data.root)
    for cs in csums:
        ch.observe_ext(cs)
    alpha = ch.sample_ext()
    q_chunks, q_shifts = [], []
    for i, (c, ld) in enumerate(zip(ordered, log_degrees)):
        nat = lambda lde: np.asarray(lde)[S.bitrev_perm(ld + 1)]  # get_evaluations_on_domain: natural order
        prep_q = nat(pk.data.ldes[pk.chip_ordering[c.name]]) if c.name in pk.chip_ordering else np.zeros((2 << ld, 1), np.uint32)
        qv = quotient_values(c, csums[i], ld, prep_q, nat(main_data.ldes[i]), nat(perm_data.ldes[i]), perm_challenges, alpha)
        qdom = S.Domain(ld).create_disjoint_domain(1 << (ld + c.log_quotient_degree))
        flat = np.ascontiguousarray(qv, np.uint32)
        deg = 1 << c.log_quotient_degree
        for dom, chunk in zip(qdom.split_domains(deg), S.Domain.split_evals(deg, flat)):
            q_chunks.append(chunk)
            q_shifts.append(dom.shift)
    quot_data = O.PcsData(q_chunks, domain_shifts=q_shifts)
    ch.observe_digest(quot_data.root)
    zeta = ch.sample_ext()
    pts = lambda ld, both: [zeta, S.Domain(ld).next_point(zeta)] if both else [zeta]
    prep_points = [pts(t.shape[0].bit_length() - 1, not lo) for t, lo in zip(pk.traces, pk.local_only)]
    main_points = [pts(ld, not c.local_only) for c, ld in zip(ordered, log_degrees)]
    perm_points = [pts(ld, True) for ld in log_degrees]
    quot_points = [[zeta] for _ in q_chunks]
    rounds = [(pk.data, prep_points), (main_data, main_points), (perm_data, perm_points), (quot_data, quot_points)]
    opened, fri_proof = S.pcs_open(cfg, rounds, ch, pow_witness)
    proof = assemble_proof(names, ordered, pk, log_degrees, main_data.root, perm_data.root, quot_data.root, csums, opened, fri_proof)
    proof["_debug"] = dict(main_data=main_data, perm_data=perm_data, quot_data=quot_data, perms=perms, perm_challenges=perm_challenges,
                           alpha=alpha, zeta=zeta, mains=mains, preps=preps, ordered=ordered)
    return proof


def assemble_proof(names, ordered, pk, log_degrees, main_root, perm_root, quot_root, csums, opened, fri_proof):
    """ShardProof layout (types.rs:32-73, prover.rs:473-552)."""
    prep_v, main_v, perm_v, quot_v = opened
    chips_out = []
    qpos = 0
    for i, c in enumerate(ordered):
        def lv(vals, both):
            loc = vals[0]
            return dict(local=loc, next=vals[1] if both else np.zeros_like(loc))
        if c.name in pk.chip_ordering:
            k = pk.chip_ordering[c.name]
            prep = lv(prep_v[k], not pk.local_only[k])
        else:
            prep = dict(local=np.zeros((0, 4), U), next=np.zeros((0, 4), U))
        deg = 1 << c.log_quotient_degree
        quotient = [quot_v[qpos + d][0] for d in range(deg)]
        qpos += deg
        chips_out.append(dict(preprocessed=prep, main=lv(main_v[i], not c.local_only), permutation=lv(perm_v[i], True), quotient=quotient,
                              cumulative_sum=csums[i], log_degree=log_degrees[i]))
    return dict(commitment=dict(main=main_root.copy(), permutation=perm_root.copy(), quotient=quot_root.copy()),
                opened_values=chips_out, opening_proof=fri_proof, chip_ordering={k: i for i, k in enumerate(names)})


# ---- Verifier::verify_shard (verifier.rs:27-216) --------------------------------------------------------------------------
def verify_shard(chips, vk, proof, ch, cfg=None):
    """vk: dict(commit, chip_information=[(name, log_height, local_only)] in pk order).  `ch` has observed the vk.
    Returns None when the proof is accepted, else an error string."""
    cfg = cfg or S.FriConfig()
    by_name = {c.name: c for c in chips}
    order = sorted(proof["chip_ordering"], key=lambda k: proof["chip_ordering"][k])
    ordered = [by_name[k] for k in order]
    ov = proof["opened_values"]
    if len(ov) != len(ordered):
        return "ChipOpeningLengthMismatch"
    log_degrees = [v["log_degree"] for v in ov]
    com = proof["commitment"]
    ch.observe_digest(com["main"])
    perm_challenges = [ch.sample_ext(), ch.sample_ext()]
    ch.observe_digest(com["permutation"])
    for v in ov:
        ch.observe_ext(v["cumulative_sum"])
    alpha = ch.sample_ext()
    ch.observe_digest(com["quotient"])
    zeta = ch.sample_ext()

    def pts(dom, vals, both):
        return [(zeta, vals["local"]), (dom.next_point(zeta), vals["next"])] if both else [(zeta, vals["local"])]

    prep_round = []
    for name, log_h, _lo in vk["chip_information"]:
        i = proof["chip_ordering"][name]
        dom = S.Domain(log_h)
        prep_round.append((dom, pts(dom, ov[i]["preprocessed"], not ordered[i].local_only)))
    main_round = [(S.Domain(ld), pts(S.Domain(ld), v["main"], not c.local_only)) for c, ld, v in zip(ordered, log_degrees, ov)]
    perm_round = [(S.Domain(ld), pts(S.Domain(ld), v["permutation"], True)) for ld, v in zip(log_degrees, ov)]
    qc_domains = []
    quot_round = []
    for c, ld, v in zip(ordered, log_degrees, ov):
        deg = 1 << c.log_quotient_degree
        doms = S.Domain(ld).create_disjoint_domain(1 << (ld + c.log_quotient_degree)).split_domains(deg)
        qc_domains.append(doms)
        for d, q in zip(doms, v["quotient"]):
            quot_round.append((d, [(zeta, q)]))
    rounds = [(vk["commit"], prep_round), (com["main"], main_round), (com["permutation"], perm_round), (com["quotient"], quot_round)]
    err = S.pcs_verify(cfg, rounds, proof["opening_proof"], ch)
    if err:
        return "InvalidOpeningArgument:" + err
    for c, ld, doms, v in zip(ordered, log_degrees, qc_domains, ov):
        if not verify_constraints(c, v, S.Domain(ld), doms, zeta, alpha, perm_challenges):
            return "OodEvaluationMismatch:" + c.name
    total = S.E_ZERO
    for v in ov:
        total = S.e_add(total, v["cumulative_sum"])
    if not S.e_eq(total, S.E_ZERO):
        return "CumulativeSumsError"
    return None


def verify_constraints(chip, opening, trace_domain, qc_domains, zeta, alpha, perm_challenges):
    sels = trace_domain.selectors_at_point(zeta)
    # recompute_quotient (verifier.rs:294-329)
    quotient = S.E_ZERO
    for i, (dom, chunk) in enumerate(zip(qc_domains, opening["quotient"])):
        zp = S.E_ONE
        for j, other in enumerate(qc_domains):
            if j != i:
                zp = S.e_mul(zp, S.e_mul(other.zp_at_point(zeta), S.e_inv(other.zp_at_point(S.e_from_base(dom.first_point())))))
        val = S.E_ZERO
        for e_i in range(4):
            mono = np.zeros(4, U); mono[e_i] = 1
            val = S.e_add(val, S.e_mul(mono, chunk[e_i]))
        quotient = S.e_add(quotient, S.e_mul(zp, val))
    # eval_constraints with every variable an extension element (opened values)
    mats = {"prep": opening["preprocessed"], "main": opening["main"]}

    def leaf(node):
        if node.op == "sel":
            return sels[node.args[0]]
        kind, offset, idx = node.args
        return mats[kind]["next" if offset else "local"][idx]

    def unflatten(v):
        v = np.asarray(v, U).reshape(-1, 4, 4)
        out = []
        for chunk in v:
            acc = S.E_ZERO
            for e_i in range(4):
                mono = np.zeros(4, U); mono[e_i] = 1
                acc = S.e_add(acc, S.e_mul(mono, chunk[e_i]))
            out.append(acc)
        return out

    perm_local, perm_next = unflatten(opening["permutation"]["local"]), unflatten(opening["permutation"]["next"])
    vals = eval_chip_constraints(chip, leaf, ExtAlg, perm_local, perm_next, perm_challenges, opening["cumulative_sum"], sels, lambda v: v)
    folded = fold(vals, alpha)
    return S.e_eq(S.e_mul(folded, sels["inv_zeroifier"]), quotient)
