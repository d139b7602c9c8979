"""oracle/stark.py — TEST INFRASTRUCTURE (CPU oracle), not product code.

numpy restatement of the generic STARK/PCS layer the reference drives from `CpuProver::open`
(reference crates/stark/src/prover.rs:242-553) and `Verifier::verify_shard`
(crates/stark/src/verifier.rs:27-216):

  * F_p^4 vector arithmetic (BinomialExtensionField<KoalaBear,4>, kb31_poseidon2.rs:21)
  * DuplexChallenger<Val, Perm, 16, 8>                     (kb31_poseidon2.rs:31)
  * TwoAdicMultiplicativeCoset domains and Lagrange selectors (quotient.rs:34-42, verifier.rs:236)
  * TwoAdicFriPcs::{open, verify}, FRI commit phase / queries / verifier (prover.rs:460-470,
    verifier.rs:186-189)

The algorithms are restated from the published Plonky3 v0.1.0 sources of the API era the reference
pins (git rev 93967fce, un-vendored, not available offline): SURVEY.md Appendix B.6-B.10.
PARITY UNPINNED: no reference golden vectors exist; what is pinned is self-consistency
(prover/verifier agreement, low-degree tests, algebraic identities) — see tests/test_oracle_stark.py.
One choice that cannot be confirmed offline is isolated in OBSERVE_OPENED_VALUES.
"""
import numpy as np

import oracle as O

P = O.P
GEN = 3
# Plonky3's TwoAdicFriPcs::open/verify of this API era write every opened value into the challenger
# before sampling the batching challenge alpha ("Write evaluations to challenger").
OBSERVE_OPENED_VALUES = True
# How a reduced opening joins the folded FRI vector when their lengths match: 0 = `folded[i] += ro[i]` (Plonky3 of the pinned API era
# as published), 1 = `folded[i] += beta^2 * ro[i]` (the later upstream rule).  Mirrors BFGPU_OPT_FRI_ROLLIN.
FRI_ROLLIN = 0
# Which valid proof-of-work witness grind() returns: 0 = the smallest, 1 = the largest (BFGPU_OPT_POW_ORDER; the reference's rayon
# find_any returns an arbitrary one, every valid witness verifies)
POW_ORDER = 0

U = np.uint64


# ---- base field helpers (vectorised, uint64) -----------------------------------------------------
def f_mul(a, b):
    return (np.asarray(a, U) * np.asarray(b, U)) % U(P)


def f_add(a, b):
    return (np.asarray(a, U) + np.asarray(b, U)) % U(P)


def f_sub(a, b):
    return (np.asarray(a, U) + U(P) - np.asarray(b, U)) % U(P)


def f_inv(a):
    return pow(int(a), P - 2, P)


def f_batch_inv(a):
    """elementwise inverse of a 1-D uint64 array (Montgomery's trick)."""
    a = np.asarray(a, U)
    n = a.shape[0]
    pref = np.ones(n + 1, U)
    for i in range(n):  # sequential prefix products; small inputs only
        pref[i + 1] = pref[i] * a[i] % U(P)
    inv = U(f_inv(pref[n]))
    out = np.zeros(n, U)
    for i in range(n - 1, -1, -1):
        out[i] = inv * pref[i] % U(P)
        inv = inv * a[i] % U(P)
    return out


def f_pow_vec(a, e):
    """a ** e elementwise (e a python int)."""
    a = np.asarray(a, U)
    r = np.ones_like(a)
    while e:
        if e & 1:
            r = r * a % U(P)
        a = a * a % U(P)
        e >>= 1
    return r


def two_adic_generator(bits):
    return pow(GEN, (P - 1) >> bits, P)


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def bitrev_perm(bits):
    n = 1 << bits
    idx = np.arange(n, dtype=np.int64)
    r = np.zeros(n, np.int64)
    for i in range(bits):
        r |= ((idx >> i) & 1) << (bits - 1 - i)
    return r


def powers(base, n, start=1):
    """[start * base^i for i < n] as uint64 (vectorised doubling)."""
    out = np.empty(n, U)
    if n == 0:
        return out
    out[0] = start % P
    filled = 1
    step = base % P
    while filled < n:
        m = min(filled, n - filled)
        out[filled:filled + m] = out[:m] * U(step) % U(P)
        step = step * step % P
        filled += m
    return out


# ---- F_p^4 = F_p[X]/(X^4 - 3): arrays of shape (..., 4) uint64 ----------------------------------
def e_from_base(a):
    a = np.asarray(a, U)
    out = np.zeros(a.shape + (4,), U)
    out[..., 0] = a
    return out


def e_const(c):
    return np.array([int(v) % P for v in c], U)


E_ZERO = np.zeros(4, U)
E_ONE = np.array([1, 0, 0, 0], U)


def e_add(a, b):
    return (np.asarray(a, U) + np.asarray(b, U)) % U(P)


def e_sub(a, b):
    return (np.asarray(a, U) + U(P) - np.asarray(b, U)) % U(P)


def e_neg(a):
    return (U(P) - np.asarray(a, U)) % U(P)


def e_scale(a, s):
    """ext * base (s broadcast over the last axis)."""
    return np.asarray(a, U) * np.asarray(s, U)[..., None] % U(P)


def e_mul(a, b):
    a, b = np.broadcast_arrays(np.asarray(a, U), np.asarray(b, U))
    t = [np.zeros(a.shape[:-1], U) for _ in range(7)]
    for i in range(4):
        for j in range(4):
            t[i + j] = (t[i + j] + a[..., i] * b[..., j] % U(P)) % U(P)
    out = np.empty(a.shape, U)
    for i in range(4):
        out[..., i] = (t[i] + (U(3) * t[i + 4] if i < 3 else U(0))) % U(P)
    return out


def e_inv(a):
    """elementwise inverse (vectorised): a = A + B X over K = F_p[Y]/(Y^2-3), Y = X^2."""
    a = np.asarray(a, U)
    a0, a1, a2, a3 = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    m = f_mul
    A2_0 = f_add(m(a0, a0), m(3, m(a2, a2)))
    A2_1 = f_mul(2, m(a0, a2))
    B2_0 = f_add(m(a1, a1), m(3, m(a3, a3)))
    B2_1 = f_mul(2, m(a1, a3))
    n0 = f_sub(A2_0, m(3, B2_1))
    n1 = f_sub(A2_1, B2_0)
    d = f_sub(m(n0, n0), m(3, m(n1, n1)))
    di = f_pow_vec(d, P - 2)
    m0, m1 = m(n0, di), f_sub(0, m(n1, di))
    out = np.empty(a.shape, U)
    out[..., 0] = f_add(m(a0, m0), m(3, m(a2, m1)))
    out[..., 2] = f_add(m(a0, m1), m(a2, m0))
    out[..., 1] = f_sub(0, f_add(m(a1, m0), m(3, m(a3, m1))))
    out[..., 3] = f_sub(0, f_add(m(a1, m1), m(a3, m0)))
    return out


def e_pow(a, e):
    r = np.broadcast_to(E_ONE, np.asarray(a).shape).copy()
    a = np.asarray(a, U)
    while e:
        if e & 1:
            r = e_mul(r, a)
        a = e_mul(a, a)
        e >>= 1
    return r


def e_powers(x, n):
    """[x^0 .. x^(n-1)] shape (n, 4)."""
    out = np.zeros((n, 4), U)
    if n:
        out[0] = E_ONE
    filled = 1
    step = np.asarray(x, U)
    while filled < n:
        m = min(filled, n - filled)
        out[filled:filled + m] = e_mul(out[:m], step)
        step = e_mul(step, step)
        filled += m
    return out


def e_sum(a, axis=0):
    return np.sum(np.asarray(a, U) % U(P), axis=axis) % U(P)  # < 2^31 * 2^33 terms is safe in uint64


def e_eq(a, b):
    return bool(np.all(np.asarray(a, U) % U(P) == np.asarray(b, U) % U(P)))


# ---- DuplexChallenger<Val, Perm, 16, 8> ------------------------------------------------------------
class Challenger:
    def __init__(self):
        self.state = np.zeros(16, np.uint32)
        self.inp = []
        self.out = []

    def clone(self):
        c = Challenger()
        c.state = self.state.copy()
        c.inp = list(self.inp)
        c.out = list(self.out)
        return c

    def _duplex(self):
        assert len(self.inp) <= 8
        for i, v in enumerate(self.inp):
            self.state[i] = v
        self.inp = []
        self.state = O.permute(self.state)
        self.out = [int(v) for v in self.state[:8]]

    def observe(self, v):
        self.out = []
        self.inp.append(int(v) % P)
        if len(self.inp) == 8:
            self._duplex()

    def observe_slice(self, vs):
        for v in np.asarray(vs).ravel():
            self.observe(int(v))

    observe_digest = observe_slice
    observe_ext = observe_slice

    def sample(self):
        if self.inp or not self.out:
            self._duplex()
        return self.out.pop()

    def sample_ext(self):
        return np.array([self.sample() for _ in range(4)], U)

    def sample_bits(self, bits):
        return self.sample() & ((1 << bits) - 1)

    def check_witness(self, bits, w):
        self.observe(w)
        return self.sample_bits(bits) == 0

    def grind(self, bits):
        """smallest witness (the reference searches in parallel with find_any: any valid one is accepted)."""
        w, step = (P - 1, -1) if POW_ORDER == 1 else (0, 1)
        while True:
            if self.clone().check_witness(bits, w):
                assert self.check_witness(bits, w)
                return w
            w += step


# ---- two-adic coset domains ---------------------------------------------------------------------------
class Domain:
    """TwoAdicMultiplicativeCoset { log_n, shift }."""

    def __init__(self, log_n, shift=1):
        self.log_n, self.shift = log_n, shift % P

    @property
    def size(self):
        return 1 << self.log_n

    def gen(self):
        return two_adic_generator(self.log_n)

    def first_point(self):
        return self.shift

    def next_point(self, x):
        return e_scale(x, self.gen())

    def create_disjoint_domain(self, min_size):
        return Domain((min_size - 1).bit_length(), self.shift * GEN % P)

    def split_domains(self, num_chunks):
        log_chunks = num_chunks.bit_length() - 1
        return [Domain(self.log_n - log_chunks, self.shift * pow(self.gen(), i, P) % P) for i in range(num_chunks)]

    @staticmethod
    def split_evals(num_chunks, evals):
        return [np.ascontiguousarray(evals[i::num_chunks]) for i in range(num_chunks)]

    def zp_at_point(self, x):
        """Z_D(x) = (x/shift)^n - 1 for ext x."""
        u = e_scale(x, f_inv(self.shift))
        return e_sub(e_pow(u, self.size), E_ONE)

    def selectors_at_point(self, x):
        u = e_scale(x, f_inv(self.shift))
        z_h = e_sub(e_pow(u, self.size), E_ONE)
        ginv = f_inv(self.gen())
        return dict(is_first_row=e_mul(z_h, e_inv(e_sub(u, E_ONE))),
                    is_last_row=e_mul(z_h, e_inv(e_sub(u, e_from_base(ginv)))),
                    is_transition=e_sub(u, e_from_base(ginv)),
                    inv_zeroifier=e_inv(z_h))

    def selectors_on_coset(self, coset):
        """self = trace domain (shift 1), coset = quotient domain; natural order vectors (uint64)."""
        assert self.shift == 1 and coset.shift != 1 and coset.log_n >= self.log_n
        rate_bits = coset.log_n - self.log_n
        s_pow_n = pow(coset.shift, self.size, P)
        evals = f_sub(f_mul(powers(two_adic_generator(rate_bits), 1 << rate_bits), s_pow_n), 1)  # Z_H on the coset, period 2^rate_bits
        xs = powers(coset.gen(), coset.size, coset.shift)
        zh = np.tile(evals, coset.size >> rate_bits)
        ginv = f_inv(self.gen())

        def single(pt):
            return f_mul(zh, f_pow_vec(f_sub(xs, pt), P - 2))

        return dict(is_first_row=single(1), is_last_row=single(ginv), is_transition=f_sub(xs, ginv),
                    inv_zeroifier=f_pow_vec(zh, P - 2))


# ---- MMCS wrappers -----------------------------------------------------------------------------------
class ExtTree:
    """ExtensionMmcs over ValMmcs: commit to an (n x w) ext matrix as its (n x 4w) base flattening."""

    def __init__(self, ext_mat):
        m = np.asarray(ext_mat, U)
        self.flat = np.ascontiguousarray(m.reshape(m.shape[0], -1), np.uint32)
        self.tree = O.Tree([self.flat])
        self.root = self.tree.root

    def open(self, index):
        rows, sib = self.tree.open_batch(index)
        return rows[0].astype(U).reshape(-1, 4), sib


# ---- interpolation ---------------------------------------------------------------------------------------
def _dot_cols(ev, vec4):
    """sum_r ev[r, c] * vec4[r]  -> (cols, 4), exact mod p (chunked to stay inside uint64)."""
    ev = np.asarray(ev, U)
    out = np.zeros((ev.shape[1], 4), U)
    for k in range(4):
        v = vec4[:, k]
        acc = np.zeros(ev.shape[1], U)
        for s in range(0, ev.shape[0], 4096):
            blk = ev[s:s + 4096] * v[s:s + 4096, None] % U(P)
            acc = (acc + blk.sum(axis=0) % U(P)) % U(P)
        out[:, k] = acc
    return out


def interpolate_coset(evals, shift, point):
    """evals: (n, w) base values of w polynomials on shift*H in NATURAL order; -> p_k(point), shape (w, 4)."""
    n = evals.shape[0]
    log_n = n.bit_length() - 1
    g = powers(two_adic_generator(log_n), n)
    xs = f_mul(g, shift)
    diff_inv = e_inv(e_sub(point, e_from_base(xs)))
    col_scale = e_scale(diff_inv, g)
    acc = _dot_cols(evals, col_scale)
    zer = e_sub(e_pow(point, n), e_from_base(pow(shift, n, P)))
    denom = pow(shift, n - 1, P) * n % P
    return e_mul(acc, e_scale(zer, f_inv(denom)))


# ---- FRI -----------------------------------------------------------------------------------------------
class FriConfig:
    def __init__(self, log_blowup=1, num_queries=84, pow_bits=16):
        self.log_blowup, self.num_queries, self.pow_bits = log_blowup, num_queries, pow_bits


def fold_matrix(beta, lo, hi):
    """TwoAdicFriGenericConfig::fold_matrix on the (n/2 x 2) matrix [lo | hi] (bit-reversed order)."""
    h = lo.shape[0]
    log_h = h.bit_length() - 1
    g_inv = f_inv(two_adic_generator(log_h + 1))
    half = f_inv(2)
    half_beta = e_scale(beta, half)
    pw = powers(g_inv, h)[bitrev_perm(log_h)]           # g_inv^{bitrev(i)}
    power = e_scale(np.broadcast_to(half_beta, (h, 4)), pw)
    a = e_add(e_from_base(np.full(h, half, U)), power)
    b = e_sub(e_from_base(np.full(h, half, U)), power)
    return e_add(e_mul(a, lo), e_mul(b, hi))


def fold_row(index, log_height, beta, e0, e1):
    """verifier-side fold of one pair (TwoAdicFriGenericConfig::fold_row)."""
    x0 = pow(two_adic_generator(log_height + 1), bitrev(index, log_height), P)
    x1 = (P - x0) % P
    # e0 + (beta - x0) (e1 - e0) / (x1 - x0)
    t = e_scale(e_mul(e_sub(beta, e_from_base(x0)), e_sub(e1, e0)), f_inv((x1 - x0) % P))
    return e_add(e0, t)


def fri_commit_phase(cfg, inputs, ch):
    inputs = list(inputs)
    folded = inputs.pop(0)
    commits, trees = [], []
    while folded.shape[0] > (1 << cfg.log_blowup):
        leaves = folded.reshape(-1, 2, 4)  # row i = (folded[2i], folded[2i+1])
        t = ExtTree(leaves)
        ch.observe_digest(t.root)
        beta = ch.sample_ext()
        folded = fold_matrix(beta, leaves[:, 0], leaves[:, 1])
        commits.append(t.root.copy())
        trees.append(t)
        if inputs and inputs[0].shape[0] == folded.shape[0]:
            ro = inputs.pop(0)
            if FRI_ROLLIN == 1:
                ro = e_mul(np.broadcast_to(e_mul(beta, beta), ro.shape), ro)
            folded = e_add(folded, ro)
    assert folded.shape[0] == (1 << cfg.log_blowup) and not inputs
    final_poly = folded[0]
    for x in folded:
        assert e_eq(x, final_poly), "FRI input is not low-degree"
    ch.observe_ext(final_poly)
    return commits, trees, final_poly


def fri_prove(cfg, inputs, ch, open_input, pow_witness=None):
    log_max_height = inputs[0].shape[0].bit_length() - 1
    commits, trees, final_poly = fri_commit_phase(cfg, inputs, ch)
    if pow_witness is None:
        pow_witness = ch.grind(cfg.pow_bits)
    else:
        assert ch.check_witness(cfg.pow_bits, pow_witness)
    queries = []
    for _ in range(cfg.num_queries):
        index = ch.sample_bits(log_max_height)
        steps = []
        for i, t in enumerate(trees):
            idx_i = index >> i
            row, sib = t.open(idx_i >> 1)
            steps.append(dict(sibling_value=row[(idx_i ^ 1) % 2], opening_proof=sib))
        queries.append(dict(index=index, input_proof=open_input(index), commit_phase_openings=steps))
    return dict(commit_phase_commits=commits, query_proofs=queries, final_poly=final_poly, pow_witness=pow_witness)


def fri_verify(cfg, proof, ch, open_input):
    betas = []
    for c in proof["commit_phase_commits"]:
        ch.observe_digest(c)
        betas.append(ch.sample_ext())
    ch.observe_ext(proof["final_poly"])
    if len(proof["query_proofs"]) != cfg.num_queries:
        return "InvalidProofShape"
    if not ch.check_witness(cfg.pow_bits, proof["pow_witness"]):
        return "InvalidPowWitness"
    log_max_height = len(proof["commit_phase_commits"]) + cfg.log_blowup
    for qp in proof["query_proofs"]:
        index = ch.sample_bits(log_max_height)
        ro = open_input(index, qp["input_proof"])
        if isinstance(ro, str):
            return ro
        folded = E_ZERO.copy()
        ro = list(ro)
        idx = index
        for k, (beta, comm, step) in enumerate(zip(betas, proof["commit_phase_commits"], qp["commit_phase_openings"])):
            log_folded_height = log_max_height - 1 - k
            if ro and ro[0][0] == log_folded_height + 1:
                r = ro.pop(0)[1]
                if FRI_ROLLIN == 1 and k > 0:
                    r = e_mul(e_mul(betas[k - 1], betas[k - 1]), r)
                folded = e_add(folded, r)
            evals = [folded, folded]
            evals[(idx ^ 1) % 2] = np.asarray(step["sibling_value"], U)
            flat = np.concatenate(evals).astype(np.uint32)
            if not O.verify_batch(comm, [(1 << log_folded_height, 8)], idx >> 1, [flat], step["opening_proof"]):
                return "CommitPhaseMmcsError"
            idx >>= 1
            folded = fold_row(idx, log_folded_height, beta, evals[0], evals[1])
        if ro:
            return "InvalidProofShape"
        if not e_eq(folded, proof["final_poly"]):
            return "FinalPolyMismatch"
    return None


# ---- TwoAdicFriPcs::open / verify -------------------------------------------------------------------------
def pcs_open(cfg, rounds, ch, pow_witness=None):
    """rounds: list of (PcsData, points) with points[i] = list of ext points for matrix i.
    Returns (opened_values[round][mat][point] -> (width, 4), fri_proof)."""
    log_global_max = max(int(l.shape[0]).bit_length() - 1 for d, _ in rounds for l in d.ldes)
    opened = []
    for data, points in rounds:
        rv = []
        for lde, pts in zip(data.ldes, points):
            h = lde.shape[0] >> cfg.log_blowup
            log_h = h.bit_length() - 1
            low = np.asarray(lde[:h])[bitrev_perm(log_h)]  # natural order evaluations on GEN*H
            mv = []
            for z in pts:
                ys = interpolate_coset(low, GEN, z)
                if OBSERVE_OPENED_VALUES:
                    ch.observe_ext(ys)
                mv.append(ys)
            rv.append(mv)
        opened.append(rv)
    alpha = ch.sample_ext()
    num_reduced = {}
    reduced = {}
    for (data, points), rv in zip(rounds, opened):
        for lde, pts, mv in zip(data.ldes, points, rv):
            hgt = lde.shape[0]
            log_height = hgt.bit_length() - 1
            if log_height not in reduced:
                reduced[log_height] = np.zeros((hgt, 4), U)
                num_reduced[log_height] = 0
            xs = f_mul(powers(two_adic_generator(log_height), hgt), GEN)[bitrev_perm(log_height)]  # x_r, bit-reversed
            apow = e_powers(alpha, lde.shape[1])
            row_red = np.zeros((hgt, 4), U)
            ld = np.asarray(lde, U)
            for k in range(lde.shape[1]):
                row_red = e_add(row_red, e_scale(np.broadcast_to(apow[k], (hgt, 4)), ld[:, k]))
            for z, ys in zip(pts, mv):
                off = e_pow(alpha, num_reduced[log_height])
                y_red = e_sum(e_mul(apow, ys))
                inv_den = e_inv(e_sub(z, e_from_base(xs)))        # 1/(z - x)
                term = e_mul(e_mul(e_sub(y_red, row_red), inv_den), off)
                reduced[log_height] = e_add(reduced[log_height], term)
                num_reduced[log_height] += lde.shape[1]
    fri_input = [reduced[k] for k in sorted(reduced, reverse=True)]

    def open_input(index):
        out = []
        for data, _ in rounds:
            log_max_h = max(int(l.shape[0]).bit_length() - 1 for l in data.ldes)
            ridx = index >> (log_global_max - log_max_h)
            rows, sib = data.tree.open_batch(ridx)
            out.append(dict(opened_values=rows, opening_proof=sib))
        return out

    proof = fri_prove(cfg, fri_input, ch, open_input, pow_witness)
    return opened, proof


def pcs_verify(cfg, rounds, proof, ch):
    """rounds: list of (commitment, [(Domain, [(point, values (w,4))...])...]).  Returns None if accepted."""
    for _, mats in rounds:
        for _, pts in mats:
            for _, vals in pts:
                if OBSERVE_OPENED_VALUES:
                    ch.observe_ext(vals)
    alpha = ch.sample_ext()
    log_global_max = len(proof["commit_phase_commits"]) + cfg.log_blowup

    def open_input(index, input_proof):
        ro = {}
        apow = {}
        for (commit, mats), bo in zip(rounds, input_proof):
            heights = [d.size << cfg.log_blowup for d, _ in mats]
            dims = [(h, len(pts[0][1])) for h, (_, pts) in zip(heights, mats)]
            log_batch_max = max(heights).bit_length() - 1
            ridx = index >> (log_global_max - log_batch_max)
            if not O.verify_batch(commit, dims, ridx, bo["opened_values"], bo["opening_proof"]):
                return "InputMmcsError"
            for (dom, pts), row, hgt in zip(mats, bo["opened_values"], heights):
                log_height = hgt.bit_length() - 1
                rr = ridx >> (log_batch_max - log_height)
                x = GEN * pow(two_adic_generator(log_height), bitrev(rr, log_height), P) % P
                ro.setdefault(log_height, E_ZERO.copy())
                apow.setdefault(log_height, E_ONE.copy())
                for z, vals in pts:
                    inv = e_inv(e_sub(e_from_base(x), z))  # 1/(x - z)
                    for p_at_x, p_at_z in zip(row, vals):
                        q = e_mul(e_sub(e_from_base(int(p_at_x)), p_at_z), inv)
                        ro[log_height] = e_add(ro[log_height], e_mul(apow[log_height], q))
                        apow[log_height] = e_mul(apow[log_height], alpha)
        return [(k, ro[k]) for k in sorted(ro, reverse=True)]

    return fri_verify(cfg, proof, ch, open_input)
