"""Independent pure-Python (big-int) restatement used to cross-check the C oracle.

Deliberately written in a different style from oracle/*.c: the Poseidon2 linear layers are
applied as explicit dense 16x16 matrices built from their mathematical definition
(SURVEY.md Appendix B.3), field inverses use pow(x, -1, p).
"""
import os
import re

P = 2130706433
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_rc():
    txt = open(os.path.join(ROOT, "oracle", "rc_16_30.h")).read()
    vals = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{8})u", txt)]
    assert len(vals) == 480
    return [vals[16 * r:16 * r + 16] for r in range(30)]


RC = load_rc()
M4 = [[2, 3, 1, 1], [1, 2, 3, 1], [1, 1, 2, 3], [3, 1, 1, 2]]
# external matrix: block-circulant circ(2*M4, M4, M4, M4)
ME = [[(2 if i // 4 == j // 4 else 1) * M4[i % 4][j % 4] for j in range(16)] for i in range(16)]


def _frac(num, log_den):
    return num * pow(1 << log_den, -1, P) % P


V = [-2 % P, 1, 2, _frac(1, 1), 3, 4, _frac(-1, 1), -3 % P, -4 % P, _frac(1, 8), _frac(1, 3), _frac(1, 24),
     _frac(-1, 8), _frac(-1, 3), _frac(-1, 4), _frac(-1, 24)]
MI = [[(1 + (V[i] if i == j else 0)) % P for j in range(16)] for i in range(16)]


def matvec(M, x):
    return [sum(M[i][j] * x[j] for j in range(16)) % P for i in range(16)]


def permute(state):
    # reference crates/stark/src/kb31_poseidon2.rs:35-50: rows 0..3 initial, rows 4..16 col 0
    # internal, rows 17..20 terminal
    s = [int(v) % P for v in state]
    s = matvec(ME, s)
    for r in range(4):
        s = [pow(s[i] + RC[r][i], 3, P) for i in range(16)]
        s = matvec(ME, s)
    for r in range(13):
        s[0] = pow(s[0] + RC[4 + r][0], 3, P)
        s = matvec(MI, s)
    for r in range(4):
        s = [pow(s[i] + RC[17 + r][i], 3, P) for i in range(16)]
        s = matvec(ME, s)
    return s


def sponge(values):
    st = [0] * 16
    vals = [int(v) for v in values]
    for off in range(0, len(vals), 8):
        chunk = vals[off:off + 8]
        st[:len(chunk)] = chunk
        st = permute(st)
    return st[:8]


def compress(l, r):
    return permute(list(l) + list(r))[:8]


def two_adic_generator(bits):
    return pow(3, (P - 1) >> bits, P)


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def interpolate_eval(evals, point):
    """Evaluate at `point` the unique poly of degree < n with p(w^i) = evals[i] (Lagrange, O(n^2))."""
    n = len(evals)
    w = two_adic_generator(n.bit_length() - 1)
    xs = [pow(w, i, P) for i in range(n)]
    acc = 0
    for i in range(n):
        num, den = 1, 1
        for j in range(n):
            if i != j:
                num = num * (point - xs[j]) % P
                den = den * (xs[i] - xs[j]) % P
        acc = (acc + evals[i] * num * pow(den, -1, P)) % P
    return acc


# ---- F_p^4 = F_p[X]/(X^4-3) as coefficient lists -------------------------------------------
def e_mul(a, b):
    t = [0] * 7
    for i in range(4):
        for j in range(4):
            t[i + j] += a[i] * b[j]
    return [(t[i] + 3 * (t[i + 4] if i < 3 else 0)) % P for i in range(4)]


def e_pow(a, e):
    r = [1, 0, 0, 0]
    while e:
        if e & 1:
            r = e_mul(r, a)
        a = e_mul(a, a)
        e >>= 1
    return r


def e_inv(a):
    return e_pow(a, P ** 4 - 2)
