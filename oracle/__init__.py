"""oracle — TEST INFRASTRUCTURE, not product code.

ctypes loader for the CPU restatement of the reference's proving hot path
(`oracle/*.c` → `oracle/libbforacle.so`).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this package.

PARITY UNPINNED: the reference (Rust + un-vendored Plonky3 @93967fce) holds no golden vectors for
this path and cannot be built offline; the oracle is pinned against independent restatements
(pure-Python big-int Poseidon2, O(n^2) DFT, open→verify round trips) — see DESIGN.md.
"""
import ctypes as C
import os
import subprocess
import numpy as np

P = 2130706433
_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class Mat(C.Structure):
    _fields_ = [("data", u32p), ("rows", C.c_uint64), ("cols", C.c_uint64)]


def build(force=False):
    """(Re)build liboracle with the committed Makefile."""
    so = os.path.join(_DIR, "libbforacle.so")
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _DIR, "libbforacle.so"], stdout=subprocess.DEVNULL)
    return so


_AIR = None


def air_lib():
    """oracle/libbfoair.so: CPU arm of the LogUp-trace and quotient phases (oracle/fast_air.cpp)."""
    global _AIR
    if _AIR is None:
        so = os.path.join(_DIR, "libbfoair.so")
        deps = [os.path.join(_DIR, "fast_air.cpp"), os.path.join(_DIR, "fast_air_packed.cpp"), os.path.join(_DIR, "packed_kb.h"),
                os.path.join(_DIR, "..", "zkvm-brainfuck_b200", "csrc", "gen_air.cuh"),
                os.path.join(_DIR, "..", "zkvm-brainfuck_b200", "csrc", "kb31.cuh")]
        if not os.path.exists(so) or any(os.path.exists(d) and os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["make", "-C", _DIR, "libbfoair.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.bfo_air_chip_info.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.bfo_air_perm_trace.argtypes = [C.c_int, u32p, u32p, C.c_uint64, u32p, u32p, u32p, u32p]
        L.bfo_air_quotient.argtypes = [C.c_int, u32p, u32p, u32p, C.c_uint64, u32p, u32p, u32p, u32p, u32p]
        L.bfo_air_perm_trace_packed.argtypes = L.bfo_air_perm_trace.argtypes
        L.bfo_air_quotient_packed.argtypes = L.bfo_air_quotient.argtypes
        L.bfo_air_packed_available.restype = C.c_int
        L.bfo_open_eval.argtypes = [u32p, C.c_uint64, C.c_uint32, u32p, u32p]
        L.bfo_open_reduce_add.argtypes = [u32p, C.c_uint64, C.c_uint32, C.c_uint32, u32p, u32p, u32p, C.c_uint64, u32p]
        L.bfo_open_eval_packed.argtypes = L.bfo_open_eval.argtypes
        L.bfo_open_reduce_add_packed.argtypes = L.bfo_open_reduce_add.argtypes
        _AIR = L
    return _AIR


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.bfo_poseidon2_permute.argtypes = [u32p]
        L.bfo_poseidon2_permute_many.argtypes = [u32p, C.c_uint64]
        L.bfo_poseidon2_constants.argtypes = [u32p, u32p, u32p]
        L.bfo_sponge_hash.argtypes = [u32p, C.c_uint64, u32p]
        L.bfo_compress.argtypes = [u32p, u32p, u32p]
        L.bfo_mmcs_commit.argtypes = [C.POINTER(Mat), C.c_int, u32p]
        L.bfo_mmcs_commit.restype = C.c_void_p
        L.bfo_tree_free.argtypes = [C.c_void_p]
        L.bfo_tree_num_layers.argtypes = [C.c_void_p]
        L.bfo_tree_layer_len.argtypes = [C.c_void_p, C.c_int]
        L.bfo_tree_layer_len.restype = C.c_uint64
        L.bfo_tree_layer.argtypes = [C.c_void_p, C.c_int]
        L.bfo_tree_layer.restype = u32p
        L.bfo_mmcs_open_batch.argtypes = [C.c_void_p, C.POINTER(Mat), C.c_int, C.c_uint64, u32p, u32p]
        L.bfo_mmcs_verify_batch.argtypes = [u32p, u64p, u64p, C.c_int, C.c_uint64, u32p, u32p]
        L.bfo_mmcs_verify_batch.restype = C.c_int
        for f in (L.bfo_coset_lde_naive, L.bfo_coset_lde_batch, L.bfo_coset_lde_batch_bitrev):
            f.argtypes = [u32p, C.c_uint64, C.c_uint64, C.c_uint, C.c_uint32, u32p]
        L.bfo_dft_batch.argtypes = [u32p, C.c_uint64, C.c_uint64]
        L.bfo_idft_batch.argtypes = [u32p, C.c_uint64, C.c_uint64]
        L.bfo_pcs_commit.argtypes = [C.POINTER(Mat), u32p, C.c_int, C.c_uint, u32p]
        L.bfo_pcs_commit.restype = C.c_void_p
        L.bfo_pcs_data_free.argtypes = [C.c_void_p]
        L.bfo_pcs_num_mats.argtypes = [C.c_void_p]
        L.bfo_pcs_lde.argtypes = [C.c_void_p, C.c_int, u64p, u64p]
        L.bfo_pcs_lde.restype = u32p
        L.bfo_pcs_tree.argtypes = [C.c_void_p]
        L.bfo_pcs_tree.restype = C.c_void_p
        L.bfo_fast_available.restype = C.c_int
        L.bfo_fast_pcs_commit.argtypes = [u32p, C.c_uint64, C.c_uint64, u32p, u32p, C.POINTER(C.c_double)]
        L.bfo_fast_pcs_commit.restype = C.c_int
        L.bfo_fast_permute_many.argtypes = [u32p, C.c_uint64]
        L.bfo_fast_permute_many.restype = C.c_int
        L.bfo_fast_fri_commit_phase.argtypes = [C.POINTER(u32p), u32p, C.c_int, u32p, C.c_int, u32p, u32p, C.POINTER(C.c_double)]
        L.bfo_fast_fri_commit_phase.restype = C.c_int
        L.bfo_set_threads.argtypes = [C.c_int]
        L.bfo_get_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(u32p)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def set_threads(n):
    lib().bfo_set_threads(int(n))


def get_threads():
    return int(lib().bfo_get_threads())


# ---- Poseidon2 ------------------------------------------------------------------------------
def poseidon2_constants():
    ei = np.zeros((4, 16), np.uint32)
    it = np.zeros(13, np.uint32)
    et = np.zeros((4, 16), np.uint32)
    lib().bfo_poseidon2_constants(_p(ei), _p(it), _p(et))
    return ei, it, et


def permute(state):
    s = _u32(state).copy()
    assert s.shape == (16,)
    lib().bfo_poseidon2_permute(_p(s))
    return s


def permute_many(states):
    s = _u32(states).copy()
    assert s.ndim == 2 and s.shape[1] == 16
    lib().bfo_poseidon2_permute_many(_p(s), s.shape[0])
    return s


def sponge_hash(values):
    v = _u32(values).ravel()
    out = np.zeros(8, np.uint32)
    lib().bfo_sponge_hash(_p(v) if v.size else None, v.size, _p(out))
    return out


def compress(left, right):
    out = np.zeros(8, np.uint32)
    l, r = _u32(left), _u32(right)
    lib().bfo_compress(_p(l), _p(r), _p(out))
    return out


# ---- Merkle ---------------------------------------------------------------------------------
def _mats(mats):
    keep = [_u32(m) for m in mats]
    arr = (Mat * len(keep))()
    for i, m in enumerate(keep):
        assert m.ndim == 2
        arr[i] = Mat(_p(m), m.shape[0], m.shape[1])
    return arr, keep


class Tree:
    """MerkleTreeMmcs prover data (digest layers) for a list of row-major matrices."""

    def __init__(self, mats):
        self._arr, self.mats = _mats(mats)
        self.root = np.zeros(8, np.uint32)
        self._h = lib().bfo_mmcs_commit(self._arr, len(self.mats), _p(self.root))
        self._owned = True

    @classmethod
    def _borrow(cls, handle, mats):
        t = cls.__new__(cls)
        t._arr, t.mats = _mats(mats)
        t._h = handle
        t._owned = False
        t.root = t.layers()[-1][0].copy()
        return t

    def layers(self):
        L = lib()
        out = []
        for l in range(L.bfo_tree_num_layers(self._h)):
            n = L.bfo_tree_layer_len(self._h, l)
            ptr = L.bfo_tree_layer(self._h, l)
            out.append(np.ctypeslib.as_array(ptr, shape=(n, 8)).copy())
        return out

    def open_batch(self, index):
        total = sum(m.shape[1] for m in self.mats)
        nl = lib().bfo_tree_num_layers(self._h) - 1
        rows = np.zeros(max(total, 1), np.uint32)
        sib = np.zeros((max(nl, 1), 8), np.uint32)
        lib().bfo_mmcs_open_batch(self._h, self._arr, len(self.mats), int(index), _p(rows), _p(sib))
        out, o = [], 0
        for m in self.mats:
            out.append(rows[o:o + m.shape[1]].copy())
            o += m.shape[1]
        return out, sib[:nl].copy()

    def __del__(self):
        if getattr(self, "_owned", False) and self._h:
            lib().bfo_tree_free(self._h)
            self._h = None


def verify_batch(root, dims, index, opened_rows, siblings):
    rows = np.array([d[0] for d in dims], np.uint64)
    cols = np.array([d[1] for d in dims], np.uint64)
    flat = _u32(np.concatenate([_u32(r).ravel() for r in opened_rows]) if opened_rows else np.zeros(0, np.uint32))
    sib = _u32(siblings).reshape(-1)
    root = _u32(root)
    rc = lib().bfo_mmcs_verify_batch(_p(root), rows.ctypes.data_as(u64p), cols.ctypes.data_as(u64p), len(dims), int(index),
                                     _p(flat), _p(sib) if sib.size else None)
    return rc == 0


# ---- DFT ------------------------------------------------------------------------------------
def _lde(fn, mat, added_bits, shift):
    m = _u32(mat)
    rows, cols = m.shape
    out = np.zeros((rows << added_bits, cols), np.uint32)
    fn(_p(m), rows, cols, added_bits, int(shift), _p(out))
    return out


def coset_lde_naive(mat, added_bits=1, shift=3):
    return _lde(lib().bfo_coset_lde_naive, mat, added_bits, shift)


def coset_lde_batch(mat, added_bits=1, shift=3):
    return _lde(lib().bfo_coset_lde_batch, mat, added_bits, shift)


def coset_lde_batch_bitrev(mat, added_bits=1, shift=3):
    return _lde(lib().bfo_coset_lde_batch_bitrev, mat, added_bits, shift)


def dft_batch(mat):
    m = _u32(mat).copy()
    lib().bfo_dft_batch(_p(m), m.shape[0], m.shape[1])
    return m


def idft_batch(mat):
    m = _u32(mat).copy()
    lib().bfo_idft_batch(_p(m), m.shape[0], m.shape[1])
    return m


# ---- PCS commit -----------------------------------------------------------------------------
class PcsData:
    """TwoAdicFriPcs::commit result: bit-reversed LDEs + Merkle tree."""

    def __init__(self, evals, domain_shifts=None, log_blowup=1):
        self._arr, self.evals = _mats(evals)
        n = len(self.evals)
        sh = _u32(domain_shifts if domain_shifts is not None else np.ones(n))
        self.root = np.zeros(8, np.uint32)
        self._h = lib().bfo_pcs_commit(self._arr, _p(sh), n, log_blowup, _p(self.root))
        self.ldes = []
        for i in range(n):
            r, c = C.c_uint64(), C.c_uint64()
            ptr = lib().bfo_pcs_lde(self._h, i, C.byref(r), C.byref(c))
            self.ldes.append(np.ctypeslib.as_array(ptr, shape=(r.value, c.value)))
        self.tree = Tree._borrow(lib().bfo_pcs_tree(self._h), self.ldes)

    def __del__(self):
        if getattr(self, "_h", None):
            self.ldes = []
            lib().bfo_pcs_data_free(self._h)
            self._h = None


# ---- tuned CPU baseline (oracle/fast_commit.c) -------------------------------------------------------------------------
def fast_available():
    return bool(lib().bfo_fast_available())


def fast_pcs_commit(mat, want_lde=False):
    """`Pcs::commit` of ONE matrix on the natural domain with the AVX-512 / Montgomery / cache-blocked implementation
    (the CPU arm bench.py times).  Returns (root, lde or None, {phase: seconds}); raises if the host has no AVX-512 or the
    shape is outside what the fast path covers (rows < 16)."""
    m = _u32(mat)
    rows, cols = m.shape
    root = np.zeros(8, np.uint32)
    lde = np.zeros((2 * rows, cols), np.uint32) if want_lde else None
    ph = (C.c_double * 3)()
    rc = lib().bfo_fast_pcs_commit(_p(m), rows, cols, _p(root), _p(lde) if want_lde else None, ph)
    if rc != 0:
        raise RuntimeError("bfo_fast_pcs_commit: unsupported shape or no AVX-512 on this host")
    return root, lde, dict(lde=ph[0], leaf_hash=ph[1], compress=ph[2])


def fast_release():
    lib().bfo_fast_release()


def fast_permute_many(states):
    s = _u32(states).copy()
    if lib().bfo_fast_permute_many(_p(s), s.shape[0]) != 0:
        raise RuntimeError("bfo_fast_permute_many: needs AVX-512 and a multiple of 16 states")
    return s


# ---- CPU arm of the AIR phases (oracle/fast_air.cpp): timing baseline for bench.py, checked against oracle/prover.py -----------
def air_chip_info(chip):
    info = (C.c_int * 4)()
    if air_lib().bfo_air_chip_info(int(chip), info) != 0:
        raise ValueError(f"unknown chip {chip}")
    return dict(main_w=info[0], prep_w=info[1], perm_w=info[2], n_constraints=info[3])


def air_packed_available():
    return bool(air_lib().bfo_air_packed_available())


def air_perm_trace(chip, main, prep, alpha, beta, timing=False, packed=None):
    """generate_permutation_trace on the CPU: (perm trace rows x 4*perm_w canonical, cumulative sum[, seconds of the C call]).
    packed: sixteen rows per step on AVX-512 (fast_air_packed.cpp); None = when the host and the shape allow, else the scalar arm."""
    import time
    inf = air_chip_info(chip)
    m = _u32(main)
    pr = _u32(prep) if inf["prep_w"] else None
    out = np.empty((m.shape[0], 4 * inf["perm_w"]), np.uint32)
    out.fill(0)  # touch the pages outside the timed call
    cs = np.zeros(4, np.uint32)
    if packed is None:
        packed = air_packed_available() and m.shape[0] >= 16 and m.shape[0] % 16 == 0
    fn = air_lib().bfo_air_perm_trace_packed if packed else air_lib().bfo_air_perm_trace
    t = time.perf_counter()
    rc = fn(int(chip), _p(m), _p(pr) if pr is not None else None, m.shape[0], _p(_u32(alpha)), _p(_u32(beta)), _p(out), _p(cs))
    dt = time.perf_counter() - t
    if rc != 0:
        raise RuntimeError("bfo_air_perm_trace failed")
    return (out, cs, dt) if timing else (out, cs)


def air_quotient(chip, main_lde, prep_lde, perm_lde, alpha_logup, beta, csum, alpha, timing=False, packed=None):
    """quotient_values on the CPU from bit-reversed-row LDEs (2n rows): (2, n, 4) canonical chunk values[, seconds of the C call]."""
    import time
    inf = air_chip_info(chip)
    ml, ql = _u32(main_lde), _u32(perm_lde)
    pl = _u32(prep_lde) if inf["prep_w"] else None
    n = ml.shape[0] // 2
    out = np.empty((2, n, 4), np.uint32)
    out.fill(0)
    if packed is None:
        packed = air_packed_available() and n >= 8
    fn = air_lib().bfo_air_quotient_packed if packed else air_lib().bfo_air_quotient
    t = time.perf_counter()
    rc = fn(int(chip), _p(ml), _p(pl) if pl is not None else None, _p(ql), n, _p(_u32(alpha_logup)), _p(_u32(beta)), _p(_u32(csum)), _p(_u32(alpha)), _p(out))
    dt = time.perf_counter() - t
    if rc != 0:
        raise RuntimeError("bfo_air_quotient failed")
    return (out, dt) if timing else out


def open_eval(lde, z, timing=False, packed=None):
    """Barycentric evaluation of every column of a committed matrix at the extension point z from its bit-reversed-row LDE (2n x w):
    (w, 4) canonical[, seconds]."""
    import time
    l = _u32(lde)
    out = np.zeros((l.shape[1], 4), np.uint32)
    if packed is None:
        packed = air_packed_available() and l.shape[0] >= 32 and l.shape[1] <= 256
    fn = air_lib().bfo_open_eval_packed if packed else air_lib().bfo_open_eval
    t = time.perf_counter()
    rc = fn(_p(l), l.shape[0] // 2, l.shape[1], _p(_u32(z)), _p(out))
    dt = time.perf_counter() - t
    if rc != 0:
        raise RuntimeError("bfo_open_eval failed")
    return (out, dt) if timing else out


def open_reduce_add(lde, zs, ys, alpha, reduced_before, ro, timing=False, packed=None):
    """ro (h x 4 canonical, updated in place) += the reduced openings of one matrix at its npts <= 2 opening points."""
    import time
    l = _u32(lde)
    zs, ys = _u32(zs).reshape(-1, 4), _u32(ys)
    assert ro.dtype == np.uint32 and ro.flags.c_contiguous and ro.shape == (l.shape[0], 4)
    if packed is None:
        packed = air_packed_available() and l.shape[0] >= 16
    fn = air_lib().bfo_open_reduce_add_packed if packed else air_lib().bfo_open_reduce_add
    t = time.perf_counter()
    rc = fn(_p(l), l.shape[0], l.shape[1], zs.shape[0], _p(zs), _p(ys), _p(_u32(alpha)), int(reduced_before), _p(ro))
    dt = time.perf_counter() - t
    if rc != 0:
        raise RuntimeError("bfo_open_reduce_add failed")
    return dt if timing else None


def fast_fri_commit_phase(inputs, betas, rollin_beta2=False):
    """FRI commit phase on the CPU with caller-supplied betas (rounds x 4): (roots (rounds, 8), final_poly (4,), {hash, fold} seconds).
    inputs: extension vectors (len x 4 canonical), strictly decreasing power-of-two lengths."""
    ins = [_u32(v) for v in inputs]
    ptrs = (u32p * len(ins))(*[_p(v) for v in ins])
    logs = np.array([v.shape[0].bit_length() - 1 for v in ins], np.uint32)
    b = _u32(betas).reshape(-1, 4)
    rounds = int(logs[0]) - 1
    assert b.shape[0] >= rounds
    roots = np.zeros((rounds, 8), np.uint32)
    fin = np.zeros(4, np.uint32)
    sec = (C.c_double * 2)()
    rc = lib().bfo_fast_fri_commit_phase(ptrs, _p(logs), len(ins), _p(b), 1 if rollin_beta2 else 0, _p(roots), _p(fin), sec)
    if rc != rounds:
        raise RuntimeError(f"bfo_fast_fri_commit_phase: {rc}")
    return roots, fin, dict(hash=sec[0], fold=sec[1])
