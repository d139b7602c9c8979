"""Worker of tests/test_verifier_fuzz.py: mutates a valid serialised proof and feeds every mutant to `bfgpu_verify_core_proof` /
`bfgpu_shard_proof_to_bincode` of the library given on the command line (libbfgpu.so, or the sanitizer build of the same headers
made from tools/verifier_host_shim.cpp).  Exit status 0 = no mutant was accepted and the process survived.

    python tests/fuzz_verifier_worker.py <lib.so> <proof.npz> <seed> <trials>
"""
import ctypes as C
import sys

import numpy as np

P = 2130706433


def load(path):
    lib = C.CDLL(path)
    u32p = C.POINTER(C.c_uint32)
    lib.bfgpu_verify_core_proof.restype = C.c_int32
    lib.bfgpu_verify_core_proof.argtypes = [u32p, C.POINTER(C.c_char_p), u32p, C.c_int32, u32p, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                            u32p, C.c_int32, C.c_char_p, C.c_uint64]
    lib.bfgpu_shard_proof_to_bincode.restype = C.c_int32
    lib.bfgpu_shard_proof_to_bincode.argtypes = [C.POINTER(C.c_char_p), u32p, C.c_int32, u32p, C.c_uint64, C.c_int, C.c_uint32, C.c_int, C.c_void_p,
                                                 C.c_uint64, C.POINTER(C.c_uint64), C.c_char_p, C.c_uint64]
    return lib, u32p


def mutate(rng, words):
    bad = words.copy()
    n_shape = 25 + 6 * 8  # commitments, chip count, per-chip (index, log_degree, cumulative sum): where the shape words live
    for _ in range(int(rng.integers(1, 4))):
        pos = int(rng.integers(0, n_shape)) if rng.random() < 0.3 else int(rng.integers(0, len(bad)))
        mode = int(rng.integers(0, 5))
        if mode == 0:
            bad[pos] = (int(bad[pos]) + 1) % P
        elif mode == 1:
            bad[pos] = int(rng.integers(0, 2 ** 32))
        elif mode == 2:
            bad[pos] = int(rng.choice([0, 1, 2, 7, 8, 9, 22, 23, 24, 25, 31, 32, 33, 64, 1 << 20, P - 1, P, P + 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF]))
        elif mode == 3:
            bad[pos] = int(bad[int(rng.integers(0, len(bad)))])
        else:
            bad[pos] = int(bad[pos]) ^ (1 << int(rng.integers(0, 32)))
    r = rng.random()
    if r < 0.08:
        bad = bad[:int(rng.integers(0, len(bad)))]
    elif r < 0.12:
        bad = np.concatenate([bad, rng.integers(0, 2 ** 32, int(rng.integers(1, 40)), dtype=np.uint64).astype(np.uint32)])
    return np.ascontiguousarray(bad, np.uint32)


def main():
    lib, u32p = load(sys.argv[1])
    z = np.load(sys.argv[2], allow_pickle=False)
    words, commit, logs = z["words"].astype(np.uint32), z["commit"].astype(np.uint32), z["logs"].astype(np.uint32)
    names = [n.encode() for n in str(z["names"]).split(",")]
    fri = [int(x) for x in z["fri"]]
    rng = np.random.default_rng(int(sys.argv[3]))
    cn = (C.c_char_p * len(names))(*names)
    err = C.create_string_buffer(256)

    def verify(w, log_blowup=fri[0]):
        return lib.bfgpu_verify_core_proof(commit.ctypes.data_as(u32p), cn, logs.ctypes.data_as(u32p), len(names), w.ctypes.data_as(u32p), w.size, 0,
                                           log_blowup, fri[1], fri[2], None, 0, err, 256)

    def bincode(w):
        n = C.c_uint64()
        rc = lib.bfgpu_shard_proof_to_bincode(cn, logs.ctypes.data_as(u32p), len(names), w.ctypes.data_as(u32p), w.size, 0, fri[0], 1, None, 0, C.byref(n), err, 256)
        if rc == 0 and n.value < (1 << 28):
            out = np.zeros(n.value, np.uint8)
            lib.bfgpu_shard_proof_to_bincode(cn, logs.ctypes.data_as(u32p), len(names), w.ctypes.data_as(u32p), w.size, 0, fri[0], 1, out.ctypes.data_as(C.c_void_p),
                                             out.size, C.byref(n), err, 256)

    assert verify(words) == 0, err.value
    bincode(words)
    accepted = 0
    for t in range(int(sys.argv[4])):
        bad = mutate(rng, words)
        if bad.size == words.size and (bad == words).all():
            continue
        if verify(bad) == 0:
            accepted += 1
            print("ACCEPTED a mutant, trial", t, flush=True)
        if t % 4 == 0:
            bincode(bad)
        if t % 16 == 0:  # hostile verifier-side parameters as well
            verify(bad, int(rng.choice([0, 2, 5, 23, 24, 25, 31, 32, 0xFFFFFFFF])))
    print("mutants", sys.argv[4], "accepted", accepted, flush=True)
    return 1 if accepted else 0


if __name__ == "__main__":
    sys.exit(main())
