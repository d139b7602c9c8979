"""Deterministic synthetic trace of the bench workload: element i = r * cols + c of the row-major
matrix is SplitMix64(seed, i) reduced to a canonical KoalaBear residue.  The same words can be produced on the host
(numpy: oracle side, golden roots under tests/golden/) and on the device (torch: bench.py, GPU tests), so the Merkle root
of the exact bench input is pinned by the CPU oracle instead of being whatever the GPU printed.

    state = seed + (i + 1) * 0x9E3779B97F4A7C15            (mod 2^64)
    z = (state ^ (state >> 30)) * 0xBF58476D1CE4E5B9;  z = (z ^ (z >> 27)) * 0x94D049BB133111EB;  z ^= z >> 31
    word  = (z >> 33) mod p,   p = 2^31 - 2^24 + 1
"""
import numpy as np

P = 2130706433
SEED = 0xB200
_G, _M1, _M2 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB


def trace_numpy(rows, cols, seed=SEED, chunk_rows=1 << 16):
    out = np.empty((rows, cols), np.uint32)
    with np.errstate(over="ignore"):
        for r0 in range(0, rows, chunk_rows):
            r1 = min(rows, r0 + chunk_rows)
            i = np.arange(r0 * cols, r1 * cols, dtype=np.uint64)
            z = np.uint64(seed) + (i + np.uint64(1)) * np.uint64(_G)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
            z = z ^ (z >> np.uint64(31))
            out[r0:r1] = ((z >> np.uint64(33)) % np.uint64(P)).astype(np.uint32).reshape(r1 - r0, cols)
    return out


def _s64(x):
    """two's-complement int64 view of an unsigned 64-bit constant (torch has no uint64 arithmetic)"""
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def trace_torch(rows, cols, seed=SEED, device="cuda", chunk_rows=1 << 17):
    """same words as trace_numpy, generated on `device` as an int32 tensor (rows, cols)"""
    import torch
    out = torch.empty((rows, cols), dtype=torch.int32, device=device)

    def lsr(z, k):  # logical shift right on int64
        return (z >> k) & ((1 << (64 - k)) - 1)

    for r0 in range(0, rows, chunk_rows):
        r1 = min(rows, r0 + chunk_rows)
        i = torch.arange(r0 * cols, r1 * cols, dtype=torch.int64, device=device)
        z = (i + 1) * _s64(_G) + _s64(seed)
        z = (z ^ lsr(z, 30)) * _s64(_M1)
        z = (z ^ lsr(z, 27)) * _s64(_M2)
        z = z ^ lsr(z, 31)
        out[r0:r1] = (lsr(z, 33) % P).to(torch.int32).reshape(r1 - r0, cols)
    return out
