"""GPU tests at BASELINE.json's full commit size (2^22 x 256, blowup 2) through size-independent properties:
determinism, Merkle open -> verify round trips against the oracle's `verify_batch`, the low-degree property of the
committed LDE (opening + FRI accepted by the oracle verifier), and consistency of the device-resident and host input
paths.  The trace is generated on the device (torch) so the test does not move 4 GiB through the host."""
import ctypes as C

import numpy as np
import pytest

import zkvm_brainfuck_b200 as bf

pytestmark = pytest.mark.gpu
P = bf.P


@pytest.fixture(scope="module")
def ctx():
    c = bf.Context()
    yield c
    c.close()


def device_commit(ctx, trace):
    """Pcs::commit on a torch int32 device tensor (row-major canonical words) -> (root, PcsProverData)."""
    mat = bf.Mat(trace.data_ptr(), trace.shape[0], trace.shape[1])
    root = np.zeros(8, np.uint32)
    h = C.c_void_p()
    ctx.set_input_space(bf.MEM_DEVICE)
    try:
        ctx.check(bf.lib().bfgpu_pcs_commit(ctx._h, C.byref(mat), None, 1, root.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(h)))
    finally:
        ctx.set_input_space(bf.MEM_HOST)
    return root, bf.PcsProverData(ctx, h)


@pytest.mark.parametrize("log_rows", [20, 22])
def test_fullsize_commit_properties(ctx, oracle, log_rows):
    import torch
    from oracle import stark as S
    free, _ = torch.cuda.mem_get_info()
    need = (1 << log_rows) * 256 * 4 * 6
    if free < need:
        pytest.skip("not enough free device memory")
    R, W = 1 << log_rows, 256
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    trace = torch.randint(0, P, (R, W), dtype=torch.int32, device="cuda", generator=g)
    torch.cuda.synchronize()
    root1, d1 = device_commit(ctx, trace)
    root2, d2 = device_commit(ctx, trace)
    assert (root1 == root2).all()
    d2.free()
    # Merkle round trips: opened LDE rows + paths verify against the root with the ORACLE's hasher
    mmcs = bf.MerkleTreeMmcs(ctx)
    rng = np.random.default_rng(5)
    for index in [0, 1, 2 * R - 1] + [int(x) for x in rng.integers(0, 2 * R, 5)]:
        rows, sib = mmcs.open_batch(index, d1.tree)
        assert sib.shape == (log_rows + 1, 8)
        assert oracle.verify_batch(root1, [(2 * R, W)], index, rows, sib)
        bad = [rows[0].copy()]
        bad[0][17] ^= 1
        assert not oracle.verify_batch(root1, [(2 * R, W)], index, bad, sib)
    # the stored rows are the LDE of the trace: row 0 of the bit-reversed LDE is every column's interpolant at x = 3,
    # i.e. sum_k c_k 3^k; check two columns against a host evaluation from the trace column (O(n) Horner on coefficients
    # obtained with the oracle's iDFT)
    rows0, _ = mmcs.open_batch(0, d1.tree)
    for c in (0, 255):
        col = trace[:, c].cpu().numpy().astype(np.uint32).reshape(-1, 1)
        coef = oracle.idft_batch(col)[:, 0].astype(np.uint64)
        pw = S.powers(3, R)
        acc = 0
        for s in range(0, R, 1 << 16):
            acc = (acc + int(np.sum(coef[s:s + (1 << 16)] * pw[s:s + (1 << 16)] % np.uint64(P)) % P)) % P
        assert int(rows0[0][c]) == acc
    # low-degree: open at a random point with a short FRI and let the oracle's Pcs::verify accept it
    ctx.set_fri_params(1, 6, 4)
    try:
        zeta = rng.integers(0, P, 4, dtype=np.uint64)
        pcs = bf.TwoAdicFriPcs(ctx)
        gch, och = bf.Challenger(ctx), S.Challenger()
        opened, proof = pcs.open([(d1, [[zeta]])], gch)
        rounds = [(root1, [(S.Domain(log_rows), [(zeta, opened[0][0][0])])])]
        assert S.pcs_verify(S.FriConfig(1, 6, 4), rounds, proof, och) is None
    finally:
        ctx.set_fri_params(1, 84, 16)
    d1.free()
    del trace
    torch.cuda.empty_cache()
