"""CPU model of the arithmetic of the tensor-core kernels (msolve_b200/csrc/f4la_tc.cuh): residues < 2^31 are split
into four 8-bit limbs, `tcgen05.mma.kind::i8` accumulates limb products in int32, the seven sums with equal weight
2^(8s), s = a + b, are recombined mod p in the epilogue.  This checks the mathematics the kernels rely on -- the
recombination is exact and the int32 accumulators cannot overflow up to the depth bound kTcMaxK = 8192 -- without
a GPU (the kernels themselves are compared with exact integer arithmetic in tests/test_gpu_tc.py)."""
import numpy as np

K_MAX = 8192            # kTcMaxK


def limbs(x):
    """uint32 array -> [4, ...] uint8 limbs, least significant first"""
    return np.stack([(x >> (8 * b)) & 0xFF for b in range(4)]).astype(np.int64)


def limb_gemm_mod_p(A, B, p):
    """A [M, K], B [K, N] residues < p < 2^31 -> (A @ B mod p, largest absolute accumulator value)"""
    la, lb = limbs(A), limbs(B)
    acc = np.zeros((7,) + (A.shape[0], B.shape[1]), dtype=np.int64)
    for a in range(4):
        for b in range(4):
            acc[a + b] += la[a] @ lb[b]           # what the 4 MMAs of a k-step add up, per weight 2^(8 (a + b))
    worst = int(acc.max())
    out = np.zeros(acc.shape[1:], dtype=object)
    for s in range(7):
        out = (out + (acc[s].astype(object) % p) * pow(2, 8 * s, p)) % p
    return out.astype(np.uint64), worst


def test_limb_recombination_is_exact():
    rng = np.random.default_rng(5)
    for p in (1073741827, 2147483629, 65521, 251):
        for M, K, N in ((5, 17, 3), (16, 200, 9), (3, 1024, 4)):
            A = rng.integers(0, p, size=(M, K), dtype=np.uint64).astype(np.uint32)
            B = rng.integers(0, p, size=(K, N), dtype=np.uint64).astype(np.uint32)
            got, worst = limb_gemm_mod_p(A, B, p)
            want = (A.astype(object) @ B.astype(object)) % p
            assert np.array_equal(got.astype(object), want), (p, M, K, N)
            assert worst < 2 ** 31


def test_accumulator_bound_at_full_depth():
    """All residues 2^31 - 1 (every limb 255 except the top one 127), depth 8192: the weight with four limb pairs
    (a + b = 3) collects at most 4 * 255^2 * 8192 = 2 130 739 200 < 2^31; two more k-blocks of 64 would not fit."""
    assert 4 * 255 * 255 * K_MAX < 2 ** 31 <= 4 * 255 * 255 * (K_MAX + 128)
    x = np.full((1, K_MAX), 2 ** 31 - 1, dtype=np.uint32)
    y = np.full((K_MAX, 1), 2 ** 31 - 1, dtype=np.uint32)
    p = 2147483629
    _, worst = limb_gemm_mod_p(x % p, y % p, p)
    assert worst < 2 ** 31
    # the true worst case over ALL limb patterns, not only over residues: every limb 255
    la = np.full((4, 1, K_MAX), 255, dtype=np.int64)
    acc3 = sum(int((la[a] @ la[3 - a].transpose(1, 0)).item()) for a in range(4))
    assert acc3 == 4 * 255 * 255 * K_MAX < 2 ** 31
