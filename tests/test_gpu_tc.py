"""Tensor-core exact product mod p (csrc/f4la_tc.cuh: tcgen05.mma.kind::i8 on 8-bit limbs, TMEM accumulators)
against an exact numpy product -- the kernel behind the verification of tall trailing blocks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def matmul_mod(A, B, p):
    """Exact (A @ B) mod p for residues < 2^31 and K <= 2^13: 16-bit halves, int64 partial products."""
    A = A.astype(np.int64); B = B.astype(np.int64)
    al, ah, bl, bh = A & 0xffff, A >> 16, B & 0xffff, B >> 16
    ll, lh, hl, hh = al @ bl, al @ bh, ah @ bl, ah @ bh          # each < 2^32 * 2^13
    w16, w32 = (1 << 16) % p, (1 << 32) % p
    mid = ((lh % p) + (hl % p)) % p
    return ((ll % p) + (mid * w16) % p + ((hh % p) * w32) % p) % p


@pytest.mark.parametrize("p", [1073741827, 2147483629, 65521, 251, 3])
def test_tc_gemm_matches_exact_product(backend, p):
    rng = np.random.default_rng(p % 977)
    for M, N, K in [(128, 64, 64), (256, 100, 70), (200, 64, 1), (130, 7, 33), (1000, 513, 333), (384, 300, 2048)]:
        A = rng.integers(0, p, size=(M, K), dtype=np.int64)
        B = rng.integers(0, p, size=(K, N), dtype=np.int64)
        A[rng.random(A.shape) < 0.3] = 0                      # zero limbs / zero entries too
        got, ms = backend.modp_gemm(A, B, p)
        want = matmul_mod(A, B, p)
        assert np.array_equal(got.astype(np.int64), want), (p, M, N, K)
        assert ms > 0


def test_tc_gemm_accumulator_bound(backend):
    """Worst case of the no-overflow bound: K = 8192 (the deepest accumulation the path is used for), every
    residue p - 1 for the largest 31-bit prime (all limbs at their maxima 255 / 255 / 255 / 127)."""
    p = 2147483629
    M, N, K = 128, 64, 8192
    A = np.full((M, K), p - 1, dtype=np.int64)
    B = np.full((K, N), p - 1, dtype=np.int64)
    got, _ = backend.modp_gemm(A, B, p)
    assert np.array_equal(got.astype(np.int64), matmul_mod(A, B, p))
    A = np.full((M, K), 0x7fffffff & ~0, dtype=np.int64) % p   # 2^31 - 1 mod p: limbs 255,255,255,127 pattern
    got, _ = backend.modp_gemm(A, B, p)
    assert np.array_equal(got.astype(np.int64), matmul_mod(A, B, p))
    from msolve_b200.backend import F4LAError
    with pytest.raises(F4LAError):
        backend.modp_gemm(np.ones((128, 8193), dtype=np.int64), np.ones((8193, 64), dtype=np.int64), p)
