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


def test_schur_tensor_path_forced_on_small_matrices(monkeypatch):
    """The non-pivot columns as a tensor-core Schur product (k_schur_*: D' = X + M * Bneg over the non-zero tiles
    of B) normally only runs on large matrices; F4LA_B200_SCHUR_MIN=0 forces it everywhere, and a small vector
    budget ($F4LA_B200_L2_MB) splits the lower rows into several launch groups.  Every mode of the pipeline
    (echelon, normal form, pivots keyed by lead, reducer bit arrays, 8/16/31-bit primes) must give the rows of the
    reference fixtures and of the oracle."""
    import golden_io
    from msolve_b200.backend import Backend
    from msolve_b200.flat import FLAG_NORMALFORM, FLAG_WANT_RBA
    from oracle import oracle
    from test_gpu_parity import random_matrix, big_sparse_matrix
    monkeypatch.setenv("F4LA_B200_SCHUR_MIN", "0")
    monkeypatch.setenv("F4LA_B200_L2_MB", "1")
    be = Backend(0)
    try:
        n = 0
        for path in golden_io.fixture_files():
            for m, ref, meta in golden_io.load_cases(path):
                out, st = be.reduce(m)
                assert out.same_rows(ref), path
                if meta["kind"] == 0 and m.nru:
                    m.flags = FLAG_NORMALFORM
                    out, _ = be.reduce(m)
                    chk, _, _ = oracle.reduce(m)
                    assert out.same_rows(chk)
                    m.flags = FLAG_WANT_RBA
                    out, _ = be.reduce(m)
                    assert out.same_rows(ref) and out.rba is not None
                    m.flags = 0
                n += 1
        assert n > 40
        rng = np.random.default_rng(31)
        for p, dt in ((1073741827, np.uint32), (2147483629, np.uint32), (65521, np.uint16), (251, np.uint8)):
            for nru, nrl, ncr, dens in [(300, 150, 70, 0.05), (1000, 520, 130, 0.02), (64, 700, 40, 0.1)]:
                m = random_matrix(rng, p, nru, nrl, ncr, dens, dt)
                out, _ = be.reduce(m)
                ref, _, _ = oracle.reduce(m)
                assert out.same_rows(ref), (p, nru, nrl, ncr)
        m = big_sparse_matrix(rng, 1073741827, 6000, 300, 400, 200)      # several k-block segments per column block
        out, st = be.reduce(m)
        ref, _, _ = oracle.reduce(m)
        assert out.same_rows(ref)
    finally:
        be.close()
