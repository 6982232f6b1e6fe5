"""CPU tests: the plain-C oracle against the reference's golden vectors, and
the numpy model of the device algorithm against the same vectors."""
import os

import numpy as np
import pytest

import golden_io
from msolve_b200 import systems
from oracle import oracle

P31, P16, P8 = 1073741827, 65521, 251


@pytest.mark.parametrize("path", golden_io.fixture_files(), ids=os.path.basename)
def test_oracle_matches_reference_fixtures(path):
    cases = golden_io.load_cases(path)
    assert cases
    for m, ref, meta in cases:
        out, units, apps = oracle.reduce(m)
        assert out.same_rows(ref), f"{path}: matrix {m.nru}x{m.nrl} flags {m.flags}"
        if m.cf_bytes == 4:      # the reference counts operations in its 31-bit reducer only
            assert abs(units - meta["units"]) <= 1e-6 * max(1.0, units)
            assert apps == int(round(meta["pivot_apps"]))


def test_oracle_mod_inverse():
    rng = np.random.default_rng(7)
    for p in (P31, P16, P8, 2147483629, 3, 2):
        for v in list(rng.integers(1, p, size=50)) + [1, p - 1]:
            v = int(v) % p
            if v == 0:
                continue
            inv = oracle.mod_inverse(v, p)
            assert (v * inv) % p == 1


def test_oracle_edge_cases():
    from msolve_b200.flat import FlatMatrix
    p = P31
    # no lower rows
    m = FlatMatrix(p, 1, 0, 1, 1, np.array([0, 2]), np.array([0, 1]), np.array([0]),
                   np.array([0, 2]), np.array([1, 5], dtype=np.uint32))
    out, _, _ = oracle.reduce(m)
    assert out.nrows == 0
    # one pivot row x0 + 5 x1, lower rows: 2*x0 + 3*x1 ; 4*x0 + 20*x1 (dependent)
    m = FlatMatrix(p, 1, 2, 1, 1, np.array([0, 2, 4, 6]), np.array([0, 1, 0, 1, 0, 1]), np.array([0, 1, 2]),
                   np.array([0, 2, 4, 6]), np.array([1, 5, 2, 3, 4, 20], dtype=np.uint32))
    out, units, apps = oracle.reduce(m)
    assert out.nrows == 1 and list(out.cols) == [1] and list(out.cf) == [1]
    # lower row identical to a pivot row -> zero
    m = FlatMatrix(p, 1, 1, 1, 1, np.array([0, 2, 4]), np.array([0, 1, 0, 1]), np.array([0, 0]),
                   np.array([0, 2]), np.array([1, 5], dtype=np.uint32))
    out, _, _ = oracle.reduce(m)
    assert out.nrows == 0


def test_device_model_matches_fixtures():
    """The mathematics of the CUDA pipeline (transposed level-scheduled solve +
    Gauss-Jordan), modelled in numpy, gives the reference's echelon rows."""
    import model_pull_solve as mp
    n = 0
    for path in golden_io.fixture_files():
        for m, ref, meta in golden_io.load_cases(path):
            if meta["kind"] != 0 or m.flags != 0 or m.nnz * max(1, m.nrl) > 400000:
                continue
            assert mp.echelon(m).same_rows(ref), path
            n += 1
    assert n >= 20


def test_system_generators_reproduce_reference_files():
    """cyclic-5 / katsura-5 / eco-6 formulas give the reference's input files
    (input_files/cyclic5-31.ms, kat6-31.ms, eco6-31.ms): same reduced GB sizes
    as recorded from the reference (tests/golden/gb_digests.json)."""
    d = golden_io.gb_digests()
    assert d["cyclic-5@1073741827"]["nelts"] == 20 and d["cyclic-5@1073741827"]["nterms"] == 232
    s = systems.cyclic(5)
    assert s.nvars == 5 and len(s.polys) == 5 and sum(len(f) for f in s.polys) == 5 + 5 + 5 + 5 + 2
    k = systems.katsura(5)
    assert k.nvars == 6 and len(k.polys) == 6
    e = systems.eco(6)
    assert e.nvars == 6 and len(e.polys) == 6
