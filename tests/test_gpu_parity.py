"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes
through the C ABI of csrc/libf4la_b200.so; the oracle / compiled reference are
only the checkers."""
import os

import numpy as np
import pytest

import golden_io
from msolve_b200 import systems
from msolve_b200.flat import FlatMatrix
from oracle import oracle
from oracle import refharness as rh

pytestmark = pytest.mark.gpu

P31, P16, P8 = 1073741827, 65521, 251


@pytest.mark.parametrize("path", golden_io.fixture_files(), ids=os.path.basename)
def test_gpu_matches_reference_fixtures(backend, path):
    n = 0
    for m, ref, meta in golden_io.load_cases(path):
        out, st = backend.reduce(m)       # F4 rounds and the final reduction step (pivots keyed by lead)
        assert out.same_rows(ref), f"{path}: {m.nru}+{m.nrl} rows x {m.ncl}+{m.ncr} cols"
        assert st["kernel_launches"] > 0
        n += 1
    assert n > 0


def test_gpu_edge_cases(backend):
    p = P31
    # no lower rows
    m = FlatMatrix(p, 1, 0, 1, 1, np.array([0, 2]), np.array([0, 1]), np.array([0]),
                   np.array([0, 2]), np.array([1, 5], dtype=np.uint32))
    out, _ = backend.reduce(m)
    assert out.nrows == 0
    # dependent lower rows -> one new pivot
    m = FlatMatrix(p, 1, 2, 1, 1, np.array([0, 2, 4, 6]), np.array([0, 1, 0, 1, 0, 1]), np.array([0, 1, 2]),
                   np.array([0, 2, 4, 6]), np.array([1, 5, 2, 3, 4, 20], dtype=np.uint32))
    out, _ = backend.reduce(m)
    ref, _, _ = oracle.reduce(m)
    assert out.same_rows(ref) and out.nrows == 1
    # a lower row equal to a pivot row reduces to zero
    m = FlatMatrix(p, 1, 1, 1, 1, np.array([0, 2, 4]), np.array([0, 1, 0, 1]), np.array([0, 0]),
                   np.array([0, 2]), np.array([1, 5], dtype=np.uint32))
    out, _ = backend.reduce(m)
    assert out.nrows == 0
    # no pivots at all: plain echelon form of the lower rows
    m = FlatMatrix(p, 0, 2, 0, 3, np.array([0, 2, 4]), np.array([0, 2, 0, 1]), np.array([0, 1]),
                   np.array([0, 2, 4]), np.array([3, 7, 6, 1], dtype=np.uint32))
    out, _ = backend.reduce(m)
    ref, _, _ = oracle.reduce(m)
    assert out.same_rows(ref) and out.nrows == 2


def test_gpu_rejects_malformed_input(backend):
    from msolve_b200.backend import F4LAError
    p = P31
    # pivot row 0 does not start with column 0
    m = FlatMatrix(p, 1, 1, 1, 1, np.array([0, 2, 4]), np.array([1, 0, 0, 1]), np.array([0, 0]),
                   np.array([0, 2]), np.array([1, 5], dtype=np.uint32))
    with pytest.raises(F4LAError):
        backend.reduce(m)


def random_matrix(rng, p, nru, nrl, ncr, density, cf_dtype):
    """Random [A B; C D] with unit upper triangular A, ragged row lengths."""
    nc = nru + ncr
    rows, cfs = [], []
    for i in range(nru):
        k = rng.integers(0, max(1, int(density * (nc - i - 1)) + 1))
        others = rng.choice(np.arange(i + 1, nc), size=min(k, nc - i - 1), replace=False) if nc - i - 1 > 0 else np.zeros(0, int)
        rng.shuffle(others)
        rows.append(np.concatenate([[i], others]))
        cfs.append(np.concatenate([[1], rng.integers(1, p, size=len(others))]))
    for r in range(nrl):
        k = int(rng.integers(0, max(2, int(density * nc) + 2)))
        c = rng.choice(nc, size=min(k, nc), replace=False)
        if len(c) == 0 and r % 3:
            c = np.array([rng.integers(0, nc)])
        c = np.sort(c)
        rows.append(c)
        cfs.append(rng.integers(1, p, size=len(c)))
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
    return FlatMatrix(p, nru, nrl, nru, ncr, row_ptr, np.concatenate(rows) if rows else np.zeros(0),
                      np.arange(nru + nrl), row_ptr.copy(), np.concatenate(cfs).astype(cf_dtype))


@pytest.mark.parametrize("p,dt", [(P31, np.uint32), (2147483629, np.uint32), (P16, np.uint16), (P8, np.uint8),
                                  (65537, np.uint32), (3, np.uint8)])
def test_gpu_vs_oracle_random_ragged(backend, p, dt):
    rng = np.random.default_rng(1234 + p % 1000)
    for nru, nrl, ncr, dens in [(40, 17, 9, 0.2), (300, 150, 70, 0.05), (257, 129, 33, 0.3), (1000, 520, 130, 0.02),
                                (64, 700, 40, 0.1)]:
        m = random_matrix(rng, p, nru, nrl, ncr, dens, dt)
        out, st = backend.reduce(m)
        ref, units, apps = oracle.reduce(m)
        assert out.same_rows(ref), (p, nru, nrl, ncr)


@pytest.mark.skipif(not rh.available(), reason="compiled reference not shipped")
@pytest.mark.parametrize("name,p", [("cyclic-7", P31), ("cyclic-7", P16), ("cyclic-7", P8), ("katsura-8", P31),
                                    ("katsura-8", P16), ("eco-9", P31), ("randquad-6", P31)])
def test_gpu_vs_reference_every_f4_step(backend, name, p):
    """Per-F4-step parity: identical reduced row echelon rows on every matrix
    the reference's F4 builds (the reference's own LA output is the checker)."""
    gb, recs = rh.record_f4(systems.by_name(name, p), p, threads=4)
    n = 0
    for r in recs:
        out, st = backend.reduce(r.matrix)
        assert out.same_rows(r.result), (name, p, r.kind, r.matrix.nru, r.matrix.nrl)
        n += 1
    assert n > 5 and any(r.kind == 1 for r in recs)


@pytest.mark.parametrize("path", [f for f in golden_io.fixture_files() if "cyclic-6" in f or "katsura-5_p65521" in f
                                  or "eco-6_p251" in f], ids=os.path.basename)
def test_gpu_normal_form_mode(backend, path):
    """nf > 0 contract (la_ff_32.c:2672-2684): every lower row comes back as its own
    normal form modulo the known pivots, not normalised, not inter-reduced."""
    from msolve_b200.flat import FLAG_NORMALFORM
    n = 0
    for m, ref, meta in golden_io.load_cases(path):
        if meta["kind"] != 0:
            continue
        m.flags = FLAG_NORMALFORM
        out, st = backend.reduce(m)
        chk, _, _ = oracle.reduce(m)
        assert out.nrows == m.nrl and out.same_rows(chk)
        if rh.available():
            again, _, _, _ = rh.reference_la(m, laopt=2, threads=1)
            assert out.same_rows(again)
        n += 1
    assert n > 3


@pytest.mark.skipif(not rh.available(), reason="compiled reference not shipped")
@pytest.mark.parametrize("name,p", [("cyclic-7", P31), ("cyclic-7", P8), ("katsura-8", P16), ("cyclic-8", P31),
                                    ("cyclic-8", P16), ("cyclic-8", P8), ("katsura-10", P31)])
def test_gpu_backend_inside_reference_f4(name, p):
    """Drop-in test: neogb's F4 with `linear_algebra` pointing at the GPU
    backend returns the reference's reduced Groebner basis bit for bit."""
    import gpu_helpers
    L = gpu_helpers.install_backend()
    try:
        gb = rh.run_f4(systems.by_name(name, p), p, threads=4, laopt=2)
        tot = gpu_helpers.neogb_totals(L)
    finally:
        gpu_helpers.uninstall_backend()
    want = golden_io.gb_digests()[f"{name}@{p}"]
    assert len(gb.lens) == want["nelts"] and int(gb.lens.sum()) == want["nterms"]
    assert gb.digest() == want["digest"]
    assert tot["calls"] > 0 and tot["kernel_launches"] > 0


@pytest.mark.skipif(not rh.available(), reason="compiled reference not shipped")
@pytest.mark.parametrize("name", ["katsura-7", "cyclic-6", "eco-8"])
def test_gpu_qq_tracer_learn_and_apply(name):
    """QQ multi-modular flow (msolve.c:1628-1632, 1815-1824): learn a trace with the
    GPU backend, replay it for further primes on the GPU, and compare every
    modular Groebner basis with the reference's own learn/apply run."""
    import gpu_helpers
    system = systems.by_name(name)
    primes = rh.primes_above(1 << 30, 9)
    # reference: CPU learn + apply
    rh.reset(); rh.set_mode(rh.MODE_REFERENCE)
    rh.qq_setup(system, threads=2)
    ref_learn = rh.qq_learn(primes[0])
    ref_apply, _ = rh.qq_apply(primes[1:], threads=2)
    assert ref_learn["rc"] == 0 and all(r["err"] == 0 for r in ref_apply)
    # GPU backend in both phases
    L = gpu_helpers.install_backend()
    try:
        rh.qq_setup(system, threads=2)
        got_learn = rh.qq_learn(primes[0])
        got_apply, wall = rh.qq_apply(primes[1:], threads=2, ngpus=1)
        tot = gpu_helpers.neogb_totals(L)
    finally:
        gpu_helpers.uninstall_backend()
    assert got_learn["rc"] == 0 and got_learn["digest"] == ref_learn["digest"]
    assert got_learn["nelts"] == ref_learn["nelts"] and got_learn["nterms"] == ref_learn["nterms"]
    for g, r in zip(got_apply, ref_apply):
        assert g["err"] == 0 and g["digest"] == r["digest"]


def test_gpu_rba_bits_match_multipliers(backend):
    """F4LA_FLAG_WANT_RBA: bit i of row r is set iff known pivot i is applied to
    row r by the reference's left-to-right reduction (checked with the oracle's
    pivot-application count and by direct recomputation on a small matrix)."""
    from msolve_b200.flat import FLAG_WANT_RBA
    path = [f for f in golden_io.fixture_files() if "cyclic-6" in f][0]
    n = 0
    for m, ref, meta in golden_io.load_cases(path):
        if meta["kind"] != 0 or m.nru == 0:
            continue
        m.flags = FLAG_WANT_RBA
        out, st = backend.reduce(m)
        assert out.same_rows(ref)
        assert out.rba is not None and out.rba.shape == (m.nrl, (m.nru + 31) // 32)
        # multipliers by forward substitution in Python (exact integers)
        p = m.fc
        piv = [dict(zip(*(x.tolist() for x in m.row(i)))) for i in range(m.nru)]
        for r in range(m.nrl):
            c, v = m.row(m.nru + r)
            acc = dict(zip(c.tolist(), v.tolist()))
            used = []
            for i in range(m.nru):
                mul = acc.get(i, 0) % p
                if mul:
                    used.append(i)
                    for cc, a in piv[i].items():
                        acc[cc] = (acc.get(cc, 0) - mul * a) % p
            bits = [i for i in range(m.nru) if (int(out.rba[r, i // 32]) >> (i % 32)) & 1]
            assert bits == used
        n += 1
        if n >= 6:
            break
    assert n > 2


@pytest.mark.skipif(not rh.available(), reason="compiled reference not shipped")
def test_gpu_interreduce_and_application_entry_points():
    """The remaining matrix-level pointers of data.h:529-595 driven on mat_t/bs_t/md_t built by
    the harness: interreduce_matrix_rows (la_ff_32.c:4254) against the reference's dispatcher,
    application_linear_algebra (return code + rows) against the reference's exact LA."""
    import ctypes as C
    import gpu_helpers
    L = gpu_helpers.install_backend()
    ir = C.cast(L.f4la_b200_interreduce_matrix_rows, C.c_void_p).value
    app = C.cast(L.f4la_b200_application_linear_algebra, C.c_void_p).value
    n_ir = n_app = n_bad = 0
    try:
        for path in golden_io.fixture_files():
            for m, ref, meta in golden_io.load_cases(path):
                if meta["kind"] == 1:
                    # all rows of the final reduction step have distinct lead columns
                    m2 = FlatMatrix(m.fc, 0, m.nru + m.nrl, 0, m.ncl + m.ncr, m.row_ptr, m.cols, m.row_cf,
                                    m.cf_ptr, m.cf, flags=0)
                    want, _, want_idx = rh.call_entry(m2, 2, None)
                    got, _, got_idx = rh.call_entry(m2, 2, ir)
                    assert got.same_rows(want) and np.array_equal(got_idx, want_idx)
                    n_ir += 1
                elif m.nru > 0:
                    want, want_ret, _ = rh.call_entry(m, 1, None)
                    got, got_ret, _ = rh.call_entry(m, 1, app)
                    assert got_ret == want_ret
                    if got_ret == 0:
                        assert got.same_rows(want) and got.same_rows(ref)
                    else:
                        assert got.nrows == 0
                        n_bad += 1
                    n_app += 1
    finally:
        gpu_helpers.uninstall_backend()
    assert n_ir >= 5 and n_app >= 20 and n_bad >= 1


def test_gpu_very_long_column_folds_accumulators(backend):
    """A non-pivot column fed by 40 000 pivot rows: more than 2^15 products go into one
    64-bit accumulator pair, which exercises the overflow fold of the hot kernel."""
    p = 2147483629                       # largest 31-bit prime: worst case for the accumulators
    rng = np.random.default_rng(5)
    nru, nrl, ncr = 40000, 70, 3
    rows, cfs = [], []
    for i in range(nru):
        rows.append(np.array([i, nru + (i % 2)]))               # lead + one entry in column nru or nru+1
        cfs.append(np.array([1, p - 1 - (i % 7)]))
    for r in range(nrl):
        c = np.sort(rng.choice(nru + ncr, size=2000, replace=False))
        rows.append(c)
        cfs.append(rng.integers(p - 1000, p, size=len(c)))      # large residues
    row_ptr = np.concatenate([[0], np.cumsum([len(x) for x in rows])])
    m = FlatMatrix(p, nru, nrl, nru, ncr, row_ptr, np.concatenate(rows), np.arange(nru + nrl), row_ptr.copy(),
                   np.concatenate(cfs).astype(np.uint32))
    out, st = backend.reduce(m)
    ref, _, _ = oracle.reduce(m)
    assert out.same_rows(ref) and out.nrows == 3


@pytest.mark.skipif(not rh.available(), reason="compiled reference not shipped")
@pytest.mark.parametrize("name", ["katsura-11", "eco-12", "randquad-10"])
def test_gpu_backend_inside_reference_f4_large(name):
    """Larger members of the BASELINE families (config 2 and 4 shapes): all five LA pointers on
    the GPU, reduced Groebner basis bit-exact against the reference's (tests/golden/gb_digests.json)."""
    import gpu_helpers
    p = P31
    L = gpu_helpers.install_backend()
    try:
        gb = rh.run_f4(systems.by_name(name, p), p, threads=8, laopt=2)
    finally:
        gpu_helpers.uninstall_backend()
    want = golden_io.gb_digests()[f"{name}@{p}"]
    assert (len(gb.lens), int(gb.lens.sum())) == (want["nelts"], want["nterms"])
    assert gb.digest() == want["digest"]


@pytest.mark.parametrize("mode", ["batched", "uncached", "columnwise"])
def test_gpu_elimination_fallbacks(monkeypatch, mode):
    """Tall trailing blocks fall back from the shared-memory-cached batches to batches that re-read
    the pivot columns from L2 ("uncached"), and to one find + one elimination pass per pivot
    ("columnwise"); force each path and check it against the reference fixtures."""
    from msolve_b200.backend import Backend
    monkeypatch.setenv("F4LA_B200_GE", mode)
    be = Backend(0)
    try:
        n = 0
        for path in golden_io.fixture_files()[:4]:
            for m, ref, meta in golden_io.load_cases(path):
                out, st = be.reduce(m)
                assert out.same_rows(ref)
                n += 1
        assert n > 20
    finally:
        be.close()


def test_gpu_compress_and_verify_paths(monkeypatch):
    """Tall trailing blocks are eliminated on random row combinations and the result is verified exactly
    (k_ge_compress / k_ge_verify).  With a small first attempt ($F4LA_B200_COMPRESS_ROWS=64) small matrices
    reach every branch: accepted, escalated (rank too close to the number of combinations), rejected
    verification (forced) and "no attempt fits" -- always the reference's rows."""
    from msolve_b200.backend import Backend
    monkeypatch.setenv("F4LA_B200_COMPRESS_ROWS", "64")
    rng = np.random.default_rng(4242)
    cases = [(P31, np.uint32, 120, 700, 40, 1), (P31, np.uint32, 120, 700, 100, 1 | 4), (P16, np.uint16, 90, 700, 400, 4),
             (P8, np.uint8, 50, 300, 30, 1), (P31, np.uint32, 0, 600, 45, 1)]
    mats = [(random_matrix(rng, p, nru, nrl, ncr, 0.1, dt), want) for p, dt, nru, nrl, ncr, want in cases]
    be = Backend(0)
    try:
        for m, want in mats:
            out, st = be.reduce(m)
            ref, _, _ = oracle.reduce(m)
            assert out.same_rows(ref)
            assert st["echelon_path"] == want, (m.nru, m.nrl, m.ncr, st["echelon_path"], st["rank"])
    finally:
        be.close()
    monkeypatch.setenv("F4LA_B200_COMPRESS_FAULT", "1")
    be = Backend(0)
    try:
        for m, want in mats[:2]:
            out, st = be.reduce(m)
            ref, _, _ = oracle.reduce(m)
            assert out.same_rows(ref)
            assert st["echelon_path"] & 2 and not st["echelon_path"] & 1
    finally:
        be.close()


def big_sparse_matrix(rng, p, nru, nrl, ncr, per_row):
    """[A B; C D] with about nru * per_row non-zeros (vectorised, for matrices beyond 10^6 non-zeros)."""
    nc = nru + ncr
    rows, cfs = [], []
    for i in range(nru):
        k = min(per_row, nc - i - 1)
        others = i + 1 + rng.choice(nc - i - 1, size=k, replace=False)
        rows.append(np.concatenate([[i], others]))
        cfs.append(np.concatenate([[1], rng.integers(1, p, size=k)]))
    for r in range(nrl):
        c = np.sort(rng.choice(nc, size=min(3 * per_row, nc), replace=False))
        rows.append(c)
        cfs.append(rng.integers(1, p, size=len(c)))
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
    return FlatMatrix(p, nru, nrl, nru, ncr, row_ptr, np.concatenate(rows), np.arange(nru + nrl), row_ptr.copy(),
                      np.concatenate(cfs).astype(np.uint32))


def test_gpu_host_copy_variants_agree(monkeypatch):
    """f4la_b200_reduce() from host buffers copies the lower rows on a second stream while the pull lists are
    built (matrices beyond 2^20 non-zeros), or takes the column indices staged piecewise beforehand
    (f4la_b200_stage_cols); both must give the rows of the plain single-copy path and of the resident path."""
    from msolve_b200.backend import Backend
    rng = np.random.default_rng(777)
    m = big_sparse_matrix(rng, P31, 6000, 300, 400, 200)
    assert m.nnz >= 1 << 20
    be = Backend(0)
    try:
        pm = be.pin(m)
        split, _ = be.reduce(pm.matrix)                      # default: split copy
        staged, _ = be.reduce_staged(pm, pieces=5)
        rm = be.upload(m)
        resident, _ = be.reduce_resident(rm)
        rm.release()
        pm.free()
    finally:
        be.close()
    monkeypatch.setenv("F4LA_B200_SPLIT_H2D", "0")
    be = Backend(0)
    try:
        single, _ = be.reduce(m)
    finally:
        be.close()
    assert single.nrows > 0
    assert split.same_rows(single) and staged.same_rows(single) and resident.same_rows(single)
    ref, _, _ = oracle.reduce(m)
    assert single.same_rows(ref)


def test_gpu_more_edge_cases(backend):
    """p = 2, empty lower rows, duplicate lower rows, a lower-row count beyond the shared-memory
    limit of the batched elimination (column-at-a-time fallback), resident-matrix API."""
    rng = np.random.default_rng(99)
    # p = 2 and p = 3
    for p in (2, 3):
        m = random_matrix(rng, p, 120, 75, 31, 0.2, np.uint8)
        out, _ = backend.reduce(m)
        ref, _, _ = oracle.reduce(m)
        assert out.same_rows(ref), p
    # empty and duplicate lower rows
    p = P16
    rows = [np.array([0, 3, 5]), np.array([1, 4]), np.array([2, 3, 4, 5])]       # pivots 0..2, cols 3..5 free
    cfs = [np.array([1, 7, 9]), np.array([1, 11]), np.array([1, 2, 3, 4])]
    lower = [np.array([], dtype=int), np.array([0, 1, 5]), np.array([0, 1, 5]), np.array([4]), np.array([], dtype=int)]
    lcf = [np.array([], dtype=int), np.array([5, 6, 7]), np.array([5, 6, 7]), np.array([9]), np.array([], dtype=int)]
    allrows, allcf = rows + lower, cfs + lcf
    row_ptr = np.concatenate([[0], np.cumsum([len(x) for x in allrows])])
    m = FlatMatrix(p, 3, 5, 3, 3, row_ptr, np.concatenate(allrows), np.arange(8), row_ptr.copy(),
                   np.concatenate(allcf).astype(np.uint16))
    out, _ = backend.reduce(m)
    ref, _, _ = oracle.reduce(m)
    assert out.same_rows(ref) and out.nrows == 2
    # 60 000 lower rows: columns do not fit 216 KB of shared memory -> per-pivot fallback
    p = P31
    nrl, ncr = 60000, 6
    lower = [np.sort(rng.choice(ncr, size=rng.integers(1, 4), replace=False)) for _ in range(nrl)]
    lcf = [rng.integers(1, p, size=len(c)) for c in lower]
    row_ptr = np.concatenate([[0], np.cumsum([len(x) for x in lower])])
    m = FlatMatrix(p, 0, nrl, 0, ncr, row_ptr, np.concatenate(lower), np.arange(nrl), row_ptr.copy(),
                   np.concatenate(lcf).astype(np.uint32))
    out, _ = backend.reduce(m)
    ref, _, _ = oracle.reduce(m)
    assert out.same_rows(ref) and out.nrows == ncr
    # resident API gives the same rows as the host-buffer API
    m = random_matrix(rng, P31, 300, 200, 64, 0.05, np.uint32)
    a, _ = backend.reduce(m)
    rm = backend.upload(m)
    b, st = backend.reduce_resident(rm)
    rm.release()
    assert a.same_rows(b) and st["hot_kernel_launches"] >= 1 and st["ms_hot_kernel"] > 0


def test_gpu_two_host_threads_two_contexts():
    """The backend is re-entrant per context (the QQ loop runs one F4 instance per host thread,
    msolve.c:1815-1824): two threads, each with its own context, reduce different matrices."""
    import threading
    from msolve_b200.backend import Backend
    rng = np.random.default_rng(3)
    mats = [random_matrix(rng, P31, 400, 260, 90, 0.04, np.uint32) for _ in range(6)]
    refs = [oracle.reduce(m)[0] for m in mats]
    errors = []

    def work(k):
        try:
            be = Backend(0)
            for rep in range(3):
                for i in range(k, len(mats), 2):
                    out, _ = be.reduce(mats[i])
                    if not out.same_rows(refs[i]):
                        errors.append((k, i))
            be.close()
        except Exception as e:      # pragma: no cover
            errors.append(repr(e))

    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors


@pytest.mark.skipif(not rh.available(), reason="compiled reference not shipped")
def test_gpu_pivot_by_lead_random_independent_rows(backend):
    """F4LA_FLAG_PIVOT_BY_LEAD on matrices that are not from an F4 run: echelonised random rows with
    distinct leads (made by the oracle in echelon mode) are scrambled by adding multiples of later
    rows' tails; the final-step contract returns the reduced form of every row in place."""
    from msolve_b200.flat import FLAG_PIVOT_BY_LEAD
    rng = np.random.default_rng(17)
    for p, dt in ((P31, np.uint32), (P8, np.uint8)):
        src = random_matrix(rng, p, 0, 60, 90, 0.3, dt)
        ech, _, _ = oracle.reduce(src)                    # reduced rows, distinct leads, decreasing
        k = ech.nrows
        assert k > 20
        # rows with distinct leads that are NOT inter-reduced: row i += c * row j (lead_j > lead_i)
        dense = np.zeros((k, 90), dtype=np.int64)
        for i in range(k):
            a, b = int(ech.row_ptr[i]), int(ech.row_ptr[i + 1])
            dense[i, ech.cols[a:b]] = ech.cf[a:b]
        for i in range(k):                                 # row 0 has the largest lead
            for j in range(i):
                dense[i] = (dense[i] + int(rng.integers(1, p)) * dense[j]) % p
        # half of the rows become "known pivots keyed by lead", half stay lower rows
        order = rng.permutation(k)
        piv, low = sorted(order[: k // 2]), sorted(order[k // 2:])
        rows = [np.nonzero(dense[i])[0] for i in piv] + [np.nonzero(dense[i])[0] for i in low]
        cfs = [dense[i][np.nonzero(dense[i])[0]] for i in piv] + [dense[i][np.nonzero(dense[i])[0]] for i in low]
        row_ptr = np.concatenate([[0], np.cumsum([len(x) for x in rows])])
        m = FlatMatrix(p, len(piv), len(low), 0, 90, row_ptr, np.concatenate(rows), np.arange(k), row_ptr.copy(),
                       np.concatenate(cfs).astype(dt), flags=FLAG_PIVOT_BY_LEAD)
        # pivot rows must be monic with the lead first: they are (echelon rows are normalised, adding
        # multiples of rows with LARGER leads does not touch the lead)
        out, _ = backend.reduce(m)
        ref, _, _ = oracle.reduce(m)
        again, _, _, _ = rh.reference_la(m, laopt=2, threads=1)
        assert out.nrows == len(low) and out.same_rows(ref) and out.same_rows(again)
    # every column is a lead (no non-pivot column at all): each row reduces to its lead term
    p, k = P31, 12
    dense = np.triu(rng.integers(1, p, size=(k, k)), 1) + np.eye(k, dtype=np.int64)
    piv, low = [11, 9, 7, 3, 2, 0], [10, 8, 6, 5, 4, 1]      # the reference's order: a row only contains leads of earlier rows
    rows = [np.nonzero(dense[i])[0] for i in piv + low]
    cfs = [dense[i][np.nonzero(dense[i])[0]] for i in piv + low]
    row_ptr = np.concatenate([[0], np.cumsum([len(x) for x in rows])])
    m = FlatMatrix(p, len(piv), len(low), 0, k, row_ptr, np.concatenate(rows), np.arange(k), row_ptr.copy(),
                   np.concatenate(cfs).astype(np.uint32), flags=FLAG_PIVOT_BY_LEAD)
    out, _ = backend.reduce(m)
    ref, _, _ = oracle.reduce(m)
    assert out.same_rows(ref)
    assert out.cols.tolist() == low and out.cf.tolist() == [1] * len(low)


def low_rank_matrix(rng, p, nru, nrl, ncr, rank, dt):
    """[A B; C D] whose lower rows are random combinations of `rank` sparse base rows."""
    base = random_matrix(rng, p, nru, rank, ncr, 0.05, dt)
    nc = nru + ncr
    dense = np.zeros((rank, nc), dtype=object)
    for i in range(rank):
        c, v = base.row(nru + i)
        dense[i, c.astype(int)] = [int(x) for x in v]
    rows, cfs = [], []
    for i in range(nru):
        c, v = base.row(i)
        rows.append(c.astype(int)); cfs.append(v.astype(np.int64))
    for r in range(nrl):
        k = rng.choice(rank, size=3, replace=False)
        row = np.zeros(nc, dtype=object)
        for kk in k:
            row = (row + int(rng.integers(1, p)) * dense[kk]) % p
        nz = np.nonzero(row)[0]
        rows.append(nz); cfs.append(np.array([int(row[j]) for j in nz], dtype=np.int64))
    row_ptr = np.concatenate([[0], np.cumsum([len(x) for x in rows])])
    return FlatMatrix(p, nru, nrl, nru, ncr, row_ptr, np.concatenate(rows), np.arange(nru + nrl), row_ptr.copy(),
                      np.concatenate(cfs).astype(dt))


def test_gpu_probabilistic_mode_random_combinations(backend):
    """F4LA_FLAG_PROBABILISTIC: 2 000 lower rows of rank <= 120 are replaced by 512 random
    combinations (rank + slack < 512: accepted), 1 800 rows of rank 600 force the escalation to the
    exact path; both give the oracle's rows, for several seeds."""
    from msolve_b200.flat import FLAG_PROBABILISTIC
    rng = np.random.default_rng(21)
    for rank, nrl in ((120, 2000), (600, 1800)):
        m = low_rank_matrix(rng, P31, 150, nrl, 700, rank, np.uint32)
        ref, _, _ = oracle.reduce(m)
        assert ref.nrows > rank // 2
        m.flags = FLAG_PROBABILISTIC
        for seed in (0, 1, 12345, 2 ** 63 + 7):
            backend.set_seed(seed)
            out, st = backend.reduce(m)
            assert out.same_rows(ref), (rank, seed)
    backend.set_seed(0)
    # small prime: the flag is ignored (exact path), as the reference's CLI does below 2^15
    m = low_rank_matrix(rng, P8, 60, 1700, 200, 40, np.uint8)
    ref, _, _ = oracle.reduce(m)
    m.flags = FLAG_PROBABILISTIC
    out, _ = backend.reduce(m)
    assert out.same_rows(ref)


@pytest.mark.skipif(not rh.available(), reason="compiled reference not shipped")
@pytest.mark.parametrize("name,laopt", [("katsura-11", 44), ("cyclic-8", 42), ("eco-12", 44), ("katsura-10", 43)])
def test_gpu_backend_probabilistic_la_options(name, laopt):
    """-l 42/43/44 through the drop-in: same reduced Groebner basis as the reference's exact run
    (the property the reference's own diff tests assert for -l 44, test/diff/diff_source.sh)."""
    import gpu_helpers
    L = gpu_helpers.install_backend()
    try:
        gb = rh.run_f4(systems.by_name(name, P31), P31, threads=8, laopt=laopt)
    finally:
        gpu_helpers.uninstall_backend()
    want = golden_io.gb_digests()[f"{name}@{P31}"]
    assert gb.digest() == want["digest"]
