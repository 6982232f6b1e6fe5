"""GPU parity tests for the BASELINE configurations themselves (run on the B200 box: pytest -m gpu).

  config 2  katsura-12 mod 2^30+3: every linear-algebra call of the run against the rows the
            unmodified reference produced for it (reference-served recording), plus the final basis;
  config 3  cyclic-8 mod 65521 / 251 with the probabilistic options (-l 44 / 42; la_ff_16.c:720,
            la_ff_8.c:868) next to the exact ones already covered in test_gpu_parity.py;
  config 4  eco-14 and random dense quadratics n = 12 (large trailing blocks);
  and the trace_linear_algebra entry point (data.h:571-576) on mat_t + trace_t.

Everything computes through the C ABI of csrc/libf4la_b200.so; the compiled reference and the
plain-C oracle are only the checkers."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_io
from msolve_b200 import systems
from msolve_b200.flat import FLAG_WANT_RBA, FlatMatrix
from oracle import refharness as rh

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rh.available(), reason="compiled reference not shipped")]

P31, P16, P8 = 1073741827, 65521, 251


def _f4_on_gpu(name, p, laopt=2, threads=None):
    import gpu_helpers
    L = gpu_helpers.install_backend()
    try:
        assert rh.hooked_pointers() == 0b11111          # all five pointers of data.h:529-595
        gb = rh.run_f4(systems.by_name(name, p), p, threads=threads or min(16, os.cpu_count() or 8), laopt=laopt)
        tot = gpu_helpers.neogb_totals(L)
    finally:
        gpu_helpers.uninstall_backend()
    return gb, tot


def test_gpu_katsura12_every_la_call_vs_reference(backend):
    """BASELINE config 2.  The reference F4 runs with its own linear algebra and records the input
    and the output of every call (MODE_RECORD: reference-served); the GPU must return the same rows
    for each of them, the final reduction step included."""
    threads = os.cpu_count() or 8
    gb, recs = rh.record_f4(systems.by_name("katsura-12", P31), P31, threads=threads)
    want = golden_io.gb_digests()[f"katsura-12@{P31}"]
    assert gb.digest() == want["digest"]               # the recording run is the committed reference run
    big = [r for r in recs if r.kind == 0 and r.matrix.nnz > 1 << 20]
    assert len(big) >= 9 and any(r.kind == 1 for r in recs)
    assert max(r.matrix.nru for r in recs) > 50000
    for r in recs:
        out, st = backend.reduce(r.matrix)
        assert out.same_rows(r.result), (r.kind, r.matrix.nru, r.matrix.nrl, r.matrix.ncr)
        assert st["kernel_launches"] > 0


@pytest.mark.parametrize("name", ["katsura-12", "eco-14", "randquad-12"])
def test_gpu_baseline_configs_inside_reference_f4(name):
    """BASELINE configs 2 and 4 through the drop-in pointers: reduced Groebner basis bit-exact against
    the unmodified reference's (tests/golden/gb_digests.json, made by tests/golden/make_digests.py)."""
    gb, tot = _f4_on_gpu(name, P31)
    want = golden_io.gb_digests()[f"{name}@{P31}"]
    assert (len(gb.lens), int(gb.lens.sum())) == (want["nelts"], want["nterms"])
    assert gb.digest() == want["digest"]
    assert tot["calls"] > 10 and tot["kernel_launches"] > 0


@pytest.mark.parametrize("p,laopt", [(P16, 44), (P16, 42), (P16, 1), (P8, 44), (P8, 42)])
def test_gpu_cyclic8_small_primes_probabilistic_options(p, laopt):
    """BASELINE config 3: cyclic-8 over 16- and 8-bit primes with -l 44 / 42 / 1.  65521 > 2^15, so the
    16-bit runs take the random-combination path of the device; for 251 the flag is ignored like in the
    reference's CLI (main.c:550-556).  Same reduced basis as the reference's exact run."""
    gb, tot = _f4_on_gpu("cyclic-8", p, laopt=laopt)
    want = golden_io.gb_digests()[f"cyclic-8@{p}"]
    assert gb.digest() == want["digest"]
    assert tot["kernel_launches"] > 0


def test_gpu_probabilistic_16bit_flat(backend):
    """F4LA_FLAG_PROBABILISTIC on a 16-bit matrix large enough for the random combinations to be used."""
    from msolve_b200.flat import FLAG_PROBABILISTIC
    from oracle import oracle
    from test_gpu_parity import low_rank_matrix
    rng = np.random.default_rng(8)
    m = low_rank_matrix(rng, P16, 140, 1900, 600, 100, np.uint16)
    ref, _, _ = oracle.reduce(m)
    m.flags = FLAG_PROBABILISTIC
    for seed in (0, 7, 2 ** 40 + 1):
        backend.set_seed(seed)
        out, st = backend.reduce(m)
        assert out.same_rows(ref), seed
    backend.set_seed(0)


def _trace_fn():
    import gpu_helpers
    L = gpu_helpers.install_backend()
    return C.cast(L.f4la_b200_trace_linear_algebra, C.c_void_p).value


def test_gpu_trace_linear_algebra_entry(backend):
    """trace_linear_algebra (data.h:571-576; la_ff_32.c:4075/3004): echelon rows, reducer bit arrays and
    the trace entry the reference's construct_trace() builds from what the backend leaves in mat_t.
      * rows: identical to the reference's own trace variant (32- and 8-bit; its 16-bit twin reads
        bs->cf_16[COEFFS] after remapping COEFFS, la_ff_16.c:927-932 vs :455-457, and crashes -- there the
        rows are compared with the exact echelon form);
      * kept rows: as many as new pivots, and replaying ONLY them gives the same rows (what the
        application phase relies on);
      * reducers: exactly the pivots with a non-zero multiplier in some kept row, and the per-row bit
        arrays (renumbered over the used reducers by construct_trace) match the multipliers."""
    import gpu_helpers
    fn = _trace_fn()
    n = n_ref = 0
    try:
        for path in golden_io.fixture_files():
            for m, ref, meta in golden_io.load_cases(path):
                if meta["kind"] != 0 or m.nru == 0:
                    continue
                got, kept, reds, bits = rh.call_trace_entry(m, fn)
                assert got.same_rows(ref), path
                if m.cf_bytes != 2:
                    want, wkept, wreds, _ = rh.call_trace_entry(m, None)
                    assert got.same_rows(want) and len(kept) == len(wkept)
                    n_ref += 1
                assert len(kept) == got.nrows and len(set(kept.tolist())) == len(kept)
                assert np.all(kept >= 0) and np.all(kept < m.nrl) and np.all(np.diff(kept) > 0)
                if got.nrows == 0:
                    continue
                # replay: the kept lower rows alone span the same new pivots
                rows = list(range(m.nru)) + [m.nru + int(k) for k in kept]
                rp = np.concatenate([[0], np.cumsum([m.row_ptr[r + 1] - m.row_ptr[r] for r in rows])])
                cols = np.concatenate([m.cols[m.row_ptr[r]:m.row_ptr[r + 1]] for r in rows])
                sub = FlatMatrix(m.fc, m.nru, len(kept), m.ncl, m.ncr, rp, cols, m.row_cf[rows], m.cf_ptr, m.cf)
                again, _ = backend.reduce(sub)
                assert again.nrows == len(kept) and again.same_rows(ref)
                # reducer bits against the multipliers (flat API, F4LA_FLAG_WANT_RBA)
                m.flags = FLAG_WANT_RBA
                full, _ = backend.reduce(m)
                m.flags = 0
                used = np.zeros(m.nru, dtype=bool)
                per_row = []
                for k in kept:
                    b = np.array([(int(full.rba[k, i // 32]) >> (i % 32)) & 1 for i in range(m.nru)], dtype=bool)
                    per_row.append(b)
                    used |= b
                assert reds.tolist() == np.nonzero(used)[0].tolist()
                for t, b in enumerate(per_row):
                    tb = [(int(bits[t, j // 32]) >> (j % 32)) & 1 for j in range(len(reds))]
                    assert tb == b[reds].astype(int).tolist()
                n += 1
    finally:
        gpu_helpers.uninstall_backend()
    assert n >= 25 and n_ref >= 20


def test_gpu_side_channels_match_reference():
    """What the wrappers leave in md_t / mat_t besides the rows (la_ff_32.c:3978-3989, :1214-1216):
    st->np, mat->np/nr/sz, num_zerored, the LA timers and the application_nr_* counters msolve prints as
    Gops/sec (msolve.c:2666-2673).  The device counts the reference's units for the reduction by the known
    pivots exactly; the reference adds the (order dependent) applications of the new pivots on top."""
    import gpu_helpers
    L = gpu_helpers.install_backend()
    la = C.cast(L.f4la_b200_linear_algebra, C.c_void_p).value
    n = 0
    try:
        for path in [f for f in golden_io.fixture_files() if "1073741827" in f]:
            for m, ref, meta in golden_io.load_cases(path):
                if meta["kind"] != 0 or m.nru == 0:
                    continue
                want, _, _ = rh.call_entry(m, 0, None, threads=1)
                ws = rh.last_side()
                got, _, _ = rh.call_entry(m, 0, la, threads=1)
                gs = rh.last_side()
                assert got.same_rows(want)
                for k in ("st_np", "mat_np", "mat_nr", "mat_sz", "num_zerored"):
                    assert gs[k] == ws[k], (k, gs[k], ws[k])
                assert gs["la_rtime"] > 0 and gs["la_ctime"] >= 0
                # units: known-pivot part exact => never above the reference's total, and the bulk of it
                assert gs["application_nr_mult"] == gs["application_nr_add"]
                assert gs["application_nr_mult"] <= ws["application_nr_mult"] * (1 + 1e-9) + 1e-9
                assert gs["application_nr_red"] <= ws["application_nr_red"]
                if ws["application_nr_mult"] > 5:       # thousands of units: matrices with real work
                    # small matrices spend half of the reference's units on the new pivots; on katsura-12 the
                    # known-pivot part is 93-100 % of the total (bench.py --per-matrix)
                    assert gs["application_nr_mult"] >= 0.3 * ws["application_nr_mult"]
                    n += 1
    finally:
        gpu_helpers.uninstall_backend()
    assert n >= 5


@pytest.mark.parametrize("name", ["katsura-7", "cyclic-6", "eco-8", "katsura-9"])
def test_gpu_qq_tape_replay_matches_reference(name):
    """Structure-cached replay of the tracer application phase (f4la_b200_neogb.h): one application run through the
    drop-in pointers is recorded, then further primes are replayed from the tape with no symbolic work.  The
    residues (coefficients of the reduced modular basis) must equal the reference's own application run for
    every prime."""
    import gpu_helpers
    from msolve_b200.backend import Backend, Tape
    system = systems.by_name(name)
    primes = rh.primes_above(1 << 30, 12)
    rh.reset(); rh.set_mode(rh.MODE_REFERENCE)
    rh.qq_setup(system, threads=2)
    learn = rh.qq_learn(primes[0])
    assert learn["rc"] == 0
    ref, _, ref_resid = rh.qq_apply(primes[1:], threads=2, residue_words=learn["nterms"])
    assert all(r["err"] == 0 for r in ref)
    L = gpu_helpers.install_backend()
    be = Backend(0)
    tape = None
    try:
        rh.qq_setup(system, threads=1)
        assert rh.qq_learn(primes[0])["rc"] == 0
        Tape.begin(L)
        got, _, got_resid = rh.qq_apply(primes[1:2], threads=1, ngpus=0, residue_words=learn["nterms"])
        tape = Tape.end(L)
        assert tape is not None and tape.calls >= 3 and tape.residue_words == learn["nterms"]
        assert np.array_equal(got_resid[0], ref_resid[0])
        for k, p in enumerate(primes[1:]):
            cf, off = rh.qq_initial_basis(p)
            assert len(off) - 1 == tape.initial_elements
            rc, resid, ms = tape.replay(be, p, cf, off)
            assert rc == 0, (name, p, rc)
            assert np.array_equal(resid, ref_resid[k]), (name, p)
        # several host threads, one context each, replaying different primes of the same tape concurrently
        import threading
        inits = [rh.qq_initial_basis(p) for p in primes[1:]]
        bes = [Backend(0) for _ in range(3)]
        bad = []

        def work(t):
            for rep in range(2):
                for k in range(t, len(inits), 3):
                    rc, resid, _ = tape.replay(bes[t], primes[1 + k], *inits[k])
                    if rc != 0 or not np.array_equal(resid, ref_resid[k]):
                        bad.append((t, k, rc))

        ts = [threading.Thread(target=work, args=(t,)) for t in range(3)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        [b.close() for b in bes]
        assert not bad, bad
        # device-resident replay: primes are only enqueued, one wait per batch; twice, so that both the warm-up replay of
        # the context and the queued ones are covered
        be2 = Backend(0)
        for rep in range(2):
            outs = np.zeros((len(inits), tape.residue_words), dtype=np.uint32)
            for k in range(len(inits)):
                tape.enqueue(be2, primes[1 + k], inits[k][0], inits[k][1], outs[k])
            rcs = tape.collect(be2)
            assert rcs == [0] * len(inits), (name, rcs)
            assert np.array_equal(outs, ref_resid), name
        # a prime whose structure differs must be reported, not silently replayed: with all coefficients zero every
        # row vanishes, so the rank of the first call already differs from the recorded one
        zeros = np.zeros_like(inits[0][0])
        out = np.zeros(tape.residue_words, dtype=np.uint32)
        tape.enqueue(be2, primes[1], zeros, inits[0][1], out)
        tape.enqueue(be2, primes[2], inits[1][0], inits[1][1], outs[1])
        rcs = tape.collect(be2)
        assert len(rcs) == 2 and rcs[0] != 0 and rcs[1] == 0, (name, rcs)
        assert np.array_equal(outs[1], ref_resid[1])
        be2.close()
    finally:
        if tape is not None:
            tape.free()
        be.close()
        gpu_helpers.uninstall_backend()
