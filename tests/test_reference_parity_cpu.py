"""CPU tests that need the compiled reference (oracle/_ref/libneogb_ref.so):
the oracle restatement against the reference itself on freshly recorded F4
runs, all three field widths, and the reference's own exact/probabilistic
equivalence (the property its diff tests assert, test/diff/diff_source.sh)."""
import pytest

import golden_io
from msolve_b200 import systems
from oracle import oracle
from oracle import refharness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="oracle/_ref/libneogb_ref.so not built")

P31, P16, P8 = 1073741827, 65521, 251


@pytest.mark.parametrize("name,p", [("cyclic-6", P31), ("cyclic-6", P16), ("cyclic-6", P8),
                                    ("katsura-6", P31), ("katsura-7", P16), ("eco-7", P8), ("randquad-5", P31)])
def test_oracle_vs_reference_recorded(name, p):
    gb, recs = rh.record_f4(systems.by_name(name, p), p, threads=1)
    assert recs
    for r in recs:
        out, units, apps = oracle.reduce(r.matrix)
        assert out.same_rows(r.result)
        # replaying the flat matrix through the reference LA gives the same rows again
        again, _, _, _ = rh.reference_la(r.matrix, laopt=2, threads=2)
        assert again.same_rows(r.result)


@pytest.mark.parametrize("name,p", [("cyclic-6", P31), ("katsura-7", P31), ("cyclic-7", P16), ("eco-8", P31)])
def test_reference_final_basis_is_independent_of_la_option(name, p):
    d = golden_io.gb_digests()
    rh.set_mode(rh.MODE_REFERENCE)
    digests = {rh.run_f4(systems.by_name(name, p), p, threads=t, laopt=l).digest()
               for l, t in ((2, 1), (44, 2), (1, 1), (42, 2))}
    assert len(digests) == 1
    key = f"{name}@{p}"
    if key in d:
        assert digests.pop() == d[key]["digest"]


def test_struct_layout_probe():
    lay = rh.struct_layout()
    assert lay[0] > 0 and lay[-7:] == [6, 5, 4, 3, 2, 1, 0]
