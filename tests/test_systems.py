"""The formula generators reproduce the reference's own input files term for term
(input_files/cyclic5-31.ms, kat6-31.ms, eco6-31.ms).  Needs /root/reference, i.e. runs in
the build container only; the parsed polynomials are also kept as a small fixture so the
check still runs elsewhere."""
import json
import os
import re

import pytest

from msolve_b200 import systems

REF_INPUTS = "/root/reference/input_files"
FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_input_files.json")
CASES = {"cyclic5-31.ms": ("cyclic", 5), "kat6-31.ms": ("katsura", 5), "eco6-31.ms": ("eco", 6)}


def parse_ms(text):
    """msolve .ms format (src/msolve/iofiles.c:1015 get_data_from_file): variables, characteristic,
    comma separated polynomials.  -> (vars, p, [ {exponent tuple: int coefficient} ])"""
    lines = [l.strip() for l in text.strip().splitlines() if l.strip()]
    names = [v.strip() for v in lines[0].split(",")]
    p = int(lines[1])
    polys = []
    for src in "".join(lines[2:]).split(","):
        if not src:
            continue
        f = {}
        for sign, term in re.findall(r"([+-]?)([^+-]+)", src):
            coef, exps = (-1 if sign == "-" else 1), [0] * len(names)
            for fac in term.split("*"):
                if re.fullmatch(r"\d+", fac):
                    coef *= int(fac)
                else:
                    v, _, e = fac.partition("^")
                    exps[names.index(v)] += int(e) if e else 1
            f[tuple(exps)] = f.get(tuple(exps), 0) + coef
        polys.append(f)
    return names, p, polys


def canonical(polys, p):
    return sorted(sorted([list(m), c % p] for m, c in f.items() if c % p) for f in polys)


def test_generators_match_reference_input_files():
    if os.path.isdir(REF_INPUTS):
        parsed = {}
        for fname in CASES:
            names, p, polys = parse_ms(open(os.path.join(REF_INPUTS, fname)).read())
            parsed[fname] = {"p": p, "nvars": len(names), "polys": canonical(polys, p)}
        if not os.path.exists(FIXTURE) or json.load(open(FIXTURE)) != parsed:
            with open(FIXTURE, "w") as f:
                json.dump(parsed, f)
    if not os.path.exists(FIXTURE):
        pytest.skip("no reference input files and no fixture")
    parsed = json.load(open(FIXTURE))
    for fname, (fam, n) in CASES.items():
        want = parsed[fname]
        s = getattr(systems, fam)(n)
        assert s.nvars == want["nvars"]
        assert canonical(s.polys, want["p"]) == want["polys"], fname


def test_encode_shapes():
    s = systems.randquad(4, 65521)
    lens, exps, cfs = s.encode(65521)
    assert len(lens) == 4 and lens.sum() == len(cfs) and len(exps) == len(cfs) * 4
    assert (cfs > 0).all() and (cfs < 65521).all()
    li, ei, ci = systems.encode_integers(systems.katsura(3))
    assert (ci < 0).any() and li.sum() == len(ci)
