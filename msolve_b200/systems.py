"""Formula-generated polynomial systems (cyclic-n, katsura-n, eco-n, random
dense quadratics) in the flat encoding neogb's ``export_f4`` takes
(reference src/neogb/f4.c:769-793): per generator a length, per term one int32
coefficient and ``nvars`` int32 exponents.

The formulas reproduce the reference's own input files term for term
(input_files/cyclic5-31.ms, kat6-31.ms, eco6-31.ms).
Terms inside a polynomial are emitted in an arbitrary order -- neogb sorts
them on import (reference src/neogb/io.c import_input_data).
"""
from __future__ import annotations

import itertools
import random
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

Monomial = Tuple[int, ...]
Poly = Dict[Monomial, int]


@dataclass
class PolySystem:
    name: str
    nvars: int
    polys: List[Poly]

    def encode(self, p: int):
        """-> (lens int32[ngens], exps int32[nterms*nvars], cfs int32[nterms]) mod p."""
        lens, exps, cfs = [], [], []
        for f in self.polys:
            terms = [(m, c % p) for m, c in f.items() if c % p != 0]
            lens.append(len(terms))
            for m, c in terms:
                exps.extend(m)
                cfs.append(c)
        return (np.asarray(lens, dtype=np.int32), np.asarray(exps, dtype=np.int32),
                np.asarray(cfs, dtype=np.int32))


def encode_integers(system: "PolySystem"):
    """Signed integer coefficients as they stand (rational input of the QQ flow)."""
    lens, exps, cfs = [], [], []
    for f in system.polys:
        lens.append(len(f))
        for m, c in f.items():
            exps.extend(m)
            cfs.append(c)
    return (np.asarray(lens, dtype=np.int32), np.asarray(exps, dtype=np.int32),
            np.asarray(cfs, dtype=np.int32))


def _add(f: Poly, m: Monomial, c: int) -> None:
    v = f.get(m, 0) + c
    if v:
        f[m] = v
    else:
        f.pop(m, None)


def _mono(nvars: int, *idx: int) -> Monomial:
    e = [0] * nvars
    for i in idx:
        e[i] += 1
    return tuple(e)


def cyclic(n: int) -> PolySystem:
    polys: List[Poly] = []
    for k in range(1, n):
        f: Poly = {}
        for i in range(n):
            _add(f, _mono(n, *[(i + j) % n for j in range(k)]), 1)
        polys.append(f)
    f = {}
    _add(f, _mono(n, *range(n)), 1)
    _add(f, _mono(n), -1)
    polys.append(f)
    return PolySystem(f"cyclic-{n}", n, polys)


def katsura(n: int) -> PolySystem:
    """katsura-n in the formula convention: n+1 variables u_0..u_n."""
    nv = n + 1
    u = lambda l: abs(l) if abs(l) <= n else None
    polys: List[Poly] = []
    f: Poly = {}
    for l in range(-n, n + 1):
        _add(f, _mono(nv, u(l)), 1)
    _add(f, _mono(nv), -1)
    polys.append(f)
    for m in range(n):
        f = {}
        for l in range(-n, n + 1):
            a, b = u(l), u(m - l)
            if a is None or b is None:
                continue
            _add(f, _mono(nv, a, b), 1)
        _add(f, _mono(nv, m), -1)
        polys.append(f)
    return PolySystem(f"katsura-{n}", nv, polys)


def eco(n: int) -> PolySystem:
    polys: List[Poly] = []
    for k in range(1, n):
        f: Poly = {}
        _add(f, _mono(n, k - 1, n - 1), 1)
        for i in range(1, n - k):
            _add(f, _mono(n, i - 1, i + k - 1, n - 1), 1)
        _add(f, _mono(n), -k)
        polys.append(f)
    f = {}
    for i in range(n - 1):
        _add(f, _mono(n, i), 1)
    _add(f, _mono(n), 1)
    polys.append(f)
    return PolySystem(f"eco-{n}", n, polys)


def randquad(n: int, p: int, seed: int = 1) -> PolySystem:
    rng = random.Random(seed)
    monos = [_mono(n, *c) for d in (0, 1, 2)
             for c in itertools.combinations_with_replacement(range(n), d)]
    polys = [{m: rng.randrange(1, p) for m in monos} for _ in range(n)]
    return PolySystem(f"randquad-{n}", n, polys)


def by_name(name: str, p: int = 1073741827) -> PolySystem:
    fam, _, arg = name.partition("-")
    n = int(arg)
    if fam == "cyclic":
        return cyclic(n)
    if fam == "katsura":
        return katsura(n)
    if fam == "eco":
        return eco(n)
    if fam == "randquad":
        return randquad(n, p)
    raise ValueError(f"unknown system family {name!r}")
