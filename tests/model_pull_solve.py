"""Numpy model of the device algorithm (NOT a product path, NOT a fallback):
used by CPU tests to check the mathematics the CUDA kernels implement --
the transposed, level-scheduled multi-right-hand-side triangular solve and the
Gauss-Jordan elimination of the trailing block -- against the oracle.

Device algorithm (msolve_b200/csrc/f4la_kernels.cu):
  V[j] (a vector over the lower rows) for every column j:
     V[j] = (column j of [C D]) - sum_{pivot rows i with an entry a_ij, i<j} a_ij * V[i]
  for j < ncl this is the multiplier of pivot j in every lower row, for
  j >= ncl it is column j-ncl of the Schur complement D' = D - C A^-1 B.
  Columns are processed level by level (level[j] = 1 + max level[i]).
"""
from __future__ import annotations

import numpy as np

from msolve_b200.flat import FlatMatrix, FlatResult


def pull_structure(m: FlatMatrix):
    """CSC of the strictly-upper part of the pivot rows: for column j the list
    of (source pivot i, coefficient).  Returns (col_ptr, src, coef, level)."""
    rp = m.row_ptr.astype(np.int64)
    nru, ncl, nc = m.nru, m.ncl, m.ncl + m.ncr
    lens = np.diff(rp)[:nru]
    rows = np.repeat(np.arange(nru, dtype=np.int64), lens)
    cols = m.cols[: rp[nru]].astype(np.int64)
    # coefficient of entry k of row r: cf[cf_ptr[row_cf[r]] + k]
    k_in_row = np.arange(rp[nru], dtype=np.int64) - np.repeat(rp[:nru], lens)
    cf = m.cf[(m.cf_ptr[m.row_cf[:nru]].astype(np.int64)).repeat(lens) + k_in_row].astype(np.int64)
    keep = cols != rows          # drop the unit diagonal (lead term, coefficient 1)
    assert np.all(cf[~keep] == 1), "pivot rows must be monic"
    rows, cols, cf = rows[keep], cols[keep], cf[keep]
    assert np.all(cols > rows) or np.all((cols >= ncl) | (cols > rows))
    order = np.argsort(cols, kind="stable")
    rows, cols, cf = rows[order], cols[order], cf[order]
    col_ptr = np.zeros(nc + 1, dtype=np.int64)
    np.add.at(col_ptr, cols + 1, 1)
    col_ptr = np.cumsum(col_ptr)
    level = np.zeros(ncl, dtype=np.int64)
    for j in range(ncl):
        s = rows[col_ptr[j]:col_ptr[j + 1]]
        if len(s):
            level[j] = level[s].max() + 1
    return col_ptr, rows, cf, level


def schur_complement(m: FlatMatrix) -> np.ndarray:
    """D' as a dense (nrl x ncr) int64 matrix, entries in [0,p)."""
    p = m.fc
    nru, nrl, ncl, nc = m.nru, m.nrl, m.ncl, m.ncl + m.ncr
    col_ptr, src, coef, level = pull_structure(m)
    V = np.zeros((nc, nrl), dtype=np.int64)
    for r in range(nrl):
        c, v = m.row(nru + r)
        V[c.astype(np.int64), r] = v.astype(np.int64)
    nlev = int(level.max()) + 1 if ncl else 0
    order = np.argsort(level, kind="stable")
    bounds = np.searchsorted(level[order], np.arange(nlev + 1))
    def solve(j):
        a, b = col_ptr[j], col_ptr[j + 1]
        if a == b:
            return
        acc = V[j].astype(object)
        for k in range(a, b):
            acc = acc - int(coef[k]) * V[src[k]].astype(object)
        V[j] = np.array([int(x) % p for x in acc], dtype=np.int64)
    for l in range(nlev):
        for j in order[bounds[l]:bounds[l + 1]]:
            solve(int(j))
    for j in range(ncl, nc):
        solve(j)
    return V[ncl:].T.copy()


def rref_mod_p(D: np.ndarray, p: int):
    """Gauss-Jordan, columns left to right.  -> (rref rows [rank x ncr], pivot cols)."""
    D = D.astype(object) % p
    nr, ncr = D.shape
    used = np.zeros(nr, dtype=bool)
    piv_rows, piv_cols = [], []
    for c in range(ncr):
        cand = np.nonzero((D[:, c] != 0) & ~used)[0]
        if len(cand) == 0:
            continue
        r = int(cand[0])
        inv = pow(int(D[r, c]), -1, p)
        D[r] = (D[r] * inv) % p
        for r2 in np.nonzero(D[:, c] != 0)[0]:
            if r2 != r:
                D[r2] = (D[r2] - D[r2, c] * D[r]) % p
        used[r] = True
        piv_rows.append(r); piv_cols.append(c)
    return D, piv_rows, piv_cols


def echelon(m: FlatMatrix) -> FlatResult:
    """Model of the whole device pipeline in echelon mode."""
    p, ncl = m.fc, m.ncl
    Dp = schur_complement(m)
    R, prow, pcol = rref_mod_p(Dp, p)
    row_ptr, cols, cf = [0], [], []
    for r, c in sorted(zip(prow, pcol), key=lambda t: -t[1]):
        nz = np.nonzero(R[r] != 0)[0]
        cols.extend((nz + ncl).tolist()); cf.extend(int(x) for x in R[r][nz])
        row_ptr.append(len(cols))
    return FlatResult(np.asarray(row_ptr, dtype=np.uint64), np.asarray(cols, dtype=np.uint32),
                      np.asarray(cf, dtype=m.cf.dtype))
