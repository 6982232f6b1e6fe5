"""Flat (CSR-like) Macaulay matrices: the host-side mirror of
``include/f4la_b200.h`` (``f4la_flat_matrix`` / ``f4la_flat_result``).

A :class:`FlatMatrix` keeps the numpy arrays alive and hands out the ctypes
struct the C ABI takes.  Column/coefficient conventions follow neogb's
``mat_t`` (reference src/neogb/data.h:203-232): rows ``0..nru-1`` are the known
pivot rows (``rr``), rows ``nru..nru+nrl-1`` the rows to be reduced (``tr``),
row ``r`` uses the shared coefficient array ``row_cf[r]``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

FLAG_NORMALFORM = 1
FLAG_PIVOT_BY_LEAD = 2
FLAG_WANT_RBA = 4
FLAG_PROBABILISTIC = 8
FLAG_COLS_STAGED = 16
FLAG_CF_STAGED = 32

_CF_DTYPE = {1: np.uint8, 2: np.uint16, 4: np.uint32}


class CFlatMatrix(C.Structure):
    _fields_ = [
        ("fc", C.c_uint32), ("cf_bytes", C.c_uint32),
        ("nru", C.c_uint32), ("nrl", C.c_uint32),
        ("ncl", C.c_uint32), ("ncr", C.c_uint32),
        ("ncf", C.c_uint32), ("flags", C.c_uint32),
        ("row_ptr", C.c_void_p), ("cols", C.c_void_p), ("row_cf", C.c_void_p),
        ("cf_ptr", C.c_void_p), ("cf", C.c_void_p),
    ]


class CFlatResult(C.Structure):
    _fields_ = [
        ("nrows", C.c_uint32), ("cf_bytes", C.c_uint32),
        ("row_ptr", C.c_void_p), ("cols", C.c_void_p), ("cf", C.c_void_p),
        ("src_row", C.c_void_p), ("rba", C.c_void_p), ("rba_words", C.c_uint32), ("reserved", C.c_uint32),
    ]


class CStats(C.Structure):
    _fields_ = [
        ("ms_total", C.c_double), ("ms_h2d", C.c_double), ("ms_prep", C.c_double),
        ("ms_reduce", C.c_double), ("ms_echelon", C.c_double), ("ms_d2h", C.c_double),
        ("units", C.c_uint64), ("pivot_applications", C.c_uint64),
        ("bytes_h2d", C.c_uint64), ("bytes_d2h", C.c_uint64),
        ("kernel_launches", C.c_uint32), ("rank", C.c_uint32),
        ("ms_hot_kernel", C.c_double), ("hot_kernel_launches", C.c_uint32), ("echelon_path", C.c_uint32),
        ("dense_macs", C.c_uint64), ("ref_applications", C.c_uint64),
        ("ms_schur", C.c_double), ("schur_macs", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _np_from_ptr(ptr, n, dtype):
    """Copy ``n`` items of ``dtype`` from a raw address into a fresh array."""
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = int(n) * np.dtype(dtype).itemsize
    buf = (C.c_ubyte * nbytes).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype, count=int(n)).copy()


@dataclass
class FlatMatrix:
    fc: int
    nru: int
    nrl: int
    ncl: int
    ncr: int
    row_ptr: np.ndarray   # uint64 [nru+nrl+1]
    cols: np.ndarray      # uint32 [nnz]
    row_cf: np.ndarray    # uint32 [nru+nrl]
    cf_ptr: np.ndarray    # uint64 [ncf+1]
    cf: np.ndarray        # uint8/16/32 [ncoef]
    flags: int = 0
    meta: dict = field(default_factory=dict)

    def __post_init__(self):
        self.row_ptr = np.ascontiguousarray(self.row_ptr, dtype=np.uint64)
        self.cols = np.ascontiguousarray(self.cols, dtype=np.uint32)
        self.row_cf = np.ascontiguousarray(self.row_cf, dtype=np.uint32)
        self.cf_ptr = np.ascontiguousarray(self.cf_ptr, dtype=np.uint64)
        self.cf = np.ascontiguousarray(self.cf)
        assert self.cf.dtype in (np.uint8, np.uint16, np.uint32)

    @property
    def cf_bytes(self) -> int:
        return self.cf.dtype.itemsize

    @property
    def ncf(self) -> int:
        return len(self.cf_ptr) - 1

    @property
    def nnz(self) -> int:
        return int(self.row_ptr[-1])

    def c_struct(self) -> CFlatMatrix:
        return CFlatMatrix(
            self.fc, self.cf_bytes, self.nru, self.nrl, self.ncl, self.ncr,
            self.ncf, self.flags,
            self.row_ptr.ctypes.data, self.cols.ctypes.data, self.row_cf.ctypes.data,
            self.cf_ptr.ctypes.data, self.cf.ctypes.data)

    @classmethod
    def from_c(cls, m: CFlatMatrix, **meta) -> "FlatMatrix":
        n = m.nru + m.nrl
        row_ptr = _np_from_ptr(m.row_ptr, n + 1, np.uint64)
        cf_ptr = _np_from_ptr(m.cf_ptr, m.ncf + 1, np.uint64)
        return cls(m.fc, m.nru, m.nrl, m.ncl, m.ncr, row_ptr,
                   _np_from_ptr(m.cols, int(row_ptr[-1]) if n else 0, np.uint32),
                   _np_from_ptr(m.row_cf, n, np.uint32), cf_ptr,
                   _np_from_ptr(m.cf, int(cf_ptr[-1]), _CF_DTYPE[m.cf_bytes]),
                   flags=m.flags, meta=dict(meta))

    # -- row access helpers (tests / analysis) ------------------------------
    def row(self, r: int):
        a, b = int(self.row_ptr[r]), int(self.row_ptr[r + 1])
        c0 = int(self.cf_ptr[self.row_cf[r]])
        return self.cols[a:b], self.cf[c0:c0 + (b - a)]

    def save(self, path: str, **extra) -> None:
        np.savez_compressed(
            path, fc=self.fc, nru=self.nru, nrl=self.nrl, ncl=self.ncl, ncr=self.ncr,
            flags=self.flags, row_ptr=self.row_ptr, cols=self.cols, row_cf=self.row_cf,
            cf_ptr=self.cf_ptr, cf=self.cf, **extra)

    @classmethod
    def load(cls, path: str) -> "FlatMatrix":
        z = np.load(path)
        return cls(int(z["fc"]), int(z["nru"]), int(z["nrl"]), int(z["ncl"]), int(z["ncr"]),
                   z["row_ptr"], z["cols"], z["row_cf"], z["cf_ptr"], z["cf"],
                   flags=int(z["flags"]))


@dataclass
class FlatResult:
    row_ptr: np.ndarray   # uint64 [nrows+1]
    cols: np.ndarray      # uint32
    cf: np.ndarray
    src_row: Optional[np.ndarray] = None
    rba: Optional[np.ndarray] = None      # uint32 [nrl, rba_words] when FLAG_WANT_RBA was set

    @property
    def nrows(self) -> int:
        return len(self.row_ptr) - 1

    @classmethod
    def from_c(cls, r: CFlatResult) -> "FlatResult":
        row_ptr = _np_from_ptr(r.row_ptr, r.nrows + 1, np.uint64)
        nnz = int(row_ptr[-1]) if len(row_ptr) else 0
        src = _np_from_ptr(r.src_row, r.nrows, np.uint32) if r.src_row else None
        return cls(row_ptr, _np_from_ptr(r.cols, nnz, np.uint32),
                   _np_from_ptr(r.cf, nnz, _CF_DTYPE[r.cf_bytes]), src)

    def attach_rba(self, r: CFlatResult, nrl: int) -> None:
        if r.rba and r.rba_words:
            self.rba = _np_from_ptr(r.rba, nrl * r.rba_words, np.uint32).reshape(nrl, r.rba_words)

    def same_rows(self, other: "FlatResult") -> bool:
        """Bit-exact equality of the echelon rows (columns and coefficients)."""
        return (np.array_equal(self.row_ptr, other.row_ptr)
                and np.array_equal(self.cols, other.cols)
                and np.array_equal(self.cf.astype(np.uint32), other.cf.astype(np.uint32)))

    def digest(self) -> str:
        import hashlib
        h = hashlib.sha256()
        h.update(self.row_ptr.tobytes()); h.update(self.cols.tobytes())
        h.update(self.cf.astype(np.uint32).tobytes())
        return h.hexdigest()
