"""ctypes binding of oracle/_ref/libf4la_oracle.so (our plain-C restatement of
the reference algorithm) -- TEST INFRASTRUCTURE ONLY, see f4la_oracle.h."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from msolve_b200.flat import CFlatMatrix, CFlatResult, FlatMatrix, FlatResult  # noqa: E402

LIB_PATH = os.path.join(_HERE, "_ref", "libf4la_oracle.so")


class _Counters(C.Structure):
    _fields_ = [("units", C.c_uint64), ("pivot_applications", C.c_uint64)]


_lib = None


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.f4la_oracle_reduce.restype = C.c_int
        L.f4la_oracle_reduce.argtypes = [C.POINTER(CFlatMatrix), C.POINTER(CFlatResult),
                                         C.POINTER(_Counters)]
        L.f4la_oracle_mod_inverse.restype = C.c_uint32
        L.f4la_oracle_mod_inverse.argtypes = [C.c_uint32, C.c_uint32]
        L.f4la_oracle_free_result.argtypes = [C.POINTER(CFlatResult)]
        _lib = L
    return _lib


def reduce(m: FlatMatrix):
    """-> (FlatResult, units, pivot_applications)"""
    cm = m.c_struct()
    res = CFlatResult()
    cnt = _Counters()
    rc = lib().f4la_oracle_reduce(C.byref(cm), C.byref(res), C.byref(cnt))
    if rc != 0:
        raise RuntimeError(f"f4la_oracle_reduce failed with code {rc}")
    out = FlatResult.from_c(res)
    lib().f4la_oracle_free_result(C.byref(res))
    return out, int(cnt.units), int(cnt.pivot_applications)


def mod_inverse(v: int, p: int) -> int:
    return int(lib().f4la_oracle_mod_inverse(v, p))
