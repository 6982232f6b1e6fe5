"""ctypes binding of oracle/_ref/libneogb_ref.so -- TEST INFRASTRUCTURE ONLY.

The shared object is the UNMODIFIED reference neogb plus oracle/ref_harness.c.
Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline /
``--impl reference`` legs may import this module; the product package
``msolve_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from msolve_b200.flat import CFlatMatrix, CFlatResult, FlatMatrix, FlatResult, _np_from_ptr  # noqa: E402

LIB_PATH = os.path.join(_HERE, "_ref", "libneogb_ref.so")
LIB_PATH_AVX512 = os.path.join(_HERE, "_ref", "libneogb_ref_avx512.so")


def host_cpu() -> dict:
    """Model name, logical / physical core counts and SIMD level of the host (for the bench lines)."""
    model, flags, cores = "unknown", set(), set()
    phys = core = None
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                k, _, v = line.partition(":")
                k, v = k.strip(), v.strip()
                if k == "model name":
                    model = v
                elif k == "flags" and not flags:
                    flags = set(v.split())
                elif k == "physical id":
                    phys = v
                elif k == "core id":
                    core = v
                elif k == "" and phys is not None:
                    cores.add((phys, core)); phys = core = None
    except OSError:
        pass
    return {"model": model, "logical_cpus": os.cpu_count() or 1, "physical_cores": len(cores) or None,
            "avx512": bool({"avx512f", "avx512bw"} <= flags), "avx2": "avx2" in flags}


def selected_lib_path() -> str:
    """The AVX-512 build of the reference when this host can run it (REFH_SIMD=avx2 forces the AVX2 one)."""
    if os.environ.get("REFH_SIMD", "") != "avx2" and os.path.exists(LIB_PATH_AVX512) and host_cpu()["avx512"]:
        return LIB_PATH_AVX512
    return LIB_PATH

MODE_REFERENCE, MODE_BACKEND, MODE_SHADOW, MODE_RECORD, MODE_RECORD_BACKEND = 0, 1, 2, 3, 4


class _CRecord(C.Structure):
    _fields_ = [
        ("inp", CFlatMatrix), ("bindex", C.c_void_p), ("mult", C.c_void_p),
        ("out", CFlatResult), ("out_bindex", C.c_void_p), ("out_mult", C.c_void_p),
        ("kind", C.c_int32), ("laopt", C.c_int32), ("ff_bits", C.c_int32),
        ("nf", C.c_int32), ("final_step", C.c_int32), ("trace_level", C.c_int32),
        ("la_seconds", C.c_double), ("units", C.c_double), ("pivot_apps", C.c_double),
    ]


class _CGb(C.Structure):
    _fields_ = [
        ("nelts", C.c_int32), ("nterms", C.c_int64),
        ("lens", C.c_void_p), ("exps", C.c_void_p), ("cfs", C.c_void_p),
        ("f4_seconds", C.c_double), ("la_seconds", C.c_double), ("la_calls", C.c_int32),
    ]


@dataclass
class Record:
    matrix: FlatMatrix
    result: FlatResult
    kind: int
    laopt: int
    ff_bits: int
    nf: int
    final_step: int
    trace_level: int
    la_seconds: float
    units: float
    pivot_apps: float


@dataclass
class GroebnerBasis:
    lens: np.ndarray
    exps: np.ndarray
    cfs: np.ndarray
    f4_seconds: float
    la_seconds: float
    la_calls: int

    def digest(self) -> str:
        import hashlib
        h = hashlib.sha256()
        for a in (self.lens, self.exps, self.cfs):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()


_lib: Optional[C.CDLL] = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(selected_lib_path(), mode=C.RTLD_GLOBAL)
        L.refh_set_mode.argtypes = [C.c_int, C.c_int, C.c_int]
        L.refh_set_backend.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.refh_num_records.restype = C.c_int
        L.refh_shadow_mismatches.restype = C.c_int64
        L.refh_shadow_calls.restype = C.c_int64
        L.refh_la_seconds.restype = C.c_double
        L.refh_la_calls.restype = C.c_int
        L.refh_get_record.restype = C.POINTER(_CRecord)
        L.refh_get_record.argtypes = [C.c_int]
        L.refh_drop_record.argtypes = [C.c_int]
        L.refh_record_size.restype = C.c_size_t
        L.refh_f4.restype = C.POINTER(_CGb)
        L.refh_f4.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                              C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.refh_reference_la.restype = C.c_int
        L.refh_reference_la.argtypes = [C.POINTER(CFlatMatrix), C.c_int32, C.c_int32,
                                        C.POINTER(CFlatResult), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.refh_free_result.argtypes = [C.POINTER(CFlatResult)]
        L.refh_struct_layout.restype = C.c_int
        L.refh_struct_layout.argtypes = [C.POINTER(C.c_uint64), C.c_int]
        assert L.refh_record_size() == C.sizeof(_CRecord), "refh_record layout drifted"
        _lib = L
    return _lib


def set_mode(mode: int, max_records: int = 0, verbose: int = 0) -> None:
    lib().refh_set_mode(mode, max_records, verbose)


def set_backend(la=None, exact_la=None, interred=None, app=None, trace=None) -> None:
    """Addresses (ints / ctypes function pointers) of the backend entry points: the five
    matrix-level pointers of data.h:529-595."""
    def addr(f):
        if f is None:
            return None
        return C.cast(f, C.c_void_p).value if not isinstance(f, int) else f
    L = lib()
    L.refh_set_backend(addr(la), addr(exact_la), addr(interred))
    L.refh_set_backend_tracer.argtypes = [C.c_void_p, C.c_void_p]
    L.refh_set_backend_tracer(addr(app), addr(trace))


def hooked_pointers() -> int:
    """Bit i set: pointer i (linear_algebra, exact_linear_algebra, interreduce_matrix_rows,
    application_linear_algebra, trace_linear_algebra) goes through the harness."""
    lib().refh_hooked_pointers.restype = C.c_int
    return int(lib().refh_hooked_pointers())


def reset() -> None:
    lib().refh_reset()


def run_f4(system, p: int, threads: int = 1, laopt: int = 2, info_level: int = 0) -> GroebnerBasis:
    lens, exps, cfs = system.encode(p)
    g = lib().refh_f4(lens.ctypes.data, exps.ctypes.data, cfs.ctypes.data, p,
                      system.nvars, len(lens), threads, laopt, info_level).contents
    n, nt = g.nelts, g.nterms
    return GroebnerBasis(_np_from_ptr(g.lens, n, np.int32),
                         _np_from_ptr(g.exps, nt * system.nvars, np.int32),
                         _np_from_ptr(g.cfs, nt, np.int32),
                         g.f4_seconds, g.la_seconds, g.la_calls)


def records(drop: bool = True) -> List[Record]:
    """Copies the recorded matrices out of the harness (and frees them there)."""
    L = lib()
    out = []
    for i in range(L.refh_num_records()):
        r = L.refh_get_record(i).contents
        out.append(Record(FlatMatrix.from_c(r.inp), FlatResult.from_c(r.out), r.kind, r.laopt,
                          r.ff_bits, r.nf, r.final_step, r.trace_level, r.la_seconds,
                          r.units, r.pivot_apps))
        if drop:
            L.refh_drop_record(i)
    return out


def record_f4(system, p: int, threads: int = 1, laopt: int = 2, max_records: int = 0):
    """Runs the reference F4 and returns (GroebnerBasis, [Record per LA call])."""
    reset()
    set_mode(MODE_RECORD, max_records)
    gb = run_f4(system, p, threads, laopt)
    recs = records()
    reset()
    set_mode(MODE_REFERENCE)
    return gb, recs


def reference_la(m: FlatMatrix, laopt: int = 2, threads: int = 1):
    """Reference LA on one flat matrix -> (FlatResult, seconds, units, pivot_apps)."""
    cm = m.c_struct()
    res = CFlatResult()
    sec, units, apps = C.c_double(), C.c_double(), C.c_double()
    rc = lib().refh_reference_la(C.byref(cm), laopt, threads, C.byref(res),
                                 C.byref(sec), C.byref(units), C.byref(apps))
    assert rc == 0
    out = FlatResult.from_c(res)
    lib().refh_free_result(C.byref(res))
    return out, sec.value, units.value, apps.value


class QQResult(C.Structure):
    _fields_ = [("digest", C.c_uint64), ("nelts", C.c_int32), ("nterms", C.c_int64), ("err", C.c_int32),
                ("seconds", C.c_double), ("la_seconds", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def qq_setup(system, threads: int = 1, laopt: int = 2, info_level: int = 0) -> None:
    """Rational input (integer coefficients of the formula) for the tracer flow."""
    from msolve_b200.systems import encode_integers
    lens, exps, cfs = encode_integers(system)
    L = lib()
    L.refh_qq_setup.restype = C.c_int
    L.refh_qq_setup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    rc = L.refh_qq_setup(lens.ctypes.data, exps.ctypes.data, cfs.ctypes.data, system.nvars, len(lens),
                         threads, laopt, info_level)
    if rc != 0:
        raise RuntimeError(f"refh_qq_setup failed: {rc}")


def qq_learn(p0: int) -> dict:
    L = lib()
    L.refh_qq_learn.restype = C.c_int
    L.refh_qq_learn.argtypes = [C.c_uint32, C.POINTER(QQResult)]
    r = QQResult()
    rc = L.refh_qq_learn(p0, C.byref(r))
    d = r.as_dict(); d["rc"] = rc
    return d


def qq_apply(primes, threads: int = 1, ngpus: int = 0, residue_words: int = 0):
    """-> (list of per-prime dicts, wall seconds[, residues uint32 [nprimes, residue_words]])
    residue_words > 0 also returns the coefficients of every modular basis (the residues msolve
    lifts by CRT); pass the learn phase's nterms."""
    L = lib()
    L.refh_qq_apply_residues.restype = C.c_int
    L.refh_qq_apply_residues.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(QQResult),
                                         C.POINTER(C.c_double), C.c_void_p, C.c_int64]
    arr = np.asarray(primes, dtype=np.uint32)
    res = (QQResult * max(1, len(arr)))()
    wall = C.c_double()
    resid = np.zeros((len(arr), residue_words), dtype=np.uint32) if residue_words else None
    rc = L.refh_qq_apply_residues(arr.ctypes.data, len(arr), threads, ngpus, res, C.byref(wall),
                                  resid.ctypes.data if resid is not None and resid.size else None, residue_words)
    if rc != 0:
        raise RuntimeError(f"refh_qq_apply failed: {rc} (learn first)")
    out = [res[i].as_dict() for i in range(len(arr))]
    return (out, wall.value, resid) if residue_words else (out, wall.value)


def residue_digest(row: np.ndarray, lens_digest: int = 0) -> int:
    """FNV-1a over the residues of one prime (to compare gathered residues across ranks)."""
    h = 1469598103934665603
    for b in np.ascontiguousarray(row, dtype=np.uint32).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def qq_initial_basis(p: int, cap: int = 1 << 20):
    """Initial basis of an application run modulo p (normalised) -> (cf uint32, off uint64 [n+1]).
    Coefficient arrays are ordered by position in the basis; the harness asserts position == COEFFS index."""
    L = lib()
    L.refh_qq_initial_basis.restype = C.c_int
    L.refh_qq_initial_basis.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_int64]
    cf = np.zeros(cap, dtype=np.uint32)
    off = np.zeros(4096, dtype=np.uint64)
    n = L.refh_qq_initial_basis(p, cf.ctypes.data, off.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("refh_qq_initial_basis failed (setup first / buffer too small)")
    return cf[:int(off[n])].copy(), off[:n + 1].copy()


def set_device_callback(fn) -> None:
    lib().refh_set_device_callback.argtypes = [C.c_void_p]
    lib().refh_set_device_callback(C.cast(fn, C.c_void_p).value if fn is not None else None)


def primes_above(start: int, n: int):
    """n successive primes > start (segmented sieve; start + window < 2^32)."""
    out = []
    lo = start + 1
    while len(out) < n:
        span = max(4096, 32 * (n - len(out)))
        hi = lo + span
        lim = int(hi ** 0.5) + 1
        small = np.ones(lim + 1, dtype=bool)
        small[:2] = False
        for q in range(2, int(lim ** 0.5) + 1):
            if small[q]:
                small[q * q::q] = False
        mark = np.ones(span, dtype=bool)
        for q in np.nonzero(small)[0]:
            q = int(q)
            first = max(q * q, (lo + q - 1) // q * q)
            mark[first - lo::q] = False
        out += [int(lo + i) for i in np.nonzero(mark)[0]]
        lo = hi
    return out[:n]


def call_entry(m: FlatMatrix, which: int, fn=None, threads: int = 1):
    """Drive a matrix-level entry point (reference dispatcher when fn is None, else the
    given function address) on a flat matrix.  -> (FlatResult, int return value, COEFFS indices)"""
    L = lib()
    L.refh_call_entry.restype = C.c_int
    L.refh_call_entry.argtypes = [C.POINTER(CFlatMatrix), C.c_int32, C.c_void_p, C.c_int32,
                                  C.POINTER(CFlatResult), C.POINTER(C.c_int32), C.c_void_p]
    cm = m.c_struct()
    res = CFlatResult()
    ret = C.c_int32(0)
    cidx = np.zeros(m.nru + m.nrl + m.ncl + m.ncr + 1, dtype=np.uint32)
    addr = None if fn is None else (fn if isinstance(fn, int) else C.cast(fn, C.c_void_p).value)
    rc = L.refh_call_entry(C.byref(cm), which, addr, threads, C.byref(res), C.byref(ret), cidx.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"refh_call_entry failed: {rc} (-2: BINDEX/MULT not carried over together)")
    out = FlatResult.from_c(res)
    L.refh_free_result(C.byref(res))
    return out, int(ret.value), cidx[:out.nrows].copy()


def call_trace_entry(m: FlatMatrix, fn=None, threads: int = 1):
    """Drive the trace_linear_algebra signature (data.h:571-576) on a flat matrix.
    -> (FlatResult, kept lower rows (flat row ids), reducers used (pivot ids), bit arrays [kept][words])"""
    L = lib()
    L.refh_call_trace_entry.restype = C.c_int
    L.refh_call_trace_entry.argtypes = [C.POINTER(CFlatMatrix), C.c_void_p, C.c_int32, C.POINTER(CFlatResult),
                                        C.c_void_p, C.c_uint64]
    cm = m.c_struct()
    res = CFlatResult()
    cap = 3 + m.nrl + m.nru + m.nrl * ((m.nru + 31) // 32 + 1)
    info = np.zeros(cap, dtype=np.uint32)
    addr = None if fn is None else (fn if isinstance(fn, int) else C.cast(fn, C.c_void_p).value)
    rc = L.refh_call_trace_entry(C.byref(cm), addr, threads, C.byref(res), info.ctypes.data, cap)
    if rc != 0:
        raise RuntimeError(f"refh_call_trace_entry failed: {rc} (-2: BINDEX/MULT not carried over together)")
    out = FlatResult.from_c(res)
    L.refh_free_result(C.byref(res))
    ntr, nrr, words = int(info[0]), int(info[1]), int(info[2])
    assert words != 0xffffffff
    kept = info[3:3 + ntr].astype(np.int64) - m.nru
    reds = info[3 + ntr:3 + ntr + nrr].astype(np.int64)
    bits = info[3 + ntr + nrr:3 + ntr + nrr + ntr * words].reshape(ntr, words).copy()
    return out, kept, reds, bits


def last_side() -> dict:
    """Side channels (md_t / mat_t fields) left by the last call_entry()."""
    buf = (C.c_double * 11)()
    lib().refh_last_side(buf)
    names = ["st_np", "mat_np", "mat_nr", "mat_sz", "num_zerored", "la_rtime", "la_ctime",
             "application_nr_mult", "application_nr_add", "application_nr_red", "call_seconds"]
    return dict(zip(names, [float(x) for x in buf]))


def struct_layout() -> List[int]:
    buf = (C.c_uint64 * 128)()
    n = lib().refh_struct_layout(buf, 128)
    return [int(buf[i]) for i in range(n)]
