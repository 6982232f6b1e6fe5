"""Host-side binding of the product library ``csrc/libf4la_b200.so``.

The library is the only compute path: if it is missing, or no sm_100 device is
usable, construction fails loudly -- there is no CPU fallback (the plain-C
restatement under ``oracle/`` is test infrastructure and is never imported
from here).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

from .flat import CFlatMatrix, CFlatResult, CStats, FlatMatrix, FlatResult

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libf4la_b200.so")

EXPORTED_SYMBOLS = [
    # include/f4la_b200.h
    "f4la_b200_version", "f4la_b200_device_count", "f4la_b200_create", "f4la_b200_destroy",
    "f4la_b200_last_error", "f4la_b200_set_seed", "f4la_b200_set_host_mode", "f4la_b200_set_grid_cap", "f4la_b200_device_alloc", "f4la_b200_device_free", "f4la_b200_device_copy", "f4la_b200_device_zero", "f4la_b200_resident_layout", "f4la_b200_replay_call", "f4la_b200_pinned_alloc", "f4la_b200_pinned_free",
    "f4la_b200_reduce", "f4la_b200_stage_cols", "f4la_b200_stage_acquire", "f4la_b200_stage_commit", "f4la_b200_upload", "f4la_b200_reduce_resident", "f4la_b200_release", "f4la_b200_resident_update",
    "f4la_b200_free_result", "f4la_b200_measure_mac_peak", "f4la_b200_modp_gemm",
    # include/f4la_b200_neogb.h
    "f4la_b200_neogb_configure", "f4la_b200_neogb_set_device", "f4la_b200_neogb_set_construct_trace", "f4la_b200_linear_algebra",
    "f4la_b200_exact_linear_algebra", "f4la_b200_application_linear_algebra", "f4la_b200_trace_linear_algebra",
    "f4la_b200_interreduce_matrix_rows", "f4la_b200_neogb_set_free_basis_elements", "f4la_b200_neogb_get_totals",
    "f4la_b200_neogb_tape_begin", "f4la_b200_neogb_tape_end", "f4la_b200_tape_free", "f4la_b200_tape_calls",
    "f4la_b200_tape_initial_elements", "f4la_b200_tape_initial_length", "f4la_b200_tape_residue_words",
    "f4la_b200_tape_replay", "f4la_b200_tape_replay_enqueue", "f4la_b200_tape_replay_collect",
]


class F4LAError(RuntimeError):
    pass


class NeogbTotals(C.Structure):
    _fields_ = [
        ("calls", C.c_uint64),
        ("ms_total", C.c_double), ("ms_flatten", C.c_double), ("ms_h2d", C.c_double),
        ("ms_prep", C.c_double), ("ms_reduce", C.c_double), ("ms_echelon", C.c_double),
        ("ms_d2h", C.c_double), ("ms_materialize", C.c_double),
        ("units", C.c_uint64), ("bytes_h2d", C.c_uint64), ("bytes_d2h", C.c_uint64),
        ("kernel_launches", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """dlopen the product library and declare the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("F4LA_B200_LIB", LIB_PATH)     # A/B builds of the same library (tools/)
    if not os.path.exists(path):
        raise F4LAError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C msolve_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    L.f4la_b200_version.restype = C.c_char_p
    L.f4la_b200_device_count.restype = C.c_int
    L.f4la_b200_create.restype = C.c_void_p
    L.f4la_b200_create.argtypes = [C.c_int]
    L.f4la_b200_destroy.argtypes = [C.c_void_p]
    L.f4la_b200_last_error.restype = C.c_char_p
    L.f4la_b200_last_error.argtypes = [C.c_void_p]
    L.f4la_b200_set_seed.argtypes = [C.c_void_p, C.c_uint64]
    L.f4la_b200_set_host_mode.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.f4la_b200_set_grid_cap.argtypes = [C.c_void_p, C.c_uint32]
    L.f4la_b200_pinned_alloc.restype = C.c_void_p
    L.f4la_b200_pinned_alloc.argtypes = [C.c_size_t]
    L.f4la_b200_pinned_free.argtypes = [C.c_void_p]
    L.f4la_b200_reduce.restype = C.c_int
    L.f4la_b200_reduce.argtypes = [C.c_void_p, C.POINTER(CFlatMatrix), C.POINTER(CFlatResult), C.POINTER(CStats)]
    L.f4la_b200_stage_cols.restype = C.c_int
    L.f4la_b200_stage_cols.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
    L.f4la_b200_upload.restype = C.c_void_p
    L.f4la_b200_upload.argtypes = [C.c_void_p, C.POINTER(CFlatMatrix)]
    L.f4la_b200_reduce_resident.restype = C.c_int
    L.f4la_b200_reduce_resident.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CFlatResult), C.POINTER(CStats)]
    L.f4la_b200_release.argtypes = [C.c_void_p, C.c_void_p]
    L.f4la_b200_free_result.argtypes = [C.POINTER(CFlatResult)]
    L.f4la_b200_modp_gemm.restype = C.c_int
    L.f4la_b200_modp_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.c_void_p, C.POINTER(C.c_double)]
    L.f4la_b200_measure_mac_peak.restype = C.c_double
    L.f4la_b200_measure_mac_peak.argtypes = [C.c_void_p, C.c_double]
    L.f4la_b200_neogb_configure.restype = C.c_int
    L.f4la_b200_neogb_configure.argtypes = [C.c_void_p]
    L.f4la_b200_neogb_set_device.argtypes = [C.c_int]
    L.f4la_b200_neogb_get_totals.argtypes = [C.POINTER(NeogbTotals), C.c_int]
    _lib = L
    return L


def pinned_copy(L: C.CDLL, a: np.ndarray) -> Tuple[np.ndarray, int]:
    """Copy ``a`` into page-locked memory; returns (view, base address to free)."""
    a = np.ascontiguousarray(a)
    nbytes = max(a.nbytes, 1)
    ptr = L.f4la_b200_pinned_alloc(nbytes)
    if not ptr:
        raise F4LAError("pinned host allocation failed")
    buf = (C.c_ubyte * nbytes).from_address(ptr)
    view = np.frombuffer(buf, dtype=a.dtype, count=a.size)
    view[...] = a.reshape(-1)
    return view, ptr


class Backend:
    """One device context (its own CUDA stream); not shared between threads."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.ctx = self.lib.f4la_b200_create(device)
        if not self.ctx:
            raise F4LAError(f"f4la_b200_create({device}) failed: no usable sm_100 GPU (no CPU fallback)")
        self.device = device

    def measure_mac_peak(self, milliseconds: float = 200.0) -> float:
        """Multiply-adds per second of the hot kernel's 96-bit accumulate with register operands."""
        return float(self.lib.f4la_b200_measure_mac_peak(self.ctx, milliseconds))

    def modp_gemm(self, A: np.ndarray, B: np.ndarray, p: int):
        """A (M x K) @ B (K x N) mod p on the tensor cores -> (C uint32 [M, N], kernel milliseconds)."""
        M, K = A.shape
        K2, N = B.shape
        assert K == K2
        a = np.asfortranarray(A, dtype=np.uint32)            # column major, leading dimension M
        b = np.ascontiguousarray(B, dtype=np.uint32)
        c = np.zeros((M, N), dtype=np.uint32, order="F")
        ms = C.c_double()
        self._check(self.lib.f4la_b200_modp_gemm(self.ctx, a.ctypes.data, b.ctypes.data, M, N, K, p, c.ctypes.data,
                                                 C.byref(ms)))
        return c, ms.value

    def set_host_mode(self, host_threads: int, blocking_sync: bool) -> None:
        self.lib.f4la_b200_set_host_mode(self.ctx, host_threads, 1 if blocking_sync else 0)

    def set_grid_cap(self, ctas: int) -> None:
        self.lib.f4la_b200_set_grid_cap(self.ctx, ctas)

    def set_seed(self, seed: int) -> None:
        self.lib.f4la_b200_set_seed(self.ctx, seed)

    def close(self) -> None:
        if self.ctx:
            self.lib.f4la_b200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            msg = self.lib.f4la_b200_last_error(self.ctx)
            raise F4LAError(f"f4la_b200 error {rc}: {msg.decode() if msg else '?'}")

    def reduce(self, m: FlatMatrix):
        """Echelon form from host buffers -> (FlatResult, stats dict)."""
        cm, res, st = m.c_struct(), CFlatResult(), CStats()
        self._check(self.lib.f4la_b200_reduce(self.ctx, C.byref(cm), C.byref(res), C.byref(st)))
        out = FlatResult.from_c(res)
        out.attach_rba(res, m.nrl)
        self.lib.f4la_b200_free_result(C.byref(res))
        return out, st.as_dict()

    def reduce_staged(self, pm: "PinnedMatrix", pieces: int = 4):
        """Same as :meth:`reduce`, with the column indices sent piecewise beforehand
        (``f4la_b200_stage_cols`` + ``F4LA_FLAG_COLS_STAGED``), the way the neogb adapter overlaps its
        gathering of the rows with the host->device copy.  ``pm`` must be pinned."""
        from .flat import FLAG_COLS_STAGED
        m = pm.matrix
        nnz = m.nnz
        step = max(1, (nnz + pieces - 1) // pieces)
        for first in range(0, nnz, step):
            self._check(self.lib.f4la_b200_stage_cols(self.ctx, m.cols.ctypes.data, first, min(step, nnz - first), nnz))
        cm, res, st = m.c_struct(), CFlatResult(), CStats()
        cm.flags |= FLAG_COLS_STAGED
        cm.cols = None                      # not read in this mode
        self._check(self.lib.f4la_b200_reduce(self.ctx, C.byref(cm), C.byref(res), C.byref(st)))
        out = FlatResult.from_c(res)
        out.attach_rba(res, m.nrl)
        self.lib.f4la_b200_free_result(C.byref(res))
        return out, st.as_dict()

    def pin(self, m: FlatMatrix) -> "PinnedMatrix":
        return PinnedMatrix(self.lib, m)

    def upload(self, m: FlatMatrix) -> "ResidentMatrix":
        cm = m.c_struct()
        h = self.lib.f4la_b200_upload(self.ctx, C.byref(cm))
        if not h:
            raise F4LAError("f4la_b200_upload failed")
        return ResidentMatrix(self, h)

    def reduce_resident(self, rm: "ResidentMatrix", want_result: bool = True):
        res, st = CFlatResult(), CStats()
        self._check(self.lib.f4la_b200_reduce_resident(self.ctx, rm.handle, C.byref(res), C.byref(st)))
        out = FlatResult.from_c(res) if want_result else None
        self.lib.f4la_b200_free_result(C.byref(res))
        return out, st.as_dict()


class ResidentMatrix:
    def __init__(self, be: Backend, handle):
        self.be, self.handle = be, handle

    def release(self) -> None:
        if self.handle:
            self.be.lib.f4la_b200_release(self.be.ctx, self.handle)
            self.handle = None


class PinnedMatrix:
    """A FlatMatrix whose arrays live in page-locked host memory."""

    def __init__(self, L: C.CDLL, m: FlatMatrix):
        self.lib = L
        self._ptrs = []
        arrs = {}
        for name in ("row_ptr", "cols", "row_cf", "cf_ptr", "cf"):
            view, ptr = pinned_copy(L, getattr(m, name))
            arrs[name] = view
            self._ptrs.append(ptr)
        self.matrix = FlatMatrix.__new__(FlatMatrix)
        self.matrix.__dict__.update(m.__dict__)
        self.matrix.__dict__.update(arrs)

    def free(self) -> None:
        for p in self._ptrs:
            self.lib.f4la_b200_pinned_free(p)
        self._ptrs = []


class Tape:
    """Structure-cached replay of the tracer application phase (include/f4la_b200_neogb.h): record one application
    run through the drop-in pointers, then replay it for other primes with no symbolic work on the host."""

    def __init__(self, lib, handle):
        self.lib, self.handle = lib, handle
        lib.f4la_b200_tape_calls.restype = C.c_uint32
        lib.f4la_b200_tape_calls.argtypes = [C.c_void_p]
        lib.f4la_b200_tape_initial_elements.restype = C.c_uint32
        lib.f4la_b200_tape_initial_elements.argtypes = [C.c_void_p]
        lib.f4la_b200_tape_initial_length.restype = C.c_uint32
        lib.f4la_b200_tape_initial_length.argtypes = [C.c_void_p, C.c_uint32]
        lib.f4la_b200_tape_residue_words.restype = C.c_uint64
        lib.f4la_b200_tape_residue_words.argtypes = [C.c_void_p]
        lib.f4la_b200_tape_replay.restype = C.c_int
        lib.f4la_b200_tape_replay.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.POINTER(C.c_double)]
        lib.f4la_b200_tape_replay_enqueue.restype = C.c_int
        lib.f4la_b200_tape_replay_enqueue.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.f4la_b200_tape_replay_collect.restype = C.c_int
        lib.f4la_b200_tape_replay_collect.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_int]
        lib.f4la_b200_tape_free.argtypes = [C.c_void_p]
        self.calls = int(lib.f4la_b200_tape_calls(handle))
        self.initial_elements = int(lib.f4la_b200_tape_initial_elements(handle))
        self.residue_words = int(lib.f4la_b200_tape_residue_words(handle))

    @staticmethod
    def begin(lib) -> None:
        lib.f4la_b200_neogb_tape_begin()

    @staticmethod
    def end(lib) -> Optional["Tape"]:
        lib.f4la_b200_neogb_tape_end.restype = C.c_void_p
        h = lib.f4la_b200_neogb_tape_end()
        return Tape(lib, h) if h else None

    def replay(self, be: Backend, p: int, init_cf: np.ndarray, init_off: np.ndarray, out: Optional[np.ndarray] = None):
        """-> (return code, residues uint32 [residue_words], milliseconds); `out` (C-contiguous uint32 row of at least
        residue_words) receives the residues in place when given"""
        init_cf = np.ascontiguousarray(init_cf, dtype=np.uint32)
        init_off = np.ascontiguousarray(init_off, dtype=np.uint64)
        if out is None:
            out = np.zeros(self.residue_words, dtype=np.uint32)
        assert out.dtype == np.uint32 and out.flags.c_contiguous and out.size >= self.residue_words
        ms = C.c_double(0.0)
        rc = self.lib.f4la_b200_tape_replay(self.handle, be.ctx, p, init_cf.ctypes.data, init_off.ctypes.data,
                                            out.ctypes.data, C.byref(ms))
        return int(rc), out, ms.value

    MAX_INFLIGHT = 16

    def enqueue(self, be: Backend, p: int, init_cf: np.ndarray, init_off: np.ndarray, out) -> None:
        """Device-resident replay: queues the prime on the backend's stream; `out` (C-contiguous uint32 row, page-locked
        for a truly asynchronous copy, or the integer address of a row in device memory) receives the residues and
        must stay alive until collect()."""
        assert init_cf.dtype == np.uint32 and init_cf.flags.c_contiguous
        assert init_off.dtype == np.uint64 and init_off.flags.c_contiguous
        if isinstance(out, int):                  # address of residue_words uint32 in device (or page-locked) memory
            dst = out
        else:
            assert out.dtype == np.uint32 and out.flags.c_contiguous and out.size >= self.residue_words
            dst = out.ctypes.data
        rc = self.lib.f4la_b200_tape_replay_enqueue(self.handle, be.ctx, p, init_cf.ctypes.data, init_off.ctypes.data, dst)
        if rc != 0:
            raise F4LAError(f"f4la_b200_tape_replay_enqueue: error {rc}")

    def collect(self, be: Backend):
        """Waits for the backend's stream; return codes of the primes enqueued since the last collect, in order."""
        rcs = (C.c_int * 64)()
        n = self.lib.f4la_b200_tape_replay_collect(self.handle, be.ctx, rcs, 64)
        if n < 0:
            raise F4LAError("f4la_b200_tape_replay_collect failed")
        return [int(rcs[i]) for i in range(n)]

    def free(self) -> None:
        if self.handle:
            self.lib.f4la_b200_tape_free(self.handle)
            self.handle = None
