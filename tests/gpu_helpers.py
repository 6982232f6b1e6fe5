"""Glue used by the GPU tests, smoke() and bench.py: plugs the product's
neogb-level entry points into the reference harness exactly like the stub in
INTEGRATION.md would plug them into neogb's function pointers."""
import ctypes as C

from msolve_b200 import backend
from oracle import refharness as rh


def install_backend(mode=rh.MODE_BACKEND, verbose=0):
    """Configure the layout descriptor and point the harness' hooks at the GPU
    linear algebra.  Returns the product library handle."""
    L = backend.load_library()
    lay = (C.c_uint32 * 64)()
    rh.lib().refh_fill_layout(lay)
    assert L.f4la_b200_neogb_configure(lay) == 0
    rh.lib().refh_construct_trace_ptr.restype = C.c_void_p
    L.f4la_b200_neogb_set_construct_trace.argtypes = [C.c_void_p]
    L.f4la_b200_neogb_set_construct_trace(rh.lib().refh_construct_trace_ptr())
    rh.set_device_callback(L.f4la_b200_neogb_set_device)
    L.f4la_b200_neogb_set_free_basis_elements.argtypes = [C.c_void_p]
    rh.lib().refh_free_basis_elements_ptr.restype = C.c_void_p
    L.f4la_b200_neogb_set_free_basis_elements(rh.lib().refh_free_basis_elements_ptr())
    # all five matrix-level pointers of data.h:529-595
    fns = [C.cast(getattr(L, n), C.c_void_p).value for n in (
        "f4la_b200_linear_algebra", "f4la_b200_exact_linear_algebra", "f4la_b200_interreduce_matrix_rows",
        "f4la_b200_application_linear_algebra", "f4la_b200_trace_linear_algebra")]
    rh.set_backend(*fns)
    rh.reset()
    rh.set_mode(mode, 0, verbose)
    return L


def uninstall_backend():
    rh.set_backend(None, None, None, None, None)
    rh.reset()
    rh.set_mode(rh.MODE_REFERENCE)


def neogb_totals(L, reset=True):
    t = backend.NeogbTotals()
    L.f4la_b200_neogb_get_totals(C.byref(t), 1 if reset else 0)
    return t.as_dict()
