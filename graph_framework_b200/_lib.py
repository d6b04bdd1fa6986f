"""ctypes loader for libgfb200.so (the C-ABI device layer + host front end).

There is no Python or CPU fallback: if the shared library is missing the import
fails, and if no B200 is visible every compute entry point fails with the
library's error text.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgfb200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `make` (or __graft_entry__.build()). "
            "The B200 back end has no fallback path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    P, I, U, SZ, D, S = (ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_size_t,
                         ctypes.c_double, ctypes.c_char_p)
    U64 = ctypes.c_uint64
    sig = {
        # gfb200.h
        "gfb_device_count": (I, []),
        "gfb_last_error": (S, []),
        "gfb_version": (S, []),
        "gfb_ctx_create": (P, [I]),
        "gfb_ctx_destroy": (None, [P]),
        "gfb_ctx_device_info": (I, [P, ctypes.c_char_p, SZ, ctypes.POINTER(I), ctypes.POINTER(I), ctypes.POINTER(I)]),
        "gfb_compile": (I, [P, S, ctypes.POINTER(S), I, S]),
        "gfb_compiled_min_blocks": (I, [P]),
        "gfb_compile_to_cubin": (I, [S, S, c_void_pp, ctypes.POINTER(SZ), ctypes.POINTER(ctypes.c_void_p)]),
        "gfb_free": (None, [P]),
        "gfb_source": (S, [P]),
        "gfb_compile_log": (S, [P]),
        "gfb_compile_options": (S, [P]),
        "gfb_buffer": (I, [P, U64, SZ, P, c_void_pp]),
        "gfb_buffer_import": (I, [P, U64, P, SZ]),
        "gfb_buffer_lookup": (I, [P, U64, c_void_pp, ctypes.POINTER(SZ)]),
        "gfb_kernel_create": (I, [P, S, ctypes.POINTER(U64), I, SZ, U, SZ, I, I, c_void_pp]),
        "gfb_kernel_run": (I, [P]),
        "gfb_kernel_launch": (I, [P, U]),
        "gfb_kernel_run_from_host": (I, [P, U, I, c_void_pp, c_void_pp, I]),
        "gfb_kernel_set_scalar": (I, [P, I, D]),
        "gfb_kernel_attributes": (I, [P, ctypes.POINTER(I), ctypes.POINTER(I), ctypes.POINTER(I), ctypes.POINTER(I)]),
        "gfb_launch_count": (U64, [P]),
        "gfb_set_max_fused_steps": (I, [P, U]),
        "gfb_flush": (I, [P]),
        "gfb_max": (I, [P, U64, SZ, c_double_p]),
        "gfb_wait": (I, [P]),
        "gfb_copy_h2d": (I, [P, U64, P, SZ]),
        "gfb_copy_d2h": (I, [P, U64, P, SZ]),
        "gfb_host_ptr": (I, [P, U64, c_void_pp]),
        "gfb_check_value": (I, [P, U64, SZ, c_double_p]),
        "gfb_snapshot_async": (I, [P, ctypes.POINTER(U64), I, SZ, P]),
        "gfb_timer_start": (I, [P]),
        "gfb_timer_stop": (I, [P, ctypes.POINTER(ctypes.c_float)]),
        "gfb_stream": (P, [P]),
        "gfb_deposit": (I, [P, P, P, P, P, SZ, P, c_double_p, c_double_p, ctypes.POINTER(I)]),
        "gfb_allreduce_sum_f64": (I, [c_void_pp, I, ctypes.POINTER(U64), SZ]),
        "gfb_measure_fp64_peak": (I, [P, c_double_p, ctypes.POINTER(ctypes.c_float)]),
        "gfb_measure_fp64_peak_registers": (I, [P, c_double_p, ctypes.POINTER(ctypes.c_float)]),
        "gfb_flush_l2": (I, [P]),
        # gfb_rays.h
        "gfb_rays_create": (P, [S, S, S, S, SZ, D, I, S]),
        "gfb_rays_destroy": (None, [P]),
        "gfb_rays_set_state": (I, [P, ctypes.POINTER(c_double_p)]),
        "gfb_rays_init": (I, [P, S, D, SZ, I]),
        "gfb_rays_compile": (I, [P]),
        "gfb_rays_step": (I, [P, SZ]),
        "gfb_rays_wait": (I, [P]),
        "gfb_rays_get_state": (I, [P, ctypes.POINTER(c_double_p), c_double_p]),
        "gfb_rays_put_state": (I, [P, ctypes.POINTER(c_double_p)]),
        "gfb_rays_step_host": (I, [P, SZ, ctypes.POINTER(c_double_p), ctypes.POINTER(c_double_p), c_double_p, I]),
        "gfb_host_alloc": (I, [SZ, c_void_pp]),
        "gfb_host_free": (I, [P]),
        "gfb_rays_trace": (I, [P, SZ, SZ, c_double_p]),
        "gfb_rays_trace_absorb": (I, [P, SZ, SZ, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                      ctypes.POINTER(ctypes.c_int)]),
        "gfb_rays_absorption_reset": (I, [P]),
        "gfb_rays_deposit_block": (I, [P, SZ, P, c_double_p, c_double_p, ctypes.POINTER(I)]),
        "gfb_rays_get_dt": (I, [P, c_double_p]),
        "gfb_rays_get_absorbed": (I, [P, ctypes.POINTER(c_double_p)]),
        "gfb_rays_profile": (I, [P, ctypes.POINTER(U64), ctypes.POINTER(SZ)]),
        "gfb_rays_set_binning": (I, [P, I, D, D, ctypes.c_uint, SZ]),
        "gfb_bin_rays": (I, [P, ctypes.c_uint64, D, D, ctypes.c_uint, ctypes.POINTER(ctypes.c_uint64), I, SZ]),
        "gfb_bin_rays_rz": (I, [P, ctypes.POINTER(ctypes.c_uint64), c_double_p, c_double_p, ctypes.POINTER(ctypes.c_uint),
                                ctypes.POINTER(ctypes.c_uint64), I, SZ]),
        "gfb_unbin_rays": (I, [P, ctypes.POINTER(ctypes.c_uint64), I, SZ]),
        "gfb_is_binned": (I, [P]),
        "gfb_bin_disorder": (I, [P, ctypes.POINTER(ctypes.c_uint64), I, c_double_p, c_double_p, ctypes.POINTER(ctypes.c_uint), SZ, c_double_p]),
        "gfb_copy_rays_d2h": (I, [P, ctypes.c_uint64, P, SZ]),
        "gfb_rays_device_ptr": (I, [P, I, c_void_pp]),
        "gfb_rays_ctx": (P, [P]),
        "gfb_rays_source": (S, [P]),
        "gfb_rays_kernel_stats": (I, [P] + [ctypes.POINTER(I)]*6),
        "gfb_rays_rhs": (I, [P, ctypes.POINTER(c_double_p)]),
        "gfb_boris_create": (P, [S, S, SZ, D, I, S]),
        "gfb_boris_destroy": (None, [P]),
        "gfb_boris_set_state": (I, [P, ctypes.POINTER(c_double_p)]),
        "gfb_boris_compile": (I, [P]),
        "gfb_boris_step": (I, [P, SZ]),
        "gfb_boris_get_state": (I, [P, ctypes.POINTER(c_double_p)]),
        "gfb_boris_set_binning": (I, [P, c_double_p, c_double_p, ctypes.POINTER(ctypes.c_uint), SZ]),
        "gfb_boris_info": (I, [P, c_double_p, c_double_p]),
        "gfb_boris_ctx": (P, [P]),
        # graph_c_binding.h
        "graph_construct_context": (P, [I, ctypes.c_bool]),
        "graph_destroy_context": (None, [P]),
        "graph_variable": (P, [P, SZ, S]),
        "graph_constant": (P, [P, D]),
        "graph_set_variable": (None, [P, P, P]),
        "graph_pseudo_variable": (P, [P, P]),
        "graph_remove_pseudo": (P, [P, P]),
        "graph_add": (P, [P, P, P]),
        "graph_sub": (P, [P, P, P]),
        "graph_mul": (P, [P, P, P]),
        "graph_div": (P, [P, P, P]),
        "graph_fma": (P, [P, P, P, P]),
        "graph_sqrt": (P, [P, P]),
        "graph_exp": (P, [P, P]),
        "graph_log": (P, [P, P]),
        "graph_erfi": (P, [P, P]),
        "graph_pow": (P, [P, P, P]),
        "graph_sin": (P, [P, P]),
        "graph_cos": (P, [P, P]),
        "graph_atan": (P, [P, P, P]),
        "graph_piecewise_1D": (P, [P, P, D, D, P, SZ]),
        "graph_piecewise_2D": (P, [P, SZ, P, D, D, P, D, D, P, SZ]),
        "graph_index_1D": (P, [P, P, P, D, D]),
        "graph_index_2D": (P, [P, P, SZ, P, D, D, P, D, D]),
        "graph_df": (P, [P, P, P]),
        "graph_get_max_concurrency": (SZ, [P]),
        "graph_set_device_number": (None, [P, SZ]),
        "graph_add_pre_item": (None, [P, c_void_pp, SZ, c_void_pp, SZ, c_void_pp, c_void_pp, SZ, P, S, SZ]),
        "graph_add_item": (None, [P, c_void_pp, SZ, c_void_pp, SZ, c_void_pp, c_void_pp, SZ, P, S, SZ]),
        "graph_add_converge_item": (None, [P, c_void_pp, SZ, c_void_pp, SZ, c_void_pp, c_void_pp, SZ, P, S, SZ, D, SZ]),
        "graph_compile": (None, [P]),
        "graph_pre_run": (None, [P]),
        "graph_run": (None, [P]),
        "graph_wait": (None, [P]),
        "graph_copy_to_device": (None, [P, P, P]),
        "graph_copy_to_host": (None, [P, P, P]),
        "graph_print": (None, [P, SZ, c_void_pp, SZ]),
        "graph_evaluate": (SZ, [P, P, c_double_p, SZ]),
        "graph_get_source": (S, [P]),
        "graph_set_fast_division": (None, [P, ctypes.c_bool]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)      # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    lib._gfb_signatures = sig
    return lib


lib = _load()


class GfbError(RuntimeError):
    pass


def check(rc, what=""):
    if rc:
        raise GfbError("%s: %s" % (what, lib.gfb_last_error().decode(errors="replace")))
