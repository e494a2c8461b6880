"""ctypes binding of libfabe13.so (the C ABI declared in include/fabe13.h and include/fabe13_cuda.h).

There is no Python or CPU fallback: if the shared library has not been built (python -m fabe_b200.build, or
__graft_entry__.build()) importing the binding raises.
"""
import ctypes
import os

from .build import LIBPATH

_vp = ctypes.c_void_p  # raw addresses: numpy .ctypes.data / torch .data_ptr()
_lib = None


class Fabe13Error(RuntimeError):
    """A CUDA error reported by libfabe13.so (every int entry point returns it; nothing hides it)."""


def signatures():
    c = ctypes
    return {
        # fabe13.h
        "fabe13_sincos": (None, [_vp, _vp, _vp, c.c_int]),
        "fabe13_get_active_implementation_name": (c.c_char_p, []),
        "fabe13_get_active_simd_width": (c.c_int, []),
        "fabe13_sin": (c.c_double, [c.c_double]),
        "fabe13_cos": (c.c_double, [c.c_double]),
        "fabe13_sinc": (c.c_double, [c.c_double]),
        "fabe13_tan": (c.c_double, [c.c_double]),
        "fabe13_cot": (c.c_double, [c.c_double]),
        "fabe13_atan": (c.c_double, [c.c_double]),
        "fabe13_asin": (c.c_double, [c.c_double]),
        "fabe13_acos": (c.c_double, [c.c_double]),
        # fabe13_cuda.h
        "fabe13_sincos_device": (c.c_int, [_vp, _vp, _vp, c.c_size_t, _vp]),
        "fabe13_sincos_device_k": (c.c_int, [_vp, _vp, _vp, _vp, c.c_size_t, _vp]),
        "fabe13_sin_device": (c.c_int, [_vp, _vp, c.c_size_t, _vp]),
        "fabe13_cos_device": (c.c_int, [_vp, _vp, c.c_size_t, _vp]),
        "fabe13_sin_host": (c.c_int, [_vp, _vp, c.c_size_t]),
        "fabe13_cos_host": (c.c_int, [_vp, _vp, c.c_size_t]),
        "fabe13_tan_device": (c.c_int, [_vp, _vp, c.c_size_t, _vp]),
        "fabe13_cot_device": (c.c_int, [_vp, _vp, c.c_size_t, _vp]),
        "fabe13_sinc_device": (c.c_int, [_vp, _vp, c.c_size_t, _vp]),
        "fabe13_tan_host": (c.c_int, [_vp, _vp, c.c_size_t]),
        "fabe13_cot_host": (c.c_int, [_vp, _vp, c.c_size_t]),
        "fabe13_sinc_host": (c.c_int, [_vp, _vp, c.c_size_t]),
        "fabe13_sincosf_device": (c.c_int, [_vp, _vp, _vp, c.c_size_t, _vp]),
        "fabe13_sincosf_host": (c.c_int, [_vp, _vp, _vp, c.c_size_t]),
        "fabe13_sincos_workspace_bytes": (c.c_size_t, [c.c_size_t]),
        "fabe13_sincos_device_ws": (c.c_int, [_vp, _vp, _vp, c.c_size_t, _vp, c.c_size_t, _vp]),
        "fabe13_sincos_multi": (c.c_int, [c.c_int, _vp, _vp, _vp, _vp, _vp]),
        "fabe13_sincos_host": (c.c_int, [_vp, _vp, _vp, c.c_size_t]),
        "fabe13_cuda_host_alloc": (_vp, [c.c_size_t]),
        "fabe13_cuda_host_free": (None, [_vp]),
        "fabe13_cuda_host_register": (c.c_int, [_vp, c.c_size_t]),
        "fabe13_cuda_host_unregister": (c.c_int, [_vp]),
        "fabe13_shard_bounds": (None, [c.c_size_t, c.c_int, c.c_int, c.POINTER(c.c_size_t), c.POINTER(c.c_size_t)]),
        "fabe13_cuda_set_option": (c.c_int, [c.c_char_p, c.c_longlong]),
        "fabe13_cuda_get_option": (c.c_int, [c.c_char_p, c.POINTER(c.c_longlong)]),
        "fabe13_cuda_device_count": (c.c_int, []),
        "fabe13_cuda_last_error": (c.c_int, []),
        "fabe13_cuda_last_error_string": (c.c_char_p, []),
        "fabe13_cuda_clear_error": (None, []),
        "fabe13_debug_host_core": (None, [_vp, _vp, _vp, _vp, c.c_size_t, c.c_int]),
        "fabe13_debug_host_sincosf": (None, [_vp, _vp, _vp, c.c_size_t, c.c_int]),
        "fabe13_debug_host_derived": (None, [_vp, _vp, c.c_size_t, c.c_int, c.c_int]),
        "fabe13_debug_host_array": (None, [_vp, _vp, _vp, c.c_size_t, c.c_int, c.c_int, c.c_int]),
        "fabe13_debug_reduce_huge": (None, [_vp, _vp, _vp, _vp, _vp, _vp, c.c_size_t]),
        "fabe13_debug_host_array_stores": (None, [_vp, _vp, _vp, c.c_size_t, c.c_int, c.c_int, c.c_int, c.c_int]),
        "fabe13_cuda_trim": (c.c_int, []),
        "fabe13_cuda_shutdown": (c.c_int, []),
        "fabe13_debug_fill_uniform_device": (c.c_int, [_vp, c.c_size_t, c.c_uint64, c.c_uint64, c.c_double, c.c_double, _vp]),
        "fabe13_debug_fill_loguniform_device": (c.c_int, [_vp, c.c_size_t, c.c_uint64, c.c_uint64, _vp]),
    }


def load():
    """Load libfabe13.so once; raise loudly if it is missing or does not export the declared ABI."""
    global _lib
    if _lib is None:
        path = os.environ.get("FABE13_LIBRARY") or LIBPATH  # override: A/B builds of kernel variants (tools/)
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: build it with `python -m fabe_b200.build` "
                "(there is no fallback implementation)"
            )
        lib = ctypes.CDLL(path)
        for name, (res, args) in signatures().items():
            fn = getattr(lib, name)  # AttributeError = the .so does not export what the headers declare
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().fabe13_cuda_last_error_string().decode()
        raise Fabe13Error(f"{what} failed: CUDA error {rc} ({msg})")
