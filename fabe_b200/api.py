"""Python face of the C ABI: array sincos on host arrays (numpy) or device tensors (torch.cuda), scalar functions,
options and the slice-sharding helper.  Everything forwards to libfabe13.so; nothing is computed in Python."""
import ctypes

import numpy as np

from . import _lib

CORE_POLY = 0
CORE_PSI = 1


def _is_torch(x):
    return type(x).__module__.split(".")[0] == "torch"


def _check_f64(x, name):
    if _is_torch(x):
        import torch

        if x.dtype != torch.float64 or not x.is_contiguous():
            raise TypeError(f"{name} must be a contiguous float64 tensor")
    else:
        if x.dtype != np.float64 or not x.flags["C_CONTIGUOUS"]:
            raise TypeError(f"{name} must be a C-contiguous float64 array")


def _check_out(out, x, name, dtype_name="float64"):
    """A caller-provided output must be the same kind of object as x, with x's element count, dtype and device:
    the library writes numel(x) elements through the raw pointer and cannot check any of this itself."""
    if _is_torch(out) != _is_torch(x):
        raise TypeError(f"{name} must be a {'torch tensor' if _is_torch(x) else 'numpy array'} like x")
    if _is_torch(x):
        import torch

        if out.dtype != getattr(torch, dtype_name) or not out.is_contiguous():
            raise TypeError(f"{name} must be a contiguous {dtype_name} tensor")
        if out.device != x.device:
            raise ValueError(f"{name} is on {out.device}, x on {x.device}")
    elif out.dtype != np.dtype(dtype_name) or not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]:
        raise TypeError(f"{name} must be a writeable C-contiguous {dtype_name} array")
    if _numel(out) != _numel(x):
        raise ValueError(f"{name} has {_numel(out)} elements, x has {_numel(x)}")


def _ptr(x):
    return x.data_ptr() if _is_torch(x) else x.ctypes.data


def _numel(x):
    return x.numel() if _is_torch(x) else x.size


def sincos(x, out_sin=None, out_cos=None, stream=None):
    """(sin(x), cos(x)) element-wise for a float64 array.

    torch CUDA tensor -> fabe13_sincos_device on the tensor's device and torch's current stream (asynchronous);
    numpy array / torch CPU tensor -> fabe13_sincos_host (synchronous; pinned memory takes the direct-copy path).
    """
    lib = _lib.load()
    _check_f64(x, "x")
    if _is_torch(x):
        import torch

        s = torch.empty_like(x) if out_sin is None else out_sin
        c = torch.empty_like(x) if out_cos is None else out_cos
        _check_out(s, x, "out_sin")
        _check_out(c, x, "out_cos")
        if x.is_cuda:
            with torch.cuda.device(x.device):
                st = torch.cuda.current_stream().cuda_stream if stream is None else stream
                rc = lib.fabe13_sincos_device(x.data_ptr(), s.data_ptr(), c.data_ptr(), x.numel(), st)
            _lib.check(rc, "fabe13_sincos_device")
            return s, c
        rc = lib.fabe13_sincos_host(x.data_ptr(), s.data_ptr(), c.data_ptr(), x.numel())
        _lib.check(rc, "fabe13_sincos_host")
        return s, c
    s = np.empty_like(x) if out_sin is None else out_sin
    c = np.empty_like(x) if out_cos is None else out_cos
    _check_out(s, x, "out_sin")
    _check_out(c, x, "out_cos")
    rc = lib.fabe13_sincos_host(x.ctypes.data, s.ctypes.data, c.ctypes.data, x.size)
    _lib.check(rc, "fabe13_sincos_host")
    return s, c


def _single(x, out, dev_fn, host_fn, what):
    lib = _lib.load()
    _check_f64(x, "x")
    if _is_torch(x):
        import torch

        o = torch.empty_like(x) if out is None else out
        _check_out(o, x, "out")
        if x.is_cuda:
            with torch.cuda.device(x.device):
                rc = getattr(lib, dev_fn)(x.data_ptr(), o.data_ptr(), x.numel(), torch.cuda.current_stream().cuda_stream)
        else:
            rc = getattr(lib, host_fn)(x.data_ptr(), o.data_ptr(), x.numel())
        _lib.check(rc, what)
        return o
    o = np.empty_like(x) if out is None else out
    _check_out(o, x, "out")
    _lib.check(getattr(lib, host_fn)(x.ctypes.data, o.ctypes.data, x.size), what)
    return o


def sin_array(x, out=None):
    """sin(x) element-wise (array form of fabe13_sin): CUDA tensor -> fabe13_sin_device, host array -> fabe13_sin_host."""
    return _single(x, out, "fabe13_sin_device", "fabe13_sin_host", "fabe13_sin")


def cos_array(x, out=None):
    """cos(x) element-wise (array form of fabe13_cos)."""
    return _single(x, out, "fabe13_cos_device", "fabe13_cos_host", "fabe13_cos")


def tan_array(x, out=None):
    """tan(x) element-wise (array form of fabe13_tan: sin/cos, NaN where cos == 0)."""
    return _single(x, out, "fabe13_tan_device", "fabe13_tan_host", "fabe13_tan")


def cot_array(x, out=None):
    """cot(x) element-wise (array form of fabe13_cot: cos/sin, NaN where sin == 0)."""
    return _single(x, out, "fabe13_cot_device", "fabe13_cot_host", "fabe13_cot")


def sinc_array(x, out=None):
    """sinc(x) = sin(x)/x element-wise, 1 for |x| < 1e-9 (array form of fabe13_sinc)."""
    return _single(x, out, "fabe13_sinc_device", "fabe13_sinc_host", "fabe13_sinc")


OP_TAN, OP_COT, OP_SINC = 3, 4, 5


def sincosf(x, out_sin=None, out_cos=None):
    """(sin(x), cos(x)) for a float32 array: torch CUDA tensor -> fabe13_sincosf_device on the current stream;
    numpy array / torch CPU tensor -> fabe13_sincosf_host (synchronous)."""
    lib = _lib.load()
    if _is_torch(x):
        import torch

        if x.dtype != torch.float32 or not x.is_contiguous():
            raise TypeError("x must be a contiguous float32 tensor")
        s = torch.empty_like(x) if out_sin is None else out_sin
        c = torch.empty_like(x) if out_cos is None else out_cos
        _check_out(s, x, "out_sin", "float32")
        _check_out(c, x, "out_cos", "float32")
        if x.is_cuda:
            with torch.cuda.device(x.device):
                rc = lib.fabe13_sincosf_device(x.data_ptr(), s.data_ptr(), c.data_ptr(), x.numel(), torch.cuda.current_stream().cuda_stream)
        else:
            rc = lib.fabe13_sincosf_host(x.data_ptr(), s.data_ptr(), c.data_ptr(), x.numel())
        _lib.check(rc, "fabe13_sincosf")
        return s, c
    if x.dtype != np.float32 or not x.flags["C_CONTIGUOUS"]:
        raise TypeError("x must be a C-contiguous float32 array")
    s = np.empty_like(x) if out_sin is None else out_sin
    c = np.empty_like(x) if out_cos is None else out_cos
    _check_out(s, x, "out_sin", "float32")
    _check_out(c, x, "out_cos", "float32")
    _lib.check(lib.fabe13_sincosf_host(x.ctypes.data, s.ctypes.data, c.ctypes.data, x.size), "fabe13_sincosf_host")
    return s, c


def host_sincosf(x, core=CORE_POLY):
    """Test hook: the FP32 entry's arithmetic on the host (widen, shared FP64 core, round once)."""
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    s = np.empty_like(x)
    c = np.empty_like(x)
    lib.fabe13_debug_host_sincosf(x.ctypes.data, s.ctypes.data, c.ctypes.data, x.size, core)
    return s, c


def sincos_k(x):
    """(sin, cos, k) for a torch CUDA float64 tensor; k is the reference's quadrant index (test hook)."""
    import torch

    lib = _lib.load()
    _check_f64(x, "x")
    if not (_is_torch(x) and x.is_cuda):
        raise TypeError("sincos_k needs a CUDA tensor")
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    k = torch.empty(x.shape, dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.fabe13_sincos_device_k(x.data_ptr(), s.data_ptr(), c.data_ptr(), k.data_ptr(), x.numel(), st)
    _lib.check(rc, "fabe13_sincos_device_k")
    return s, c, k


def sincos_legacy(x, s, c):
    """The reference's own entry point, void fabe13_sincos(in, sin_out, cos_out, int n), on host or device memory."""
    lib = _lib.load()
    _check_f64(x, "x")
    _check_out(s, x, "s")
    _check_out(c, x, "c")
    if _numel(x) > 0x7FFFFFFF:
        raise ValueError("fabe13_sincos takes an int n (reference fabe13.h:49): at most 2^31 - 1 elements")
    lib.fabe13_cuda_clear_error()
    lib.fabe13_sincos(_ptr(x), _ptr(s), _ptr(c), _numel(x))
    _lib.check(lib.fabe13_cuda_last_error(), "fabe13_sincos")


def sin(x):
    return _lib.load().fabe13_sin(float(x))


def cos(x):
    return _lib.load().fabe13_cos(float(x))


def sinc(x):
    return _lib.load().fabe13_sinc(float(x))


def tan(x):
    return _lib.load().fabe13_tan(float(x))


def cot(x):
    return _lib.load().fabe13_cot(float(x))


def implementation_name():
    return _lib.load().fabe13_get_active_implementation_name().decode()


def simd_width():
    return _lib.load().fabe13_get_active_simd_width()


def set_option(key, value):
    rc = _lib.load().fabe13_cuda_set_option(key.encode(), int(value))
    if rc != 0:
        raise ValueError(f"fabe13_cuda_set_option({key!r}, {value}) rejected (rc={rc})")


def get_option(key):
    v = ctypes.c_longlong(0)
    rc = _lib.load().fabe13_cuda_get_option(key.encode(), ctypes.byref(v))
    if rc != 0:
        raise KeyError(key)
    return v.value


def shard_bounds(n, world, rank):
    """[begin, end) of the contiguous slice `rank` owns; interior cuts are multiples of 4 elements (32 bytes)."""
    b = ctypes.c_size_t(0)
    e = ctypes.c_size_t(0)
    _lib.load().fabe13_shard_bounds(n, world, rank, ctypes.byref(b), ctypes.byref(e))
    return b.value, e.value


class _PinnedBlock:
    """Owner of one fabe13_cuda_host_alloc block; exposes the buffer protocol so numpy views keep it alive."""

    def __init__(self, lib, nbytes):
        self._lib = lib
        self._p = lib.fabe13_cuda_host_alloc(max(int(nbytes), 8))
        if not self._p:
            raise MemoryError(f"fabe13_cuda_host_alloc({nbytes}) failed")
        self.buf = (ctypes.c_char * int(nbytes)).from_address(self._p)

    def release(self):
        if self._p:
            self._lib.fabe13_cuda_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class _PinnedNdarray(np.ndarray):
    """ndarray whose views carry a reference to the pinned block they look into (through .base chains and this
    attribute), so the page-locked memory is freed only when the last view is gone."""

    _fabe13_owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._fabe13_owner = getattr(obj, "_fabe13_owner", None)


class PinnedArray:
    """A float64 numpy array in page-locked host memory from fabe13_cuda_host_alloc (fast host path).

    `.array` (and every view or slice of it) keeps the block alive: `a = PinnedArray(n).array` is safe, the memory is
    returned when the last such array is garbage-collected.  free() drops this wrapper's own reference only."""

    def __init__(self, n):
        lib = _lib.load()
        n = int(n)
        block = _PinnedBlock(lib, n * 8)
        arr = np.frombuffer(block.buf, dtype=np.float64, count=n).view(_PinnedNdarray)
        arr._fabe13_owner = block
        self.array = arr

    def free(self):
        """Drop this wrapper's reference; the block is freed once no view of .array is left."""
        self.array = None


def fill_uniform_device(t, seed, lo, hi, first_index=0):
    """Bench/test hook: fill a CUDA float64 tensor with the synthetic uniform inputs of SURVEY.md section 8d
    (bit-identical to oracle fabe13_oracle_fill_uniform)."""
    import torch

    lib = _lib.load()
    with torch.cuda.device(t.device):
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.fabe13_debug_fill_uniform_device(t.data_ptr(), t.numel(), seed, first_index, lo, hi, st)
    _lib.check(rc, "fabe13_debug_fill_uniform_device")
    return t


def fill_loguniform_device(t, seed, first_index=0):
    """Bench/test hook: |x| log-uniform over [1, 2^1024) with random sign (BASELINE config 4)."""
    import torch

    lib = _lib.load()
    with torch.cuda.device(t.device):
        st = torch.cuda.current_stream().cuda_stream
        rc = lib.fabe13_debug_fill_loguniform_device(t.data_ptr(), t.numel(), seed, first_index, st)
    _lib.check(rc, "fabe13_debug_fill_loguniform_device")
    return t


def host_derived(x, op, core=CORE_POLY):
    """Test hook: tan / cot / sinc (op = OP_TAN / OP_COT / OP_SINC) evaluated by the shared core on the host."""
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib.fabe13_debug_host_derived(x.ctypes.data, out.ctypes.data, x.size, op, core)
    return out


def host_core(x, core=CORE_POLY):
    """Test hook: the shared per-element core evaluated on the host (fabe13_debug_host_core). Not a product path."""
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    s = np.empty_like(x)
    c = np.empty_like(x)
    k = np.empty(x.shape, dtype=np.int64)
    lib.fabe13_debug_host_core(x.ctypes.data, s.ctypes.data, c.ctypes.data, k.ctypes.data, x.size, core)
    return s, c, k
