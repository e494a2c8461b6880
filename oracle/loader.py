"""ctypes access to the test-only CPU checkers in oracle/ (TEST INFRASTRUCTURE -- see fabe13_oracle.c header).

May be imported by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) only.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libfabe13_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_vp = ctypes.c_void_p


def build(ref=True):
    """make the restatement (always) and oracle/_ref (only where the reference source is mounted)."""
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if ref:
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


class Oracle:
    """The C restatement of the reference's SIMD-lane algorithm."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        lib = ctypes.CDLL(ORACLE_SO)
        lib.fabe13_oracle_sincos.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_size_t]
        lib.fabe13_oracle_k.argtypes = [_vp, _vp, ctypes.c_size_t]
        lib.fabe13_oracle_libm_sincos.argtypes = [_vp, _vp, _vp, ctypes.c_size_t]
        lib.fabe13_oracle_scalar_core.argtypes = [ctypes.c_double, _vp, _vp]
        lib.fabe13_oracle_fill_uniform.argtypes = [_vp, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint64,
                                                   ctypes.c_double, ctypes.c_double]
        lib.fabe13_oracle_fill_loguniform.argtypes = [_vp, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint64]
        self.lib = lib

    def sincos(self, x, want_k=True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        s = np.empty_like(x)
        c = np.empty_like(x)
        k = np.empty(x.shape, dtype=np.int64) if want_k else None
        self.lib.fabe13_oracle_sincos(x.ctypes.data, s.ctypes.data, c.ctypes.data,
                                      k.ctypes.data if want_k else None, x.size)
        return (s, c, k) if want_k else (s, c)

    def k(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        k = np.empty(x.shape, dtype=np.int64)
        self.lib.fabe13_oracle_k(x.ctypes.data, k.ctypes.data, x.size)
        return k

    def libm(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        s = np.empty_like(x)
        c = np.empty_like(x)
        self.lib.fabe13_oracle_libm_sincos(x.ctypes.data, s.ctypes.data, c.ctypes.data, x.size)
        return s, c

    def scalar_core(self, x):
        s = ctypes.c_double()
        c = ctypes.c_double()
        self.lib.fabe13_oracle_scalar_core(float(x), ctypes.byref(s), ctypes.byref(c))
        return s.value, c.value

    def fill_uniform(self, n, seed, lo, hi, first_index=0):
        out = np.empty(n, dtype=np.float64)
        self.lib.fabe13_oracle_fill_uniform(out.ctypes.data, n, seed, first_index, lo, hi)
        return out

    def fill_loguniform(self, n, seed, first_index=0):
        out = np.empty(n, dtype=np.float64)
        self.lib.fabe13_oracle_fill_loguniform(out.ctypes.data, n, seed, first_index)
        return out


def cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def ref_variant(kind):
    """Path of the best oracle/_ref build of `kind` ('strict' or 'buildsh') this host can execute, or None."""
    flags = cpu_flags()
    order = []
    if kind == "buildsh":
        host = os.path.join(REF_DIR, "build_host.txt")
        if os.path.exists(host):
            with open(host) as f:
                built_on = set(f.read().split())
            if built_on and built_on <= flags:  # -march=native code only where every ISA extension exists
                order.append("own")     # built by the reference's own build.sh (oracle/build_ref_with_buildsh.sh)
                order.append("native")  # same flags, compiled directly by oracle/Makefile
    if {"avx512f", "avx512dq"} <= flags:
        order.append("avx512")
    if {"avx2", "fma"} <= flags:
        order.append("avx2")
    for v in order:
        p = os.path.join(REF_DIR, f"libfabe13_ref_{kind}_{v}.so")
        if os.path.exists(p):
            return p
    return None


class Reference:
    """The unmodified reference library (oracle/_ref), called through its own fabe13_sincos."""

    def __init__(self, kind="strict", path=None):
        self.path = path or ref_variant(kind)
        if self.path is None:
            raise FileNotFoundError(f"no oracle/_ref build of kind {kind!r} usable on this host")
        lib = ctypes.CDLL(self.path)
        lib.fabe13_sincos.argtypes = [_vp, _vp, _vp, ctypes.c_int]
        lib.fabe13_get_active_implementation_name.restype = ctypes.c_char_p
        lib.fabe13_sin.restype = ctypes.c_double
        lib.fabe13_sin.argtypes = [ctypes.c_double]
        lib.fabe13_cos.restype = ctypes.c_double
        lib.fabe13_cos.argtypes = [ctypes.c_double]
        self.lib = lib
        self.name = lib.fabe13_get_active_implementation_name().decode()
        self.width = lib.fabe13_get_active_simd_width()

    def sincos_bulk(self, x):
        """Every element through the SIMD bulk path: n padded to a multiple of 8 and >= 32."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = x.size
        m = max(32, (n + 7) // 8 * 8)
        xp = np.zeros(m, dtype=np.float64)
        xp[:n] = x.ravel()
        s = np.empty(m, dtype=np.float64)
        c = np.empty(m, dtype=np.float64)
        self.lib.fabe13_sincos(xp.ctypes.data, s.ctypes.data, c.ctypes.data, m)
        return s[:n].reshape(x.shape), c[:n].reshape(x.shape)

    def sincos_raw(self, x, s, c):
        """fabe13_sincos exactly as a user would call it (no padding); used for timing."""
        self.lib.fabe13_sincos(x.ctypes.data, s.ctypes.data, c.ctypes.data, x.size)
