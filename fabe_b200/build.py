"""Build libfabe13.so (the C-ABI library with the sm_100a kernels) in-tree with nvcc.

The .so lands in fabe_b200/lib/ so that it travels with the repository snapshot to the GPU box; nothing is
JIT-compiled at import time.  `python -m fabe_b200.build` rebuilds it.
"""
import os
import platform
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libfabe13.so")

SOURCES = ["fabe13_kernels.cu", "fabe13_shim.cu", "fabe13_hostcore.cpp"]
HEADERS = ["fabe13_core.h", "fabe13_constants.h", "fabe13_ph_remainders.h", "fabe13_internal.h", "fabe13_hostcore.h"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libfabe13.so")


def _stale():
    if not os.path.exists(LIBPATH):
        return True
    t = os.path.getmtime(LIBPATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps += [os.path.join(HERE, "..", "include", f) for f in ("fabe13.h", "fabe13_cuda.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, out=None):
    """Compile csrc/*.cu for sm_100a into lib/libfabe13.so (or `out`).  Returns the path.

    FABE13_BUILD_DEFINES="A=1 B" in the environment adds -DA=1 -DB (used by tools/ for A/B builds of kernel variants;
    pair it with `out` and load the result through FABE13_LIBRARY)."""
    if out is None and not force and not _stale():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    target = out or LIBPATH
    if out:
        os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    cmd = [
        _nvcc(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-lineinfo", "-std=c++17",
        "-shared", "-cudart", "static",
        # host side: keep * and + as separate roundings (the host core must match the device bit for bit); -O3 so that
        # the host array loop (fabe13_hostcore.cpp) is vectorised; -mfma only where the flag exists (x86: fma() must
        # inline to the instruction; aarch64 has fused multiply-add in its base ISA and gcc rejects the flag there)
        "-Xcompiler", "-fPIC,-O3,-ffp-contract=off,-fvisibility=hidden" + (",-mfma" if platform.machine() in ("x86_64", "AMD64") else ""),
        "-Xptxas", "-v" if verbose else "-O3",
    ] + [f"-D{d}" for d in os.environ.get("FABE13_BUILD_DEFINES", "").split() if d] + [
        "-o", target,
    ] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libfabe13.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return target


if __name__ == "__main__":
    outs = [a for a in sys.argv[1:] if not a.startswith("-")]
    print(build_library(force=True, verbose="-v" in sys.argv, out=outs[0] if outs else None))
