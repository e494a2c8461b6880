"""fabe_b200 -- B200-native (sm_100a) FP64 array sincos behind the FABE13 C API.

The product is libfabe13.so (fabe_b200/lib/, built from fabe_b200/csrc/ by fabe_b200/build.py); this package
is its ctypes binding.  See DESIGN.md and INTEGRATION.md.
"""
from ._lib import Fabe13Error  # noqa: F401
from .api import (  # noqa: F401
    CORE_POLY,
    CORE_PSI,
    OP_COT,
    OP_SINC,
    OP_TAN,
    PinnedArray,
    cos,
    cos_array,
    cot,
    cot_array,
    fill_loguniform_device,
    fill_uniform_device,
    get_option,
    host_core,
    host_derived,
    host_sincosf,
    implementation_name,
    set_option,
    shard_bounds,
    simd_width,
    sin,
    sin_array,
    sinc,
    sinc_array,
    sincos,
    sincos_k,
    sincosf,
    sincos_legacy,
    tan,
    tan_array,
)
