"""torch.ops.fabe13.* -- the array entry points as PyTorch custom operators (the reference's
roadmap item README.md:227).  Importing this module registers

    torch.ops.fabe13.sincos(Tensor x) -> (Tensor sin, Tensor cos)
    torch.ops.fabe13.sin(Tensor x) -> Tensor        torch.ops.fabe13.cos(Tensor x) -> Tensor
    torch.ops.fabe13.tan(Tensor x) -> Tensor        torch.ops.fabe13.cot(Tensor x) -> Tensor
    torch.ops.fabe13.sinc(Tensor x) -> Tensor       (unnormalised: sin(x)/x, 1 for |x| < 1e-9, as the reference's)
    torch.ops.fabe13.sincosf(Tensor x) -> (Tensor sin, Tensor cos)      float32 tensors

for float64 tensors: CUDA tensors go to fabe13_sincos_device on the current stream, CPU tensors to fabe13_sincos_host.
Non-contiguous inputs are made contiguous first.  The operators carry fake (meta) implementations, so they trace
under torch.compile / export; they define no autograd formula (the C library is forward-only).
"""
import torch

from . import api

_DEFINED = False


def _prep(x):
    if x.dtype != torch.float64:
        raise TypeError("fabe13 operators take float64 tensors")
    return x.contiguous()


def register():
    global _DEFINED
    if _DEFINED:
        return
    _DEFINED = True

    @torch.library.custom_op("fabe13::sincos", mutates_args=())
    def sincos(x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        xc = _prep(x)
        s, c = api.sincos(xc)
        return s.view(x.shape), c.view(x.shape)

    @sincos.register_fake
    def _(x):
        return torch.empty_like(x, memory_format=torch.contiguous_format), torch.empty_like(x, memory_format=torch.contiguous_format)

    @torch.library.custom_op("fabe13::sin", mutates_args=())
    def sin(x: torch.Tensor) -> torch.Tensor:
        return api.sin_array(_prep(x)).view(x.shape)

    @sin.register_fake
    def _(x):
        return torch.empty_like(x, memory_format=torch.contiguous_format)

    @torch.library.custom_op("fabe13::cos", mutates_args=())
    def cos(x: torch.Tensor) -> torch.Tensor:
        return api.cos_array(_prep(x)).view(x.shape)

    @cos.register_fake
    def _(x):
        return torch.empty_like(x, memory_format=torch.contiguous_format)

    def single(name, fn):
        op = torch.library.custom_op(f"fabe13::{name}", mutates_args=())(fn)
        op.register_fake(lambda x: torch.empty_like(x, memory_format=torch.contiguous_format))

    def tan(x: torch.Tensor) -> torch.Tensor:
        return api.tan_array(_prep(x)).view(x.shape)

    def cot(x: torch.Tensor) -> torch.Tensor:
        return api.cot_array(_prep(x)).view(x.shape)

    def sinc(x: torch.Tensor) -> torch.Tensor:
        return api.sinc_array(_prep(x)).view(x.shape)

    @torch.library.custom_op("fabe13::sincosf", mutates_args=())
    def sincosf(x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        if x.dtype != torch.float32:
            raise TypeError("fabe13::sincosf takes float32 tensors")
        s, c = api.sincosf(x.contiguous())
        return s.view(x.shape), c.view(x.shape)

    @sincosf.register_fake
    def _(x):
        return torch.empty_like(x, memory_format=torch.contiguous_format), torch.empty_like(x, memory_format=torch.contiguous_format)

    single("tan", tan)
    single("cot", cot)
    single("sinc", sinc)


register()
