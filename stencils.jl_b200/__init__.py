"""stencils_b200 — B200-native stencil-mapping engine behind the API of rafaqz/Stencils.jl.

Host-side mirror of the reference's Julia API (src/Stencils.jl:11-22 export list) over the C ABI in
include/stencils_b200.h. The sweep itself runs in hand-written sm_100a CUDA kernels
(csrc/ -> lib/libstencils_b200.so); there is no CPU or PyTorch fallback.
"""
from . import _abi
from ._abi import ArgumentError, SB200Error
