"""stencils_b200 — B200-native stencil-mapping engine behind the API of rafaqz/Stencils.jl.

Host-side mirror of the reference's Julia API (src/Stencils.jl:11-22 export list) over the C ABI in
include/stencils_b200.h. The sweep itself runs in hand-written sm_100a CUDA kernels
(csrc/ -> lib/libstencils_b200.so); there is no CPU or PyTorch fallback.

Julia's mutating `f!` functions are spelled `f_` here (mapstencil! -> mapstencil_).
"""
from . import _abi
from ._abi import ArgumentError, SB200Error
from .stencils import (Annulus, AngledCross, BackSlash, Cardinal, Circle, Cross, Diamond, ForwardSlash, Horizontal,
                       Kernel, Layered, Moore, NamedStencil, Ordinal, Positional, Rectangle, Stencil, Vertical, VonNeumann,
                       Window, layer, center, diameter, distance_zones, distances, indices, merge, neighbors, offsets, radius)
from .array import (AbstractStencilArray, BoundaryCondition, Conditional, Halo, Padding, Reflect, Remove,
                    StencilArray, SwitchingStencilArray, Use, Wrap, boundary, colmajor_empty, dest, padding, padval,
                    source, stencil, switch)
from .ops import (Diffusion, Life, LinearCombination, ScatterCenterWeights, ScatterWeights, gatherstencil, gatherstencil_, iterate_,
                  kernelproduct, mapstencil, mapstencil_, maximum, mean, minimum, scatterstencil_, update_boundary_)
from .ops import sum  # noqa: A004  (Julia's `sum` applied to a stencil)

__all__ = [n for n in dir() if not n.startswith("_")]
