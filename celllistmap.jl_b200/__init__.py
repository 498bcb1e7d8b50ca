"""celllistmap.jl_b200 -- B200-native cutoff-pair engine behind CellListMap.jl's operator interface.

Only what the hot path needs lives here: `csrc/` (hand-written sm_100a CUDA kernels + the C ABI of
include/clm_b200.h), `_capi.py` (ctypes binding of that ABI), `api.py` (host-side mirror of the reference's
ParticleSystem / pairwise! / neighborlist interface), `slab.py` (multi-GPU slab decomposition), and
`julia/` (the `ccall` host package for a Julia box).

The directory name is not a Python identifier; import it through the root shim `celllistmap_b200`.
"""
from . import _capi
from ._capi import ClmError, Handle, nl_dtype, SO_PATH
from .api import *  # noqa: F401,F403
from .api import __all__ as _api_all

__all__ = ["ClmError", "Handle", "nl_dtype", "SO_PATH"] + list(_api_all)
