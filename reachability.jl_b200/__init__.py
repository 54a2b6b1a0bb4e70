"""reachability.jl_b200 -- B200-native (sm_100a) linear-flowpipe hot path behind the Reachability.jl API.

Only what the path needs: ``csrc/`` (CUDA kernels + the C ABI of include/reach.h, built into
``libreach_sm100a.so``), the ctypes binding, and a host-side mirror of the reference's operator interface
(`solve`, `BFFPSV18`, `GLGM06`, `discretize`, result containers) so that parity tests read like the reference's.
"""
from ._lib import ReachError, LIB_PATH, EXPORTED_SYMBOLS  # noqa: F401
from .backend import Context, BoxPlan, Flowpipe, Ensemble  # noqa: F401

from . import api  # noqa: F401,E402

__version__ = "0.1.0"
