"""sidm_b200 - B200-native hot path of junkoda/sidm-nbody (tree gravity + SIDM scatter).

The compute lives in libsidm_b200.so (hand-written CUDA for sm_100a, C ABI in
include/sidm_b200.h); this package is only the binding and the host-side mirror of the
reference's entry points.  There is no CPU fallback.
"""
from . import capi, ic  # noqa: F401
from .capi import B200Error, PARTICLE_DTYPE  # noqa: F401
from .hotpath import HotPath, make_params  # noqa: F401
