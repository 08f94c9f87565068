"""
B200-native SwiFTly facet<->subgrid distributed Fourier transform hot path.

Drop-in for the compute core of ``ska_sdp_exec_swiftly``: the same
``SwiftlyConfig`` / ``SwiftlyForward`` / ``SwiftlyBackward`` API surface and the
eight-primitive core interface, implemented as hand-written sm_100a CUDA kernels
behind a C ABI (``include/swiftly_b200.h``).  No CPU fallback.
"""

from .core import SwiftlyCoreB200  # noqa: F401

__version__ = "0.1.0"
