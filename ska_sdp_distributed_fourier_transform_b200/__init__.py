"""
B200-native SwiFTly facet<->subgrid distributed Fourier transform hot path.

Drop-in for the compute core of ``ska_sdp_exec_swiftly``: the same
``SwiftlyConfig`` / ``SwiftlyForward`` / ``SwiftlyBackward`` API surface and the
eight-primitive core interface, implemented as hand-written sm_100a CUDA kernels
behind a C ABI (``include/swiftly_b200.h``).  No CPU fallback.
"""

from .api import (  # noqa: F401
    FacetConfig,
    SubgridConfig,
    SwiftlyBackward,
    SwiftlyConfig,
    SwiftlyForward,
    make_full_facet_cover,
    make_full_subgrid_cover,
)
from .api_helper import check_facet, check_subgrid, make_facet, make_subgrid  # noqa: F401
from .core import SwiftlyCoreB200  # noqa: F401
from .distributed import (  # noqa: F401
    SwiftlyBackwardSharded,
    SwiftlyForwardSharded,
    partition_facets,
)
from .fourier_algorithm import make_facet_from_sources, make_subgrid_from_sources  # noqa: F401
from .swift_configs import SWIFT_CONFIGS  # noqa: F401

__all__ = [
    "FacetConfig",
    "SubgridConfig",
    "SwiftlyConfig",
    "SwiftlyForward",
    "SwiftlyBackward",
    "SwiftlyCoreB200",
    "SwiftlyForwardSharded",
    "SwiftlyBackwardSharded",
    "partition_facets",
    "SWIFT_CONFIGS",
    "check_facet",
    "check_subgrid",
    "make_subgrid",
    "make_facet",
    "make_full_facet_cover",
    "make_full_subgrid_cover",
    "make_facet_from_sources",
    "make_subgrid_from_sources",
]

__version__ = "0.1.0"
