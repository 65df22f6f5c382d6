"""varlociraptor_b200 — B200-native per-locus Bayesian posterior engine (varlociraptor `call variants` inner loop).

The compute path is hand-written sm_100a CUDA behind the C-ABI in include/vlr_engine.h
(varlociraptor_b200/csrc); this Python package is the harness-side mirror of the reference's
operator interface (scenario flattening, SoA batch packing, ctypes binding, sharding).
"""
from . import abi  # noqa: F401
from .scenario import Scenario  # noqa: F401
from .batch import LocusBatch  # noqa: F401

__all__ = ["abi", "Scenario", "LocusBatch"]
