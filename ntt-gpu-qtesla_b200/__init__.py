"""qtesla_b200 — B200-native batched negacyclic polynomial multiplication for qTESLA.

Host-side mirror of the reference's operator surface (benlwk/ntt-gpu-qTESLA, main.cuh:52-71) over
the C ABI of include/qtesla_b200.h.  The directory name carries a hyphen, so import it through
`qtesla_b200_loader.load()` at the repository root.

There is no CPU path in this package: constructing an Engine without the built CUDA library or
without a GPU raises.
"""
from .engine import (Engine, MultiEngine, QtError, lib, get_params, get_table, device_count, polymul_host_multi,
                     SET_I, SET_III, SET_P_I, SET_P_III, SET_NAMES,
                     TABLE_BITREV, TABLE_PHI, TABLE_INVPHI, TABLE_TF0, TABLE_TI0,
                     RING_2P32M1, RING_MODQ, RING_2P32M1_LIFT_Q, LIB_PATH)
from . import harness  # noqa: F401
from . import sharding  # noqa: F401
from . import numa  # noqa: F401

__all__ = ["Engine", "MultiEngine", "QtError", "lib", "get_params", "get_table", "device_count", "polymul_host_multi",
           "SET_I", "SET_III", "SET_P_I", "SET_P_III", "SET_NAMES", "harness", "sharding", "numa", "LIB_PATH",
           "TABLE_BITREV", "TABLE_PHI", "TABLE_INVPHI", "TABLE_TF0", "TABLE_TI0", "RING_2P32M1", "RING_MODQ", "RING_2P32M1_LIFT_Q"]
