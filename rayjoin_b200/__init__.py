"""rayjoin_b200 -- B200-native LSI / PIP / polygon-overlay engine.

Host-side mirror of RayJoin's operator interface over the C ABI in
include/rjb200.h (librjb200.so, hand-written CUDA for sm_100a).
"""
from .capi import (Context, LSI, PIP, MapOverlay, PlanarGraph, RjbError, load_from,  # noqa: F401
                   read_pgraph, serialize_pgraph, deserialize_pgraph, load_library,
                   MODE_GRID, MODE_LBVH, MODE_BRUTE, NO_HIT, XSECT_DTYPE)
