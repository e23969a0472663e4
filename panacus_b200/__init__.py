"""panacus_b200 -- B200-native implementation of panacus's counting hot path
(coverage histogram, ordered / permuted growth, group intersections) behind a C ABI
(include/panacus_b200.h, libpanacus_b200.so).  This package is the thin Python plumbing around it."""
from ._native import LIB_PATH, NativeLibraryMissing, PgxError, lib  # noqa: F401
from .abacus import (Comm, DeviceAbacus, Threshold, similarity_shard_bounds, curve_from_fused, growth_cutoffs, pack_bits, pinned_empty,  # noqa: F401
                     quorum_thresholds, row_words)

__version__ = "0.1.0"
