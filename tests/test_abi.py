"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol that
include/panacus_b200.h declares, reports errors through status codes, and the host-side helpers
(thresholds, packing) behave like the reference's."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import panacus_b200 as pb
from panacus_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "panacus_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pgx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = pb.lib()
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/panacus_b200.h but not exported"
    assert sorted(_native.EXPORTS) == declared


def test_version_and_row_words():
    L = pb.lib()
    assert b"sm_100a" in L.pgx_version()
    for G in (1, 63, 64, 65, 128, 129, 192, 256, 1000, 1024, 1025):
        assert L.pgx_row_words(G) == pb.row_words(G)
        w = (G + 63) // 64
        assert pb.row_words(G) >= w and (pb.row_words(G) == 1 or pb.row_words(G) % 2 == 0)
    assert L.pgx_fused_out_words(1024, 3) == 2 * 1025 + 3 * 1024


def test_errors_are_status_codes_not_crashes():
    L = pb.lib()
    h = C.c_void_p()
    # no CUDA device here (or a bad argument on a GPU box): must return a negative status + message
    rc = L.pgx_abacus_create(C.byref(h), 0, 10, 0)
    assert rc < 0 and L.pgx_last_error()
    assert L.pgx_hist(None, None, None, None) == -1
    assert L.pgx_ordered_growth(None, 0, None, None, None, 0, None) == -1
    assert L.pgx_similarity(None, 0, 0, 0, None, None) == -1
    assert L.pgx_similarity_upper(None, 0, 0, 0, None, None) == -1
    L.pgx_abacus_destroy(None)  # no-op


def test_threshold_semantics():
    T = pb.Threshold
    assert T.relative(0.5).to_absolute(5) == 3 and T.absolute(2).to_absolute(5) == 2
    assert T.absolute(2).to_relative(4) == 0.5 and T.relative(0.25).to_relative(9) == 0.25
    assert [T.relative(x).get_string() for x in (0.0, 0.5, 1.0)] == ["0", "0.5", "1"]
    assert T.absolute(7).get_string() == "7"
    c, thr = pb.growth_cutoffs(10, T.absolute(0), T.relative(0.3))
    assert c == 1
    assert list(thr) == [math.ceil((g + 1.0) * 0.3) for g in range(10)]
    assert not pb.quorum_thresholds(8, 0.0).any()


def test_pack_bits_layout():
    bits = np.zeros((3, 70), dtype=np.uint8)
    bits[1, 0] = bits[1, 63] = bits[2, 64] = bits[2, 69] = 1
    bm = pb.pack_bits(bits)
    assert bm.shape == (3, 2) and bm.dtype == np.uint64
    assert bm[1, 0] == np.uint64(1) | (np.uint64(1) << np.uint64(63)) and bm[1, 1] == 0
    assert bm[2, 0] == 0 and bm[2, 1] == np.uint64(1) | (np.uint64(1) << np.uint64(5))
