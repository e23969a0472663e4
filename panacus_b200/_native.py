"""Loader for libpanacus_b200.so (the C ABI declared in include/panacus_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded, every hot-path
entry point raises.  Build it with ``python -c 'import __graft_entry__ as g; g.build()'`` or
``make -C panacus_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpanacus_b200.so")

# every symbol include/panacus_b200.h declares (tests check that the .so exports all of them)
EXPORTS = [
    "pgx_version", "pgx_last_error", "pgx_device_count", "pgx_device_warmup", "pgx_row_words",
    "pgx_abacus_create", "pgx_abacus_destroy", "pgx_abacus_set_stream", "pgx_abacus_shape",
    "pgx_abacus_upload", "pgx_abacus_adopt_device", "pgx_abacus_scatter", "pgx_abacus_build", "pgx_abacus_build_u32", "pgx_host_alloc", "pgx_host_free", "pgx_abacus_clear",
    "pgx_abacus_download", "pgx_abacus_copy_rows", "pgx_abacus_csr_rows", "pgx_abacus_csr_fill", "pgx_hist", "pgx_ordered_growth", "pgx_hist_ordered_growth",
    "pgx_permuted_growth", "pgx_similarity", "pgx_similarity_upper", "pgx_fused_out_words", "pgx_fused_pass_async",
    "pgx_launch_count", "pgx_last_launch_info", "pgx_abacus_set_timing", "pgx_kernel_time_ms",
    "pgx_exchange_export", "pgx_exchange_connect", "pgx_exchange_disconnect", "pgx_exchange_status",
    "pgx_comm_unique_id", "pgx_comm_create", "pgx_comm_create_all", "pgx_comm_destroy", "pgx_comm_info",
    "pgx_abacus_broadcast", "pgx_exchange_connect_comm", "pgx_hist_ordered_growth_sharded",
    "pgx_permuted_growth_sharded", "pgx_similarity_sharded", "pgx_similarity_shard_bounds",
]
EXCHANGE_HANDLE_BYTES = 64
COMM_ID_BYTES = 128

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: the CUDA extension is not built (run __graft_entry__.build()); "
            "panacus_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    u64p, u32p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    vp = C.c_void_p
    L.pgx_version.restype = C.c_char_p
    L.pgx_version.argtypes = []
    L.pgx_last_error.restype = C.c_char_p
    L.pgx_last_error.argtypes = []
    L.pgx_device_count.restype = C.c_int
    L.pgx_device_count.argtypes = [C.POINTER(C.c_int)]
    L.pgx_device_warmup.restype = C.c_int
    L.pgx_device_warmup.argtypes = [C.c_int]
    L.pgx_row_words.restype = C.c_uint32
    L.pgx_row_words.argtypes = [C.c_uint32]
    L.pgx_abacus_create.restype = C.c_int
    L.pgx_abacus_create.argtypes = [C.POINTER(vp), C.c_int, C.c_uint64, C.c_uint32]
    L.pgx_abacus_destroy.restype = None
    L.pgx_abacus_destroy.argtypes = [vp]
    L.pgx_abacus_set_stream.restype = C.c_int
    L.pgx_abacus_set_stream.argtypes = [vp, vp]
    L.pgx_abacus_shape.restype = C.c_int
    L.pgx_abacus_shape.argtypes = [vp, u64p, u32p, u32p]
    L.pgx_abacus_upload.restype = C.c_int
    L.pgx_abacus_upload.argtypes = [vp, vp, C.c_uint32, vp]
    L.pgx_abacus_adopt_device.restype = C.c_int
    L.pgx_abacus_adopt_device.argtypes = [vp, vp, vp]
    L.pgx_abacus_scatter.restype = C.c_int
    L.pgx_abacus_scatter.argtypes = [vp, vp, C.c_uint64, C.c_uint32, vp]
    L.pgx_abacus_build.restype = C.c_int
    L.pgx_abacus_build.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp]
    L.pgx_abacus_build_u32.restype = C.c_int
    L.pgx_abacus_build_u32.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp]
    L.pgx_host_alloc.restype = C.c_int
    L.pgx_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.pgx_host_free.restype = None
    L.pgx_host_free.argtypes = [vp]
    L.pgx_abacus_clear.restype = C.c_int
    L.pgx_abacus_clear.argtypes = [vp]
    L.pgx_abacus_copy_rows.restype = C.c_int
    L.pgx_abacus_copy_rows.argtypes = [vp, vp, C.c_uint64]
    L.pgx_abacus_download.restype = C.c_int
    L.pgx_abacus_download.argtypes = [vp, vp, C.c_uint32]
    L.pgx_abacus_csr_rows.restype = C.c_int
    L.pgx_abacus_csr_rows.argtypes = [vp, vp, u64p]
    L.pgx_abacus_csr_fill.restype = C.c_int
    L.pgx_abacus_csr_fill.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, vp, vp, vp, vp]
    L.pgx_hist.restype = C.c_int
    L.pgx_hist.argtypes = [vp, vp, vp, vp]
    L.pgx_ordered_growth.restype = C.c_int
    L.pgx_ordered_growth.argtypes = [vp, C.c_uint32, vp, vp, vp, C.c_int, vp]
    L.pgx_hist_ordered_growth.restype = C.c_int
    L.pgx_hist_ordered_growth.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, C.c_int, vp]
    L.pgx_permuted_growth.restype = C.c_int
    L.pgx_permuted_growth.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, vp, C.c_int, vp]
    L.pgx_similarity.restype = C.c_int
    L.pgx_similarity.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, vp, vp]
    L.pgx_similarity_upper.restype = C.c_int
    L.pgx_similarity_upper.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, vp, vp]
    L.pgx_fused_out_words.restype = C.c_size_t
    L.pgx_fused_out_words.argtypes = [C.c_uint32, C.c_uint32]
    L.pgx_fused_pass_async.restype = C.c_int
    L.pgx_fused_pass_async.argtypes = [vp, C.c_int, C.c_int, C.c_uint32, vp, vp, C.c_int, vp]
    L.pgx_exchange_export.restype = C.c_int
    L.pgx_exchange_export.argtypes = [vp, vp]
    L.pgx_exchange_connect.restype = C.c_int
    L.pgx_exchange_connect.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
    L.pgx_exchange_disconnect.restype = C.c_int
    L.pgx_exchange_disconnect.argtypes = [vp]
    L.pgx_exchange_status.restype = C.c_int
    L.pgx_exchange_status.argtypes = [vp]
    L.pgx_comm_unique_id.restype = C.c_int
    L.pgx_comm_unique_id.argtypes = [vp]
    L.pgx_comm_create.restype = C.c_int
    L.pgx_comm_create.argtypes = [C.POINTER(vp), C.c_int, C.c_uint32, C.c_uint32, vp]
    L.pgx_comm_create_all.restype = C.c_int
    L.pgx_comm_create_all.argtypes = [C.POINTER(vp), C.c_uint32, C.POINTER(C.c_int)]
    L.pgx_comm_destroy.restype = None
    L.pgx_comm_destroy.argtypes = [vp]
    L.pgx_comm_info.restype = C.c_int
    L.pgx_comm_info.argtypes = [vp, u32p, u32p, C.POINTER(C.c_int)]
    L.pgx_abacus_broadcast.restype = C.c_int
    L.pgx_abacus_broadcast.argtypes = [vp, vp, C.c_uint32, C.c_int]
    L.pgx_exchange_connect_comm.restype = C.c_int
    L.pgx_exchange_connect_comm.argtypes = [vp, vp]
    L.pgx_hist_ordered_growth_sharded.restype = C.c_int
    L.pgx_hist_ordered_growth_sharded.argtypes = [vp, vp, vp, vp, C.c_uint32, vp, vp, C.c_int, vp]
    L.pgx_permuted_growth_sharded.restype = C.c_int
    L.pgx_permuted_growth_sharded.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp, C.c_int, vp]
    L.pgx_similarity_sharded.restype = C.c_int
    L.pgx_similarity_sharded.argtypes = [vp, vp, C.c_int, vp, vp]
    L.pgx_similarity_shard_bounds.restype = C.c_int
    L.pgx_similarity_shard_bounds.argtypes = [C.c_uint32, C.c_uint32, u32p]
    L.pgx_abacus_set_timing.restype = C.c_int
    L.pgx_abacus_set_timing.argtypes = [vp, C.c_int]
    L.pgx_kernel_time_ms.restype = C.c_int
    L.pgx_kernel_time_ms.argtypes = [vp, C.POINTER(C.c_float), u32p]
    L.pgx_launch_count.restype = C.c_uint64
    L.pgx_launch_count.argtypes = [vp]
    L.pgx_last_launch_info.restype = C.c_int
    L.pgx_last_launch_info.argtypes = [vp, C.c_char_p, C.c_size_t]
    _lib = L
    return L


class PgxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pgx error {code}: {msg}")
        self.code = code


def check(rc: int) -> None:
    if rc != 0:
        raise PgxError(rc, lib().pgx_last_error().decode(errors="replace"))
