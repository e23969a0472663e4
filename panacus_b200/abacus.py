"""Host-side mirror of the reference interface for the counting hot path, on top of the C ABI.

Names follow the reference (marschall-lab/panacus @ 395ba41):

  * ``Threshold``            src/util.rs:327-364
  * ``DeviceAbacus``         stands in for AbacusByTotal + AbacusByGroup (src/graph_broker/abacus.rs:486-503,
                             790-800): the item x group incidence lives on the GPU as a bitmap
  * ``DeviceAbacus.hist``    Hist::from_abacus -> construct_hist / construct_hist_bps (abacus.rs:746-787)
  * ``DeviceAbacus.calc_growth``  AbacusByGroup::calc_growth (abacus.rs:989-1032), f64 result like the reference
  * ``quorum_thresholds``    the per-column integer cutoff ``ceil((c[k] + 1) * q)`` of abacus.rs:1010

This module only marshals arrays; all counting runs in libpanacus_b200.so.  It never imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _native


@dataclass(frozen=True)
class Threshold:
    """src/util.rs:327-364.  kind: 'relative' (f64 in [0,1]) or 'absolute' (usize)."""
    kind: str
    value: float

    @staticmethod
    def relative(v: float) -> "Threshold":
        return Threshold("relative", float(v))

    @staticmethod
    def absolute(v: int) -> "Threshold":
        return Threshold("absolute", int(v))

    def to_absolute(self, n: int) -> int:
        """src/util.rs:351-356: Absolute(c) -> c, Relative(c) -> ceil(n * c) as usize"""
        if self.kind == "absolute":
            return int(self.value)
        v = math.ceil(float(n) * self.value)
        return max(0, int(v))

    def to_relative(self, n: int) -> float:
        """src/util.rs:358-363"""
        if self.kind == "relative":
            return self.value
        return float(int(self.value)) / float(n)

    def get_string(self) -> str:
        """src/util.rs:344-349 (Rust `{}` formatting of usize / f64)"""
        if self.kind == "absolute":
            return str(int(self.value))
        v = self.value
        if v == math.floor(v) and abs(v) < 1e16:
            return str(int(v))
        return repr(v)


def row_words(n_groups: int) -> int:
    """u64 words per device bitmap row (include/panacus_b200.h: pgx_row_words)."""
    w = (n_groups + 63) // 64
    return 1 if w <= 1 else (w + 1) // 2 * 2


def quorum_thresholds(n_groups: int, q: float) -> np.ndarray:
    """thr[g] = ceil((g as f64 + 1.0) * q) as usize, evaluated in f64 exactly like abacus.rs:1010."""
    q = max(0.0, float(q))
    return np.array([max(0, int(math.ceil((float(g) + 1.0) * q))) for g in range(n_groups)], dtype=np.uint32)


def growth_cutoffs(n_groups: int, t_coverage: Threshold, t_quorum: Threshold):
    """(cov_abs, quorum_thr[G]) handed to the GPU for one (coverage, quorum) pair (abacus.rs:997-998)."""
    c = max(1, t_coverage.to_absolute(n_groups))
    q = max(0.0, t_quorum.to_relative(n_groups))
    return c, quorum_thresholds(n_groups, q)


def pack_bits(bits: np.ndarray) -> np.ndarray:
    """bool/uint8 [(N+1), G] -> u64 [(N+1), row_words(G)], bit g%64 of word g/64 (row 0 = dummy item)."""
    bits = np.asarray(bits, dtype=np.uint8)
    n_rows, G = bits.shape
    Wp = row_words(G)
    padded = np.zeros((n_rows, Wp * 64), dtype=np.uint8)
    padded[:, :G] = bits
    return np.packbits(padded, axis=1, bitorder="little").view(np.uint64).reshape(n_rows, Wp)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _PinnedBlock:
    """Owner of one pgx_host_alloc block; the numpy views made from it keep it alive."""

    def __init__(self, nbytes: int):
        self._L = _native.lib()
        self.ptr = C.c_void_p()
        _native.check(self._L.pgx_host_alloc(C.byref(self.ptr), int(nbytes)))
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                self._L.pgx_host_free(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array in page-locked host memory (pgx_host_alloc): uploads from / downloads into it are single DMA copies."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    block = _PinnedBlock(max(n * dtype.itemsize, 1))
    buf = (C.c_ubyte * block.nbytes).from_address(block.ptr.value)
    buf._pinned_block = block  # ctypes array -> numpy base chain keeps the block alive
    arr = np.frombuffer(buf, dtype=dtype, count=n)
    return arr.reshape(shape)


class Comm:
    """One rank's NCCL communicator behind the C ABI (pgx_comm_*): one per GPU, one process (or thread) per GPU."""

    def __init__(self, handle, rank: int, world: int, device: int):
        self._L = _native.lib()
        self._h, self.rank, self.world, self.device = handle, int(rank), int(world), int(device)

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(_native.COMM_ID_BYTES)
        _native.check(_native.lib().pgx_comm_unique_id(buf))
        return buf.raw

    @staticmethod
    def create(device: int, rank: int, world: int, unique_id: bytes) -> "Comm":
        if len(unique_id) != _native.COMM_ID_BYTES:
            raise ValueError("unique_id must be PGX_COMM_ID_BYTES long")
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, len(unique_id))
        _native.check(_native.lib().pgx_comm_create(C.byref(h), int(device), int(rank), int(world), buf))
        return Comm(h, rank, world, device)

    @staticmethod
    def from_torch_distributed(device: int, group=None) -> "Comm":
        """Rank 0 draws the NCCL id, torch.distributed carries it to the other ranks (any backend)."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        raw = bytearray(Comm.unique_id() if rank == 0 else bytes(_native.COMM_ID_BYTES))
        t = torch.frombuffer(raw, dtype=torch.uint8).clone()
        if dist.get_backend(group) == "nccl":
            t = t.to(torch.device("cuda", device))
        dist.broadcast(t, src=0, group=group)
        return Comm.create(device, rank, world, bytes(t.cpu().numpy().tobytes()))

    @staticmethod
    def create_all(devices: Sequence[int]):
        """Single-process variant (ncclCommInitAll): one Comm per device, to be driven from one thread each."""
        n = len(devices)
        hs = (C.c_void_p * n)()
        devs = (C.c_int * n)(*[int(d) for d in devices])
        _native.check(_native.lib().pgx_comm_create_all(hs, n, devs))
        return [Comm(C.c_void_p(hs[i]), i, n, int(devices[i])) for i in range(n)]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.pgx_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def similarity_shard_bounds(n_groups: int, world: int) -> np.ndarray:
    """Row-block boundaries (2 * world + 1) of pgx_similarity_sharded; rank r owns blocks r and 2 * world - 1 - r."""
    out = np.zeros(2 * world + 1, dtype=np.uint32)
    _native.check(_native.lib().pgx_similarity_shard_bounds(int(n_groups), int(world), out.ctypes.data_as(C.POINTER(C.c_uint32))))
    return out


class DeviceAbacus:
    """GPU-resident item x group incidence bitmap (+ bp weights) and the hot-path queries on it."""

    def __init__(self, n_items: int, n_groups: int, device: int = 0):
        self._L = _native.lib()
        self._h = C.c_void_p()
        _native.check(self._L.pgx_abacus_create(C.byref(self._h), int(device), int(n_items), int(n_groups)))
        self.n_items, self.n_groups, self.device = int(n_items), int(n_groups), int(device)
        self.row_words = row_words(n_groups)
        self._keep = []  # adopted device tensors kept alive

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.pgx_abacus_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- data ------------------------------------------------------------------------------------
    def upload(self, bitmap: Optional[np.ndarray], weight: Optional[np.ndarray] = None):
        """bitmap: u64 [(N+1), >= ceil(G/64)] node-major; weight: u32 [N+1] (bp lengths) or None."""
        hw = 0
        if bitmap is not None:
            bitmap = np.ascontiguousarray(bitmap, dtype=np.uint64)
            if bitmap.ndim != 2 or bitmap.shape[0] != self.n_items + 1:
                raise ValueError("bitmap must have n_items + 1 rows")
            hw = bitmap.shape[1]
        if weight is not None:
            weight = np.ascontiguousarray(weight, dtype=np.uint32)
            if weight.shape != (self.n_items + 1,):
                raise ValueError("weight must have n_items + 1 entries")
        _native.check(self._L.pgx_abacus_upload(self._h, _ptr(bitmap), hw, _ptr(weight)))

    def adopt_device(self, d_bitmap_ptr: int, d_weight_ptr: Optional[int] = None, keepalive=None):
        """Use device buffers in place (bitmap with pgx_row_words(G) words per row)."""
        _native.check(self._L.pgx_abacus_adopt_device(self._h, C.c_void_p(d_bitmap_ptr),
                                                      C.c_void_p(d_weight_ptr) if d_weight_ptr else None))
        self._keep = [keepalive]

    def set_stream(self, cuda_stream_ptr: Optional[int]):
        _native.check(self._L.pgx_abacus_set_stream(self._h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def clear(self):
        _native.check(self._L.pgx_abacus_clear(self._h))

    def scatter(self, items: np.ndarray, group_id: int, exclude: Optional[np.ndarray] = None):
        """OR bit `group_id` into the rows of `items` (one path's ItemTable slice, src/util.rs:80-93)."""
        items = np.ascontiguousarray(items, dtype=np.uint64)
        ex = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.uint8)
        if ex is not None and ex.shape != (self.n_items + 1,):
            raise ValueError("exclude must have n_items + 1 entries")
        _native.check(self._L.pgx_abacus_scatter(self._h, _ptr(items), items.size, int(group_id), _ptr(ex)))

    def build(self, items: np.ndarray, id_prefsum: np.ndarray, path_group: np.ndarray, exclude: Optional[np.ndarray] = None):
        """Whole ItemTable -> bitmap in one call; path_group[p] = group id of path p, or -1 if not counted.
        items: u64 (the reference's ItemIdSize) or u32 (half the PCIe bytes); page-locked arrays (pinned_empty) are
        read by DMA in place."""
        items = np.asarray(items)
        fn = self._L.pgx_abacus_build
        if items.dtype == np.uint32:
            fn = self._L.pgx_abacus_build_u32
            items = np.ascontiguousarray(items)
        else:
            items = np.ascontiguousarray(items, dtype=np.uint64)
        id_prefsum = np.ascontiguousarray(id_prefsum, dtype=np.uint64)
        path_group = np.ascontiguousarray(path_group, dtype=np.int64)
        if path_group.size + 1 != id_prefsum.size:
            raise ValueError("id_prefsum must have one more entry than path_group")
        ex = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.uint8)
        if ex is not None and ex.shape != (self.n_items + 1,):
            raise ValueError("exclude must have n_items + 1 entries")
        _native.check(fn(self._h, _ptr(items), items.size, _ptr(id_prefsum), path_group.size, _ptr(path_group), _ptr(ex)))

    def copy_rows_from(self, src: "DeviceAbacus", src_first_item: int):
        """Items src_first_item .. of `src` (possibly on another GPU) become this handle's items 1 .. n_items."""
        _native.check(self._L.pgx_abacus_copy_rows(self._h, src._h, int(src_first_item)))

    def download(self) -> np.ndarray:
        out = np.zeros((self.n_items + 1, self.row_words), dtype=np.uint64)
        _native.check(self._L.pgx_abacus_download(self._h, _ptr(out), self.row_words))
        return out

    def csr(self, items=None, id_prefsum=None, path_group=None, exclude=None, values: bool = True):
        """AbacusByGroup {r, c, v} (abacus.rs:790-799, 859-986) -> (r u64[N+2], c u64[nnz], v u32[nnz] | None).
        v (occurrence counts) needs the ItemTable the bitmap was built from."""
        r = np.zeros(self.n_items + 2, dtype=np.uint64)
        nnz = C.c_uint64(0)
        _native.check(self._L.pgx_abacus_csr_rows(self._h, _ptr(r), C.byref(nnz)))
        c = np.zeros(nnz.value, dtype=np.uint64)
        v = None
        if values and items is not None:
            items = np.ascontiguousarray(items, dtype=np.uint64)
            id_prefsum = np.ascontiguousarray(id_prefsum, dtype=np.uint64)
            path_group = np.ascontiguousarray(path_group, dtype=np.int64)
            if path_group.size + 1 != id_prefsum.size:
                raise ValueError("id_prefsum must have one more entry than path_group")
            ex = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.uint8)
            v = np.zeros(nnz.value, dtype=np.uint32)
            _native.check(self._L.pgx_abacus_csr_fill(self._h, _ptr(items), items.size, _ptr(id_prefsum), path_group.size,
                                                      _ptr(path_group), _ptr(ex), _ptr(c), _ptr(v)))
        else:
            _native.check(self._L.pgx_abacus_csr_fill(self._h, None, 0, None, 0, None, None, _ptr(c), None))
        return r, c, v

    # -- hot path --------------------------------------------------------------------------------
    def hist(self, count: bool = True, weight: bool = False, countable: bool = False):
        """-> (hist_count u64[G+1] | None, hist_weight u64[G+1] | None, countable u32[N+1] | None)"""
        G1 = self.n_groups + 1
        hc = np.zeros(G1, dtype=np.uint64) if count else None
        hw = np.zeros(G1, dtype=np.uint64) if weight else None
        ct = np.zeros(self.n_items + 1, dtype=np.uint32) if countable else None
        _native.check(self._L.pgx_hist(self._h, _ptr(hc), _ptr(hw), _ptr(ct)))
        return hc, hw, ct

    @staticmethod
    def _cutoffs(cov_abs, quorum_thr, G):
        cov = np.ascontiguousarray(cov_abs, dtype=np.uint32).reshape(-1)
        thr = None
        if quorum_thr is not None:
            thr = np.ascontiguousarray(quorum_thr, dtype=np.uint32).reshape(-1)
            if thr.size != cov.size * G:
                raise ValueError("quorum_thr must have n_thresholds * n_groups entries")
        return cov, thr

    def ordered_growth(self, cov_abs: Sequence[int], quorum_thr=None, col_order=None, weighted: bool = False) -> np.ndarray:
        """-> u64 [T, G] exact integer curves (AbacusByGroup::calc_growth for T threshold pairs)."""
        G = self.n_groups
        cov, thr = self._cutoffs(cov_abs, quorum_thr, G)
        order = None if col_order is None else np.ascontiguousarray(col_order, dtype=np.uint32)
        if order is not None and order.shape != (G,):
            raise ValueError("col_order must have n_groups entries")
        curve = np.zeros((cov.size, G), dtype=np.uint64)
        _native.check(self._L.pgx_ordered_growth(self._h, cov.size, _ptr(cov), _ptr(thr), _ptr(order),
                                                 int(bool(weighted)), _ptr(curve)))
        return curve

    def hist_ordered_growth(self, cov_abs, quorum_thr=None, weighted: bool = False, hist_count=True, hist_weight=False):
        """one pass: -> (hist_count | None, hist_weight | None, curve u64[T, G])"""
        G = self.n_groups
        cov, thr = self._cutoffs(cov_abs, quorum_thr, G)
        hc = np.zeros(G + 1, dtype=np.uint64) if hist_count else None
        hw = np.zeros(G + 1, dtype=np.uint64) if hist_weight else None
        curve = np.zeros((cov.size, G), dtype=np.uint64)
        _native.check(self._L.pgx_hist_ordered_growth(self._h, _ptr(hc), _ptr(hw), cov.size, _ptr(cov), _ptr(thr),
                                                      int(bool(weighted)), _ptr(curve)))
        return hc, hw, curve

    def permuted_growth(self, orders: np.ndarray, cov_abs, quorum_thr=None, weighted: bool = False,
                        out: Optional[np.ndarray] = None) -> np.ndarray:
        """orders: u32 [P, G], each row a permutation of the groups -> u64 [P, T, G] (written into `out` when given, e.g. a
        pinned_empty array: the result is then copied by DMA straight into it)"""
        G = self.n_groups
        orders = np.ascontiguousarray(orders, dtype=np.uint32)
        if orders.ndim != 2 or orders.shape[1] != G:
            raise ValueError("orders must be [P, n_groups]")
        cov, thr = self._cutoffs(cov_abs, quorum_thr, G)
        curves = out if out is not None else np.zeros((orders.shape[0], cov.size, G), dtype=np.uint64)
        if curves.shape != (orders.shape[0], cov.size, G) or curves.dtype != np.uint64 or not curves.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous u64 [P, T, G] array")
        _native.check(self._L.pgx_permuted_growth(self._h, orders.shape[0], _ptr(orders), cov.size, _ptr(cov), _ptr(thr),
                                                  int(bool(weighted)), _ptr(curves)))
        return curves

    def similarity(self, weighted: bool = False, row_begin: int = 0, row_end: Optional[int] = None, upper: bool = False,
                   out_inter: Optional[np.ndarray] = None):
        """-> (inter u64[rows, G], len u64[G]); integer part of Similarity::set_table.  upper: only the columns
        >= row_begin are computed (the others are zero): the matrix is symmetric (pgx_similarity_upper)."""
        G = self.n_groups
        row_end = G if row_end is None else int(row_end)
        inter = out_inter if out_inter is not None else np.zeros((max(row_end - row_begin, 0), G), dtype=np.uint64)
        if inter.shape != (max(row_end - row_begin, 0), G) or inter.dtype != np.uint64 or not inter.flags.c_contiguous:
            raise ValueError("out_inter must be a C-contiguous u64 [rows, G] array")
        ln = np.zeros(G, dtype=np.uint64)
        fn = self._L.pgx_similarity_upper if upper else self._L.pgx_similarity
        _native.check(fn(self._h, int(bool(weighted)), int(row_begin), row_end, _ptr(inter), _ptr(ln)))
        return inter, ln

    # -- reference-shaped convenience --------------------------------------------------------------
    def calc_growth(self, t_coverage: Threshold, t_quorum: Threshold, count: str = "node") -> np.ndarray:
        """AbacusByGroup::calc_growth(t_coverage, t_quorum, node_lens) -> Vec<f64> (abacus.rs:989-1032)."""
        c, thr = growth_cutoffs(self.n_groups, t_coverage, t_quorum)
        curve = self.ordered_growth([c], thr, weighted=(count == "bp"))
        return curve[0].astype(np.float64)

    # -- device-resident / async -------------------------------------------------------------------
    def fused_out_words(self, n_thresholds: int) -> int:
        return int(self._L.pgx_fused_out_words(self.n_groups, int(n_thresholds)))

    def fused_pass_async(self, d_out_ptr: int, cov_abs, quorum_thr=None, weighted=False, hist_count=True,
                         hist_weight=False):
        cov, thr = self._cutoffs(cov_abs, quorum_thr, self.n_groups)
        _native.check(self._L.pgx_fused_pass_async(self._h, int(bool(hist_count)), int(bool(hist_weight)), cov.size,
                                                   _ptr(cov), _ptr(thr), int(bool(weighted)), C.c_void_p(d_out_ptr)))

    # -- sharded over an NCCL communicator (collective: every rank calls with its own handle) ------------------
    def broadcast(self, comm: "Comm", root: int = 0, with_weights: bool = False):
        _native.check(self._L.pgx_abacus_broadcast(self._h, comm._h, int(root), int(bool(with_weights))))

    def hist_ordered_growth_sharded(self, comm: "Comm", cov_abs, quorum_thr=None, weighted=False, hist_count=True,
                                    hist_weight=False):
        """Item-range shards + ncclAllReduce: -> (hist_count | None, hist_weight | None, curve u64[T, G]) of the whole graph."""
        G = self.n_groups
        cov, thr = self._cutoffs(cov_abs, quorum_thr, G)
        hc = np.zeros(G + 1, dtype=np.uint64) if hist_count else None
        hw = np.zeros(G + 1, dtype=np.uint64) if hist_weight else None
        curve = np.zeros((cov.size, G), dtype=np.uint64)
        _native.check(self._L.pgx_hist_ordered_growth_sharded(self._h, comm._h, _ptr(hc), _ptr(hw), cov.size, _ptr(cov),
                                                              _ptr(thr), int(bool(weighted)), _ptr(curve)))
        return hc, hw, curve

    def permuted_growth_sharded(self, comm: "Comm", orders: np.ndarray, cov_abs, quorum_thr=None, weighted=False,
                                out: Optional[np.ndarray] = None) -> np.ndarray:
        """Order p on rank p % world, ncclAllGather of the device-resident curves -> u64 [P, T, G] on every rank."""
        G = self.n_groups
        orders = np.ascontiguousarray(orders, dtype=np.uint32)
        if orders.ndim != 2 or orders.shape[1] != G:
            raise ValueError("orders must be [P, n_groups]")
        cov, thr = self._cutoffs(cov_abs, quorum_thr, G)
        curves = out if out is not None else np.zeros((orders.shape[0], cov.size, G), dtype=np.uint64)
        _native.check(self._L.pgx_permuted_growth_sharded(self._h, comm._h, orders.shape[0], _ptr(orders), cov.size,
                                                          _ptr(cov), _ptr(thr), int(bool(weighted)), _ptr(curves)))
        return curves

    def similarity_sharded(self, comm: "Comm", weighted: bool = False, out_inter: Optional[np.ndarray] = None):
        """Upper-triangle row blocks per rank, ncclAllGather, device-side assembly -> (inter u64[G, G], len u64[G])."""
        G = self.n_groups
        inter = out_inter if out_inter is not None else np.zeros((G, G), dtype=np.uint64)
        ln = np.zeros(G, dtype=np.uint64)
        _native.check(self._L.pgx_similarity_sharded(self._h, comm._h, int(bool(weighted)), _ptr(inter), _ptr(ln)))
        return inter, ln

    def exchange_connect_comm(self, comm: "Comm"):
        _native.check(self._L.pgx_exchange_connect_comm(self._h, comm._h))

    def exchange_status(self):
        _native.check(self._L.pgx_exchange_status(self._h))

    # -- fused multi-GPU exchange (item-range sharding) ----------------------------------------------
    def exchange_export(self) -> bytes:
        buf = C.create_string_buffer(_native.EXCHANGE_HANDLE_BYTES)
        _native.check(self._L.pgx_exchange_export(self._h, buf))
        return buf.raw

    def exchange_connect(self, rank: int, world: int, all_handles: bytes):
        if len(all_handles) != world * _native.EXCHANGE_HANDLE_BYTES:
            raise ValueError("all_handles must hold one handle per rank")
        buf = C.create_string_buffer(all_handles, len(all_handles))
        _native.check(self._L.pgx_exchange_connect(self._h, int(rank), int(world), buf))

    def exchange_disconnect(self):
        _native.check(self._L.pgx_exchange_disconnect(self._h))

    def set_timing(self, enable: bool = True):
        """Bracket the hot-path kernels of every call with CUDA events on the handle's stream (for measurements)."""
        _native.check(self._L.pgx_abacus_set_timing(self._h, int(bool(enable))))

    def kernel_time_ms(self):
        """-> (summed device time in ms, number of kernel sections) recorded since the last query."""
        ms, n = C.c_float(0), C.c_uint32(0)
        _native.check(self._L.pgx_kernel_time_ms(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    @property
    def launch_count(self) -> int:
        return int(self._L.pgx_launch_count(self._h))

    def last_launch_info(self) -> str:
        buf = C.create_string_buffer(512)
        _native.check(self._L.pgx_last_launch_info(self._h, buf, 512))
        return buf.value.decode()


def curve_from_fused(out: np.ndarray, n_groups: int, n_thresholds: int):
    """Split a fused-layout u64 buffer into (hist_count, hist_weight, curves[T, G])."""
    G1 = n_groups + 1
    hc, hw = out[:G1].copy(), out[G1:2 * G1].copy()
    d = out[2 * G1:2 * G1 + n_thresholds * n_groups].reshape(n_thresholds, n_groups)
    return hc, hw, np.cumsum(d, axis=1, dtype=np.uint64)
