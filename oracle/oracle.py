"""ctypes front-end of the CPU parity oracle (oracle/panacus_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py.  Nothing under panacus_b200/ imports this module.

The C library restates the reference statement by statement (citations in panacus_oracle.h);
this file only marshals numpy arrays and adds
  * ``bitmap_to_item_table``: turns a node-major incidence bitmap into the reference's ItemTable
    (src/util.rs:80-93) with one path per group, so the same synthetic input can be fed to the
    reference algorithm and to the GPU path;
  * ``ordered_growth_bitmap_rule``: an independent numpy formulation of
    AbacusByGroup::calc_growth (src/graph_broker/abacus.rs:989-1032) used to cross-check the
    literal CSR walk (the reference has no golden vector for ordered growth);
  * the TSV writers of src/io.rs:460-604 / src/analyses/{hist,growth}.rs for whole-file diffs.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from typing import Iterable, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpanacus_oracle.so")
_lib = None


class _Threshold(C.Structure):
    _fields_ = [("kind", C.c_int), ("rel", C.c_double), ("abs_", C.c_uint64)]


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc); returns the .so path."""
    src = os.path.join(_HERE, "panacus_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    u64p, u32p, u8p, f64p, f32p = (C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_uint8), C.POINTER(C.c_double),
                                   C.POINTER(C.c_float))
    L.po_threshold_to_absolute.restype = C.c_uint64
    L.po_threshold_to_absolute.argtypes = [_Threshold, C.c_uint64]
    L.po_threshold_to_relative.restype = C.c_double
    L.po_threshold_to_relative.argtypes = [_Threshold, C.c_uint64]
    L.po_abacus_by_total.restype = None
    L.po_abacus_by_total.argtypes = [C.c_uint64, u64p, u64p, u64p, u64p, C.c_uint64, u8p, u32p]
    L.po_construct_hist.restype = None
    L.po_construct_hist.argtypes = [u32p, C.c_uint64, C.c_uint64, u64p]
    L.po_construct_hist_bps.restype = None
    L.po_construct_hist_bps.argtypes = [u32p, u32p, C.c_uint64, C.c_uint64, u64p, u64p, C.c_uint64, u64p]
    L.po_csr_build.restype = C.c_int
    L.po_csr_build.argtypes = [C.c_uint64, u64p, u64p, u64p, u64p, C.c_uint64, u8p, u64p,
                               C.POINTER(u64p), C.POINTER(u32p)]
    L.po_free.restype = None
    L.po_free.argtypes = [C.c_void_p]
    L.po_calc_growth.restype = None
    L.po_calc_growth.argtypes = [u64p, u64p, C.c_uint64, C.c_uint64, _Threshold, _Threshold, C.c_int,
                                 u32p, u64p, f64p]
    L.po_choose.restype = C.c_double
    L.po_choose.argtypes = [C.c_uint64, C.c_uint64]
    for name in ("po_growth_union", "po_growth_core"):
        getattr(L, name).restype = None
        getattr(L, name).argtypes = [u64p, C.c_uint64, _Threshold, f64p]
    L.po_growth_quorum.restype = None
    L.po_growth_quorum.argtypes = [u64p, C.c_uint64, _Threshold, _Threshold, f64p]
    L.po_hist_calc_growth.restype = C.c_uint64
    L.po_hist_calc_growth.argtypes = [u64p, C.c_uint64, _Threshold, _Threshold, f64p]
    L.po_similarity.restype = C.c_int
    L.po_similarity.argtypes = [u64p, u64p, C.c_uint64, C.c_uint64, C.c_int, u32p, u64p, u64p, f32p]
    _lib = L
    return L


# ---- marshalling helpers -------------------------------------------------------------------

def _p(a: np.ndarray | None, ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def absolute(v: int) -> _Threshold:
    """Threshold::Absolute(v)"""
    return _Threshold(1, 0.0, int(v))


def relative(v: float) -> _Threshold:
    """Threshold::Relative(v)"""
    return _Threshold(0, float(v), 0)


def _thr(t) -> _Threshold:
    if isinstance(t, _Threshold):
        return t
    kind, v = t
    return absolute(v) if kind in ("abs", "A", 1) else relative(v)


# ---- reference algorithm wrappers ----------------------------------------------------------

def abacus_by_total(n_items, items, id_prefsum, order_path, order_group, exclude=None) -> np.ndarray:
    items, id_prefsum = _u64(items), _u64(id_prefsum)
    order_path, order_group = _u64(order_path), _u64(order_group)
    ex = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.uint8)
    countable = np.zeros(n_items + 1, dtype=np.uint32)
    lib().po_abacus_by_total(n_items, _p(items, C.c_uint64), _p(id_prefsum, C.c_uint64),
                             _p(order_path, C.c_uint64), _p(order_group, C.c_uint64),
                             len(order_path), _p(ex, C.c_uint8), _p(countable, C.c_uint32))
    return countable


def construct_hist(countable, n_groups) -> np.ndarray:
    countable = _u32(countable)
    hist = np.zeros(n_groups + 1, dtype=np.uint64)
    lib().po_construct_hist(_p(countable, C.c_uint32), len(countable) - 1, n_groups, _p(hist, C.c_uint64))
    return hist


def construct_hist_bps(countable, node_lens, n_groups, uncovered: dict | None = None) -> np.ndarray:
    countable, node_lens = _u32(countable), _u32(node_lens)
    uncovered = uncovered or {}
    ids = _u64(list(uncovered.keys()))
    vals = _u64(list(uncovered.values()))
    hist = np.zeros(n_groups + 1, dtype=np.uint64)
    lib().po_construct_hist_bps(_p(countable, C.c_uint32), _p(node_lens, C.c_uint32),
                                len(countable) - 1, n_groups, _p(ids, C.c_uint64),
                                _p(vals, C.c_uint64), len(ids), _p(hist, C.c_uint64))
    return hist


def csr_build(n_items, items, id_prefsum, order_path, order_group, exclude=None):
    """-> (r[N+2], c[nnz], v[nnz])  (AbacusByGroup r/c/v, abacus.rs:790-799)"""
    items, id_prefsum = _u64(items), _u64(id_prefsum)
    order_path, order_group = _u64(order_path), _u64(order_group)
    ex = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.uint8)
    r = np.zeros(n_items + 2, dtype=np.uint64)
    cp, vp = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)()
    rc = lib().po_csr_build(n_items, _p(items, C.c_uint64), _p(id_prefsum, C.c_uint64),
                            _p(order_path, C.c_uint64), _p(order_group, C.c_uint64), len(order_path),
                            _p(ex, C.c_uint8), _p(r, C.c_uint64), C.byref(cp), C.byref(vp))
    if rc != 0:
        raise MemoryError("po_csr_build")
    nnz = int(r[-1])
    c = np.ctypeslib.as_array(cp, shape=(max(nnz, 1),))[:nnz].copy()
    v = np.ctypeslib.as_array(vp, shape=(max(nnz, 1),))[:nnz].copy()
    lib().po_free(cp)
    lib().po_free(vp)
    return r, c, v


def calc_growth(r, c, n_groups, t_coverage, t_quorum, count_bp=False, node_lens=None,
                uncovered=None) -> np.ndarray:
    """AbacusByGroup::calc_growth -> f64[G]"""
    r, c = _u64(r), _u64(c)
    n_items = len(r) - 2
    nl = None if node_lens is None else _u32(node_lens)
    unc = None if uncovered is None else _u64(uncovered)
    res = np.zeros(n_groups, dtype=np.float64)
    lib().po_calc_growth(_p(r, C.c_uint64), _p(c, C.c_uint64), n_items, n_groups, _thr(t_coverage),
                         _thr(t_quorum), int(bool(count_bp)), _p(nl, C.c_uint32),
                         _p(unc, C.c_uint64), _p(res, C.c_double))
    return res


def choose(n, k) -> float:
    return lib().po_choose(n, k)


def growth_union(hist, t_cov) -> np.ndarray:
    hist = _u64(hist)
    n = len(hist) - 1
    out = np.zeros(n, dtype=np.float64)
    lib().po_growth_union(_p(hist, C.c_uint64), n, _thr(t_cov), _p(out, C.c_double))
    return out


def growth_core(hist, t_cov) -> np.ndarray:
    hist = _u64(hist)
    n = len(hist) - 1
    out = np.zeros(n, dtype=np.float64)
    lib().po_growth_core(_p(hist, C.c_uint64), n, _thr(t_cov), _p(out, C.c_double))
    return out


def growth_quorum(hist, t_cov, t_quorum) -> np.ndarray:
    hist = _u64(hist)
    n = len(hist) - 1
    out = np.zeros(n, dtype=np.float64)
    lib().po_growth_quorum(_p(hist, C.c_uint64), n, _thr(t_cov), _thr(t_quorum), _p(out, C.c_double))
    return out


def hist_calc_growth(hist, t_cov, t_quorum) -> np.ndarray:
    """Hist::calc_growth (dispatch union / core / quorum), hist.rs:51-66."""
    hist = _u64(hist)
    n = len(hist) - 1
    out = np.zeros(max(n, 1), dtype=np.float64)
    k = lib().po_hist_calc_growth(_p(hist, C.c_uint64), n, _thr(t_cov), _thr(t_quorum), _p(out, C.c_double))
    return out[:k]


def similarity(r, c, n_groups, count_bp=False, node_lens=None, with_table=True):
    """-> (inter[G,G] u64, len[G] u64, table[G,G] f32 | None)"""
    r, c = _u64(r), _u64(c)
    n_items = len(r) - 2
    nl = None if node_lens is None else _u32(node_lens)
    inter = np.zeros((n_groups, n_groups), dtype=np.uint64)
    ln = np.zeros(n_groups, dtype=np.uint64)
    table = np.zeros((n_groups, n_groups), dtype=np.float32) if with_table else None
    rc = lib().po_similarity(_p(r, C.c_uint64), _p(c, C.c_uint64), n_items, n_groups,
                             int(bool(count_bp)), _p(nl, C.c_uint32), _p(inter, C.c_uint64),
                             _p(ln, C.c_uint64), _p(table, C.c_float))
    if rc != 0:
        table = None  # reference panics: a group without any item
    return inter, ln, table


# ---- threshold parsing (src/graph_broker/hist.rs:207-323) -----------------------------------

def parse_thresholds(quorum: str, coverage: str):
    """ThresholdContainer::parse_params -> (coverage thresholds, quorum thresholds)"""
    def parse(s: str, absolute_required: bool):
        out = []
        for i, el in enumerate(s.split(",")):
            el = el.strip()
            if absolute_required:
                if not (el.isdigit()):
                    raise ValueError(f'threshold "{s}" ({i + 1}. element in list) is required to be integer, but isn\'t.')
                out.append(absolute(int(el)))
            else:
                try:
                    t = float(el)
                except ValueError:
                    raise ValueError(f'threshold "{s}" ({i + 1}. element in list) is required to be float, but isn\'t.')
                if not (0.0 <= t <= 1.0):
                    raise ValueError(f'relative threshold "{s}" ({i + 1}. element in list) must be within [0,1].')
                out.append(relative(t))
        return out

    if not quorum:
        raise ValueError("quorum threshold setting requires at least one element, but none is given")
    q = parse(quorum, False)
    if not coverage:
        raise ValueError("coverage threshold setting requires at least one element, but none is given")
    c = parse(coverage, True)
    if len(q) != len(c):
        if len(q) == 1:
            q = [q[0]] * len(c)
        elif len(c) == 1:
            c = [c[0]] * len(q)
        else:
            raise ValueError("number of coverage and quorum threshold must match, or either one must have a single value")
    return c, q


def rust_f64_display(x: float) -> str:
    """Rust `{}` for f64: shortest round-trip digits, never scientific; integers without '.0'
    under `{:0}`... (format!("{:0}", 5.0f64) prints "5"; NaN prints "NaN")."""
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    if x == math.floor(x) and abs(x) < 1e16:
        return str(int(x))
    r = repr(x)
    if "e" in r or "E" in r:
        from decimal import Decimal
        r = format(Decimal(r), "f")
    return r


def threshold_string(t: _Threshold) -> str:
    """Threshold::get_string (src/util.rs:344-349)"""
    return str(int(t.abs_)) if t.kind == 1 else rust_f64_display(t.rel)


# ---- TSV writers (src/io.rs:460-518, 557-604; src/analyses/hist.rs:24-54; growth.rs:33-102) ---

def write_table(headers: Sequence[Sequence[str]], columns: Sequence[Sequence[float]], start_index: int = 0) -> str:
    res = []
    n = len(headers[0]) if headers else 0
    for i in range(n):
        res.append("\t".join(h[i] for h in headers) + "\n")
    n = len(columns[0]) if columns else 0
    for i in range(n):
        row = str(i + start_index)
        for col in columns:
            v = col[i]
            row += "\t" + rust_f64_display(v if math.isnan(v) else math.floor(v))
        res.append(row + "\n")
    return "".join(res)


def write_ordered_table(headers, columns, index: Sequence[str]) -> str:
    res = []
    n = len(headers[0]) if headers else 0
    for i in range(n):
        res.append("\t".join(h[i] for h in headers) + "\n")
    n = len(columns[0]) if columns else 0
    for i in range(1, n):
        row = index[i - 1]
        for col in columns:
            v = col[i]
            row += "\t" + rust_f64_display(v if math.isnan(v) else math.floor(v))
        res.append(row + "\n")
    return "".join(res)


def hist_table(hists: Sequence[tuple[str, Iterable[int]]]) -> str:
    """Body of analyses::hist::generate_table after the comment lines."""
    headers = [["panacus", "count", "", ""]]
    cols = []
    for count, h in hists:
        cols.append([float(x) for x in h])
        headers.append(["hist", count, "", ""])
    return write_table(headers, cols)


def growth_table(hists: Sequence[tuple[str, Iterable[int]]], cov, quo, add_hist=False) -> str:
    """Body of analyses::growth::generate_table after the comment lines."""
    headers = [["panacus", "count", "coverage", "quorum"]]
    cols = []
    if add_hist:
        for count, h in hists:
            cols.append([float(x) for x in h])
            headers.append(["hist", count, "", ""])
    for count, h in hists:
        for c, q in zip(cov, quo):
            g = hist_calc_growth(h, c, q)
            cols.append([math.nan] + list(g))
            headers.append(["growth", count, threshold_string(c), threshold_string(q)])
    return write_table(headers, cols)


def ordered_growth_table(count: str, groups: Sequence[str], curves: Sequence[Sequence[float]], cov, quo) -> str:
    """Body of io::write_ordered_histgrowth_table after the comment lines."""
    headers = [["panacus", "count", "coverage", "quorum"]]
    cols = []
    for curve, c, q in zip(curves, cov, quo):
        cols.append([math.nan] + list(curve))
        headers.append(["ordered-growth", count, threshold_string(c), threshold_string(q)])
    return write_ordered_table(headers, cols, list(groups))


def coverage_line_table(hists: Sequence[tuple[str, Iterable[int]]]) -> str:
    """Body of analyses::coverage_line::generate_table after the comment lines (coverage_line.rs:23-57):
    every hist without its row 0, index starting at 1."""
    headers = [["panacus", "count", "", ""]]
    cols = []
    for count, h in hists:
        cols.append([float(x) for x in list(h)[1:]])
        headers.append(["hist", count, "", ""])
    return write_table(headers, cols, 1)


# ---- synthetic-input helpers ---------------------------------------------------------------

def bitmap_to_item_table(bitmap: np.ndarray, n_groups: int):
    """node-major bitmap [(N+1), W] u64 (row 0 = dummy) -> ItemTable with one path per group.

    Returns (items u64[S], id_prefsum u64[G+1], order_path, order_group).  Path g visits, in
    ascending id order, every item whose bit g is set; this is the ItemTable the reference's GFA
    parser would produce for a GFA with P lines ``g: id+,id+,...`` (src/util.rs:80-93).
    """
    bitmap = np.ascontiguousarray(bitmap, dtype=np.uint64)
    n_rows, W = bitmap.shape
    bits = np.unpackbits(bitmap.view(np.uint8).reshape(n_rows, W * 8), axis=1, bitorder="little")[:, :n_groups]
    items = []
    prefsum = np.zeros(n_groups + 1, dtype=np.uint64)
    for g in range(n_groups):
        ids = np.nonzero(bits[:, g])[0].astype(np.uint64)
        items.append(ids)
        prefsum[g + 1] = prefsum[g] + np.uint64(len(ids))
    items = np.concatenate(items) if items else np.zeros(0, dtype=np.uint64)
    order = np.arange(n_groups, dtype=np.uint64)
    return items, prefsum, order, order.copy()


def quorum_thresholds(n_groups: int, q: float) -> np.ndarray:
    """thr[g] = ceil((g + 1) * q) as usize, in f64 exactly as abacus.rs:1010."""
    return np.array([int(math.ceil((float(g) + 1.0) * q)) for g in range(n_groups)], dtype=np.uint32)


def ordered_growth_bitmap_rule(bitmap: np.ndarray, n_groups: int, cov_abs: int, q: float,
                               weights: np.ndarray | None = None) -> np.ndarray:
    """Independent formulation of calc_growth on the bitmap (integer result, u64[G]).

    Item i counts at column j iff popc(row) >= C, a set bit <= j exists, and with
    last = highest set bit <= j: popc(row & mask(<= j)) >= ceil((last + 1) * q).
    """
    bitmap = np.ascontiguousarray(bitmap, dtype=np.uint64)
    n_rows, W = bitmap.shape
    bits = np.unpackbits(bitmap.view(np.uint8).reshape(n_rows, W * 8), axis=1, bitorder="little")[:, :n_groups].astype(np.int64)
    bits[0, :] = 0
    C_ = max(1, int(cov_abs))
    q = max(0.0, float(q))
    thr = quorum_thresholds(n_groups, q).astype(np.int64)
    total = bits.sum(axis=1)
    cnt = np.cumsum(bits, axis=1)
    col = np.arange(n_groups, dtype=np.int64)[None, :]
    last = np.maximum.accumulate(np.where(bits > 0, col, -1), axis=1)
    has = last >= 0
    need = thr[np.clip(last, 0, n_groups - 1)]
    ok = has & (cnt >= need) & (total[:, None] >= C_)
    w = np.ones(n_rows, dtype=np.uint64) if weights is None else np.asarray(weights, dtype=np.uint64)
    return (ok.astype(np.uint64) * w[:, None]).sum(axis=0, dtype=np.uint64)
