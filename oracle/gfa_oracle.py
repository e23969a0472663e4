"""Pure-Python restatement of the reference's GFA front end, for SMALL fixtures only.

TEST INFRASTRUCTURE ONLY (see oracle/panacus_oracle.h).  It turns a GFA file into exactly the
structures the reference hands to its counting loops, so that the C oracle can be driven from the
reference's own fixtures and compared with the reference's golden vectors:

  * GraphStorage::parse_nodes_gfa / parse_edge_gfa        src/graph_broker/graph.rs:276-375
  * PathSegment::from_str / id / clear_coords             src/graph_broker/graph.rs:495-616
  * parse_path_identifier / parse_walk_identifier         src/graph_broker/util.rs:368-410
  * GraphMask::load_groups / get_path_order               src/graph_broker/abacus.rs:242-347
  * parse_bed_to_path_segments / parse_groups             src/io.rs:35-147
  * ItemTable construction for node/bp and edge counts    src/graph_broker/util.rs:208-366, 724-790
  * subset / exclude handling (update_tables, ActiveTable, IntervalContainer,
    quantify_uncovered_bps)                               src/graph_broker/util.rs:569-721,
                                                          src/util.rs:117-310, abacus.rs:1187-1229
"""
from __future__ import annotations

import gzip
import os
import re
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

PATHID_PANSN = re.compile(r"^([^#]+)(#[^#]+)?(#[^#].*)?$")
PATHID_COORDS = re.compile(r"^(.+):([0-9]+)-([0-9]+)$")

USIZE_MAX = (1 << 64) - 1


@dataclass(frozen=True)
class PathSegment:
    sample: str
    haplotype: Optional[str] = None
    seqid: Optional[str] = None
    start: Optional[int] = None
    end: Optional[int] = None

    @staticmethod
    def from_str(s: str) -> "PathSegment":
        """graph.rs:495-549"""
        sample, hap, seqid, start, end = s, None, None, None, None
        m = PATHID_PANSN.match(s)
        if m:
            segs = [g for g in m.groups() if g is not None]
            if len(segs) == 3:
                sample = segs[0]
                hap = segs[1][1:]
                cc = PATHID_COORDS.match(segs[2][1:])
                if cc is None:
                    seqid = segs[2][1:]
                else:
                    seqid, start, end = cc.group(1), int(cc.group(2)), int(cc.group(3))
            elif len(segs) == 2:
                sample = segs[0]
                cc = PATHID_COORDS.match(segs[1][1:])
                if cc is None:
                    hap = segs[1][1:]
                else:
                    hap, start, end = cc.group(1), int(cc.group(2)), int(cc.group(3))
            elif len(segs) == 1:
                cc = PATHID_COORDS.match(segs[0])
                if cc is not None:
                    sample, start, end = cc.group(1), int(cc.group(2)), int(cc.group(3))
        return PathSegment(sample, hap, seqid, start, end)

    @staticmethod
    def from_str_start_end(s: str, start: int, end: int) -> "PathSegment":
        p = PathSegment.from_str(s)
        return PathSegment(p.sample, p.haplotype, p.seqid, start, end)

    def id(self) -> str:
        """graph.rs:558-579"""
        if self.haplotype is not None:
            return f"{self.sample}#{self.haplotype}" + (f"#{self.seqid}" if self.seqid is not None else "")
        if self.seqid is not None:
            return f"{self.sample}#*#{self.seqid}"
        return self.sample

    def clear_coords(self) -> "PathSegment":
        return PathSegment(self.sample, self.haplotype, self.seqid, None, None)

    def coords(self):
        if self.start is not None and self.end is not None:
            return (self.start, self.end)
        return None

    def __str__(self) -> str:
        c = self.coords()
        return f"{self.id()}:{c[0]}-{c[1]}" if c else self.id()


def _open(path: str):
    return gzip.open(path, "rb") if path.endswith(".gz") else open(path, "rb")


def canonical_edge(u: int, o1: str, v: int, o2: str):
    """Edge::canonical, graph.rs:142-148 ('+' forward, '-' backward)."""
    flip = {"+": "-", "-": "+"}
    if u > v or (u == v and o1 == "-"):
        return (v, flip[o2], u, flip[o1])
    return (u, o1, v, o2)


@dataclass
class Graph:
    node2id: dict = field(default_factory=dict)
    node_lens: list = field(default_factory=lambda: [0])
    path_segments: list = field(default_factory=list)
    path_steps: list = field(default_factory=list)  # per P/W line: [(node_id, '+'|'-')]
    edge2id: dict = field(default_factory=dict)

    @property
    def node_count(self) -> int:
        return len(self.node2id)

    @property
    def edge_count(self) -> int:
        return len(self.edge2id)


def parse_gfa(path: str, with_edges: bool = True) -> Graph:
    g = Graph()
    lines = _open(path).read().split(b"\n")
    # pass 1: S / P / W (graph.rs:308-375)
    for buf in lines:
        if not buf:
            continue
        if buf[:1] == b"S":
            cols = buf.split(b"\t")
            name = cols[1]
            if name in g.node2id:
                raise ValueError(f"Segment with ID {name!r} occurs multiple times in GFA")
            g.node2id[name] = len(g.node2id) + 1
            seq = cols[2] if len(cols) > 2 else b""
            seq = seq.split(b"\r")[0]
            g.node_lens.append(len(seq))
        elif buf[:1] == b"P":
            cols = buf.split(b"\t")
            g.path_segments.append(PathSegment.from_str(cols[1].decode()))
        elif buf[:1] == b"W":
            cols = buf.split(b"\t")
            st = None if cols[4] == b"*" else int(cols[4])
            en = None if cols[5] == b"*" else int(cols[5])
            g.path_segments.append(PathSegment(cols[1].decode(), cols[2].decode(), cols[3].decode(), st, en))
    # pass 1b: L (graph.rs:276-306)
    if with_edges:
        for buf in lines:
            if buf[:1] == b"L":
                cols = buf.split(b"\t")
                e = canonical_edge(g.node2id[cols[1]], cols[2].decode(), g.node2id[cols[3]], cols[4].decode())
                if e not in g.edge2id:
                    g.edge2id[e] = len(g.edge2id) + 1
    # pass 2: steps of P / W lines (util.rs:1093-1142, 916-931)
    for buf in lines:
        if buf[:1] == b"P":
            cols = buf.split(b"\t")
            field_ = cols[2].split(b"\r")[0]
            steps = []
            for tok in field_.split(b","):
                if not tok:
                    continue
                steps.append((g.node2id[tok[:-1]], chr(tok[-1])))
            g.path_steps.append(steps)
        elif buf[:1] == b"W":
            cols = buf.split(b"\t")
            field_ = cols[6].split(b"\r")[0]
            steps = []
            for m in re.finditer(rb"([<>])([^<>]+)", field_):
                steps.append((g.node2id[m.group(2)], "+" if m.group(1) == b">" else "-"))
            g.path_steps.append(steps)
    return g


# ---- grouping / ordering (abacus.rs:242-347) -------------------------------------------------

def parse_bed_to_path_segments(path: str, use_block_info: bool = True) -> list:
    """io.rs:35-115"""
    segs = []
    for i, line in enumerate(open(path).read().splitlines()):
        fields = line.split("\t")
        name = fields[0]
        if name.startswith("browser ") or name.startswith("track ") or name.startswith("#"):
            continue
        if len(fields) == 1:
            segs.append(PathSegment.from_str(name))
        elif len(fields) >= 3:
            start, end = int(fields[1]), int(fields[2])
            if use_block_info and len(fields) == 12:
                sizes = [int(s) for s in fields[10].split(",") if s.strip().isdigit()]
                starts = [int(s) for s in fields[11].split(",") if s.strip().isdigit()]
                for size, off in zip(sizes, starts):
                    segs.append(PathSegment.from_str_start_end(name, start + off, start + off + size))
            else:
                segs.append(PathSegment.from_str_start_end(name, start, end))
        else:
            raise ValueError(f"error in line {i + 1}: row must have either 1, 3, or 12 columns, but has 2")
    return segs


def load_groups(g: Graph, groupby_file: str = "", groupby_haplotype=False, groupby_sample=False) -> dict:
    """GraphMask::load_groups, abacus.rs:242-308 -> {PathSegment(no coords): group name}"""
    if groupby_haplotype:
        return {p.clear_coords(): f"{p.sample}#{p.haplotype or ''}" for p in g.path_segments}
    if groupby_sample:
        return {p.clear_coords(): p.sample for p in g.path_segments}
    if groupby_file:
        res = {}
        for i, line in enumerate(open(groupby_file).read().split("\n")):
            if line == "" :
                continue
            line = line.rstrip("\r")
            cols = line.split("\t")
            if len(cols) != 2:
                raise ValueError(f"error in line {i + 1}: table must have exactly two columns")
            p = PathSegment.from_str(cols[0]).clear_coords()
            if p in res and res[p] != cols[1]:
                raise ValueError(f"path {p} cannot be assigned to more than one group")
            res.setdefault(p, cols[1])
        for p in g.path_segments:
            res.setdefault(p.clear_coords(), p.id())
        return res
    return {p.clear_coords(): p.id() for p in g.path_segments}


def complement_with_group_assignments(coords, groups: dict):
    """abacus.rs:152-201"""
    if coords is None:
        return None
    group2paths: dict = {}
    for p, grp in groups.items():
        group2paths.setdefault(grp, []).append(p)
    out = []
    for p in coords:
        if p.clear_coords() in groups:
            out.append(p)
        elif p.id() in group2paths:
            if p.coords() is not None:
                raise ValueError(f'invalid coordinate "{p}": group identifiers are not allowed to have start/stop information!')
            out.extend(group2paths[p.id()])
        # unknown path/group: silently dropped
    return out


def load_coord_list(text: str, paths: list):
    """abacus.rs:212-240"""
    if not text:
        return None
    if os.path.isfile(text):
        return parse_bed_to_path_segments(text, True)
    rx = re.compile(text)
    return [p for p in paths if rx.search(str(p))]


@dataclass
class Mask:
    groups: dict
    include: Optional[list]
    exclude: Optional[list]
    order: Optional[list]


def make_mask(g: Graph, groupby_file="", groupby_haplotype=False, groupby_sample=False,
              subset="", exclude="", order: Optional[str] = None) -> Mask:
    """GraphMask::from_datamgr, abacus.rs:55-150 (order-file consistency checks omitted)."""
    groups = load_groups(g, groupby_file, groupby_haplotype, groupby_sample)
    inc = complement_with_group_assignments(load_coord_list(subset, g.path_segments), groups)
    exc = complement_with_group_assignments(load_coord_list(exclude, g.path_segments), groups)
    od = None
    if order:
        od = complement_with_group_assignments(parse_bed_to_path_segments(order, True), groups)
    return Mask(groups, inc, exc, od)


def get_path_order(mask: Mask, path_segments: list):
    """GraphMask::get_path_order, abacus.rs:310-347 -> [(path index, group name)]"""
    group_to_paths: dict = {}
    for i, p in enumerate(path_segments):
        grp = mask.groups[p.clear_coords()]
        group_to_paths.setdefault(grp, []).append((i, grp))
    if mask.order is not None:
        order = list(mask.order)
    elif mask.include is not None:
        order = list(mask.include)
    else:
        ex = set(mask.exclude or [])
        order = [p for p in path_segments if p not in ex]
    out = []
    for p in order:
        out.extend(group_to_paths.pop(mask.groups[p.clear_coords()], []))
    return out


def path_order_arrays(mask: Mask, g: Graph):
    """-> (order_path u64[], order_group u64[], group names) as item_table_to_abacus uses them
    (abacus.rs:555-569)."""
    names: list = []
    op, og = [], []
    for path_id, grp in get_path_order(mask, g.path_segments):
        if not names or names[-1] != grp:
            names.append(grp)
        op.append(path_id)
        og.append(len(names) - 1)
    return np.array(op, dtype=np.uint64), np.array(og, dtype=np.uint64), names


# ---- subset / exclude machinery (src/util.rs:117-325, graph_broker/util.rs:569-790) ------------

def intersects(v, el) -> bool:
    """src/util.rs:370-383: inclusive overlap test (touching intervals intersect)."""
    return any(s <= el[1] and e >= el[0] for s, e in v)


def is_contained(v, el) -> bool:
    """src/util.rs:385-398"""
    return any(s <= el[0] and e >= el[1] for s, e in v)


class IntervalContainer:
    """src/util.rs:199-310: per-item sorted, merged interval lists."""

    def __init__(self):
        self.map: dict = {}

    def add(self, sid: int, start: int, end: int):
        v = self.map.setdefault(sid, [])
        v.append((start, end))
        v.sort()
        i = 1
        while i < len(v):
            if v[i - 1][1] >= v[i][0]:
                x = v.pop(i)
                v[i - 1] = (v[i - 1][0], max(v[i - 1][1], x[1]))
            else:
                i += 1

    def get(self, sid):
        return self.map.get(sid)

    def contains(self, sid) -> bool:
        return sid in self.map

    def remove(self, sid):
        self.map.pop(sid, None)

    def keys(self):
        return list(self.map.keys())

    def total_coverage(self, sid, exclude) -> int:
        """src/util.rs:257-298, literal (including its -1/+1 interval arithmetic, wrapping usize)."""
        iv = self.map.get(sid)
        if iv is None:
            return 0
        if exclude is None:
            return sum(b - a for a, b in iv)
        ex = exclude
        res = 0
        i = 0
        for (start, end) in iv:
            while i < len(ex) and ex[i][1] <= start:
                i += 1
            if i < len(ex) and ex[i][0] < end:
                res += min((ex[i][0] - 1) & USIZE_MAX, end) - start
                if ex[i][1] < end:
                    res += end - ex[i][1] + 1
            else:
                res += end - start
        return res & USIZE_MAX


class ActiveTable:
    """src/util.rs:117-197"""

    def __init__(self, size: int, with_annotation: bool):
        self.items = [False] * size
        self.annotation = IntervalContainer() if with_annotation else None

    def activate(self, sid: int):
        self.items[sid] = True

    def activate_n_annotate(self, sid: int, item_len: int, start: int, end: int):
        m = self.annotation
        if end - start == item_len:
            self.items[sid] = True
            m.remove(sid)
        else:
            if start <= end:
                m.add(sid, start, end)
            if m.get(sid)[0] == (0, item_len):
                m.remove(sid)
                self.items[sid] = True

    def get_active_intervals(self, sid: int, item_len: int):
        if self.items[sid]:
            return [(0, item_len)]
        if self.annotation is not None:
            return list(self.annotation.get(sid) or [])
        return []


def build_subpath_map(segments) -> dict:
    """abacus.rs:354-383"""
    res: dict = {}
    for x in segments:
        res.setdefault(x.id(), set()).add(x.coords() if x.coords() else (0, USIZE_MAX))
    out = {}
    for pid, cs in res.items():
        v = sorted(cs)
        i = 1
        while i < len(v):
            if v[i - 1][1] >= v[i][0]:
                x = v.pop(i)
                v[i - 1] = (v[i - 1][0], max(v[i - 1][1], x[1]))
            else:
                i += 1
        out[pid] = v
    return out


@dataclass
class Tables:
    items: np.ndarray
    id_prefsum: np.ndarray
    exclude: Optional[np.ndarray]      # u8[N+1] or None
    uncovered: dict                    # item id -> uncovered bps (quantify_uncovered_bps)
    n_items: int


def _update_tables(g: Graph, steps, include_coords, exclude_coords, offset, items_out,
                   subset_covered, exclude_table):
    """update_tables, graph_broker/util.rs:569-721"""
    i = j = 0
    p = offset
    if not steps:
        return
    for sid, o in steps:
        l = g.node_lens[sid]
        stop_here = False
        while i < len(include_coords) and include_coords[i][0] < p + l and not stop_here:
            if include_coords[i][1] > p:
                a = include_coords[i][0] - p if include_coords[i][0] > p else 0
                if include_coords[i][1] < p + l:
                    i += 1
                    b = include_coords[i - 1][1] - p
                else:
                    stop_here = True
                    b = l
                if o == "-":
                    a, b = l - b, l - a
                items_out.append(sid)
                if subset_covered is not None:
                    if b - a == l:
                        if subset_covered.contains(sid):
                            subset_covered.remove(sid)
                    else:
                        subset_covered.add(sid, a, b)
            else:
                i += 1
        stop_here = False
        while j < len(exclude_coords) and exclude_coords[j][0] < p + l and not stop_here:
            if exclude_coords[j][1] > p:
                a = exclude_coords[j][0] - p if exclude_coords[j][0] > p else 0
                if exclude_coords[j][1] < p + l:
                    j += 1
                    b = exclude_coords[j - 1][1] - p
                else:
                    stop_here = True
                    b = l
                if o == "-":
                    a, b = l - b, l - a
                if exclude_table is not None:
                    if exclude_table.annotation is not None:
                        exclude_table.activate_n_annotate(sid, l, a, b)
                    else:
                        exclude_table.activate(sid)
            else:
                j += 1
        if i >= len(include_coords) and j >= len(exclude_coords):
            break
        p += l


def _update_tables_edgecount(g: Graph, steps, include_coords, exclude_coords, offset, items_out,
                             exclude_table):
    """update_tables_edgecount, graph_broker/util.rs:723-790"""
    i = j = 0
    p = offset
    if steps:
        p += g.node_lens[steps[0][0]]
    for (s1, o1), (s2, o2) in zip(steps, steps[1:]):
        while i < len(include_coords) and include_coords[i][1] <= p:
            i += 1
        while j < len(exclude_coords) and exclude_coords[j][1] <= p:
            j += 1
        l = g.node_lens[s2]
        eid = g.edge2id[canonical_edge(s1, o1, s2, o2)]
        if i < len(include_coords) and include_coords[i][0] < p + l:
            items_out.append(eid)
        if exclude_table is not None and j < len(exclude_coords) and exclude_coords[j][0] < p + l:
            exclude_table.activate(eid)
        elif i >= len(include_coords) and j >= len(exclude_coords):
            break
        p += l


def item_tables(g: Graph, mask: Mask, count: str) -> Tables:
    """parse_gfa_paths_walks, graph_broker/util.rs:208-366 (count in node|bp|edge)."""
    n_items = g.edge_count if count == "edge" else g.node_count
    subset_covered = IntervalContainer() if (count == "bp" and mask.include is not None) else None
    exclude_table = ActiveTable(n_items + 1, count == "bp") if mask.exclude is not None else None
    include_map = build_subpath_map(mask.include) if mask.include is not None else {}
    exclude_map = build_subpath_map(mask.exclude) if mask.exclude is not None else {}
    complete = [(0, USIZE_MAX)]
    items: list = []
    prefsum = [0]
    for seg, steps in zip(g.path_segments, g.path_steps):
        inc = complete if mask.include is None else include_map.get(seg.id(), [])
        exc = [] if mask.exclude is None else exclude_map.get(seg.id(), [])
        start, end = seg.coords() or (0, USIZE_MAX)
        if mask.include is not None and not intersects(inc, (start, end)) and not intersects(exc, (start, end)):
            prefsum.append(len(items))
            continue
        if (count != "edge" and (mask.include is None or is_contained(inc, (start, end)))
                and (mask.exclude is None or is_contained(exc, (start, end)))):
            # parse_path_seq_update_tables, util.rs:1186-1248: every step is added; if exclude
            # coords apply to this path, all of its items are flagged as excluded
            first = len(items)
            items.extend(s for s, _ in steps)
            if exc and exclude_table is not None:
                for s in items[first:]:
                    exclude_table.items[s] = True
        elif count == "edge":
            _update_tables_edgecount(g, steps, inc, exc, start, items, exclude_table)
        else:
            _update_tables(g, steps, inc, exc, start, items, subset_covered, exclude_table)
        prefsum.append(len(items))
    # quantify_uncovered_bps, abacus.rs:1187-1229
    uncovered = {}
    if subset_covered is not None:
        for sid in subset_covered.keys():
            if exclude_table is None or not exclude_table.items[sid]:
                l = g.node_lens[sid]
                ex = exclude_table.get_active_intervals(sid, l) if exclude_table is not None else None
                covered = subset_covered.total_coverage(sid, ex)
                if covered <= l:
                    uncovered[sid] = l - covered
    ex_arr = None if exclude_table is None else np.array(exclude_table.items, dtype=np.uint8)
    return Tables(np.array(items, dtype=np.uint64), np.array(prefsum, dtype=np.uint64), ex_arr,
                  uncovered, n_items)


# ---- AbacusByGroup::to_tsv (abacus.rs:1056-1178): the `table` analysis ------------------------------------------

def abacus_by_group_to_tsv(g: Graph, count: str, total: bool, groups, r, c, v, uncovered=None) -> str:
    """Statement-by-statement restatement, including the edge branch indexing the occurrence counts with the
    group index (`v[j as usize]`, abacus.rs:1166; IndexError where Rust would panic).  v = None <=> self.v is None."""
    uncovered = uncovered or {}
    id2node = [b""] * (g.node_count + 1)
    for node, nid in g.node2id.items():
        id2node[nid] = node
    out = []
    n_groups = len(groups)
    if count in ("node", "bp"):
        out.append("node" + ("\ttotal" if total else "".join("\t" + grp for grp in groups)) + "\n")
        for i in range(1, len(r) - 1):  # tuple_windows().enumerate(), first entry ignored
            start, end = int(r[i]), int(r[i + 1])
            bp = int(g.node_lens[i]) - int(uncovered.get(i, 0)) if count == "bp" else 1
            row = id2node[i].decode()
            if total:
                row += "\t%d" % (end - start)
            else:
                k = start
                for j in range(n_groups):
                    if k == end or j < int(c[k]):
                        row += "\t0"
                    elif j == int(c[k]):
                        row += "\t%d" % (bp if v is None else int(v[k]) * bp)
                        k += 1
            out.append(row + "\n")
    elif count == "edge":
        if not g.edge2id:
            return ""
        sym = {"+": ">", "-": "<"}
        id2edge = [None] * (g.edge_count + 1)
        for edge, eid in g.edge2id.items():
            id2edge[eid] = edge
        out.append("edge" + ("\ttotal" if total else "".join("\t" + grp for grp in groups)) + "\n")
        for i in range(1, len(r) - 1):
            start, end = int(r[i]), int(r[i + 1])
            u, o1, w, o2 = id2edge[i]
            row = sym[o1] + id2node[u].decode() + sym[o2] + id2node[w].decode()
            if total:
                row += "\t%d" % (end - start)
            else:
                k = start
                for j in range(n_groups):
                    if k == end or j < int(c[k]):
                        row += "\t0"
                    elif j == int(c[k]):
                        row += "\t1" if v is None else "\t%d" % int(v[j])  # sic: v[j], not v[k]
                        k += 1
            out.append(row + "\n")
    else:
        raise ValueError("inadmissible count type")
    return "".join(out)
