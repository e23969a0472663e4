"""CPU-only: the C++ host front end (GFA parsing, PanSN path names, grouping, ordering, subset / exclude
bookkeeping, ItemTable construction -- SURVEY 8a rows a1/a2 and 8f-2) against the oracle's restatement of
the reference (oracle/gfa_oracle.py), through `panacus debug-tables` (no GPU involved)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import gfa_oracle as go

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(ROOT, "panacus_b200", "bin", "panacus")


def debug_tables(gfa, count, flags):
    r = subprocess.run([BIN, "debug-tables", gfa, "-c", count, *flags], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    d = {}
    for line in r.stdout.rstrip("\n").split("\n"):
        key, *vals = line.split("\t")
        d[key] = vals
    return d


def check(gfa, count, flags, kw):
    _check(gfa, count, flags, kw)
    if count != "edge":
        # the lean parse (node ids straight into one flat u32 table: what hist / growth / similarity runs use) must hand
        # over exactly the same tables -- also with a subset list (whole paths in or out: applied to the flat table; BED
        # intervals that cut a path: the front end falls back to the general parse) and with an exclude list (general parse)
        _check(gfa, count, list(flags) + ["--lean"], kw)


def _check(gfa, count, flags, kw):
    got = debug_tables(gfa, count, flags)
    g = go.parse_gfa(gfa)
    mask = go.make_mask(g, **kw)
    t = go.item_tables(g, mask, count)
    op, og, names = go.path_order_arrays(mask, g)
    assert got["groups"] == names, (count, flags)
    assert got["path_order"] == [f"{int(p)}:{int(q)}" for p, q in zip(op, og)]
    assert int(got["n_items"][0]) == t.n_items
    assert [int(x) for x in got["id_prefsum"]] == [int(x) for x in t.id_prefsum]
    assert [int(x) for x in got["items"]] == [int(x) for x in t.items]
    want_ex = [] if t.exclude is None else [int(i) for i in np.nonzero(t.exclude)[0]]
    assert [int(x) for x in got["exclude"]] == want_ex
    assert got["uncovered"] == [f"{k}:{v}" for k, v in sorted(t.uncovered.items())]
    assert [int(x) for x in got["node_lens"]] == list(g.node_lens)
    assert got["paths"] == [str(p) for p in g.path_segments]


B = lambda name: os.path.join(GOLDEN, name)
FLAG_SETS = [
    ([], {}),
    (["-S"], {"groupby_sample": True}),
    (["-H"], {"groupby_haplotype": True}),
    (["-g", B("test_groups.txt")], {"groupby_file": B("test_groups.txt")}),
    (["-s", B("inclusion.bed1")], {"subset": B("inclusion.bed1")}),
    (["-s", B("inclusion.bed3")], {"subset": B("inclusion.bed3")}),
    (["-s", B("inclusion_chm13.bed1")], {"subset": B("inclusion_chm13.bed1")}),
    (["-s", B("inclusion_sub.bed1"), "-S"], {"subset": B("inclusion_sub.bed1"), "groupby_sample": True}),
    (["-e", B("exclusion.bed3")], {"exclude": B("exclusion.bed3")}),
    (["-s", B("inclusion.bed3"), "-e", B("exclusion.bed3"), "-H"],
     {"subset": B("inclusion.bed3"), "exclude": B("exclusion.bed3"), "groupby_haplotype": True}),
    (["-s", "HG00"], {"subset": "HG00"}),          # not a file: a regex over the path names (abacus.rs:212-240)
    (["-e", "grch38"], {"exclude": "grch38"}),
]


@pytest.mark.parametrize("flags,kw", FLAG_SETS)
def test_chrM_front_end(flags, kw):
    for count in ("node", "bp", "edge"):
        check(B("chrM_test.gfa"), count, flags, kw)


@pytest.mark.parametrize("gfa", ["t_groups.gfa", "cdbg.gfa"])
def test_small_fixture_front_end(gfa):
    for count in ("node", "bp", "edge"):
        check(B(gfa), count, [], {})
        check(B(gfa), count, ["-S"], {"groupby_sample": True})


def test_order_file(tmp_path):
    order = tmp_path / "order.txt"
    order.write_text("HG00621\nchm13\nHG00438\ngrch38\n")
    check(B("chrM_test.gfa"), "node", ["-S", "-O", str(order)], {"groupby_sample": True, "order": str(order)})
    frag = tmp_path / "frag.txt"   # groups must not be interleaved in an order file (abacus.rs:114-127)
    g = go.parse_gfa(B("chrM_test.gfa"))
    names = [str(p) for p in g.path_segments]
    frag.write_text("\n".join(names) + "\n")
    r = subprocess.run([BIN, "debug-tables", B("chrM_test.gfa"), "-g", B("test_groups.txt"), "-O", str(frag)],
                       capture_output=True, text=True)
    mask_ok = True
    try:
        go.make_mask(g, groupby_file=B("test_groups.txt"), order=str(frag))
    except Exception:
        mask_ok = False
    assert (r.returncode == 0) or ("fragmented" in r.stderr) or not mask_ok


def test_generated_gfa_with_walks(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("tc", os.path.join(ROOT, "tests", "test_cli.py"))
    tc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tc)
    gfa = str(tmp_path / "syn.gfa")
    tc.write_synthetic_gfa(gfa, n_nodes=1500, n_samples=5, seed=4)
    for count in ("node", "bp", "edge"):
        check(gfa, count, [], {})
        check(gfa, count, ["-H"], {"groupby_haplotype": True})
        check(gfa, count, ["-S", "-e", "smp1"], {"groupby_sample": True, "exclude": "smp1"})


@pytest.mark.parametrize("name,expect", [
    ("a#1#chr:5-9", "a#1#chr:5-9"), ("a#1", "a#1"), ("x", "x"), ("a#1#h1#more", "a#1#h1#more"), ("a##b", "a##b"),
    ("s:10-20", "s:10-20"), ("a#2:3-7", "a#2:3-7"),
])
def test_path_segment_names(tmp_path, name, expect):
    """PathSegment::from_str / Display (graph.rs:495-616) on PanSN corner cases, vs the oracle's regex version."""
    gfa = tmp_path / "p.gfa"
    gfa.write_text(f"S\t1\tACGT\nS\t2\tAC\nP\t{name}\t1+,2+\t*\n")
    got = debug_tables(str(gfa), "node", [])
    seg = go.PathSegment.from_str(name)
    assert got["paths"] == [str(seg)] == [expect]
    assert got["groups"] == [seg.id()]


# ---- the `table` writer (AbacusByGroup::to_tsv, abacus.rs:1056-1178) fed with the oracle's r / c / v -----------------

@pytest.mark.parametrize("gfa", ["chrM_test.gfa", "t_groups.gfa", "cdbg.gfa"])
def test_table_writer_matches_oracle(gfa, tmp_path):
    from oracle import oracle as po
    path = B(gfa)
    sets = [([], {}), (["-S"], {"groupby_sample": True})]
    if gfa == "chrM_test.gfa":
        sets += [(["-s", B("inclusion.bed3")], {"subset": B("inclusion.bed3")}),
                 (["-e", B("exclusion.bed3"), "-H"], {"exclude": B("exclusion.bed3"), "groupby_haplotype": True})]
    for flags, kw in sets:
        for count in ("node", "bp", "edge"):
            g = go.parse_gfa(path)
            mask = go.make_mask(g, **kw)
            t = go.item_tables(g, mask, count)
            op, og, names = go.path_order_arrays(mask, g)
            r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
            csr = tmp_path / "csr.txt"
            csr.write_text("\n".join("\t".join(str(int(x)) for x in a) for a in (r, c, v)) + "\n")
            for total in (False, True):
                try:
                    want = go.abacus_by_group_to_tsv(g, count, total, names, r, c, v, t.uncovered)
                except IndexError:
                    want = None  # the reference panics here (v[j] with j >= nnz, abacus.rs:1166)
                p = subprocess.run([BIN, "debug-table-tsv", path, "-c", count, "--csr", str(csr), *flags,
                                    *(["--total"] if total else [])], capture_output=True, text=True, timeout=120)
                if want is None:
                    assert p.returncode != 0 and "index out of bounds" in p.stderr
                else:
                    assert p.returncode == 0, p.stderr
                    assert p.stdout == want, (gfa, flags, count, total)


# ---- GFA parser variants: numeric segment names (direct table), other names (hash map), walks, gzip, threads ---------

def _write_gfa(path, names, paths, walks=(), links=()):
    with open(path, "w") as f:
        f.write("H\tVN:Z:1.1\n")
        for i, nm in enumerate(names):
            f.write(f"S\t{nm}\t{'ACGT'[i % 4] * (1 + i % 5)}\n")
        for u, ou, v, ov in links:
            f.write(f"L\t{u}\t{ou}\t{v}\t{ov}\t0M\n")
        for pname, steps in paths:
            f.write(f"P\t{pname}\t" + ",".join(f"{n}{o}" for n, o in steps) + "\t*\n")
        for (sample, hap, seq), steps in walks:
            f.write(f"W\t{sample}\t{hap}\t{seq}\t0\t100\t" + "".join((">" if o == "+" else "<") + n for n, o in steps) + "\n")


@pytest.mark.parametrize("naming", ["dense", "sparse", "leading_zero", "alpha"])
@pytest.mark.parametrize("threads", ["1", "3"])
def test_parser_naming_schemes(naming, threads, tmp_path):
    rng = np.random.default_rng(len(naming))
    n = 300
    if naming == "dense":
        names = [str(i) for i in rng.permutation(np.arange(1, n + 1))]          # direct table, ids != values
    elif naming == "sparse":
        names = [str(int(v)) for v in rng.choice(10 ** 12, n, replace=False)]    # too sparse for a table -> hash map
    elif naming == "leading_zero":
        names = [f"{i:05d}" for i in range(1, n + 1)]                            # not canonical decimals -> hash map
    else:
        names = [f"s{i}" for i in range(n)]
    def steps(k):
        return [(names[int(j)], "+-"[int(o)]) for j, o in zip(rng.integers(0, n, k), rng.integers(0, 2, k))]
    paths = [(f"smp{p // 2}#{p % 2 + 1}#ctg{p}", steps(int(rng.integers(1, 400)))) for p in range(7)]
    walks = [((f"w{p}", "1", f"chr{p}"), steps(int(rng.integers(1, 200)))) for p in range(3)]
    gfa = str(tmp_path / "g.gfa")
    _write_gfa(gfa, names, paths, walks)
    check(gfa, "node", ["-t", threads], {})
    check(gfa, "bp", ["-S", "-t", threads], {"groupby_sample": True})
    # the same file gzip-compressed
    import gzip, shutil
    with open(gfa, "rb") as fi, gzip.open(gfa + ".gz", "wb") as fo:
        shutil.copyfileobj(fi, fo)
    check(gfa + ".gz", "node", ["-t", threads], {})


def test_parser_errors(tmp_path):
    gfa = str(tmp_path / "bad.gfa")
    _write_gfa(gfa, ["1", "2", "3"], [("a#1#x", [("1", "+"), ("9", "-")])])
    r = subprocess.run([BIN, "debug-tables", gfa], capture_output=True, text=True)
    assert r.returncode != 0 and "unknown node 9" in r.stderr
    _write_gfa(gfa, ["1", "2", "7"], [("a#1#x", [("1", "+"), ("07", "-")])])  # "07" is not the segment "7"
    r = subprocess.run([BIN, "debug-tables", gfa], capture_output=True, text=True)
    assert r.returncode != 0 and "unknown node 07" in r.stderr
    _write_gfa(gfa, ["1", "2", "2"], [("a#1#x", [("1", "+")])])
    r = subprocess.run([BIN, "debug-tables", gfa], capture_output=True, text=True)
    assert r.returncode != 0 and "occurs multiple times" in r.stderr
    _write_gfa(gfa, ["s1", "s2"], [("a#1#x", [("s1", "+"), ("s3", "-")])])
    r = subprocess.run([BIN, "debug-tables", gfa, "-t", "2"], capture_output=True, text=True)
    assert r.returncode != 0 and "unknown node s3" in r.stderr


@pytest.mark.parametrize("count", ["node", "bp", "edge"])
def test_threaded_front_end_matches_oracle(count):
    """-t N: P / W lines parsed and (without subset / exclude lists) translated to item ids on N threads"""
    check(B("chrM_test.gfa"), count, ["-t", "4"], {})
    check(B("chrM_test.gfa"), count, ["-t", "4", "-S"], {"groupby_sample": True})
    check(B("chrM_test.gfa"), count, ["-t", "4", "-e", B("exclusion.bed3")], {"exclude": B("exclusion.bed3")})


def test_lean_parse_on_awkward_files(tmp_path):
    """CRLF line ends, no trailing newline, empty lines, empty step tokens, tags after the sequence, walks, many threads
    on a file that is cut into several slices: default and lean parse agree with each other and with the oracle."""
    rng = np.random.default_rng(9)
    n = 5000
    names = [str(i) for i in range(1, n + 1)]
    lines = ["H\tVN:Z:1.1", ""]
    for i, nm in enumerate(names):
        lines.append(f"S\t{nm}\t{'ACGT'[i % 4] * (1 + i % 7)}" + ("\tLN:i:9\tSN:Z:x" if i % 11 == 0 else ""))
    for p in range(40):
        k = int(rng.integers(1, 3000))
        st = [f"{int(j)}{'+-'[int(o)]}" for j, o in zip(rng.integers(1, n + 1, k), rng.integers(0, 2, k))]
        lines.append(f"P\ts{p // 4}#{p % 2 + 1}#c{p}\t" + ",".join(st) + ("," if p % 5 == 0 else "") + "\t*")
    lines.append("W\tw0\t1\tchrW\t0\t10\t" + "".join(f">{int(j)}" if j % 2 else f"<{int(j)}" for j in rng.integers(1, n + 1, 500)))
    for name, text in (("lf.gfa", "\n".join(lines) + "\n"), ("crlf.gfa", "\r\n".join(lines) + "\r\n"), ("noeol.gfa", "\n".join(lines))):
        gfa = str(tmp_path / name)
        with open(gfa, "w", newline="") as f:
            f.write(text)
        for threads in ("1", "5"):
            a = debug_tables(gfa, "bp", ["-S", "-t", threads])
            b = debug_tables(gfa, "bp", ["-S", "-t", threads, "--lean"])
            assert a == b, (name, threads)
        if name == "lf.gfa":
            check(gfa, "node", ["-t", "5"], {})
    # errors surface from the lean parse as well
    bad = str(tmp_path / "bad.gfa")
    _write_gfa(bad, ["1", "2", "3"], [("a#1#x", [("1", "+"), ("9", "-")])])
    r = subprocess.run([BIN, "debug-tables", bad, "--lean"], capture_output=True, text=True)
    assert r.returncode != 0 and "unknown node 9" in r.stderr
    _write_gfa(bad, ["1", "2", "2"], [("a#1#x", [("1", "+")])])
    r = subprocess.run([BIN, "debug-tables", bad, "--lean", "-t", "3"], capture_output=True, text=True)
    assert r.returncode != 0 and "occurs multiple times" in r.stderr


# ---- seeded fuzz: random graphs, naming schemes, line orders, P / W mixes, grouping flags, subset / exclude lists --------

def _fuzz_case(seed, d):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    scheme = int(rng.integers(0, 4))
    if scheme == 0:
        names = [str(i) for i in rng.permutation(np.arange(1, n + 1))]
    elif scheme == 1:
        names = [str(int(v)) for v in rng.choice(10 ** 9, n, replace=False) + 1]
    elif scheme == 2:
        names = [f"{i:04d}" for i in range(1, n + 1)]
    else:
        names = [f"n{i}x" for i in range(n)]
    seg = [f"S\t{nm}\t{'ACGT'[i % 4] * int(rng.integers(1, 9))}" + ("\tRC:i:3" if rng.random() < .2 else "") for i, nm in enumerate(names)]
    pw, used, pnames = [], [], []
    for p in range(int(rng.integers(1, 12))):
        k = int(rng.integers(1, 300))
        idx, ori = rng.integers(0, n, k), rng.integers(0, 2, k)
        used += [(int(idx[i]), int(ori[i]), int(idx[i + 1]), int(ori[i + 1])) for i in range(k - 1)]
        if rng.random() < .7:
            nm = [f"s{p // 3}#{p % 2 + 1}#c{p}", f"s{p // 3}#{p % 2 + 1}#c{p}:{10 * p}-{10 * p + k}", f"plain{p}", f"s{p // 2}#{p % 2}"][int(rng.integers(0, 4))]
            pw.append(f"P\t{nm}\t" + ",".join(f"{names[int(j)]}{'-+'[int(o)]}" for j, o in zip(idx, ori)) + "\t*")
            pnames.append(nm)
        else:
            pw.append(f"W\tw{p // 2}\t{p % 2 + 1}\tchr{p}\t{5 * p}\t{5 * p + k}\t" + "".join(("<>"[int(o)]) + names[int(j)] for j, o in zip(idx, ori)))
            pnames.append(f"w{p // 2}#{p % 2 + 1}#chr{p}:{5 * p}-{5 * p + k}")
    for _ in range(int(rng.integers(0, 50))):
        u, v = rng.integers(0, n, 2)
        used.append((int(u), int(rng.integers(0, 2)), int(v), int(rng.integers(0, 2))))
    links = []
    for u, ou, v, ov in used:  # every edge a path walks has its L line (the reference panics otherwise), some twice
        links.append(f"L\t{names[u]}\t{'-+'[ou]}\t{names[v]}\t{'-+'[ov]}\t0M")
        if rng.random() < .3:
            links.append(links[-1])
    body = seg + pw + links
    if rng.random() < .5:  # the record types may come in any order
        body = [body[i] for i in rng.permutation(len(body))]
    gfa = os.path.join(d, f"f{seed}.gfa")
    with open(gfa, "w") as f:
        f.write("\n".join(["H\tVN:Z:1.0"] + body) + ("\n" if rng.random() < .8 else ""))
    count = ["node", "bp", "edge"][int(rng.integers(0, 3))]
    flags, kw = [([], {}), (["-S"], {"groupby_sample": True}), (["-H"], {"groupby_haplotype": True})][int(rng.integers(0, 3))]
    flags, kw = flags + ["-t", str(int(rng.integers(1, 6)))], dict(kw)
    for kind in ("subset", "exclude"):  # 1-column lists (names with / without coordinates) or BED3 intervals
        r = rng.random()
        if r < .35:
            continue
        rows = []
        for nm in pnames:
            if rng.random() < .5:
                continue
            base, _, coords = nm.partition(":")
            if r < .65:
                rows.append(nm if rng.random() < .5 else base)
            else:
                off = int(coords.split("-")[0]) if "-" in coords else 0
                a = off + int(rng.integers(0, 40))
                rows.append(f"{base}\t{a}\t{a + int(rng.integers(0, 200))}")
        if not rows and kind == "subset":
            rows.append(pnames[0])
        if rows:
            lp = os.path.join(d, f"f{seed}.{kind}")
            with open(lp, "w") as f:
                f.write("\n".join(rows) + "\n")
            flags += ["-s" if kind == "subset" else "-e", lp]
            kw[kind] = lp
    return gfa, count, flags, kw


@pytest.mark.parametrize("block", range(4))
def test_front_end_fuzz(block, tmp_path):
    """Whatever the file looks like, the C++ front end (default and lean parse) hands over the oracle's tables."""
    for seed in range(block * 15, block * 15 + 15):
        gfa, count, flags, kw = _fuzz_case(seed, str(tmp_path))
        check(gfa, count, flags, kw)


# ---- bgzip-compressed input: blocks inflated on the worker threads --------------------------------------------------------

def _bgzf(data, block=65280, eof=True):
    """BGZF as bgzip writes it: gzip members of <= 64 KiB with a 'BC' extra field holding the member size - 1"""
    import struct
    import zlib
    out = bytearray()

    def member(chunk):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        d = c.compress(chunk) + c.flush()
        out.extend(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(d) + 8 - 1))
        out.extend(d + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    for i in range(0, len(data), block):
        member(data[i:i + block])
    if eof:
        member(b"")
    return bytes(out)


def test_bgzf_input(tmp_path):
    import gzip
    rng = np.random.default_rng(21)
    n = 3000
    names = [str(i) for i in range(1, n + 1)]
    paths = [(f"s{p // 2}#{p % 2 + 1}#c{p}", [(names[int(j)], "+-"[int(o)]) for j, o in zip(rng.integers(0, n, 4000), rng.integers(0, 2, 4000))])
             for p in range(12)]
    plain = str(tmp_path / "g.gfa")
    _write_gfa(plain, names, paths)
    data = open(plain, "rb").read()
    want = debug_tables(plain, "bp", ["-S"])
    variants = {
        "blocks.gfa.gz": _bgzf(data, block=4096),                        # ~60 blocks, lines straddle them
        "noeof.gfa.gz": _bgzf(data, eof=False),
        "one.gfa.gz": _bgzf(data, block=1 << 30) if len(data) < 60000 else _bgzf(data, block=60000),
        "mixed.gfa.gz": _bgzf(data[:50000], eof=False) + gzip.compress(data[50000:]),  # not BGZF throughout: serial inflate
        "plain.gfa.gz": gzip.compress(data),
    }
    for name, blob in variants.items():
        f = str(tmp_path / name)
        with open(f, "wb") as fo:
            fo.write(blob)
        for threads in ("1", "4"):
            assert debug_tables(f, "bp", ["-S", "-t", threads]) == want, (name, threads)
            assert debug_tables(f, "bp", ["-S", "-t", threads, "--lean"]) == want, (name, threads)
    check(str(tmp_path / "blocks.gfa.gz"), "node", ["-t", "3"], {})      # and against the oracle's parser
    # a damaged block is an error, not silently different text
    blob = bytearray(variants["blocks.gfa.gz"])
    blob[len(blob) // 2] ^= 0x5A
    bad = str(tmp_path / "bad.gfa.gz")
    with open(bad, "wb") as fo:
        fo.write(bytes(blob))
    r = subprocess.run([BIN, "debug-tables", bad, "-t", "4"], capture_output=True, text=True)
    assert r.returncode != 0 and r.stderr.strip()
    # only the end-of-file marker: an empty graph, like an empty file
    empty = str(tmp_path / "empty.gfa.gz")
    with open(empty, "wb") as fo:
        fo.write(_bgzf(b""))
    r = subprocess.run([BIN, "debug-parse", empty], capture_output=True, text=True)
    assert r.returncode == 0 and "nodes\t0" in r.stdout
