"""The `panacus` CLI of the C++ host layer (panacus_b200/bin/panacus) against the oracle's TSV, from the
`panacus` header row on (the comment lines carry argv, like the reference's own R check skips them,
test/integrated_test.R:20-24)."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import gfa_oracle as go
from oracle import oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(ROOT, "panacus_b200", "bin", "panacus")


def run_cli(*args, expect_ok=True):
    assert os.path.exists(BIN), "build the host layer first (__graft_entry__.build())"
    r = subprocess.run([BIN, *args], capture_output=True, text=True, timeout=300)
    if expect_ok:
        assert r.returncode == 0, r.stderr
    return r


def run_many(arglists, tmp_path):
    """several commands in ONE process (`panacus batch`): the CUDA context is created once"""
    f = tmp_path / "batch.txt"
    f.write_text("".join(" ".join(a) + "\n" for a in arglists))
    r = subprocess.run([BIN, "batch", str(f)], capture_output=True, text=True, timeout=600)
    parts = r.stdout.split("## batch ")[1:]
    assert len(parts) == len(arglists), (r.stdout[-500:], r.stderr[-500:])
    outs = []
    for i, part in enumerate(parts):
        head, _, rest = part.partition("\n")
        assert head == f"{i} rc=0", (head, rest[:300])
        outs.append(rest)
    return outs


def body(text):
    """drop the comment lines and the trailing blank line"""
    lines = [l for l in text.split("\n") if not l.startswith("#")]
    while lines and lines[-1] == "":
        lines.pop()
    return "\n".join(lines) + "\n"


# ---- CPU: `growth <hist.tsv>` needs no GPU; pins the product's closed-form growth to the reference goldens ----

def test_growth_from_hist_tsv_chr22_golden(tmp_path):
    d = json.load(open(os.path.join(GOLDEN, "chr22_histgrowth.json")))
    counts = ["bp", "node", "edge"]
    hist_tsv = tmp_path / "chr22.hist.tsv"
    n = len(d["hist"]["bp"]["values"])
    with open(hist_tsv, "w") as f:
        f.write("# panacus hist chr22 (golden, docs/chr22.hprc-v1.0-pggb.histgrowth.html:267-269)\n")
        f.write("panacus\t" + "\t".join("hist" for _ in counts) + "\n")
        f.write("count\t" + "\t".join(counts) + "\n")
        f.write("\t" * len(counts) + "\n" + "\t" * len(counts) + "\n")
        for i in range(n):
            f.write(str(i) + "\t" + "\t".join(str(d["hist"][c]["values"][i]) for c in counts) + "\n")
    g = d["growth"]["bp"]
    cov = ",".join(str(int(c)) for c in g["coverage"])
    quo = ",".join(repr(float(q)) if float(q) not in (0.0, 1.0) else str(int(q)) for q in g["quorum"])
    out = run_cli("growth", str(hist_tsv), "-l", cov, "-q", quo).stdout
    rows = [l.split("\t") for l in body(out).strip().split("\n")]
    assert rows[0][0] == "panacus" and rows[1][0] == "count"
    T = len(g["coverage"])
    for ci, c in enumerate(counts):
        for t in range(T):
            col = 1 + ci * T + t
            assert rows[0][col] == "growth" and rows[1][col] == c
            got = [r[col] for r in rows[4:]]
            assert got[0] == "NaN"
            assert [int(x) for x in got[1:]] == [int(v) for v in d["growth"][c]["curves"][t]], (c, t)
    # the comment lines of the input table are carried over, then "# argv" (growth.rs:40-51)
    assert out.startswith("# panacus hist chr22")


def test_growth_from_hist_tsv_matches_oracle_table():
    out = run_cli("growth", os.path.join(GOLDEN, "t_groups.hist.tsv"), "-q", "0,0.5,1", "-l", "1,1,2", "-a").stdout
    cov, quo = po.parse_thresholds("0,0.5,1", "1,1,2")
    assert body(out) == po.growth_table([("node", [5, 0, 10, 0, 0, 0, 0])], cov, quo, add_hist=True)


@pytest.mark.parametrize("threads", ["1", "3", "8"])
def test_closed_form_growth_bits_match_oracle_for_any_thread_count(tmp_path, threads):
    """The host's closed-form growth (threads over coverage classes, log2 table) against the oracle's
    statement-by-statement port of hist.rs:51-187: every f64 bit (the TSV only shows floor()), for union, core and
    quorum pairs, on random hists with empty bins and on the reference's own KAT hists (hist.rs:352-398)."""
    rng = np.random.default_rng(12)
    hists = [[0, 5, 3, 2], [0, 5, 3, 2, 3, 5, 0, 4, 2, 1]]
    for n in (7, 45, 130):
        h = rng.integers(0, 10**6, n + 1)
        h[rng.random(n + 1) < 0.2] = 0
        h[0] = 0
        hists.append([int(x) for x in h])
    pairs = [("1", "0"), ("2", "0"), ("1", "1"), ("1", "0.5"), ("2", "0.9"), ("1", "0.1"), ("3", "0.33")]
    cov_s, quo_s = ",".join(c for c, _ in pairs), ",".join(q for _, q in pairs)
    cov, quo = po.parse_thresholds(quo_s, cov_s)
    for k, h in enumerate(hists):
        f = tmp_path / f"h{k}.tsv"
        f.write_text("panacus\thist\ncount\tnode\n\t\n\t\n" + "".join(f"{i}\t{v}\n" for i, v in enumerate(h)))
        out = run_cli("debug-growth", str(f), "-l", cov_s, "-q", quo_s, "-t", threads).stdout.strip().split("\n")
        assert len(out) == len(pairs)
        for t, line in enumerate(out):
            got = [float.fromhex(x) for x in line.split("\t")[1:]]
            want = po.hist_calc_growth(np.array(h, dtype=np.uint64), cov[t], quo[t])
            assert len(got) == len(want) - 1 or len(got) == len(want)
            want = [float(x) for x in (want[1:] if len(want) == len(got) + 1 else want)]
            assert [x.hex() for x in got] == [x.hex() for x in want], (k, pairs[t], threads)


def test_cli_errors():
    r = run_cli("growth", os.path.join(GOLDEN, "t_groups.hist.tsv"), "-q", "1.5", expect_ok=False)
    assert r.returncode != 0 and "within [0,1]" in r.stderr
    r = run_cli("growth", os.path.join(GOLDEN, "t_groups.hist.tsv"), "-q", "0,0.5", "-l", "1,2,3", expect_ok=False)
    assert r.returncode != 0 and "must match" in r.stderr
    r = run_cli("growth", os.path.join(GOLDEN, "t_groups.hist.tsv"), "-S", expect_ok=False)
    assert r.returncode != 0 and "graph mode" in r.stderr
    r = run_cli("frobnicate", "x", expect_ok=False)
    assert r.returncode == 2


# ---- GPU: full subcommands on the reference's fixtures --------------------------------------------------------------

def oracle_tables(gfa, count, **kw):
    g = go.parse_gfa(os.path.join(GOLDEN, gfa))
    mask = go.make_mask(g, **kw)
    t = go.item_tables(g, mask, count)
    op, og, names = go.path_order_arrays(mask, g)
    return g, t, op, og, names


def oracle_hist(gfa, count, **kw):
    g, t, op, og, names = oracle_tables(gfa, count, **kw)
    countable = po.abacus_by_total(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    if count == "bp":
        return po.construct_hist_bps(countable, g.node_lens, len(names), t.uncovered)
    return po.construct_hist(countable, len(names))


FLAG_SETS = [
    ([], {}),
    (["-S"], {"groupby_sample": True}),
    (["-H"], {"groupby_haplotype": True}),
    (["-g", os.path.join(GOLDEN, "test_groups.txt")], {"groupby_file": os.path.join(GOLDEN, "test_groups.txt")}),
    (["-s", os.path.join(GOLDEN, "inclusion.bed1")], {"subset": os.path.join(GOLDEN, "inclusion.bed1")}),
    (["-s", os.path.join(GOLDEN, "inclusion.bed3")], {"subset": os.path.join(GOLDEN, "inclusion.bed3")}),
    (["-e", os.path.join(GOLDEN, "exclusion.bed3")], {"exclude": os.path.join(GOLDEN, "exclusion.bed3")}),
    (["-S", "-s", os.path.join(GOLDEN, "inclusion_sub.bed1"), "-e", os.path.join(GOLDEN, "exclusion.bed3")],
     {"groupby_sample": True, "subset": os.path.join(GOLDEN, "inclusion_sub.bed1"),
      "exclude": os.path.join(GOLDEN, "exclusion.bed3")}),
]


@pytest.mark.gpu
@pytest.mark.parametrize("flags,kw", FLAG_SETS)
def test_hist_chrM(flags, kw, tmp_path):
    gfa = os.path.join(GOLDEN, "chrM_test.gfa")
    counts = ["node", "bp", "edge", "all"]
    outs = run_many([["hist", gfa, "-c", c, *flags] for c in counts], tmp_path)
    for count, out in zip(counts[:3], outs):
        assert body(out) == po.hist_table([(count, oracle_hist("chrM_test.gfa", count, **kw))]), (count, flags)
        assert out.split("\n")[1].startswith("# version")
    want = po.hist_table([(c, oracle_hist("chrM_test.gfa", c, **kw)) for c in ("node", "bp", "edge")])
    assert body(outs[3]) == want


@pytest.mark.gpu
def test_hist_chrM_reference_golden():
    """BASELINE.json configs[0] family: the printed bp / node / edge hists are the reference's KATs."""
    kats = json.load(open(os.path.join(GOLDEN, "kats.json")))["chrM_groupby_sample"]
    out = run_cli("hist", os.path.join(GOLDEN, "chrM_test.gfa"), "-S", "-c", "all").stdout
    rows = [l.split("\t") for l in body(out).strip().split("\n")][4:]
    assert [int(r[1]) for r in rows] == kats["node"]
    assert [int(r[2]) for r in rows] == kats["bp"]
    assert [int(r[3]) for r in rows] == kats["edge"]


@pytest.mark.gpu
@pytest.mark.parametrize("gfa", ["chrM_test.gfa", "t_groups.gfa", "cdbg.gfa"])
def test_histgrowth_and_growth(gfa, tmp_path):
    path = os.path.join(GOLDEN, gfa)
    cov, quo = po.parse_thresholds("0,0.5,1", "0,1,2")
    counts = [c for c in ("node", "bp", "edge") if not (gfa == "cdbg.gfa" and c == "edge")]
    outs = run_many([["histgrowth", path, "-c", c, "-q", "0,0.5,1", "-l", "0,1,2", "-a"] for c in counts]
                    + [["growth", path, "-q", "0,0.5,1", "-l", "0,1,2"]], tmp_path)
    for count, out in zip(counts, outs):
        want = po.growth_table([(count, oracle_hist(gfa, count))], cov, quo, add_hist=True)
        assert body(out) == want, (gfa, count)
    out = outs[-1]
    assert body(out) == po.growth_table([("node", oracle_hist(gfa, "node"))], cov, quo)
    assert not any(l.startswith("# version") for l in out.split("\n"))  # growth.rs:48-51: no version line


@pytest.mark.gpu
@pytest.mark.parametrize("flags,kw", FLAG_SETS)
def test_ordered_histgrowth_chrM(flags, kw, tmp_path):
    cov, quo = po.parse_thresholds("0,0.5,0.9,1", "1,2,1,1")
    counts = ("node", "bp", "edge")
    outs = run_many([["ordered-histgrowth", os.path.join(GOLDEN, "chrM_test.gfa"), "-c", c, "-q", "0,0.5,0.9,1",
                      "-l", "1,2,1,1", *flags] for c in counts], tmp_path)
    for count, out in zip(counts, outs):
        g, t, op, og, names = oracle_tables("chrM_test.gfa", count, **kw)
        r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
        unc = None
        if t.uncovered:
            unc = np.zeros(t.n_items + 1, dtype=np.uint64)
            for k, val in t.uncovered.items():
                unc[k] = val
        curves = [po.calc_growth(r, c, len(names), cc, qq, count_bp=(count == "bp"), node_lens=g.node_lens, uncovered=unc)
                  for cc, qq in zip(cov, quo)]
        want = po.ordered_growth_table(count, names, curves, cov, quo)
        assert body(out) == want, (count, flags)


@pytest.mark.gpu
def test_ordered_histgrowth_with_order_file(tmp_path):
    order = tmp_path / "order.txt"
    order.write_text("HG00621\nchm13\nHG00438\ngrch38\n")
    kw = {"groupby_sample": True, "order": str(order)}
    g, t, op, og, names = oracle_tables("chrM_test.gfa", "node", **kw)
    assert names == ["HG00621", "chm13", "HG00438", "grch38"]
    r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    cov, quo = po.parse_thresholds("0", "1")
    want = po.ordered_growth_table("node", names, [po.calc_growth(r, c, 4, cov[0], quo[0])], cov, quo)
    out = run_cli("ordered-histgrowth", os.path.join(GOLDEN, "chrM_test.gfa"), "-S", "-O", str(order)).stdout
    assert body(out) == want
    # t_groups in file order: the SURVEY appendix A example
    out = run_cli("ordered-histgrowth", os.path.join(GOLDEN, "t_groups.gfa")).stdout
    assert body(out) == ("panacus\tordered-growth\ncount\tnode\ncoverage\t1\nquorum\t0\n"
                         "y#1\t2\ny#2\t5\ny#3\t8\ny#4\t9\ny#5\t10\nx\t10\n")


@pytest.mark.gpu
def test_similarity_values():
    for count in ("node", "bp"):
        g, t, op, og, names = oracle_tables("chrM_test.gfa", count, groupby_sample=True)
        r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
        inter, ln, table = po.similarity(r, c, len(names), count_bp=(count == "bp"), node_lens=g.node_lens)
        out = run_cli("similarity", os.path.join(GOLDEN, "chrM_test.gfa"), "-S", "-c", count, "--no-cluster=1").stdout
        rows = [l.split("\t") for l in body(out).strip().split("\n")]
        assert rows[0] == ["group"] + names
        for i, name in enumerate(names):
            assert rows[1 + i][0] == name
            assert [np.float32(x) for x in rows[1 + i][1:]] == [table[i, j] for j in range(len(names))]
        # clustered output: same values, rows and columns permuted consistently
        out = run_cli("similarity", os.path.join(GOLDEN, "chrM_test.gfa"), "-S", "-c", count).stdout
        rows = [l.split("\t") for l in body(out).strip().split("\n")]
        perm = [names.index(x) for x in rows[0][1:]]
        assert sorted(perm) == list(range(len(names))) and [r[0] for r in rows[1:]] == rows[0][1:]
        for i, pi in enumerate(perm):
            assert [np.float32(x) for x in rows[1 + i][1:]] == [table[pi, pj] for pj in perm]


# ---- GPU: a generated GFA (P and W lines, PanSN names, links) end to end ---------------------------------------------

def write_synthetic_gfa(path, n_nodes=3000, n_samples=7, seed=1):
    rng = np.random.default_rng(seed)
    lens = np.where(rng.random(n_nodes) < 0.5, 1, 1 + np.floor(np.exp(rng.normal(2.0, 1.5, n_nodes))).astype(int))
    lines = ["H\tVN:Z:1.1"]
    for i in range(n_nodes):
        lines.append(f"S\ts{i + 1}\t" + "ACGT"[i % 4] * int(min(lens[i], 400)))
    walks, links = [], set()
    for s in range(n_samples):
        for hap in (1, 2):
            for contig in range(2):
                n = int(rng.integers(50, 600))
                start = int(rng.integers(0, n_nodes - 1))
                nodes = [start]
                for _ in range(n - 1):
                    step = int(rng.choice([1, 1, 1, 2, 3, -1]))
                    nodes.append(int(np.clip(nodes[-1] + step, 0, n_nodes - 1)))
                ori = ["+" if rng.random() < 0.85 else "-" for _ in nodes]
                for (a, oa), (b, ob) in zip(zip(nodes, ori), zip(nodes[1:], ori[1:])):
                    links.add((a, oa, b, ob))
                walks.append((f"smp{s}", hap, f"ctg{contig}", nodes, ori))
    for a, oa, b, ob in sorted(links):
        lines.append(f"L\ts{a + 1}\t{oa}\ts{b + 1}\t{ob}\t0M")
    for k, (smp, hap, ctg, nodes, ori) in enumerate(walks):
        if k % 3 == 0:  # W line
            walk = "".join((">" if o == "+" else "<") + f"s{n + 1}" for n, o in zip(nodes, ori))
            lines.append(f"W\t{smp}\t{hap}\t{ctg}\t0\t{len(nodes)}\t{walk}")
        else:
            steps = ",".join(f"s{n + 1}{o}" for n, o in zip(nodes, ori))
            lines.append(f"P\t{smp}#{hap}#{ctg}\t{steps}\t*")
    open(path, "w").write("\n".join(lines) + "\n")


@pytest.mark.gpu
def test_generated_gfa_end_to_end(tmp_path):
    gfa = str(tmp_path / "synthetic.gfa")
    write_synthetic_gfa(gfa)
    cov, quo = po.parse_thresholds("0,0.3,0.8", "1,2,3")
    cmds = [["hist", gfa, "-c", "all", "-H"], ["histgrowth", gfa, "-c", "bp", "-S", "-q", "0,0.3,0.8", "-l", "1,2,3"],
            ["ordered-histgrowth", gfa, "-c", "bp", "-S", "-q", "0,0.3,0.8", "-l", "1,2,3"],
            ["ordered-histgrowth", gfa, "-c", "edge", "-q", "0,0.3,0.8", "-l", "1,2,3"]]
    outs = run_many(cmds, tmp_path)

    def tables(count, **kw):
        g = go.parse_gfa(gfa)
        mask = go.make_mask(g, **kw)
        t = go.item_tables(g, mask, count)
        op, og, names = go.path_order_arrays(mask, g)
        return g, t, op, og, names

    hists = []
    for count in ("node", "bp", "edge"):
        g, t, op, og, names = tables(count, groupby_haplotype=True)
        ct = po.abacus_by_total(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
        hists.append((count, po.construct_hist_bps(ct, g.node_lens, len(names)) if count == "bp"
                      else po.construct_hist(ct, len(names))))
    assert body(outs[0]) == po.hist_table(hists)
    g, t, op, og, names = tables("bp", groupby_sample=True)
    ct = po.abacus_by_total(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    assert body(outs[1]) == po.growth_table([("bp", po.construct_hist_bps(ct, g.node_lens, len(names)))], cov, quo)
    r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    curves = [po.calc_growth(r, c, len(names), cc, qq, count_bp=True, node_lens=g.node_lens) for cc, qq in zip(cov, quo)]
    assert body(outs[2]) == po.ordered_growth_table("bp", names, curves, cov, quo)
    g, t, op, og, names = tables("edge")
    r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    curves = [po.calc_growth(r, c, len(names), cc, qq) for cc, qq in zip(cov, quo)]
    assert body(outs[3]) == po.ordered_growth_table("edge", names, curves, cov, quo)
    assert len(names) == 28  # 7 samples x 2 haplotypes x 2 contigs, one group per path


@pytest.mark.gpu
def test_packed_abacus_cache_round_trip(tmp_path):
    """--save-abacus writes <prefix>.<count>.pabm; feeding the .pabm back gives byte-identical tables
    (skip GFA parsing on repeated runs)."""
    gfa = os.path.join(GOLDEN, "chrM_test.gfa")
    prefix = str(tmp_path / "chrM")
    sub = os.path.join(GOLDEN, "inclusion.bed3")
    first = run_many([["hist", gfa, "-c", "all", "-S", "-s", sub, "--save-abacus", prefix],
                      ["ordered-histgrowth", gfa, "-c", "bp", "-S", "-s", sub, "-q", "0,0.5", "-l", "1,2"],
                      ["histgrowth", gfa, "-c", "edge", "-S", "-s", sub, "-q", "0,1", "-a"],
                      ["similarity", gfa, "-c", "node", "-S", "-s", sub]], tmp_path)
    for c in ("node", "bp", "edge"):
        assert os.path.exists(f"{prefix}.{c}.pabm")
    again = run_many([["hist", f"{prefix}.bp.pabm", "-c", "bp"],
                      ["ordered-histgrowth", f"{prefix}.bp.pabm", "-c", "bp", "-q", "0,0.5", "-l", "1,2"],
                      ["histgrowth", f"{prefix}.edge.pabm", "-c", "edge", "-q", "0,1", "-a"],
                      ["similarity", f"{prefix}.node.pabm", "-c", "node"]], tmp_path)
    all_rows = [l.split("\t") for l in body(first[0]).strip().split("\n")]
    bp_rows = [l.split("\t") for l in body(again[0]).strip().split("\n")]
    assert [r[2] for r in all_rows] == [r[1] for r in bp_rows]      # the bp column, incl. the uncovered_bps patch
    for k in (1, 2, 3):
        assert body(first[k]) == body(again[k])
    r = run_cli("hist", f"{prefix}.bp.pabm", "-c", "node", expect_ok=False)
    assert r.returncode != 0 and "holds count type" in r.stderr
    r = run_cli("hist", f"{prefix}.bp.pabm", "-c", "bp", "-S", expect_ok=False)
    assert r.returncode != 0 and "cannot be combined" in r.stderr


# ---- table / coverage-line: AbacusByGroup::to_tsv and CoverageLine over the device-built CSR ------------

TABLE_SETS = [
    ("chrM_test.gfa", [], {}),
    ("chrM_test.gfa", ["-S"], {"groupby_sample": True}),
    ("chrM_test.gfa", ["-H", "-e", os.path.join(GOLDEN, "exclusion.bed3")],
     {"groupby_haplotype": True, "exclude": os.path.join(GOLDEN, "exclusion.bed3")}),
    ("chrM_test.gfa", ["-s", os.path.join(GOLDEN, "inclusion.bed3")], {"subset": os.path.join(GOLDEN, "inclusion.bed3")}),
    ("t_groups.gfa", ["-g", os.path.join(GOLDEN, "test_groups.txt")], {"groupby_file": os.path.join(GOLDEN, "test_groups.txt")}),
    ("cdbg.gfa", [], {}),
]


@pytest.mark.gpu
def test_table_matches_oracle(tmp_path):
    cmds, wants = [], []
    for gfa, flags, kw in TABLE_SETS:
        for count in ("node", "bp", "edge"):
            g, t, op, og, names = oracle_tables(gfa, count, **kw)
            r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
            for total in (False, True):
                try:
                    want = go.abacus_by_group_to_tsv(g, count, total, names, r, c, v, t.uncovered)
                except IndexError:
                    continue  # the reference panics on this input (v[j] out of bounds, abacus.rs:1166)
                cmds.append(["table", os.path.join(GOLDEN, gfa), "-c", count, *flags, *(["-a"] if total else [])])
                wants.append(want)
    outs = run_many(cmds, tmp_path)
    for cmd, out, want in zip(cmds, outs, wants):
        assert out.split("\n")[1].startswith("# version"), cmd
        assert body(out) == want, cmd
    # -O changes the column order of the table like it does for ordered-histgrowth
    order = tmp_path / "order.txt"
    g, t, op, og, names = oracle_tables("chrM_test.gfa", "node")
    order.write_text("\n".join(reversed(names)) + "\n")
    g, t, op, og, names2 = oracle_tables("chrM_test.gfa", "node", order=str(order))
    assert names2 == list(reversed(names))
    r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
    out = run_cli("table", os.path.join(GOLDEN, "chrM_test.gfa"), "-O", str(order)).stdout
    assert body(out) == go.abacus_by_group_to_tsv(g, "node", False, names2, r, c, v, t.uncovered)


TABLE_YAML = """- graph: {chrM}
  grouping: Haplotype
  analyses:
    - !Hist {{count_type: Bp}}
    - !CoverageLine {{count_type: Bp, reference: x}}
    - !Table {{count_type: Bp, total: false}}
    - !Table
      count_type: Node
      total: true
"""


@pytest.mark.gpu
def test_report_table_and_coverage_line(tmp_path):
    chrM = os.path.join(GOLDEN, "chrM_test.gfa")
    y = tmp_path / "t.yaml"
    y.write_text(TABLE_YAML.format(chrM=chrM))
    rep = run_cli("report", str(y)).stdout
    sections = [sec.split("\n", 1) for sec in rep.split("## run ")[1:]]
    assert [h for h, _ in sections] == [f"{chrM} analysis Hist", f"{chrM} analysis CoverageLine", f"{chrM} analysis Table",
                                        f"{chrM} analysis Table"]
    hist = oracle_hist("chrM_test.gfa", "bp", groupby_haplotype=True)
    assert body(sections[0][1]) == po.hist_table([("bp", hist)])
    assert body(sections[1][1]) == po.coverage_line_table([("bp", hist)])
    for sec, count, total in ((sections[2], "bp", False), (sections[3], "node", True)):
        g, t, op, og, names = oracle_tables("chrM_test.gfa", count, groupby_haplotype=True)
        r, c, v = po.csr_build(t.n_items, t.items, t.id_prefsum, op, og, t.exclude)
        assert body(sec[1]) == go.abacus_by_group_to_tsv(g, count, total, names, r, c, v, t.uncovered)


# ---- report: the YAML front end (commands/report.rs; tables as TSV, no HTML) ----------------------------------------

REPORT_YAML = """# example in the style of src/commands/report.rs:55-62
- graph: {chrM}
  name: chrM            # optional
  subset: ""
  grouping: Sample
  nice: false
  analyses:
    - !Hist
      count_type: Bp
    - !Growth
      coverage: 1,1,2
      quorum: 0,0.9,0
    - !OrderedGrowth {{coverage: "1,2", quorum: "0,0.5", order: null, count_type: Node}}
    - !Info
- graph: {tg}
  analyses:
    - !Hist {{count_type: Node}}
    - !Hist {{count_type: Edge}}
    - !Growth {{add_hist: true}}
"""


def _write_report(tmp_path):
    y = tmp_path / "report.yaml"
    y.write_text(REPORT_YAML.format(chrM=os.path.join(GOLDEN, "chrM_test.gfa"), tg=os.path.join(GOLDEN, "t_groups.gfa")))
    return str(y)


def test_report_dry_run(tmp_path):
    out = run_cli("report", _write_report(tmp_path), "--dry-run").stdout
    lines = out.strip().split("\n")
    assert lines[0] == "## run chrM analysis Hist" and "--count bp" in lines[1] and "--groupby-sample" in lines[1]
    assert "histgrowth" in lines[3] and "--coverage 1,1,2" in lines[3] and "--quorum 0,0.9,0" in lines[3] and "--count bp" in lines[3]
    assert "ordered-histgrowth" in lines[5] and "--count node" in lines[5] and "--order" not in lines[5]
    assert "not supported" in lines[7]
    # two different !Hist count types in one run -> every histogram of the run is built (graph_broker.rs:150-160)
    assert "--count all" in lines[9] and "--count all" in lines[11] and "--hist" in lines[13]


@pytest.mark.gpu
def test_report_runs_the_same_tables_as_the_subcommands(tmp_path):
    chrM, tg = os.path.join(GOLDEN, "chrM_test.gfa"), os.path.join(GOLDEN, "t_groups.gfa")
    rep = run_cli("report", _write_report(tmp_path)).stdout
    sections = [sec.split("\n", 1) for sec in rep.split("## run ")[1:]]
    assert [h for h, _ in sections] == ["chrM analysis Hist", "chrM analysis Growth", "chrM analysis OrderedGrowth",
                                        "chrM analysis Info", f"{tg} analysis Hist", f"{tg} analysis Hist",
                                        f"{tg} analysis Growth"]
    direct = run_many([["hist", chrM, "-S", "-c", "bp"], ["histgrowth", chrM, "-S", "-c", "bp", "-l", "1,1,2", "-q", "0,0.9,0"],
                       ["ordered-histgrowth", chrM, "-S", "-c", "node", "-l", "1,2", "-q", "0,0.5"],
                       ["hist", tg, "-c", "all"], ["histgrowth", tg, "-c", "all", "-a"]], tmp_path)
    assert body(sections[0][1]) == body(direct[0])
    assert body(sections[1][1]) == body(direct[1])
    assert body(sections[2][1]) == body(direct[2])
    assert body(sections[4][1]) == body(direct[3]) == body(sections[5][1])
    assert body(sections[6][1]) == body(direct[4])
