#!/usr/bin/env python3
"""Regenerates tests/golden/ from the read-only reference checkout (run in the build container).

  * copies the reference's small GFA / BED / group fixtures (test data, not source code);
  * extracts the chr22 HPRC pggb hist + growth arrays embedded in the example report
    docs/chr22.hprc-v1.0-pggb.histgrowth.html:267-274 (panacus v0.2.2 output) into
    chr22_histgrowth.json;
  * records the known-answer vectors of the reference's own unit tests
    (src/graph_broker/hist.rs:341-398, src/graph_broker/abacus.rs:1498-1630,
    tests/test_files/t_groups.hist.tsv) in kats.json.
"""
import json
import os
import re
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

for rel in ["test/chrM_test.gfa", "test/cdbg.gfa", "test/test_groups.txt",
            "tests/test_files/t_groups.gfa", "tests/test_files/t_groups.hist.tsv",
            "test/bed_chrM/exclusion.bed3", "test/bed_chrM/inclusion.bed1",
            "test/bed_chrM/inclusion.bed3", "test/bed_chrM/inclusion_chm13.bed1",
            "test/bed_chrM/inclusion_sub.bed1"]:
    dst = os.path.join(OUT, os.path.basename(rel))
    shutil.copyfile(os.path.join(REF, rel), dst)
    os.chmod(dst, 0o644)

html = open(os.path.join(REF, "docs/chr22.hprc-v1.0-pggb.histgrowth.html")).read()
hists = {}
for m in re.finditer(r"new Hist\('(\w+)', (\[[^\]]*\]), (\[[^\]]*\])\)", html):
    hists[m.group(1)] = {"index": json.loads(m.group(2)), "values": json.loads(m.group(3))}
growths = {}
for m in re.finditer(r"new Growth\('(\w+)', (\[[^\]]*\]), (\[[^\]]*\]), (\[[^\]]*\]), (\[\[.*?\]\])\)", html):
    growths[m.group(1)] = {"index": json.loads(m.group(2)), "coverage": json.loads(m.group(3)),
                           "quorum": json.loads(m.group(4)), "curves": json.loads(m.group(5))}
assert set(hists) == {"bp", "node", "edge"} and set(growths) == set(hists), (hists.keys(), growths.keys())
json.dump({"source": "docs/chr22.hprc-v1.0-pggb.histgrowth.html:267-274", "hist": hists, "growth": growths},
          open(os.path.join(OUT, "chr22_histgrowth.json"), "w"))

kats = {
    "source": "src/graph_broker/hist.rs:341-398; src/graph_broker/abacus.rs:1498-1630; tests/test_files/t_groups.hist.tsv",
    "union": {"hist": [0, 5, 3, 2], "coverage": 0, "expect": [5.666666666666667, 8.333333333333334, 10.0]},
    "core": {"hist": [0, 5, 3, 2], "coverage": 0, "expect": [5.666666666666666, 3.0, 2.0]},
    "quorum": {"hist": [0, 5, 3, 2, 3, 5, 0, 4, 2, 1], "coverage": 0, "quorum": 0.9,
               "expect": [11.88888888888889, 7.027777777777777, 4.761904761904761, 3.4444444444444438,
                          2.5952380952380953, 2.0, 1.5555555555555545, 1.2222222222222217, 1.0]},
    "chrM_groupby_sample": {"groups": ["chm13", "grch38", "HG00438", "HG00621"],
                            "node": [0, 39, 29, 41, 45], "edge": [0, 80, 59, 66, 0],
                            "bp": [0, 616, 31, 601, 15949]},
    "t_groups_node_hist": [5, 0, 10, 0, 0, 0, 0],
}
# full per-item coverage vectors of the (commented-out, still valid) chrM tests, abacus.rs:1480-1633
src = open(os.path.join(REF, "src/graph_broker/abacus.rs")).read()
for m in re.finditer(r"fn test_abacus_by_total_from_chr_m_(\w+)\(\)(.*?)countable: vec!\[(.*?)\],\n", src, re.S):
    body = re.sub(r"//", " ", m.group(3))
    vals = [int(x) for x in re.findall(r"\b\d+\b", body.replace("CountSize::MAX", ""))]
    kats["chrM_groupby_sample"]["countable_" + m.group(1)] = vals  # items 1..N (sentinel [0] dropped)
assert {"countable_node", "countable_edge", "countable_bp"} <= set(kats["chrM_groupby_sample"]), kats["chrM_groupby_sample"].keys()
json.dump(kats, open(os.path.join(OUT, "kats.json"), "w"), indent=1)
print("golden fixtures written to", OUT)
