"""Hierarchical clustering of the similarity table (panacus_b200/host/cluster.cpp; kodama::linkage in the reference,
src/analyses/similarity.rs:165-181, a crate that is not in the reference tree).  The restatement of the published
algorithms (MST-linkage, NN-chain, generic linkage: Muellner, arXiv:1109.2378) is cross-checked on tie-free random
inputs against scipy.cluster.hierarchy.linkage, an independent implementation of the same algorithms with the same
labelling convention (cluster of sorted step i = n + i, smaller label first): identical merge structure and heights for
all seven methods in f64, identical structure in f32.  Kodama's own tie-breaking cannot be checked here: the row order
of the similarity TSV stays "parity unpinned"."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "panacus_b200", "bin", "panacus")
METHODS = ["single", "complete", "average", "weighted", "ward", "centroid", "median"]


def run_linkage(cond, n, method, tmp_path, f32=False):
    f = tmp_path / f"cond_{method}_{n}.txt"
    f.write_text(f"{n}\n" + " ".join(repr(float(x)) for x in cond) + "\n")
    args = [CLI, "debug-linkage", str(f), "-m", method] + (["--f32=1"] if f32 else [])
    out = subprocess.run(args, capture_output=True, text=True, check=True).stdout
    rows = [line.split() for line in out.strip().splitlines()]
    return [(int(a), int(b), float(d)) for a, b, d in rows]


@pytest.mark.parametrize("method", METHODS)
def test_linkage_matches_scipy_on_tie_free_inputs(method, tmp_path):
    hierarchy = pytest.importorskip("scipy.cluster.hierarchy")
    distance = pytest.importorskip("scipy.spatial.distance")
    rng = np.random.default_rng(hash(method) % 1000)
    for n in (2, 3, 5, 17, 64, 150):
        pts = rng.random((n, 6))
        cond = distance.pdist(pts)  # euclidean, like the reference's calculate_distances
        Z = hierarchy.linkage(cond, method=method)
        got = run_linkage(cond, n, method, tmp_path)
        assert len(got) == n - 1
        assert [(a, b) for a, b, _ in got] == [(int(r[0]), int(r[1])) for r in Z], (method, n)
        assert np.allclose([d for _, _, d in got], Z[:, 2], rtol=1e-9, atol=1e-12), (method, n)
        got32 = run_linkage(cond.astype(np.float32), n, method, tmp_path, f32=True)
        if n <= 17:  # (f32 rounding may reorder near-equal heights on the larger inputs)
            assert [(a, b) for a, b, _ in got32] == [(a, b) for a, b, _ in got], (method, n)


def test_leaf_order_and_degenerate_inputs(tmp_path):
    # one observation: no steps; two: one step naming both
    assert run_linkage([], 1, "centroid", tmp_path) == []
    assert run_linkage([0.25], 2, "centroid", tmp_path) == [(0, 1, 0.25)]
    # all distances equal: ties everywhere -- the dendrogram must still be a valid one (every label used once)
    n = 9
    for method in METHODS:
        steps = run_linkage([1.0] * (n * (n - 1) // 2), n, method, tmp_path)
        used = sorted(x for a, b, _ in steps for x in (a, b))
        assert used == list(range(2 * n - 2)), method
        assert all(a < b for a, b, _ in steps), method
