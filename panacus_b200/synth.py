"""Deterministic synthetic pangenome incidence tables for tests and bench.py.

Item classes mimic a pangenome's U-shaped coverage: 35 % "core" (each group contains the item with
p = 0.97), 45 % "rare" (present in 1 + Geom(0.5) uniformly chosen groups), 20 % "shell"
(p ~ U(0.05, 0.95) per item).  Item lengths (bp weights): 50 % length 1, otherwise
1 + floor(exp(N(2.0, 1.5))) capped at 1e5.  Row 0 is the reference's dummy item (all zero).

``numpy_table`` builds small tables on the host; ``torch_table`` builds large ones directly in GPU
memory (chunked) in the device layout of include/panacus_b200.h.
"""
from __future__ import annotations

import numpy as np

from .abacus import pack_bits, row_words

SEED_BASE = 0x5EED0000


def numpy_table(n_items: int, n_groups: int, seed: int = SEED_BASE, variant: str = "ushape"):
    """-> (bits uint8 [(N+1), G], bitmap u64 [(N+1), row_words(G)], weight u32 [N+1])"""
    rng = np.random.default_rng(seed)
    N, G = n_items, n_groups
    bits = np.zeros((N + 1, G), dtype=np.uint8)
    if variant == "dense":
        bits[1:] = rng.random((N, G)) < 0.5
    elif variant == "sparse":
        bits[1:] = rng.random((N, G)) < 0.02
    else:
        cls = rng.random(N)
        core = cls < 0.35
        rare = (cls >= 0.35) & (cls < 0.80)
        p = np.where(core, 0.97, rng.uniform(0.05, 0.95, N))
        b = rng.random((N, G)) < p[:, None]
        k = np.minimum(rng.geometric(0.5, N), G)
        rr = np.zeros((N, G), dtype=bool)
        idx = np.nonzero(rare)[0]
        for i in idx:
            rr[i, rng.choice(G, size=k[i], replace=False)] = True
        b[rare] = rr[rare]
        bits[1:] = b
    weight = np.ones(N + 1, dtype=np.uint32)
    long_ = rng.random(N) >= 0.5
    ln = 1 + np.floor(np.exp(rng.normal(2.0, 1.5, N))).astype(np.int64)
    weight[1:] = np.where(long_, np.minimum(ln, 100000), 1).astype(np.uint32)
    weight[0] = 0
    return bits, pack_bits(bits), weight


def torch_table(n_items: int, n_groups: int, seed: int = SEED_BASE, device="cuda", chunk_items: int = 1 << 18):
    """Large table generated on `device`: -> (bitmap int64 tensor [(N+1), row_words(G)] (u64 bit patterns),
    weight int32 tensor [N+1] (u32 bit patterns)).  Same class mix as numpy_table (not the same bits)."""
    import torch

    N, G = n_items, n_groups
    Wp = row_words(G)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    bitmap = torch.zeros((N + 1, Wp), dtype=torch.int64, device=device)
    weight = torch.ones(N + 1, dtype=torch.int32, device=device)
    # bit weights of a 64-bit word, as int64 bit patterns (bit 63 = sign bit)
    shifts = torch.arange(64, device=device, dtype=torch.int64)
    for start in range(1, N + 1, chunk_items):
        n = min(chunk_items, N + 1 - start)
        cls = torch.rand(n, generator=gen, device=device)
        core = cls < 0.35
        rare = (cls >= 0.35) & (cls < 0.80)
        p = torch.where(core, torch.full_like(cls, 0.97),
                        0.05 + 0.90 * torch.rand(n, generator=gen, device=device))
        b = torch.rand((n, G), generator=gen, device=device) < p[:, None]
        # rare items: k = 1 + Geom(0.5) groups at uniform positions (collisions allowed, >= 1 group)
        k = torch.clamp(1 + torch.floor(-torch.log2(torch.rand(n, generator=gen, device=device).clamp_min(1e-12))).long(),
                        max=min(G, 8))
        pos = torch.randint(0, G, (n, 8), generator=gen, device=device)
        sel = torch.arange(8, device=device)[None, :] < k[:, None]
        rr = torch.zeros((n, G), dtype=torch.bool, device=device)
        rows = torch.arange(n, device=device)[:, None].expand(n, 8)
        rr[rows[sel], pos[sel]] = True
        b = torch.where(rare[:, None], rr, b)
        if G < Wp * 64:
            b = torch.nn.functional.pad(b, (0, Wp * 64 - G))
        words = (b.view(n, Wp, 64).to(torch.int64) << shifts).sum(dim=2)  # disjoint bits: sum == OR
        bitmap[start:start + n] = words
        ln = 1 + torch.floor(torch.exp(2.0 + 1.5 * torch.randn(n, generator=gen, device=device))).clamp(max=99999)
        is_long = torch.rand(n, generator=gen, device=device) >= 0.5
        weight[start:start + n] = torch.where(is_long, ln, torch.ones_like(ln)).to(torch.int32)
    weight[0] = 0
    return bitmap, weight


def random_orders(n_orders: int, n_groups: int, seed: int = SEED_BASE) -> np.ndarray:
    """Fisher-Yates permutations; permutation p is seeded seed + p."""
    out = np.zeros((n_orders, n_groups), dtype=np.uint32)
    for p in range(n_orders):
        out[p] = np.random.default_rng(seed + p).permutation(n_groups).astype(np.uint32)
    return out
