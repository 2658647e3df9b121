"""Deterministic synthetic inputs of the shapes BASELINE.json names (there is no network for the
real datasets; the reference loads them through DGL, experiments/experiment_setup.py:153-181).

* :func:`citation_graph`  — Cora-/PubMed-shaped: E/2 unique undirected pairs, no self loops,
  inserted in BOTH directions into an ``nx.DiGraph`` in node order (how DGL citation graphs arrive,
  experiment_setup.py:173-178), so the real networkx ``graph2adj`` walk is exercised and the
  symmetrised list holds every entry twice (nnz = 2E with duplicates).
* :func:`powerlaw_edges`  — arxiv-/products-shaped array-native edge lists: power-law source
  degrees (weight ∝ (rank+r0)^-a, r0 solved for the target maximum degree) and multi-scale
  locality (log-uniform |u−v|, i.e. equal edge mass at every distance scale, the Kleinberg /
  hierarchical-community model).  ``ordering="local"`` keeps that node order;
  ``ordering="random"`` applies a random relabelling (the worst case for gather locality).
* :func:`rmat_edges`      — R-MAT (a,b,c,d = .57,.19,.19,.05) for the SpMM sweep.
All generators are torch-based and run on CPU or GPU from an explicit seed.
"""
from __future__ import annotations

import math

import numpy as np
import torch

SHAPES = {
    # name: (nodes, edges as G.edges() lists them, feature width, classes)
    "cora": (2708, 10556, 1433, 7),
    "pubmed": (19717, 88648, 500, 3),
    "arxiv": (169343, 1166243, 128, 40),
    "products": (2449029, 61859140, 100, 47),
}
POWERLAW = {"arxiv": dict(a=0.8, max_degree=13000), "products": dict(a=0.5, max_degree=17000)}


def citation_graph(n, n_edges, seed=0):
    """networkx DiGraph with ``n_edges`` directed edges = ``n_edges/2`` distinct undirected pairs
    in both directions, nodes added in id order, edges added per source node in node order."""
    import networkx as nx
    rng = np.random.default_rng(seed)
    need = n_edges // 2
    pairs = set()
    while len(pairs) < need:
        u = rng.integers(0, n, size=2 * (need - len(pairs)) + 16)
        v = rng.integers(0, n, size=u.size)
        for a, b in zip(u.tolist(), v.tolist()):
            if a != b:
                pairs.add((min(a, b), max(a, b)))
                if len(pairs) == need:
                    break
    und = np.array(sorted(pairs), dtype=np.int64)
    both = np.concatenate([und, und[:, ::-1]], axis=0)
    both = both[np.lexsort((both[:, 1], both[:, 0]))]
    G = nx.DiGraph()
    G.add_nodes_from(range(n))
    G.add_edges_from(both.tolist())
    return G


def citation_features(n, width, seed=1):
    """Row-normalised sparse-ish Bernoulli(0.01) bag-of-words features, fp32."""
    rng = np.random.default_rng(seed)
    X = (rng.random((n, width)) < 0.01).astype(np.float32)
    X[np.arange(n), rng.integers(0, width, size=n)] = 1.0  # no empty rows
    return X / X.sum(axis=1, keepdims=True)


def _solve_r0(n, n_edges, a, max_out_degree):
    """r0 such that the heaviest source expects ``max_out_degree`` of the ``n_edges`` draws."""
    ranks = np.arange(n, dtype=np.float64)
    lo, hi = 1e-3, 1e6
    for _ in range(80):
        r0 = math.sqrt(lo * hi)
        w = (ranks + r0) ** (-a)
        top = n_edges * w[0] / w.sum()
        if top > max_out_degree:
            lo = r0
        else:
            hi = r0
    return math.sqrt(lo * hi)


def powerlaw_edges(n, n_edges, seed=0, a=0.5, max_degree=17000, ordering="local", device="cpu"):
    """int64 [n_edges, 2] edge list (the ``graph2indices`` form); no self loops, duplicates kept."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    # the symmetrised degree of the top node ≈ its out-degree + ~mean in-degree
    r0 = _solve_r0(n, n_edges, a, max(1.0, max_degree - n_edges / n))
    w = (torch.arange(n, dtype=torch.float64, device=dev) + r0) ** (-a)
    w = w[torch.randperm(n, generator=g, device=dev)]  # hubs anywhere in the id space
    cdf = torch.cumsum(w, 0)
    cdf = cdf / cdf[-1]
    u = torch.searchsorted(cdf, torch.rand(n_edges, generator=g, device=dev, dtype=torch.float64)).clamp_(max=n - 1)
    half = max(2, n // 2)
    d = torch.exp(torch.rand(n_edges, generator=g, device=dev, dtype=torch.float64) * math.log(half)).long().clamp_(1, half - 1 if half > 2 else 1)
    sign = torch.randint(0, 2, (n_edges,), generator=g, device=dev) * 2 - 1
    v = torch.remainder(u + sign * d, n)
    if n > 1:
        same = v == u
        v = torch.where(same, torch.remainder(u + 1, n), v)
    edges = torch.stack([u, v], dim=1)
    if ordering == "random":
        relabel = torch.randperm(n, generator=g, device=dev)
        edges = relabel[edges]
    elif ordering != "local":
        raise ValueError("ordering must be 'local' or 'random'")
    return edges.contiguous()


def shaped_edges(name, seed=0, ordering="local", device="cpu", scale=1.0):
    """Edge list of a named BASELINE shape (optionally scaled down by ``scale`` for tests)."""
    n, e, _, _ = SHAPES[name]
    n, e = max(2, int(n * scale)), max(1, int(e * scale))
    p = POWERLAW[name]
    return n, powerlaw_edges(n, e, seed, a=p["a"], max_degree=max(4, int(p["max_degree"] * min(1.0, scale * 4))),
                             ordering=ordering, device=device)


def rmat_edges(scale, n_edges, seed=0, abcd=(0.57, 0.19, 0.19, 0.05), device="cpu", noise=0.1):
    """R-MAT with per-level parameter noise; returns (n = 2**scale, int64 [n_edges,2])."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    u = torch.zeros(n_edges, dtype=torch.int64, device=dev)
    v = torch.zeros(n_edges, dtype=torch.int64, device=dev)
    a, b, c, d = abcd
    for level in range(scale):
        jitter = 1.0 + noise * (torch.rand(4, generator=g, device=dev) * 2 - 1)
        pa, pb, pc, pd = (torch.tensor([a, b, c, d], device=dev) * jitter).tolist()
        s = pa + pb + pc + pd
        pa, pb, pc = pa / s, pb / s, pc / s
        r = torch.rand(n_edges, generator=g, device=dev)
        bit_u = r >= pa + pb
        bit_v = ((r >= pa) & (r < pa + pb)) | (r >= pa + pb + pc)
        u = (u << 1) | bit_u.long()
        v = (v << 1) | bit_v.long()
    return 1 << scale, torch.stack([u, v], dim=1).contiguous()


def features(n, width, seed=1, device="cpu"):
    g = torch.Generator(device=torch.device(device)).manual_seed(seed)
    return torch.randn((n, width), generator=g, device=device, dtype=torch.float32)
