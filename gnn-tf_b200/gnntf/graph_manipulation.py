"""``graph2adj`` and friends (gnntf/core/gnn/graph_manipulation.py:5-31).

The networkx walk that fixes node ids and edge order is necessarily host Python (it defines the
reference's index order: ``enumerate(G)`` ids, ``G.edges()`` order); everything after it —
symmetrisation, value duplication, the COO tensors and the CSR view — is the GPU builder
(``gnntf_csr_build``).  :func:`edges2adj` is the array-native entry for graphs networkx cannot
hold (ogbn-products-scale edge lists).
"""
from __future__ import annotations

import numpy as np
import torch

from .sparse import SparseAdjacency


def create_nx_graph(nodes, edges):
    """graph_manipulation.py:5-12."""
    import networkx as nx
    graph = nx.DiGraph()
    if nodes is not None:
        graph.add_nodes_from(nodes)
    for u, v in edges:
        graph.add_edge(u, v)
    return graph


def adj2graph(nodes, adj):
    """graph_manipulation.py:15-16."""
    return create_nx_graph(nodes, adj.indices.cpu().numpy())


def graph2indices(G):
    """graph_manipulation.py:19-21 — ids follow ``enumerate(G)``, pairs follow ``G.edges()``."""
    node2id = {u: idx for idx, u in enumerate(G)}
    return [[node2id[u], node2id[v]] for u, v in G.edges()]


def edges2adj(edges, weights=None, num_nodes=None, directed=False):
    """Array-native ``graph2adj``: ``edges`` is the ``graph2indices`` list as an [E,2] integer
    array/tensor, ``weights`` the per-edge ``weight`` attribute (``None`` = 1.)."""
    if not torch.cuda.is_available():
        raise RuntimeError("gnntf_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda")
    if not isinstance(edges, torch.Tensor):
        edges = torch.as_tensor(np.asarray(edges, dtype=np.int64).reshape(-1, 2))
    edges = edges.to(device=dev, dtype=torch.int64).reshape(-1, 2).contiguous()
    if weights is not None:
        if not isinstance(weights, torch.Tensor):
            weights = torch.as_tensor(np.asarray(weights, dtype=np.float32))
        weights = weights.to(device=dev, dtype=torch.float32).contiguous()
        if weights.numel() != edges.shape[0]:
            raise Exception("one weight per edge is required")
    if num_nodes is None:
        num_nodes = int(edges.max().item()) + 1 if edges.numel() else 0
    if edges.numel():
        lo, hi = int(edges.min().item()), int(edges.max().item())
        if lo < 0 or hi >= num_nodes:
            raise Exception(f"edge endpoint out of range [0, {num_nodes}): min {lo}, max {hi}")
    return SparseAdjacency(edges, weights, num_nodes, directed)


def csr2adj(indptr, indices, data=None, directed=True):
    """Adjacency from CSR arrays (the ``.npz`` format of the reference's loaders,
    experiments/experiment_setup.py:273-282: ``adj_indptr``/``adj_indices``/``adj_data``).  The CSR
    rows are expanded to the edge list ``[[row, col], ...]`` in CSR order, then handled exactly like
    ``graph2adj``'s list.  ``directed=True`` (default) takes the matrix as stored; ``False`` appends the
    reversed list as ``graph2adj`` does."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    n = indptr.shape[0] - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    edges = np.stack([rows, indices], axis=1)
    weights = None if data is None else np.asarray(data, dtype=np.float32)
    return edges2adj(edges, weights, n, directed)


def graph2adj(G, directed=False):
    """graph_manipulation.py:24-31.  Returns an object with the SparseTensor fields the reference
    exposes (``indices`` int64 [nnz,2] in the reference order, ``values``, ``dense_shape``/``shape``)."""
    indices = graph2indices(G)
    values = [edge[2].get("weight", 1.) for edge in G.edges(data=True)]          # :27
    uniform = all(v == 1. for v in values)
    return edges2adj(np.asarray(indices, dtype=np.int64).reshape(-1, 2),
                     None if uniform else np.asarray(values, dtype=np.float32), len(G), directed)
