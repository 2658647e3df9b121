"""Locality-restoring node order, computed on the GPU with the path's own kernels.

The SpMM of the propagation path is bound by L2 misses of the gathered feature rows:
what it costs depends on how far apart, in memory, the endpoints of an edge are.  The order of
``enumerate(G)`` (graph_manipulation.py:19-21) is whatever the user happened to build; this module
finds a better INTERNAL order and the rest of the package permutes features in and results out, so
every externally visible index (``adj.indices``, node ids in tasks) stays exactly the reference's.

Algorithm (a minimum-linear-arrangement heuristic on a ring):
  1. spectral seed — block power iteration (3 vectors + deflation of the trivial eigenvector
     sqrt(deg)) on the normalised adjacency with ``gnntf_spmm_f32``; the angle of the two leading
     non-trivial Ritz vectors places every node on a circle (exact for graphs whose expected
     adjacency is circulant, a good coarse layout otherwise);
  2. robust refinement — ``gnntf_arrange_sweep_f32``: every node moves to the 1/(|offset|+eps)-
     weighted circular mean of its neighbours (iteratively re-weighted least absolute deviation, so
     the many near neighbours decide and the far ones are ignored), angles re-ranked between sweeps,
     eps annealed;
  3. the final ranking is the permutation.
On a randomly relabelled products-shaped graph this recovers the locality of the generator's
native order (median edge span 154 vs 156 positions at 1/50 scale).
"""
from __future__ import annotations

import math

import torch

from . import _native as nat


def _uniformize(theta):
    """Replace angles by their ranks spread evenly over [0, 2*pi)."""
    n = theta.numel()
    order = torch.argsort(theta)
    ranks = torch.empty(n, dtype=torch.float32, device=theta.device)
    ranks[order] = torch.arange(n, dtype=torch.float32, device=theta.device)
    return ranks * (2.0 * math.pi / n), order


def spectral_angles(A, power_iters=100, seed=0):
    """Angle of the two leading non-trivial eigenvectors of the symmetric-normalised adjacency."""
    from .ops import spmm_raw
    n, dev = A.base.n, A.val.device
    g = torch.Generator(device=dev).manual_seed(seed)
    X = torch.randn((n, 4), generator=g, device=dev, dtype=torch.float32)
    s = torch.sqrt(A.deg.clamp_min(0))
    s = s / s.norm().clamp_min(1e-30)
    struct = A.struct(4)
    Y = torch.empty_like(X)
    for it in range(power_iters):
        spmm_raw(struct, n, X, out=Y)
        X = 0.5 * (Y + X)                         # shift: eigenvalues into [0, 1]
        X -= s[:, None] * (s @ X)[None, :]        # deflate the trivial eigenvector
        if it % 5 == 4 or it == power_iters - 1:
            X, _ = torch.linalg.qr(X)
            X = X.contiguous()
    spmm_raw(struct, n, X, out=Y)
    T = X.T @ Y
    w, V = torch.linalg.eigh(0.5 * (T + T.T))
    R = X @ V.flip(1)                             # Ritz vectors, largest eigenvalue first
    return torch.atan2(R[:, 1], R[:, 0]) % (2.0 * math.pi)


def arrangement_order(adj, power_iters=100, sweeps=30, seed=0):
    """Permutation ``order`` (new position -> old node id) of a :class:`SparseAdjacency`."""
    L = nat.lib()
    n = adj.n
    if n < 3 or adj.csr.nnz == 0:
        return torch.arange(n, dtype=torch.int64, device=adj.csr.row_ptr.device)
    A = adj.normalized("symmetric")
    theta, order = _uniformize(spectral_angles(A, power_iters, seed))
    out = torch.empty_like(theta)
    eps, eps_min = 2.0 * math.pi / 50.0, 2.0 * (2.0 * math.pi / n)
    for _ in range(sweeps):
        nat.check(L.gnntf_arrange_sweep_f32(nat.ptr(adj.csr.row_ptr), nat.ptr(adj.csr.col_idx), nat.ptr(theta),
                                            eps, nat.ptr(out), n, nat.stream_ptr()), "arrange_sweep")
        theta, order = _uniformize(out)
        eps = max(eps * 0.8, eps_min)
    return order
