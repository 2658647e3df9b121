"""GPU parity at the FULL sizes of BASELINE.json's configs, against the C oracle (oracle_c.c through
oracle/oracle_big.py: the reference's op sequence, row-parallel over the stable CSR — bit-identical
to the sequential COO loop, see tests/test_oracle.py).

  configs[1]  GCN 2-layer, PubMed shape (19,717 x 500 -> 64 -> 3): eval forward, masked training
              forward, backward                                                    gcn.py:87-89,108-113
  configs[2]  APPNP K=10, arxiv shape, F = 128 and F = 40 (class width)            filter.py:17-22
  configs[3]  APPNP, products shape (nnz = 123,718,280), F = 100: builder bit-exact, one fused
              step over ALL rows, and the K = 10 propagation
  configs[4]  R-MAT scale 20 with hub rows far above the split threshold

The comparisons use the oracle's own CSR and its own normalised values (never the GPU's).
"""
import numpy as np
import pytest
import torch

import gnntf_oracle as oracle
import oracle_big
import synthetic

pytestmark = pytest.mark.gpu


def _gnntf():
    import gnntf
    return gnntf


def _np(t):
    return t.detach().cpu().numpy()


def _unsplit_mask(adj):
    m = np.ones(adj.n, bool)
    m[_np(adj.csr.long_row)] = False
    return m


def _check_builder_bit_exact(adj, big):
    assert adj.csr.nnz == big.nnz
    assert np.array_equal(_np(adj.csr.row_ptr).astype(np.int64), big.row_ptr), "row_ptr"
    assert np.array_equal(_np(adj.csr.col_idx), big.col), "col_idx"
    assert np.array_equal(_np(adj.csr.coo_pos).astype(np.int64), big.coo_pos), "coo_pos"


# ------------------------------------------------------------------------------------------
# configs[2]: arxiv shape
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F", [128, 40])
def test_arxiv_full_size_k10_vs_c_oracle(F):
    gnntf = _gnntf()
    n, edges = synthetic.shaped_edges("arxiv", seed=0, device="cuda")
    adj = gnntf.edges2adj(edges, None, n)
    assert adj.csr.nnz == 2 * synthetic.SHAPES["arxiv"][1] == 2332486
    big = oracle_big.BigOracle(_np(edges), None, n)
    _check_builder_bit_exact(adj, big)
    A = adj.normalized("symmetric")
    assert np.array_equal(_np(A.val), big.val), "normalised values (CSR order) differ from the oracle's bits"
    H0 = synthetic.features(n, F, seed=1, device="cuda")
    H0h = _np(H0)
    # one fused step: rows that are not split are bit-identical
    unsplit = _unsplit_mask(adj)
    step = _np(gnntf.appnp_step(A, H0, H0, 0.1))
    expect = big.step(H0h, H0h, 0.1)
    assert np.array_equal(step[unsplit], expect[unsplit])
    oracle_big.check_against(big, step, expect, big.step(H0h, H0h, 0.1, acc64=True), ~unsplit,
                             f"arxiv F={F} one step ({int((~unsplit).sum())} split rows)")
    # the K = 10 propagation
    out = _np(gnntf.appnp_propagate(A, H0, 0.1, 10))
    oracle_big.check_against(big, out, big.propagate(H0h, 0.1, 10), big.propagate(H0h, 0.1, 10, acc64=True), ~unsplit,
                             f"arxiv F={F} K=10")


# ------------------------------------------------------------------------------------------
# configs[3]: products shape, full size
# ------------------------------------------------------------------------------------------
def test_products_full_size_builder_step_and_k10_vs_c_oracle():
    gnntf = _gnntf()
    n, edges = synthetic.shaped_edges("products", seed=0, device="cuda")
    assert n == 2449029 and edges.shape[0] == 61859140
    adj = gnntf.edges2adj(edges, None, n)
    edges_h = _np(edges)
    del edges
    big = oracle_big.BigOracle(edges_h, None, n)
    del edges_h
    assert big.nnz == 123718280
    _check_builder_bit_exact(adj, big)                      # indices, nnz, order at nnz = 123.7 M
    A = adj.normalized("symmetric")
    assert np.array_equal(_np(A.val), big.val)
    assert np.array_equal(_np(A.dinv), big.D)
    F = 100
    H0 = synthetic.features(n, F, seed=1, device="cuda")
    H = synthetic.features(n, F, seed=2, device="cuda")
    H0h, Hh = _np(H0), _np(H)
    unsplit = _unsplit_mask(adj)
    assert (~unsplit).sum() > 1000                            # thousands of split hub rows
    step = _np(gnntf.appnp_step(A, H, H0, 0.1))              # every row of one fused step
    expect = big.step(Hh, H0h, 0.1)
    assert np.array_equal(step[unsplit], expect[unsplit]), "un-split rows must match the oracle bit for bit"
    oracle_big.check_against(big, step, expect, big.step(Hh, H0h, 0.1, acc64=True), ~unsplit, "products F=100 one step")
    del step, expect, H, Hh
    out = _np(gnntf.appnp_propagate(A, H0, 0.1, 10))
    oracle_big.check_against(big, out, big.propagate(H0h, 0.1, 10), big.propagate(H0h, 0.1, 10, acc64=True), ~unsplit,
                             "products F=100 K=10")


# ------------------------------------------------------------------------------------------
# configs[4]: R-MAT with hub rows
# ------------------------------------------------------------------------------------------
def test_rmat_scale20_hub_rows_vs_c_oracle():
    gnntf = _gnntf()
    n, edges = synthetic.rmat_edges(20, 8_000_000, seed=0, device="cuda")
    adj = gnntf.edges2adj(edges, None, n)
    big = oracle_big.BigOracle(_np(edges), None, n)
    _check_builder_bit_exact(adj, big)
    deg = np.diff(big.row_ptr)
    assert deg.max() > 20000 and adj.csr.n_long > 1000, (deg.max(), adj.csr.n_long)   # hub rows >> 256
    A = adj.normalized("symmetric")
    assert np.array_equal(_np(A.val), big.val)
    unsplit = _unsplit_mask(adj)
    for F in (16, 64):
        H = synthetic.features(n, F, seed=3, device="cuda")
        got = _np(gnntf.sparse_dense_matmul(A, H))
        expect = big.spmm(_np(H))
        assert np.array_equal(got[unsplit], expect[unsplit])
        oracle_big.check_against(big, got, expect, big.spmm(_np(H), acc64=True), ~unsplit, f"R-MAT scale 20 SpMM F={F}")


# ------------------------------------------------------------------------------------------
# configs[1]: GCN 2-layer on the PubMed shape
# ------------------------------------------------------------------------------------------
def _pubmed():
    gnntf = _gnntf()
    n, e, width, classes = synthetic.SHAPES["pubmed"]
    G = synthetic.citation_graph(n, e, seed=0)
    X = synthetic.citation_features(n, width, seed=1)
    gnntf.set_seed(0)
    adj = gnntf.graph2adj(G)
    arch = gnntf.GCN(adj, X, num_classes=classes)
    arch.reset()
    idx, val, _ = oracle.graph2adj(G)
    assert np.array_equal(_np(adj.indices), idx)
    return gnntf, arch, adj, X, idx, val, n


def test_pubmed_gcn_eval_forward_vs_oracle():
    """The reference aggregates first ((ÂX)W, gcn.py:88-89); the layer here evaluates Â(XW) when it
    narrows (500 -> 64).  Both orders are compared with the oracle's (ÂX)W."""
    gnntf, arch, adj, X, idx, val, n = _pubmed()
    arch.training_mode(False)
    out = _np(arch(arch.features))
    Ws = [w.numpy() for w in arch.vars()][0::2]
    bs = [w.numpy() for w in arch.vars()][1::2]
    assert Ws[0].shape == (500, 64) and Ws[1].shape == (64, 3)
    expect = oracle.gcn_forward(idx, val, n, X, Ws, bs)
    oracle.assert_close(out, expect, what="PubMed GCN eval forward", floor=0.05)
    expect64 = oracle.gcn_forward(idx, val, n, X, Ws, bs, dtype=np.float64)
    oracle.assert_close(out, expect64, what="PubMed GCN eval forward vs fp64", floor=0.05)


def test_pubmed_gcn_training_forward_and_backward_vs_oracle():
    """Training mode with injected masks (TF's Philox stream cannot be reproduced): per-layer edge
    keep-masks in COO order (layered.py:47-50, gcn.py:88) and the hidden layer's feature dropout
    mask (gcn.py:89); forward value and the gradients of every variable against the oracle's VJP."""
    gnntf, arch, adj, X, idx, val, n = _pubmed()
    from gnntf.gnn import MaskedAdjacency
    rng = np.random.default_rng(7)
    nnz = idx.shape[0]
    edge_keep = rng.random(nnz) >= 0.5                     # hidden layer: graph_dropout = 0.5 (gcn.py:112)
    feat_keep = rng.random((n, 64)) >= 0.5                 # hidden layer: dropout = 0.5
    g_out = rng.standard_normal((n, 3)).astype(np.float32)
    edge_keep_t = torch.from_numpy(edge_keep).cuda()
    feat_keep_t = torch.from_numpy(feat_keep).cuda()
    arch.training_mode(True)
    arch.sparse_dropout = lambda G, p=0.5: G if p == 0 else MaskedAdjacency(G, edge_keep_t, float(p))
    arch.dropout_mask = lambda shape, p, device=None: None if p == 0 else feat_keep_t     # every feature dropout of the stack
    out = arch(arch.features)
    out.backward(torch.from_numpy(g_out).cuda())
    ws = arch.vars()
    W1, b1, W2, b2 = (w.numpy() for w in ws)
    # oracle forward, fp32 in the reference's order: (Â_k X) W + b, relu, dropout
    t = np.float32
    v1 = oracle.sparse_dropout(val, 0.5, edge_keep)
    _, nv1, _ = oracle.get_adjacency(idx, v1, n)
    _, nv2, _ = oracle.get_adjacency(idx, val, n)          # output layer: graph_dropout = 0 (gcn.py:113,79)
    AX = oracle.spmm_coo(idx, nv1, X.astype(t))
    Z1 = AX @ W1 + b1
    H1 = np.where(feat_keep, np.maximum(Z1, 0) * oracle.dropout_scale(0.5), t(0)).astype(t)
    AH1 = oracle.spmm_coo(idx, nv2, H1)
    Z2 = AH1 @ W2 + b2
    expect = np.maximum(Z2, 0)
    oracle.assert_close(_np(out), expect, what="PubMed GCN training forward", floor=0.05)
    # oracle backward (fp64 accumulation of the same graph)
    d = np.float64
    gZ2 = g_out.astype(d) * (Z2 > 0)
    gW2 = AH1.astype(d).T @ gZ2
    gb2 = gZ2.sum(0, keepdims=True)
    gAH1 = gZ2 @ W2.astype(d).T
    gH1 = oracle.spmm_coo_T(idx, nv2, gAH1, dtype=d)
    gZ1 = gH1 * feat_keep * float(oracle.dropout_scale(0.5)) * (Z1 > 0)
    gW1 = AX.astype(d).T @ gZ1
    gb1 = gZ1.sum(0, keepdims=True)
    for name, w, ref in (("dW2", ws[3 - 1], gW2), ("db2", ws[3], gb2), ("dW1", ws[0], gW1), ("db1", ws[1], gb1)):
        oracle.assert_close(_np(w.var.grad), ref, what=f"PubMed GCN {name}", floor=0.05)


# ------------------------------------------------------------------------------------------
# Training-mode propagation at products scale: memory
# ------------------------------------------------------------------------------------------
def test_products_training_forward_backward_memory():
    """APPNP's K = 10 training propagation at the products shape and class width (47 -> 48 columns):
    one edge mask per iteration (filter.py:18), forward + backward, WITHOUT K materialised adjacencies.
    Round 1 kept val + val_T per iteration (10 x 0.99 GB + masks); now only the masks survive."""
    gnntf = _gnntf()
    n, edges = synthetic.shaped_edges("products", seed=0, device="cuda")
    adj = gnntf.edges2adj(edges, None, n)
    del edges
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    base = torch.cuda.memory_allocated()
    torch.cuda.reset_peak_memory_stats()
    F, K = 48, 10
    g = torch.Generator(device="cuda").manual_seed(3)
    masks = [torch.rand(adj.n_graph, device="cuda", generator=g) >= 0.5 for _ in range(K)]
    H0 = synthetic.features(n, F, seed=1, device="cuda").requires_grad_(True)
    out = gnntf.ops.appnp_propagate_masked(adj, masks, 0.5, H0, 0.1)
    out.backward(torch.ones_like(out))
    torch.cuda.synchronize()
    peak_total = torch.cuda.max_memory_allocated() / 2**30
    peak_extra = (torch.cuda.max_memory_allocated() - base) / 2**30
    print(f"products training K=10 F={F}: adjacency {base / 2**30:.2f} GiB, peak total {peak_total:.2f} GiB, propagation {peak_extra:.2f} GiB")
    assert peak_total < 12.0, (peak_total, peak_extra)
    assert torch.isfinite(H0.grad).all()
    # the backward of a linear map: <out, 1> differentiated w.r.t. H0 is the transposed operator applied to 1;
    # check the adjoint identity <P(H0), G> = <H0, P^T(G)> with G = 1 on the full-size graph
    lhs = out.double().sum().item()
    rhs = (H0.detach().double() * H0.grad.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0), (lhs, rhs)
