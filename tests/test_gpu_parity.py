"""GPU parity: the CUDA path (through the C-ABI / ctypes shim) against the CPU oracle on the same
seeded inputs.  Integer / index work is bit-exact; fp32 work is held to the north_star tolerance
of 1e-5 relative (oracle.assert_close: |x−y| ≤ 1e-5·max(|y|, floor·‖y‖∞), floor stated per call
where it is not the default).  Rows that are not split accumulate in the oracle's order with the
oracle's roundings, so the un-split SpMM / fused step is additionally checked BIT FOR BIT.
Full-size BASELINE configs live in tests/test_gpu_fullsize.py."""
import numpy as np
import pytest
import torch

import gnntf_oracle as oracle
import synthetic

pytestmark = pytest.mark.gpu


def _gnntf():
    import gnntf
    return gnntf


def _kat_graph():
    import networkx as nx
    G = nx.DiGraph()
    for u in ["c", "a", "b", "d", "iso"]:
        G.add_node(u)
    G.add_edge("a", "b")
    G.add_edge("b", "a")
    G.add_edge("c", "d", weight=2.5)
    G.add_edge("d", "d")
    G.add_edge("a", "c")
    return G


def _random_edges(n, e, seed, self_loops=True):
    rng = np.random.default_rng(seed)
    edges = rng.integers(0, n, size=(e, 2)).astype(np.int64)
    if not self_loops:
        edges = edges[edges[:, 0] != edges[:, 1]]
    w = rng.random(edges.shape[0]).astype(np.float32) + 0.25
    return edges, w


def _np(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------
# (a) graph2adj / CSR builder — bit-exact
# ------------------------------------------------------------------------------------------
def test_kat1_graph2adj_indices_and_values():
    gnntf = _gnntf()
    adj = gnntf.graph2adj(_kat_graph())
    assert _np(adj.indices).tolist() == [[0, 3], [1, 2], [1, 0], [2, 1], [3, 3], [3, 0], [2, 1], [0, 1], [1, 2], [3, 3]]
    assert _np(adj.values).tolist() == [2.5, 1, 1, 1, 1, 2.5, 1, 1, 1, 1]
    assert adj.dense_shape == (5, 5) and adj.shape == (5, 5)
    assert adj.indices.dtype == torch.int64
    row_ptr, col_idx, coo_pos, _ = oracle.csr_from_coo(_np(adj.indices), 5)
    assert _np(adj.csr.row_ptr).tolist() == row_ptr.tolist()
    assert _np(adj.csr.col_idx).tolist() == col_idx.tolist()
    assert _np(adj.csr.coo_pos).tolist() == coo_pos.tolist()


@pytest.mark.parametrize("n,e,directed", [(1, 0, False), (7, 0, False), (1, 3, False), (50, 400, False),
                                          (50, 400, True), (1000, 20000, False), (4097, 100000, False),
                                          (300, 70000, False)])
def test_builder_bit_exact_vs_oracle(n, e, directed):
    gnntf = _gnntf()
    edges, w = _random_edges(n, e, seed=n + e)
    adj = gnntf.edges2adj(edges, w, n, directed=directed)
    idx, val, shape = oracle.graph2adj_arrays(edges, w, n, directed)
    assert adj.csr.nnz == idx.shape[0] == (e if directed else 2 * e)
    assert np.array_equal(_np(adj.indices), idx)
    assert np.array_equal(_np(adj.values), val)
    row_ptr, col_idx, coo_pos, _ = oracle.csr_from_coo(idx, n, directed)
    assert np.array_equal(_np(adj.csr.row_ptr).astype(np.int64), row_ptr)
    assert np.array_equal(_np(adj.csr.col_idx), col_idx)
    assert np.array_equal(_np(adj.csr.coo_pos).astype(np.int64), coo_pos)
    assert np.array_equal(_np(adj.raw_val), val[coo_pos])


def test_builder_networkx_path_matches_oracle_cora_shape():
    gnntf = _gnntf()
    n, e, _, _ = synthetic.SHAPES["cora"]
    G = synthetic.citation_graph(n, e, seed=0)
    adj = gnntf.graph2adj(G)
    idx, val, _ = oracle.graph2adj(G)
    assert adj.csr.nnz == 2 * e == 21112
    assert np.array_equal(_np(adj.indices), idx)
    assert np.array_equal(_np(adj.values), val)


def test_builder_out_of_range_edge_raises():
    gnntf = _gnntf()
    with pytest.raises(Exception):
        gnntf.edges2adj(np.array([[0, 5]]), None, 3)


# ------------------------------------------------------------------------------------------
# get_adjacency — normalisation
# ------------------------------------------------------------------------------------------
def test_kat1_get_adjacency():
    gnntf = _gnntf()
    adj = gnntf.graph2adj(_kat_graph())
    A = adj.normalized("symmetric")
    np.testing.assert_allclose(_np(A.deg), [3.5, 3, 2, 4.5, 0], rtol=0, atol=0)
    np.testing.assert_allclose(_np(A.dinv), np.float32([0.5345225, 0.57735026, 0.70710677, 0.47140455, 0]), rtol=2e-7)
    expect = np.float32([0.6299408, 0.40824828, 0.30860668, 0.40824828, 0.22222225,
                         0.6299408, 0.40824828, 0.30860668, 0.40824828, 0.22222225])
    np.testing.assert_allclose(_np(A.values), expect, rtol=3e-7)


def test_kat2_masked_adjacency_is_not_symmetric():
    gnntf = _gnntf()
    adj = gnntf.graph2adj(_kat_graph())
    keep = torch.tensor([1, 0, 1, 1, 0, 1, 1, 0, 1, 1], dtype=torch.uint8)
    A = adj.normalized("symmetric", keep_mask=keep.cuda(), rate=0.5)
    np.testing.assert_allclose(_np(A.deg), [7, 4, 2, 7, 0], rtol=0, atol=0)
    expect = np.float32([0.7142858, 0, 0.3779645, 0.70710677, 0, 0.7142858, 0.70710677, 0, 0.70710677, 0.28571433])
    np.testing.assert_allclose(_np(A.values), expect, rtol=3e-7)
    # transposed values: dense check against the oracle
    idx = _np(adj.indices)
    dense = np.zeros((5, 5), np.float64)
    np.add.at(dense, (idx[:, 0], idx[:, 1]), _np(A.values).astype(np.float64))
    csr_t, val_t = A.transposed()
    dense_t = np.zeros((5, 5), np.float64)
    rows = np.repeat(np.arange(5), np.diff(_np(csr_t.row_ptr)))
    np.add.at(dense_t, (rows, _np(csr_t.col_idx)), _np(val_t).astype(np.float64))
    assert abs(dense[1, 0] - 0.3779645) < 1e-6 and dense[0, 1] == 0
    np.testing.assert_allclose(dense_t, dense.T, rtol=1e-7)


@pytest.mark.parametrize("mode", ["symmetric", "bipartite", "none"])
@pytest.mark.parametrize("eye", ["none", "before", "after"])
@pytest.mark.parametrize("directed,masked", [(False, False), (False, True), (True, False), (True, True)])
def test_normalize_vs_oracle(mode, eye, directed, masked):
    gnntf = _gnntf()
    n, e = 300, 4000
    edges, w = _random_edges(n, e, seed=11)
    edges[:40, 1] = edges[:40, 0]            # self loops
    edges = edges[edges[:, 0] < n - 10]       # isolated nodes at the end
    edges = edges[edges[:, 1] < n - 10]
    w = w[: edges.shape[0]]
    adj = gnntf.edges2adj(edges, w, n, directed=directed)
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n, directed)
    keep, rate = None, 0.0
    if masked:
        rate = 0.5
        keep = (np.random.default_rng(2).random(idx.shape[0]) >= rate)
        val_m = oracle.sparse_dropout(val, rate, keep)
    else:
        val_m = val
    A = adj.normalized(mode, eye, keep_mask=None if keep is None else torch.from_numpy(keep).cuda(), rate=rate)
    idx_o, val_o, D = oracle.get_adjacency(idx, val_m, n, mode, eye)
    assert np.array_equal(_np(A.indices), idx_o)
    oracle.assert_close(_np(A.values), val_o, what=f"normalised values {mode}/{eye}")
    if D is not None:
        oracle.assert_close(_np(A.dinv), D, what="D")
        assert np.all(_np(A.dinv)[n - 10:] == 0) or eye == "before"
    # CSR-order values are the COO values permuted
    assert np.array_equal(_np(A.val), _np(A.values)[_np(A.csr.coo_pos)])
    if not directed or eye == "none":
        csr_t, val_t = A.transposed()
        H = np.random.default_rng(3).standard_normal((n, 5)).astype(np.float32)
        got = _np(gnntf.ops.spmm_raw(csr_t.struct(val_t, 5), n, torch.from_numpy(H).cuda()))
        oracle.assert_close(got, oracle.spmm_coo_T(idx_o, val_o, H, n_cols=n), what="transposed SpMM")


def test_invalid_normalization_raises():
    gnntf = _gnntf()
    adj = gnntf.graph2adj(_kat_graph())
    with pytest.raises(Exception, match="Invalid matrix normalization"):
        adj.normalized("laplacian")


# ------------------------------------------------------------------------------------------
# (b) SpMM
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F", [1, 3, 4, 7, 8, 16, 31, 40, 47, 64, 100, 128, 130, 256, 500, 1433])
def test_spmm_vs_oracle_feature_widths(F):
    gnntf = _gnntf()
    n, e = 257, 3000
    edges, w = _random_edges(n, e, seed=F)
    adj = gnntf.edges2adj(edges, w, n)
    A = adj.normalized("symmetric")
    H = np.random.default_rng(F + 1).standard_normal((n, F)).astype(np.float32)
    got = _np(gnntf.sparse_dense_matmul(A, torch.from_numpy(H).cuda()))
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n)
    _, nv, _ = oracle.get_adjacency(idx, val, n)
    oracle.assert_close(got, oracle.spmm_coo(idx, nv, H), what=f"SpMM F={F}")


@pytest.mark.parametrize("F", [1, 4, 7, 8, 40, 48, 100, 128, 130, 500])
def test_spmm_unsplit_rows_are_bit_identical_to_the_oracle(F):
    """Same accumulation order (stable CSR = COO order inside a row) and the same roundings (one
    multiply, one add per term, like TF-CPU without FMA contraction and like the oracle): every
    row that is not split must equal the oracle bit for bit — SpMM and the fused PPR step.
    Weights are dyadic so the degree sums are exact in any order (the normalisation kernels are
    IEEE-exact per operation: sqrt, divide, two rounded multiplies)."""
    gnntf = _gnntf()
    rng = np.random.default_rng(100 + F)
    n, e = 613, 9000
    edges = rng.integers(0, n, size=(e, 2)).astype(np.int64)
    edges[:400, 0] = 3                                            # one split row (deg > 256)
    w = rng.choice(np.float32([0.5, 1.0, 1.5, 2.0, 2.5]), size=e)
    adj = gnntf.edges2adj(edges, w, n)
    assert adj.csr.n_long >= 1
    A = adj.normalized("symmetric")
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n)
    _, nv, D = oracle.get_adjacency(idx, val, n)
    assert np.array_equal(_np(A.values), nv), "normalised values differ from the oracle's bits"
    H = rng.standard_normal((n, F)).astype(np.float32)
    H0 = rng.standard_normal((n, F)).astype(np.float32)
    unsplit = np.ones(n, bool)
    unsplit[_np(adj.csr.long_row)] = False
    got = _np(gnntf.sparse_dense_matmul(A, torch.from_numpy(H).cuda()))
    expect = oracle.spmm_coo(idx, nv, H)
    assert np.array_equal(got[unsplit], expect[unsplit])
    oracle.assert_close(got[~unsplit], oracle.spmm_coo(idx, nv, H, dtype=np.float64)[~unsplit], what=f"split rows F={F}",
                        floor=oracle.FLOOR_REORDERED, norm=float(np.abs(expect).max()))
    step = _np(gnntf.appnp_step(A, torch.from_numpy(H).cuda(), torch.from_numpy(H0).cuda(), 0.1))
    expect = oracle.ppr_iteration(idx, nv, H, H0, 0.1)
    assert np.array_equal(step[unsplit], expect[unsplit])


@pytest.mark.parametrize("F", [7, 16, 40, 100, 128, 500])
def test_spmm_long_rows_and_empty_rows(F):
    """Hub rows far above the split threshold (exercise the piece + fixed-order reduce path),
    rows of every short length, and empty rows."""
    gnntf = _gnntf()
    n = 3000
    rng = np.random.default_rng(5)
    hub = np.stack([np.zeros(2500, np.int64), rng.integers(1, n - 100, 2500)], 1)       # deg ≈ 2500+
    hub2 = np.stack([np.full(700, 17, np.int64), rng.integers(1, n - 100, 700)], 1)
    rest = rng.integers(1, n - 100, size=(8000, 2)).astype(np.int64)
    edges = np.concatenate([hub, hub2, rest])
    w = rng.random(edges.shape[0]).astype(np.float32)
    adj = gnntf.edges2adj(edges, w, n)
    assert adj.csr.n_long >= 2 and adj.csr.n_chunks >= 12
    A = adj.normalized("symmetric")
    H = rng.standard_normal((n, F)).astype(np.float32)
    got = _np(gnntf.sparse_dense_matmul(A, torch.from_numpy(H).cuda()))
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n)
    _, nv, _ = oracle.get_adjacency(idx, val, n)
    expect = oracle.spmm_coo(idx, nv, H, dtype=np.float64)
    oracle.assert_close(got, expect, what=f"SpMM long rows F={F}")
    assert np.all(got[n - 100:] == 0)


def test_spmm_strided_and_unaligned_operands():
    gnntf = _gnntf()
    n, e, F = 500, 6000, 100
    edges, w = _random_edges(n, e, seed=9)
    adj = gnntf.edges2adj(edges, w, n)
    A = adj.normalized("symmetric")
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n)
    _, nv, _ = oracle.get_adjacency(idx, val, n)
    big = torch.randn((n, 160), device="cuda")
    for view in (big[:, :F], big[:, 1:F + 1], big[:, 3:F + 4]):   # ld 160; aligned, off by 4 B, off by 12 B (odd F)
        got = _np(gnntf.ops.spmm_raw(A.struct(view.shape[1]), n, view))
        oracle.assert_close(got, oracle.spmm_coo(idx, nv, _np(view)), what="strided SpMM")


def test_spmm_empty_graph_and_zero_rows():
    gnntf = _gnntf()
    adj = gnntf.edges2adj(np.zeros((0, 2), np.int64), None, 6)
    H = torch.randn((6, 9), device="cuda")
    got = gnntf.sparse_dense_matmul(adj.normalized("symmetric"), H)
    assert got.shape == (6, 9) and torch.all(got == 0)


def test_spmm_dimension_mismatch_raises():
    gnntf = _gnntf()
    adj = gnntf.graph2adj(_kat_graph())
    with pytest.raises(Exception):
        gnntf.sparse_dense_matmul(adj.normalized(), torch.zeros((4, 3), device="cuda"))


# ------------------------------------------------------------------------------------------
# (c) APPNP step / K-step loop / backward
# ------------------------------------------------------------------------------------------
def test_kat1_appnp_k10():
    gnntf = _gnntf()
    adj = gnntf.graph2adj(_kat_graph())
    H0 = np.array([[(3 * i + j) / 7 - 1 for j in range(3)] for i in range(5)], np.float32)
    out = _np(gnntf.appnp_propagate(adj.normalized("symmetric"), torch.from_numpy(H0).cuda(), 0.1, 10))
    expect = np.array([[-0.386717034, -0.242003374, -0.097289714], [-0.37323383, -0.233458343, -0.093682856],
                       [-0.301202958, -0.181881658, -0.062560359], [-0.311411903, -0.151985331, 0.007441241],
                       [0.071428571, 0.085714286, 0.1]])
    oracle.assert_close(out, expect, what="KAT-1 H10")


@pytest.mark.parametrize("F", [7, 40, 47, 100, 128])
@pytest.mark.parametrize("K", [0, 1, 2, 10])
def test_appnp_propagate_vs_oracle(F, K):
    gnntf = _gnntf()
    n, e = 1200, 15000
    edges, w = _random_edges(n, e, seed=F + K)
    adj = gnntf.edges2adj(edges, w, n)
    H0 = np.random.default_rng(1).standard_normal((n, F)).astype(np.float32)
    out = _np(gnntf.appnp_propagate(adj.normalized("symmetric"), torch.from_numpy(H0).cuda(), 0.1, K))
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n)
    expect = oracle.appnp_propagate(idx, val, n, H0, 0.1, K)[-1] if K else H0
    oracle.assert_close(out, expect, what=f"APPNP F={F} K={K}")


def test_appnp_step_with_feature_dropout_and_relu():
    gnntf = _gnntf()
    n, e, F = 400, 5000, 24
    edges, w = _random_edges(n, e, seed=4)
    adj = gnntf.edges2adj(edges, w, n)
    A = adj.normalized("symmetric")
    rng = np.random.default_rng(8)
    H, H0 = rng.standard_normal((n, F)).astype(np.float32), rng.standard_normal((n, F)).astype(np.float32)
    keep = rng.random((n, F)) >= 0.3
    got = _np(gnntf.appnp_step(A, torch.from_numpy(H).cuda(), torch.from_numpy(H0).cuda(), 0.1,
                               torch.from_numpy(keep).cuda(), 0.3, relu=True))
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n)
    _, nv, _ = oracle.get_adjacency(idx, val, n)
    expect = oracle.ppr_iteration(idx, nv, H, H0, 0.1, feat_keep=keep, p_feat=0.3, training=True, activation=oracle.relu)
    oracle.assert_close(got, expect, what="fused step with dropout+relu")


@pytest.mark.parametrize("masked", [False, True])
def test_appnp_training_forward_and_backward_vs_oracle(masked):
    """Per-iteration edge masks (filter.py:18) and the fused VJP."""
    gnntf = _gnntf()
    n, e, F, K, a = 600, 7000, 12, 5, 0.1
    edges, w = _random_edges(n, e, seed=21)
    adj = gnntf.edges2adj(edges, w, n)
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n)
    rng = np.random.default_rng(6)
    H0 = rng.standard_normal((n, F)).astype(np.float32)
    g = rng.standard_normal((n, F)).astype(np.float32)
    if masked:
        keeps = [rng.random(idx.shape[0]) >= 0.5 for _ in range(K)]
        adjs = [adj.normalized("symmetric", keep_mask=torch.from_numpy(k).cuda(), rate=0.5) for k in keeps]
        nvs = [oracle.get_adjacency(idx, oracle.sparse_dropout(val, 0.5, k), n)[1] for k in keeps]
        expect = oracle.appnp_propagate(idx, val, n, H0, a, K, 0.5, keeps, training=True)[-1]
    else:
        adjs = adj.normalized("symmetric")
        nvs = [oracle.get_adjacency(idx, val, n)[1]] * K
        expect = oracle.appnp_propagate(idx, val, n, H0, a, K)[-1]
    H0_t = torch.from_numpy(H0).cuda().requires_grad_(True)
    out = gnntf.appnp_propagate(adjs, H0_t, a, K)
    oracle.assert_close(_np(out), expect, what="training forward")
    out.backward(torch.from_numpy(g).cuda())
    oracle.assert_close(_np(H0_t.grad), oracle.appnp_propagate_bwd(idx, nvs, g, a), what="dH0")


def test_sqrt_degree_is_a_fixed_point_arxiv_shape():
    """Size-independent property at a full BASELINE size: for the symmetric normalisation
    Â·√deg = √deg, so H0 = √deg ⊗ c is a fixed point of every PPR step."""
    gnntf = _gnntf()
    n, edges = synthetic.shaped_edges("arxiv", seed=0, device="cuda")
    adj = gnntf.edges2adj(edges, None, n)
    assert adj.csr.nnz == 2 * synthetic.SHAPES["arxiv"][1]
    A = adj.normalized("symmetric")
    c = torch.linspace(-1, 1, 40, device="cuda")
    H0 = torch.sqrt(A.deg)[:, None] * c[None, :]
    out = gnntf.appnp_propagate(A, H0, 0.1, 10)
    err = (out - H0).abs().max().item() / H0.abs().max().item()
    assert err < 1e-5, err


def test_linearity_and_row_sample_products_tenth():
    """products-shaped at 1/10 scale: linearity of the K-step map, and 200 sampled SpMM rows
    recomputed on the CPU from the CSR arrays."""
    gnntf = _gnntf()
    n, edges = synthetic.shaped_edges("products", seed=0, device="cuda", scale=0.1)
    adj = gnntf.edges2adj(edges, None, n)
    A = adj.normalized("symmetric")
    F = 100
    X, Y = synthetic.features(n, F, 1, "cuda"), synthetic.features(n, F, 2, "cuda")
    pX, pY, pXY = (gnntf.appnp_propagate(A, t, 0.1, 10) for t in (X, Y, X + 2 * Y))
    rel = ((pXY - (pX + 2 * pY)).abs().max() / pXY.abs().max()).item()
    assert rel < 1e-5, rel
    P = _np(gnntf.sparse_dense_matmul(A, X))
    row_ptr, col, val, Xh = _np(A.csr.row_ptr), _np(A.csr.col_idx), _np(A.val), _np(X)
    rows = np.random.default_rng(0).integers(0, n, 200).tolist() + _np(adj.csr.long_row)[:5].tolist()
    for r in rows:
        s, t = row_ptr[r], row_ptr[r + 1]
        expect = (val[s:t, None].astype(np.float64) * Xh[col[s:t]].astype(np.float64)).sum(0)
        oracle.assert_close(P[r], expect, what=f"row {r}")


# ------------------------------------------------------------------------------------------
# Drop-in API: APPNP / GCN architectures, train / predict
# ------------------------------------------------------------------------------------------
def _oracle_mlp(X, Ws, bs):
    H = X
    for i, (W, b) in enumerate(zip(Ws, bs)):
        H = H @ W + b
        if i < len(Ws) - 1:
            H = np.maximum(H, 0)
    return H


def test_appnp_architecture_eval_forward_vs_oracle():
    gnntf = _gnntf()
    gnntf.set_seed(0)
    n, e, width, classes = 600, 3000, 50, 7
    G = synthetic.citation_graph(n, e, seed=3)
    X = synthetic.citation_features(n, width, seed=4)
    arch = gnntf.APPNP(gnntf.graph2adj(G), X, num_classes=classes)
    arch.reset()
    arch.training_mode(False)
    out = _np(arch(arch.features))
    Ws = [w.numpy() for w in arch.vars()][0::2]
    bs = [w.numpy() for w in arch.vars()][1::2]
    H0 = _oracle_mlp(X, Ws, bs).astype(np.float32)
    idx, val, _ = oracle.graph2adj(G)
    expect = oracle.appnp_propagate(idx, val, n, H0, 0.1, 10)[-1]
    oracle.assert_close(out, expect, what="APPNP eval forward")


def test_gcn_architecture_eval_forward_vs_oracle():
    gnntf = _gnntf()
    gnntf.set_seed(1)
    n, e, width, classes = 500, 2600, 60, 3
    G = synthetic.citation_graph(n, e, seed=5)
    X = synthetic.citation_features(n, width, seed=6)
    arch = gnntf.GCN(gnntf.graph2adj(G), X, num_classes=classes)
    arch.reset()
    arch.training_mode(False)
    out = _np(arch(arch.features))
    Ws = [w.numpy() for w in arch.vars()][0::2]
    bs = [w.numpy() for w in arch.vars()][1::2]
    idx, val, _ = oracle.graph2adj(G)
    expect = oracle.gcn_forward(idx, val, n, X, Ws, bs)
    oracle.assert_close(out, expect, what="GCN eval forward")
    assert (out >= 0).all()  # relu on the output layer, gcn.py:113


def test_train_predict_roundtrip_learns_planted_labels():
    gnntf = _gnntf()
    gnntf.set_seed(0)
    n, classes, width = 900, 4, 32
    rng = np.random.default_rng(0)
    labels = rng.integers(0, classes, n)
    # homophilous graph + label-correlated features
    import networkx as nx
    G = nx.DiGraph()
    G.add_nodes_from(range(n))
    for u in range(n):
        same = np.flatnonzero(labels == labels[u])
        for v in rng.choice(same, 4):
            if u != v:
                G.add_edge(u, int(v))
                G.add_edge(int(v), u)
    X = rng.standard_normal((n, width)).astype(np.float32) * 2.0
    X[np.arange(n), labels] += 1.5
    order = rng.permutation(n)
    train, valid, test = order[:200], order[200:400], order[400:]
    for make in (lambda: gnntf.APPNP(gnntf.graph2adj(G), X, num_classes=classes),
                 lambda: gnntf.GCN(gnntf.graph2adj(G), X, num_classes=classes)):
        arch = make()
        arch.train(train=gnntf.NodeClassification(train, labels[train]),
                   valid=gnntf.NodeClassification(valid, labels[valid]), patience=20, epochs=150)
        prediction = arch.predict(gnntf.NodeClassification(test))
        accuracy = gnntf.acc(prediction, labels[test])
        assert accuracy > 0.8, accuracy
        assert not arch.is_training()


def test_host_buffer_entry_matches_device_entry():
    gnntf = _gnntf()
    n, e, F = 2000, 30000, 47
    edges, w = _random_edges(n, e, seed=13)
    adj = gnntf.edges2adj(edges, w, n)
    A = adj.normalized("symmetric")
    H0 = torch.randn((n, F)).pin_memory()
    out_host = gnntf.appnp_propagate_host(A, H0, 0.1, 10)
    out_dev = gnntf.appnp_propagate(A, H0.cuda(), 0.1, 10)
    assert torch.equal(out_host, out_dev.cpu())


@pytest.mark.parametrize("n_batches,F,K", [(1, 47, 10), (2, 40, 3), (5, 100, 10), (7, 16, 1)])
def test_host_batched_pipeline_matches_single_calls(n_batches, F, K):
    """gnntf_appnp_propagate_host_batched_f32 (three-stream software pipeline, two device slots): every
    result is bit-equal to the device entry on the same input, for odd / even batch counts, repeated
    input pointers and a caller-provided workspace; the call before and after on the same stream stay ordered."""
    gnntf = _gnntf()
    n, e = 30000, 400000
    edges, w = _random_edges(n, e, seed=17)
    adj = gnntf.edges2adj(edges, w, n)
    A = adj.normalized("symmetric")
    ins = [torch.randn((n, F)).pin_memory() for _ in range(min(n_batches, 3))]
    seq = [ins[b % len(ins)] for b in range(n_batches)]                      # pointers repeat from batch 3 on
    work = torch.full((5, n, F), float("nan"), device="cuda")
    outs = gnntf.appnp_propagate_host_batched(A, seq, 0.1, K, work=work)
    assert len(outs) == n_batches
    for H, got in zip(seq, outs):
        assert torch.equal(got, gnntf.appnp_propagate(A, H.cuda(), 0.1, K).cpu())
    again = gnntf.appnp_propagate_host_batched(A, seq, 0.1, K)                # own workspace, own outputs
    assert all(torch.equal(x, y) for x, y in zip(outs, again))
    assert gnntf.appnp_propagate_host_batched(A, [], 0.1, K) == []
    with pytest.raises(ValueError):
        gnntf.appnp_propagate_host_batched(A, [seq[0].cuda()], 0.1, K)


# ------------------------------------------------------------------------------------------
# Row-sharded path: all ranks emulated in ONE process on one GPU (the exchange is a device copy),
# native pack + interior/boundary step kernels with row_map.  The NCCL exchange itself is covered
# by bench.py --gpus N on a multi-GPU box and by the gloo tests on CPU.
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_propagator_emulated_ranks(world):
    gnntf = _gnntf()
    from gnntf import dist as gdist
    n, edges = synthetic.shaped_edges("arxiv", seed=0, device="cuda", scale=0.2)
    adj = gnntf.edges2adj(edges, None, n)
    A = adj.normalized("symmetric")
    F, K, a = 40, 10, 0.1
    H0 = synthetic.features(n, F, 1, "cuda")
    expect = gnntf.appnp_propagate(A, H0, a, K)
    csr = A.csr
    bounds = gdist.partition_bounds(csr.row_ptr, world)
    first = [gdist.build_shard_plan(csr.row_ptr, csr.col_idx, A.val, r, 1 if world == 1 else world, peer_wants=lambda d: torch.empty(0))
             for r in range(world)]          # pass 1: learn every rank's halo
    plans = [gdist.build_shard_plan(csr.row_ptr, csr.col_idx, A.val, r, world,
                                    peer_wants=lambda d, r=r: gdist.wanted_rows(first[d].halo_cols, bounds, r))
             for r in range(world)]
    assert sum(p.n_local for p in plans) == n and any(p.n_halo > 0 for p in plans)
    props = []

    def exchange(me, send_buf, halo_out):  # deliver every peer's packed rows for `me`
        pass
    props = [gdist.ShardedPropagator(adj, A, F, r, world, plan=plans[r], exchange=exchange, halves=1) for r in range(world)]
    # lock-step emulation of the K steps: pack on every rank, route, then compute on every rank
    L, nat = gnntf._native.lib(), gnntf._native
    bufs = [(p.buf[0], p.buf[1]) for p in props]
    for p in props:
        p.H0.copy_(H0[p.lo:p.hi])
        p.buf[0][:p.n_local].copy_(p.H0)
    for _ in range(K):
        packed = []
        for p, (src, dst) in zip(props, bufs):
            if p.send_buf.shape[0]:
                nat.check(L.gnntf_halo_pack_f32(nat.ptr(src), F, nat.ptr(p.plan.send_idx), p.send_buf.shape[0],
                                                nat.ptr(p.send_buf), F, F, nat.stream_ptr()))
            packed.append(torch.split(p.send_buf, p.plan.send_counts))
        for r, (p, (src, dst)) in enumerate(zip(props, bufs)):
            parts = [packed[o][r] for o in range(world)]                 # from owner o to me, in owner order
            if p.n_halo:
                src[p.n_local:].copy_(torch.cat(parts))
        for p, (src, dst) in zip(props, bufs):
            p._exchange = lambda *args: None
            p._step(src, dst, a)
        bufs = [(dst, src) for (src, dst) in bufs]
    got = torch.cat([src[:p.n_local] for p, (src, dst) in zip(props, bufs)])
    oracle.assert_close(_np(got), _np(expect), what="sharded vs single-GPU propagation", floor=oracle.FLOOR_REORDERED)
    assert props[0].launches_per_propagation(K) >= 2 * K
    assert props[0].owned.nnz + props[0].halo_part.nnz == props[0].nnz_local and props[0].halo_part.nnz > 0


def _emulated_plans(gdist, A, world):
    csr = A.csr
    bounds = gdist.partition_bounds(csr.row_ptr, world)
    first = [gdist.build_shard_plan(csr.row_ptr, csr.col_idx, A.val, r, world, peer_wants=lambda d: torch.empty(0))
             for r in range(world)]          # pass 1: learn every rank's halo
    return [gdist.build_shard_plan(csr.row_ptr, csr.col_idx, A.val, r, world,
                                   peer_wants=lambda d, r=r: gdist.wanted_rows(first[d].halo_cols, bounds, r))
            for r in range(world)]


@pytest.mark.parametrize("n_peers,F,rotate_frac", [(2, 100, 0.5), (3, 40, 0.3), (4, 48, 0.0), (5, 52, 0.9), (8, 100, 0.37),
                                                   (4, 7, 0.5), (3, 47, 0.2)])
def test_halo_push_kernel_on_one_gpu(n_peers, F, rotate_frac):
    """gnntf_halo_push_f32 with the peer pointers aimed at buffers on THIS device: every destination
    receives exactly its slice of the send list (float4 path for F % 4 == 0, scalar otherwise), for
    any rotation of the starting row, and nothing outside its halo region is touched."""
    gnntf = _gnntf()
    nat, L = gnntf._native, gnntf._native.lib()
    rng = np.random.default_rng(n_peers * 100 + F)
    n_src = 5000
    H = torch.from_numpy(rng.standard_normal((n_src, F)).astype(np.float32)).cuda()
    counts = [0] + [int(x) for x in rng.integers(0, 3000, n_peers - 1)]            # slot 0 plays "myself": nothing sent
    if n_peers > 2:
        counts[2] = 0                                                               # a peer that needs nothing (NULL pointer)
    send_idx = torch.from_numpy(rng.integers(0, n_src, sum(counts)).astype(np.int32)).cuda()
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    row0 = rng.integers(0, 50, n_peers).astype(np.int64)
    bufs = [torch.full((int(row0[d]) + counts[d] + 7, F), -7.0, device="cuda") for d in range(n_peers)]
    ptrs = torch.tensor([0 if counts[d] == 0 else bufs[d].data_ptr() for d in range(n_peers)], dtype=torch.int64, device="cuda")
    n_send = int(off[-1])
    rotate = int(rotate_frac * n_send) if n_send else 0
    off_d, row0_d = torch.from_numpy(off).cuda(), torch.from_numpy(row0).cuda()     # (kept alive across the launch)
    nat.check(L.gnntf_halo_push_f32(nat.ptr(H), F, nat.ptr(send_idx), nat.ptr(off_d), nat.ptr(ptrs), nat.ptr(row0_d),
                                    n_peers, n_send, rotate, F, F, nat.stream_ptr()))
    torch.cuda.synchronize()
    for d in range(n_peers):
        lo, hi = int(off[d]), int(off[d + 1])
        r0 = int(row0[d])
        assert torch.equal(bufs[d][r0:r0 + counts[d]], H[send_idx[lo:hi].long()])
        assert torch.all(bufs[d][:r0] == -7.0) and torch.all(bufs[d][r0 + counts[d]:] == -7.0)


@pytest.mark.parametrize("copy", [False, True], ids=["push", "copy_engine"])
@pytest.mark.parametrize("world,F,K", [(2, 40, 10), (2, 100, 1), (3, 52, 3), (4, 48, 10), (8, 100, 3), (4, 7, 2)])
def test_sharded_push_loop_with_emulated_peers(world, F, K, copy):
    """The full ShardedPropagator(push=True) loop — fused pack+send+signal kernel, epoch flags in (peer)
    memory, flag-wait kernels, two-pass step — with every rank emulated in this process on one GPU and
    the peer tables aimed at the other ranks' buffers.  Several propagations back to back (odd K too:
    the buffer-reuse hazard ADVICE r1 describes), then a sharded plain SpMM.  ``copy``: the copy-engine form
    of the exchange (all-gather layout, gnntf_peer_copy_signal: one DMA per peer + a 4-byte DMA of the epoch)."""
    gnntf = _gnntf()
    from gnntf import dist as gdist
    n, edges = synthetic.shaped_edges("arxiv", seed=0, device="cuda", scale=0.15)
    adj = gnntf.edges2adj(edges, None, n)
    A = adj.normalized("symmetric")
    plans = _emulated_plans(gdist, A, world)
    assert sum(p.n_local for p in plans) == n and any(p.n_halo > 0 for p in plans)
    props = [gdist.ShardedPropagator(adj, A, F, r, world, plan=plans[r], push=True, peers="local", copy=copy)
             for r in range(world)]
    assert all(p.push and p.copy == copy for p in props)
    gdist.connect_local(props)
    for call in range(3):
        H0 = synthetic.features(n, F, 10 + call, "cuda")
        expect = gnntf.appnp_propagate(A, H0, 0.1, K)
        outs = gdist.propagate_lockstep(props, [H0[p.lo:p.hi] for p in props], 0.1, K)
        got = torch.cat(outs)
        oracle.assert_close(_np(got), _np(expect), what=f"push loop world={world} F={F} K={K} call {call}",
                            floor=oracle.FLOOR_REORDERED)
    assert all(int(p._epoch.item()) == 3 * K for p in props)                 # epochs advanced once per step
    flags = torch.stack([p._flags.tensor for p in props]).cpu()             # [rank, 2, source]
    for r in range(world):
        for q in range(world):
            if q != r:
                assert int(flags[r, 0, q]) == 3 * K and int(flags[r, 1, q]) == 3 * K, (r, q, flags[r])
    H = synthetic.features(n, F, 99, "cuda")
    outs = gdist.propagate_lockstep(props, [H[p.lo:p.hi] for p in props], spmm_only=True)
    oracle.assert_close(_np(torch.cat(outs)), _np(gnntf.sparse_dense_matmul(A, H)), what=f"sharded SpMM world={world} F={F}",
                        floor=oracle.FLOOR_REORDERED)
    for p in props:
        p.close()


def test_sharded_host_pipeline_world1_matches_device_run():
    """ShardedPropagator.propagate_host_batched / propagate_host_timed (upload, K steps and read-back of
    consecutive inputs on three streams): every host result equals the device-side propagation of its input."""
    gnntf = _gnntf()
    from gnntf import dist as gdist
    n, edges = synthetic.shaped_edges("arxiv", seed=0, device="cuda", scale=0.2)
    adj = gnntf.edges2adj(edges, None, n)
    A = adj.normalized("symmetric")
    prop = gdist.ShardedPropagator(adj, A, 40, 0, 1)
    ins = [torch.randn((n, 40)).pin_memory() for _ in range(3)]
    seq = [ins[b % 3] for b in range(5)]
    outs = [torch.empty((n, 40)).pin_memory() for _ in range(5)]
    prop.propagate_host_batched(seq, outs, 0.1, 10)
    for h_in, h_out in zip(seq, outs):
        assert torch.equal(h_out, gnntf.appnp_propagate(A, h_in.cuda(), 0.1, 10).cpu())
    rec = prop.propagate_host_timed(ins[1], 0.1, 10, reps=3)
    assert torch.equal(rec["host_out"], gnntf.appnp_propagate(A, ins[1].cuda(), 0.1, 10).cpu())
    assert rec["seconds"] > 0 and rec["single_call_seconds"] > 0 and rec["h2d"] == n * 40 * 4


def test_sharded_propagator_column_halves_world1_matches():
    """The two-half software pipeline (used when world > 1) on one rank: same result as one chain."""
    gnntf = _gnntf()
    from gnntf import dist as gdist
    n, edges = synthetic.shaped_edges("arxiv", seed=0, device="cuda", scale=0.1)
    adj = gnntf.edges2adj(edges, None, n)
    A = adj.normalized("symmetric")
    H0 = synthetic.features(n, 100, 1, "cuda")
    expect = gnntf.appnp_propagate(A, H0, 0.1, 10)
    for halves in (1, 2):
        prop = gdist.ShardedPropagator(adj, A, 100, 0, 1, halves=halves)
        assert len(prop.parts) == halves and sum(p["F"] for p in prop.parts) == 100
        got = prop.propagate(H0, 0.1, 10)
        oracle.assert_close(_np(got), _np(expect), what=f"halves={halves}")


@pytest.mark.parametrize("F", [40, 100, 128])
def test_spmm_short_rows_mapping_and_accumulate(F):
    """Very short rows (the shape of the halo-column pass of a shard) and gnntf_spmm_acc_f32
    (C += scale * A.B), the second pass of a sharded step."""
    import ctypes
    gnntf = _gnntf()
    nat = gnntf._native
    n, e = 5000, 7000
    edges, w = _random_edges(n, e, seed=F)
    adj = gnntf.edges2adj(edges, w, n)
    assert adj.csr.nnz < 12 * n
    A = adj.normalized("symmetric")
    rng = np.random.default_rng(3)
    H = rng.standard_normal((n, F)).astype(np.float32)
    C0 = rng.standard_normal((n, F)).astype(np.float32)
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n)
    _, nv, _ = oracle.get_adjacency(idx, val, n)
    P = oracle.spmm_coo(idx, nv, H)
    oracle.assert_close(_np(gnntf.sparse_dense_matmul(A, torch.from_numpy(H).cuda())), P, what="short rows SpMM")
    C = torch.from_numpy(C0).cuda()
    s = A.struct(F)
    nat.check(nat.lib().gnntf_spmm_acc_f32(ctypes.byref(s), nat.ptr(torch.from_numpy(H).cuda()), F, nat.ptr(C), F, F, 0.9,
                                           nat.stream_ptr()))
    oracle.assert_close(_np(C), C0 + np.float32(0.9) * P, what="accumulate pass")


# ------------------------------------------------------------------------------------------
# Locality-restoring internal node order (gnntf/reorder.py)
# ------------------------------------------------------------------------------------------
def test_reordered_adjacency_gives_identical_results_and_keeps_external_indices():
    gnntf = _gnntf()
    n, edges = synthetic.shaped_edges("arxiv", seed=0, ordering="random", device="cuda", scale=0.2)
    w = torch.rand(edges.shape[0], device="cuda") + 0.5
    adj = gnntf.edges2adj(edges, w, n)
    radj = adj.reordered(power_iters=30, sweeps=10)
    assert radj.perm is not None and torch.equal(torch.sort(radj.perm).values, torch.arange(n, device="cuda"))
    assert torch.equal(radj.indices, adj.indices) and torch.equal(radj.values, adj.values)   # what the user sees
    assert radj.reordered() is radj
    H0 = synthetic.features(n, 47, 1, "cuda").requires_grad_(True)
    H1 = H0.detach().clone().requires_grad_(True)
    g = synthetic.features(n, 47, 2, "cuda")
    out_a = gnntf.appnp_propagate(adj.normalized("symmetric"), H0, 0.1, 10)
    out_b = gnntf.appnp_propagate(radj.normalized("symmetric"), H1, 0.1, 10)
    oracle.assert_close(_np(out_b), _np(out_a), rtol=1e-6, what="reordered forward")
    out_a.backward(g)
    out_b.backward(g)
    oracle.assert_close(_np(H1.grad), _np(H0.grad), rtol=1e-6, what="reordered backward")
    X = synthetic.features(n, 24, 3, "cuda")
    oracle.assert_close(_np(gnntf.sparse_dense_matmul(radj.normalized("symmetric"), X)),
                        _np(gnntf.sparse_dense_matmul(adj.normalized("symmetric"), X)), rtol=1e-6, what="reordered SpMM")
    # masked (training-mode) normalisation: COO order is unchanged, so the same mask means the same matrix
    keep = torch.rand(adj.n_graph, device="cuda") >= 0.5
    oracle.assert_close(_np(radj.normalized("symmetric", keep_mask=keep, rate=0.5).values),
                        _np(adj.normalized("symmetric", keep_mask=keep, rate=0.5).values), rtol=1e-6, what="masked values")


def test_arrangement_recovers_hidden_locality():
    """A randomly relabelled graph with multi-scale locality: after reordering, edges span far
    fewer positions (the property the SpMM's L2 hit rate depends on)."""
    gnntf = _gnntf()
    n, edges = synthetic.shaped_edges("products", seed=0, ordering="random", device="cuda", scale=0.02)
    adj = gnntf.edges2adj(edges, None, n)
    radj = adj.reordered()

    def median_span(e):
        d = (e[:, 0] - e[:, 1]).abs()
        return torch.minimum(d, n - d).float().median().item()
    before, after = median_span(edges), median_span(radj.edges)
    n_loc, e_loc = synthetic.shaped_edges("products", seed=0, ordering="local", device="cuda", scale=0.02)
    native = median_span(e_loc)
    assert after < before / 20 and after < 3 * native, (before, after, native)


@pytest.mark.parametrize("shape,F", [("cora", 7), ("cora", 64), ("pubmed", 3), ("pubmed", 64), ("pubmed", 100), ("pubmed", 500)])
def test_persistent_k_step_kernel_equals_step_by_step(shape, F):
    """Graphs whose step is a single wave of CTAs and that have no split rows run all K iterations in
    ONE cooperative launch (appnp_persistent_kernel, grid barrier between steps).  The result must be
    bit-identical to K separate fused-step launches and, through them, to the oracle."""
    gnntf = _gnntf()
    n, e, _, _ = synthetic.SHAPES[shape]
    G = synthetic.citation_graph(n, e, seed=0)
    adj = gnntf.graph2adj(G)
    assert adj.csr.n_long == 0
    A = adj.normalized("symmetric")
    H0 = synthetic.features(n, F, seed=1, device="cuda")
    for K in (2, 3, 10):
        fused = gnntf.appnp_propagate(A, H0, 0.1, K)
        H = H0
        for _ in range(K):
            H = gnntf.appnp_step(A, H, H0, 0.1)
        assert torch.equal(fused, H), f"K={K}"
    idx, val, _ = oracle.graph2adj(G)
    expect = oracle.appnp_propagate(idx, val, n, _np(H0), 0.1, 10)[-1]
    assert np.array_equal(_np(fused), expect), "no split rows, unit weights: bit-identical to the oracle"


@pytest.mark.parametrize("shape,F,K", [("cora", 8, 10), ("cora", 8, 1), ("cora", 12, 3), ("cora", 64, 2), ("pubmed", 4, 10),
                                       ("pubmed", 16, 3), ("cora", 128, 2)])
def test_cluster_resident_kernel_equals_step_by_step(shape, F, K):
    """gnntf_appnp_propagate_cluster_f32: graph, features and teleport rows resident in the shared memory of one
    thread-block cluster for all K steps, neighbour rows gathered through distributed shared memory.  Every
    cluster size / CTA size that can hold the shape must be bit-identical to K separate fused-step launches."""
    gnntf = _gnntf()
    from gnntf import ops
    n, e, _, _ = synthetic.SHAPES[shape]
    G = synthetic.citation_graph(n, e, seed=0)
    adj = gnntf.graph2adj(G)
    A = adj.normalized("symmetric")
    H0 = synthetic.features(n, F, seed=1, device="cuda")
    H = H0
    for _ in range(K):
        H = gnntf.appnp_step(A, H, H0, 0.1)
    ran = 0
    for C in (0, 1, 2, 4, 8, 16):
        for threads in (0, 512, 1024):
            got = ops.propagate_cluster_raw(A, H0, 0.1, K, C, threads, out=torch.full_like(H0, float("nan")))
            if got is None:
                continue
            ran += 1
            assert torch.equal(got, H), f"cluster of {C} CTAs x {threads} threads"
    assert ran >= 3 or (shape, F) == ("pubmed", 16), ran       # PubMed x 16 floats: no cluster size holds it
    # shapes that cannot be resident are declined, not mangled
    big_n, big_edges = synthetic.shaped_edges("arxiv", seed=0, device="cuda", scale=0.5)
    big = gnntf.edges2adj(big_edges, None, big_n).normalized("symmetric")
    assert ops.propagate_cluster_raw(big, synthetic.features(big_n, 64, 1, "cuda"), 0.1, 3) is None


def test_cluster_resident_kernel_hub_rows_and_spilled_slices():
    """Weighted graph with a few hub rows just under the split threshold and very uneven slices: the CTA
    owning the hubs holds more entries than its shared-memory slice, so its tail is read from global memory."""
    gnntf = _gnntf()
    from gnntf import ops
    rng = np.random.default_rng(5)
    n = 6000
    hubs = np.stack([np.repeat(np.arange(3), 120), rng.integers(3, n, 360)], 1)
    rest = rng.integers(0, n, (9000, 2))
    edges = np.concatenate([hubs, rest[rest[:, 0] != rest[:, 1]]]).astype(np.int64)
    w = rng.uniform(0.5, 2.0, len(edges)).astype(np.float32)
    adj = gnntf.edges2adj(torch.from_numpy(edges).cuda(), torch.from_numpy(w).cuda(), n)
    assert adj.csr.n_long == 0
    A = adj.normalized("symmetric")
    for F in (8, 32):
        H0 = synthetic.features(n, F, seed=2, device="cuda")
        H = H0
        for _ in range(3):
            H = gnntf.appnp_step(A, H, H0, 0.1)
        for C in (2, 8, 16):
            got = ops.propagate_cluster_raw(A, H0, 0.1, 3, C, 0)
            if got is not None:
                assert torch.equal(got, H), (F, C)
        assert torch.equal(gnntf.appnp_propagate(A, H0, 0.1, 3), H)


def test_propagation_is_bitwise_deterministic():
    """No float atomics on the undirected path: two runs (and a rebuilt adjacency) give identical bits,
    including rows split into pieces and the backward pass."""
    gnntf = _gnntf()
    n, edges = synthetic.shaped_edges("arxiv", seed=0, device="cuda", scale=0.5)
    H0 = synthetic.features(n, 100, 1, "cuda")
    g = synthetic.features(n, 100, 2, "cuda")
    outs, grads = [], []
    for _ in range(2):
        adj = gnntf.edges2adj(edges, None, n)
        assert adj.csr.n_long > 0
        A = adj.normalized("symmetric")
        for _ in range(2):
            h = H0.clone().requires_grad_(True)
            out = gnntf.appnp_propagate(A, h, 0.1, 10)
            out.backward(g)
            outs.append(out.detach().clone())
            grads.append(h.grad.clone())
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    assert all(torch.equal(grads[0], x) for x in grads[1:])


def test_csr2adj_matches_edges2adj():
    import scipy.sparse as sp
    gnntf = _gnntf()
    rng = np.random.default_rng(1)
    M = sp.random(300, 300, density=0.02, format="csr", random_state=3, dtype=np.float32)
    adj = gnntf.csr2adj(M.indptr, M.indices, M.data)
    assert adj.directed and adj.csr.nnz == M.nnz
    H = rng.standard_normal((300, 12)).astype(np.float32)
    got = _np(gnntf.sparse_dense_matmul(adj.normalized("none"), torch.from_numpy(H).cuda()))
    oracle.assert_close(got, M @ H, what="csr2adj SpMM")


def test_arrange_sweep_kernel_vs_numpy():
    """gnntf_arrange_sweep_f32 against the same formula in NumPy (fp64): weighted circular mean of the
    neighbours' positions with weights 1/(|offset| + eps); isolated nodes keep their position."""
    gnntf = _gnntf()
    nat = gnntf._native
    n, e = 700, 5000
    edges, _ = _random_edges(n, e, seed=31)
    edges = edges[edges[:, 0] < n - 5]
    edges = edges[edges[:, 1] < n - 5]                      # the last nodes are isolated
    adj = gnntf.edges2adj(edges, None, n)
    rng = np.random.default_rng(2)
    theta = (rng.random(n) * 2 * np.pi).astype(np.float32)
    eps = 0.05
    out = torch.empty(n, device="cuda")
    nat.check(nat.lib().gnntf_arrange_sweep_f32(nat.ptr(adj.csr.row_ptr), nat.ptr(adj.csr.col_idx),
                                                nat.ptr(torch.from_numpy(theta).cuda()), eps, nat.ptr(out), n,
                                                nat.stream_ptr()))
    rp, col = _np(adj.csr.row_ptr), _np(adj.csr.col_idx)
    t64 = theta.astype(np.float64)
    expect = t64.copy()
    for i in range(n):
        nb = col[rp[i]:rp[i + 1]]
        if nb.size:
            d = t64[nb] - t64[i]
            d -= 2 * np.pi * np.rint(d / (2 * np.pi))
            w = 1.0 / (np.abs(d) + eps)
            expect[i] = (t64[i] + (w * d).sum() / w.sum()) % (2 * np.pi)
    diff = np.abs(_np(out).astype(np.float64) - expect)
    diff = np.minimum(diff, 2 * np.pi - diff)               # compare on the circle
    assert diff.max() < 2e-5, diff.max()
    assert np.array_equal(_np(out)[n - 5:], theta[n - 5:])


# ------------------------------------------------------------------------------------------
# §8 f-1 / f-2: fused dense tail, cached Â·X, fused node-classification loss
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("act", ["identity", "relu", "leaky_relu"])
@pytest.mark.parametrize("F", [3, 64, 100])
def test_bias_act_dropout_forward_and_backward_vs_numpy(act, F):
    gnntf = _gnntf()
    rng = np.random.default_rng(F)
    n = 301
    Z = rng.standard_normal((n, F)).astype(np.float32)
    b = rng.standard_normal((1, F)).astype(np.float32)
    keep = rng.random((n, F)) >= 0.4
    g = rng.standard_normal((n, F)).astype(np.float32)
    f = {"identity": lambda x: x, "relu": oracle.relu, "leaky_relu": oracle.leaky_relu}[act]
    pre = Z + b
    scale = oracle.dropout_scale(0.4)
    expect = np.where(keep, f(pre) * scale, np.float32(0)).astype(np.float32)
    Zt = torch.from_numpy(Z).cuda().requires_grad_(True)
    bt = torch.from_numpy(b).cuda().requires_grad_(True)
    out = gnntf.bias_act_dropout(Zt, bt, act, torch.from_numpy(keep).cuda(), 0.4)
    assert np.array_equal(_np(out), expect)                       # same operations, same roundings
    out.backward(torch.from_numpy(g).cuda())
    slope = {"identity": np.ones_like(pre), "relu": (pre > 0).astype(np.float32),
             "leaky_relu": np.where(pre > 0, 1.0, 0.2).astype(np.float32)}[act]
    dZ = g * keep * scale * slope
    oracle.assert_close(_np(Zt.grad), dZ, what=f"dZ {act}")
    oracle.assert_close(_np(bt.grad), dZ.sum(0, keepdims=True), what=f"dbias {act}", floor=oracle.FLOOR_REORDERED)
    # no dropout, no bias
    out2 = gnntf.bias_act_dropout(torch.from_numpy(Z).cuda(), None, act)
    assert np.array_equal(_np(out2), f(Z))


def test_node_cross_entropy_vs_oracle_and_torch():
    gnntf = _gnntf()
    rng = np.random.default_rng(3)
    n, C, m = 5000, 47, 1234
    logits = (rng.standard_normal((n, C)) * 3).astype(np.float32)
    nodes = rng.choice(n, m, replace=False).astype(np.int64)
    labels = rng.integers(0, C, m).astype(np.int64)
    lt = torch.from_numpy(logits).cuda().requires_grad_(True)
    task = gnntf.NodeClassification(nodes, labels)
    loss = task.loss(lt)
    expect = oracle.node_classification_loss(logits.astype(np.float64), nodes, labels, dtype=np.float64)
    assert abs(float(loss) - float(expect)) <= 1e-5 * abs(float(expect))
    loss.backward()
    ref = torch.from_numpy(logits).cuda().requires_grad_(True)
    rows = ref.index_select(0, torch.from_numpy(nodes).cuda())
    torch.nn.functional.cross_entropy(torch.log_softmax(rows, 1), torch.from_numpy(labels).cuda()).backward()
    oracle.assert_close(_np(lt.grad), _np(ref.grad), what="node CE gradient", floor=oracle.FLOOR_REORDERED)
    assert task.evaluate(lt.detach()) == pytest.approx(float((logits[nodes].argmax(1) == labels).mean()))


def test_gcn_first_layer_aggregation_is_cached_in_eval_mode():
    gnntf = _gnntf()
    gnntf.set_seed(2)
    n, e, width, classes = 400, 2400, 30, 4
    G = synthetic.citation_graph(n, e, seed=8)
    X = synthetic.citation_features(n, width, seed=9)
    arch = gnntf.GCN(gnntf.graph2adj(G), X, num_classes=classes, latent_dims=[48])   # widening layer: (ÂX)W order
    arch.reset()
    arch.training_mode(False)
    out1 = arch(arch.features)
    A = arch.get_adjacency(0)
    assert getattr(A, "_agg_cache", None) is not None and A._agg_cache[0] is arch.features
    calls = []
    orig = gnntf.ops.sparse_dense_matmul
    gnntf.gnn.ops.sparse_dense_matmul = lambda a, h: (calls.append(h.shape), orig(a, h))[1]
    try:
        out2 = arch(arch.features)
    finally:
        gnntf.gnn.ops.sparse_dense_matmul = orig
    assert torch.equal(out1, out2) and calls == [(n, 4)], calls   # only the second layer propagates (at its output width)
    Ws = [w.numpy() for w in arch.vars()][0::2]
    bs = [w.numpy() for w in arch.vars()][1::2]
    idx, val, _ = oracle.graph2adj(G)
    oracle.assert_close(_np(out1), oracle.gcn_forward(idx, val, n, X, Ws, bs), what="GCN (cached ÂX) eval forward")


# ------------------------------------------------------------------------------------------
# §8 f-4: GCNII / NGCF / spectral-preserving variants / Structural on the same op
# ------------------------------------------------------------------------------------------
def _small_graph(seed, n=500, e=2600, width=40):
    G = synthetic.citation_graph(n, e, seed=seed)
    X = synthetic.citation_features(n, width, seed=seed + 1)
    return G, X, n


@pytest.mark.parametrize("spectral", [False, True])
def test_gcnii_eval_forward_vs_oracle(spectral):
    gnntf = _gnntf()
    gnntf.set_seed(4)
    G, X, n = _small_graph(11)
    layer_type = gnntf.GCNIISpectralPreservingLayer if spectral else gnntf.GCNIILayer
    arch = gnntf.GCNII(gnntf.graph2adj(G), X, num_classes=5, latent_dims=[32], iterations=8, layer_type=layer_type)
    arch.reset()
    rng = np.random.default_rng(0)
    for layer in arch.layers():          # the reference initialises the convolution weights to zero (gcn.py:11): perturb them
        if isinstance(layer, gnntf.GCNIILayer):
            layer.W.data.copy_(torch.from_numpy((rng.standard_normal(tuple(layer.W.shape)) * 0.2).astype(np.float32)))
            if spectral:
                layer.bias.data.copy_(torch.from_numpy((rng.standard_normal(tuple(layer.bias.shape)) * 0.1).astype(np.float32)))
    arch.training_mode(False)
    out = _np(arch(arch.features))
    layers = arch.layers()
    dense = [l for l in layers if isinstance(l, gnntf.Dense)]
    conv = [l for l in layers if isinstance(l, gnntf.GCNIILayer)]
    assert len(conv) == 8 and [c.k for c in conv] == list(range(8))
    idx, val, _ = oracle.graph2adj(G)
    expect = oracle.gcnii_forward(idx, val, n, X, [_np(dense[0].W)], [_np(dense[0].b)], [_np(c.W) for c in conv],
                                  _np(dense[1].W), _np(dense[1].b), a=0.1, l=0.5,
                                  conv_bias=[_np(c.bias) for c in conv] if spectral else None)
    oracle.assert_close(out, expect, what=f"GCNII eval forward spectral={spectral}", floor=oracle.FLOOR_REORDERED)


def test_gcnii_training_step_runs_and_gradients_reach_every_variable():
    gnntf = _gnntf()
    gnntf.set_seed(5)
    G, X, n = _small_graph(13)
    arch = gnntf.GCNII(gnntf.graph2adj(G), X, num_classes=4, latent_dims=[16], iterations=4)
    labels = np.random.default_rng(1).integers(0, 4, n)
    arch.train(gnntf.NodeClassification(np.arange(0, 200), labels[:200]), gnntf.NodeClassification(np.arange(200, 300), labels[200:300]),
               epochs=3, patience=3)
    assert all(torch.isfinite(w.var).all() for w in arch.vars())
    assert any(float(w.var.abs().sum()) > 0 for w in arch.vars() if w.normalization == "zero" and w.var.shape[0] == w.var.shape[1])


def test_ngcf_eval_forward_vs_oracle():
    gnntf = _gnntf()
    gnntf.set_seed(6)
    G, X, n = _small_graph(15, width=12)
    arch = gnntf.NGCF(gnntf.graph2adj(G), X, num_classes=12, dropout=0.1)
    arch.reset()
    arch.training_mode(False)
    out = _np(arch(arch.features))
    assert out.shape == (3 * n, 12)                        # Concatenate stacks along axis 0 (layers.py:100)
    ng = [l for l in arch.layers() if isinstance(l, gnntf.NGCFLayer)]
    assert len(ng) == 3 and all(l.adjacency.mode == "bipartite" for l in ng)
    idx, val, _ = oracle.graph2adj(G)
    expect = oracle.ngcf_forward(idx, val, n, X, [(_np(l.W1), _np(l.b1), _np(l.W2), _np(l.b2)) for l in ng])
    oracle.assert_close(out, expect, what="NGCF eval forward", floor=oracle.FLOOR_REORDERED)
    rows = np.linalg.norm(out, axis=1)
    assert np.allclose(rows[rows > 0], 1.0, atol=1e-5)     # l2_normalize (gcn.py:135)


def test_gcn_spectral_preserving_layer_and_structural_preprocessor():
    gnntf = _gnntf()
    gnntf.set_seed(7)
    G, X, n = _small_graph(17, width=10)
    arch = gnntf.GCN(gnntf.graph2adj(G), X, num_classes=3, latent_dims=[16], layer_type=gnntf.GCNSpectralPreservingLayer,
                     preprocessor=gnntf.Structural(dims=6))
    arch.reset()
    with torch.no_grad():
        for w in arch.vars():
            if w.var.shape[0] == 1:
                w.var.uniform_(-0.3, 0.3)                   # non-zero biases so that the "- b" term matters
    arch.training_mode(False)
    out = _np(arch(arch.features))
    st = arch.layers()[0]
    assert isinstance(st, gnntf.Structural) and st.output_shape == (n, 16)
    Xp = np.concatenate([_np(st.embeddings2), X], axis=1)
    idx, val, _ = oracle.graph2adj(G)
    _, nv, _ = oracle.get_adjacency(idx, val, n)
    H = Xp.astype(np.float32)
    for layer in arch.layers()[1:]:
        W, b = _np(layer.W), _np(layer.b)
        H = 2 * (oracle.relu(oracle.spmm_coo(idx, nv, H) @ W + b) - b)       # gcn.py:104-105, eval mode
    oracle.assert_close(out, H, what="GCN spectral-preserving + Structural", floor=oracle.FLOOR_REORDERED)


# ------------------------------------------------------------------------------------------
# Training mode without K materialised adjacencies
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F,K", [(7, 3), (48, 10), (100, 2)])
def test_masked_propagation_recomputed_from_masks_equals_materialised_adjacencies(F, K):
    gnntf = _gnntf()
    rng = np.random.default_rng(F + K)
    n, e = 900, 9000
    edges, w = _random_edges(n, e, seed=F)
    adj = gnntf.edges2adj(edges, w, n)
    masks = [torch.from_numpy(rng.random(adj.n_graph) >= 0.5).cuda() for _ in range(K)]
    H0a = torch.from_numpy(rng.standard_normal((n, F)).astype(np.float32)).cuda().requires_grad_(True)
    H0b = H0a.detach().clone().requires_grad_(True)
    g = torch.from_numpy(rng.standard_normal((n, F)).astype(np.float32)).cuda()
    out_a = gnntf.ops.appnp_propagate_masked(adj, masks, 0.5, H0a, 0.1)
    out_b = gnntf.appnp_propagate([adj.normalized("symmetric", keep_mask=m, rate=0.5) for m in masks], H0b, 0.1, K)
    assert torch.equal(out_a, out_b)                                     # same kernels on the same values
    out_a.backward(g)
    out_b.backward(g)
    oracle.assert_close(_np(H0a.grad), _np(H0b.grad), what=f"masked backward F={F} K={K}", floor=oracle.FLOOR_REORDERED)
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n)
    keeps = [_np(m) for m in masks]
    expect = oracle.appnp_propagate(idx, val, n, _np(H0a.detach()), 0.1, K, 0.5, keeps, training=True)[-1]
    oracle.assert_close(_np(out_a), expect, what=f"masked forward vs oracle F={F} K={K}")
    nvs = [oracle.get_adjacency(idx, oracle.sparse_dropout(val, 0.5, k), n)[1] for k in keeps]
    oracle.assert_close(_np(H0a.grad), oracle.appnp_propagate_bwd(idx, nvs, _np(g), 0.1), what="masked dH0 vs oracle",
                        floor=oracle.FLOOR_REORDERED)
