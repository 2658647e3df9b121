"""CPU suite for the oracle itself: the NumPy restatement against the survey-derived known-answer
tests (KAT-1 / KAT-2, SURVEY.md §8c — the reference ships no golden vectors), against the
committed fixtures under tests/golden/, against independent scipy.sparse arithmetic, against its
own fp64 twin, and the C restatement against the NumPy one."""
import ctypes
import json
import os

import numpy as np
import pytest

import gnntf_oracle as oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def kat_graph():
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


def test_kat1_graph2adj_and_normalisation():
    G = kat_graph()
    assert oracle.graph2indices(G) == [[0, 3], [1, 2], [1, 0], [2, 1], [3, 3]]
    idx, val, shape = oracle.graph2adj(G)
    assert idx.tolist() == [[0, 3], [1, 2], [1, 0], [2, 1], [3, 3], [3, 0], [2, 1], [0, 1], [1, 2], [3, 3]]
    assert val.tolist() == [2.5, 1, 1, 1, 1, 2.5, 1, 1, 1, 1] and shape == (5, 5)
    assert oracle.column_sums(idx, val, 5).tolist() == [3.5, 3, 2, 4.5, 0]
    _, nv, D = oracle.get_adjacency(idx, val, 5)
    np.testing.assert_allclose(D, np.float32([0.5345225, 0.57735026, 0.70710677, 0.47140455, 0]), rtol=1e-7)
    np.testing.assert_allclose(nv, np.float32([0.6299408, 0.40824828, 0.30860668, 0.40824828, 0.22222225] * 2), rtol=1e-7)


def test_kat1_appnp_k10_fp64_and_fp32_drift():
    idx, val, _ = oracle.graph2adj(kat_graph())
    H0 = np.array([[(3 * i + j) / 7 - 1 for j in range(3)] for i in range(5)])
    h64 = oracle.appnp_propagate(idx, val, 5, H0, 0.1, 10, dtype=np.float64)[-1]
    expect = np.array([[-0.386717034, -0.242003374, -0.097289714], [-0.37323383, -0.233458343, -0.093682856],
                       [-0.301202958, -0.181881658, -0.062560359], [-0.311411903, -0.151985331, 0.007441241],
                       [0.071428571, 0.085714286, 0.1]])
    np.testing.assert_allclose(h64, expect, atol=2e-9)
    h32 = oracle.appnp_propagate(idx, val, 5, H0, 0.1, 10)[-1]
    assert np.abs(h32 - h64).max() < 2e-6


def test_kat2_masked_normalisation_is_not_symmetric():
    idx, val, _ = oracle.graph2adj(kat_graph())
    mv = oracle.sparse_dropout(val, 0.5, [1, 0, 1, 1, 0, 1, 1, 0, 1, 1])
    assert mv.tolist() == [5, 0, 2, 2, 0, 5, 2, 0, 2, 2]
    assert oracle.column_sums(idx, mv, 5).tolist() == [7, 4, 2, 7, 0]
    _, nv, D = oracle.get_adjacency(idx, mv, 5)
    np.testing.assert_allclose(D, np.float32([0.3779645, 0.5, 0.70710677, 0.3779645, 0]), rtol=1e-7)
    np.testing.assert_allclose(nv, np.float32([0.7142858, 0, 0.3779645, 0.70710677, 0, 0.7142858, 0.70710677, 0,
                                               0.70710677, 0.28571433]), rtol=2e-7)
    dense = np.zeros((5, 5))
    np.add.at(dense, (idx[:, 0], idx[:, 1]), nv)
    assert dense[1, 0] > 0.37 and dense[0, 1] == 0 and abs(dense[2, 1] - 1.4142135) < 1e-6


def test_golden_fixtures():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        kat = json.load(f)
    for case in kat["cases"]:
        idx = np.array(case["indices"], dtype=np.int64).reshape(-1, 2)
        val = np.array(case["values"], dtype=np.float32)
        n = case["n"]
        if case.get("keep") is not None:
            val = oracle.sparse_dropout(val, case["rate"], case["keep"])
        _, nv, _ = oracle.get_adjacency(idx, val, n, case["normalized"], case["add_eye"])
        np.testing.assert_allclose(nv, np.array(case["norm_values"], np.float32), rtol=3e-7, err_msg=case["name"])
        H0 = np.array(case["H0"], np.float32)
        out = oracle.appnp_propagate(idx, val, n, H0, case["alpha"], case["K"])[-1] if case["normalized"] == "symmetric" and case["add_eye"] == "none" else None
        if out is not None:
            oracle.assert_close(out, np.array(case["H_K"]), what=case["name"])


def test_sparse_ops_against_scipy():
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    n, e, F = 400, 5000, 17
    edges = rng.integers(0, n, (e, 2))
    w = rng.random(e).astype(np.float32) + 0.5
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n)
    A = sp.coo_matrix((val.astype(np.float64), (idx[:, 0], idx[:, 1])), shape=(n, n)).tocsr()  # coalesces duplicates
    deg = np.asarray(A.sum(axis=0)).ravel()
    np.testing.assert_allclose(oracle.column_sums(idx, val, n, np.float64), deg, rtol=1e-12)
    d = np.where(deg > 0, 1 / np.sqrt(np.where(deg > 0, deg, 1)), 0)
    Ahat = sp.diags(d) @ A @ sp.diags(d)
    _, nv, _ = oracle.get_adjacency(idx, val, n, dtype=np.float64)
    H = rng.standard_normal((n, F))
    np.testing.assert_allclose(oracle.spmm_coo(idx, nv, H, dtype=np.float64), Ahat @ H, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(oracle.spmm_coo_T(idx, nv, H, n, dtype=np.float64), Ahat.T @ H, rtol=1e-10, atol=1e-12)
    oracle.assert_close(oracle.spmm_coo(idx, nv.astype(np.float32), H.astype(np.float32)), Ahat @ H)
    # K-step loop == the closed-form polynomial
    a, K = 0.1, 6
    expect, P = np.zeros_like(H), H.copy()
    for j in range(K):
        expect += a * (1 - a) ** j * P
        P = Ahat @ P
    expect += (1 - a) ** K * P
    got = oracle.appnp_propagate(idx, val, n, H, a, K, dtype=np.float64)[-1]
    np.testing.assert_allclose(got, expect, rtol=1e-10, atol=1e-12)


def test_backward_is_the_adjoint():
    """<appnp(H0), G> == <H0, appnp_bwd(G)> with per-iteration masked (non-symmetric) adjacencies."""
    rng = np.random.default_rng(1)
    n, e, F, K, a = 120, 900, 5, 4, 0.15
    idx, val, _ = oracle.graph2adj_arrays(rng.integers(0, n, (e, 2)), None, n)
    keeps = [rng.random(idx.shape[0]) >= 0.5 for _ in range(K)]
    H0, G = rng.standard_normal((n, F)), rng.standard_normal((n, F))
    out = oracle.appnp_propagate(idx, val, n, H0, a, K, 0.5, keeps, training=True, dtype=np.float64)[-1]
    nvs = [oracle.get_adjacency(idx, oracle.sparse_dropout(val, 0.5, k), n, dtype=np.float64)[1] for k in keeps]
    dH0 = oracle.appnp_propagate_bwd(idx, nvs, G, a, dtype=np.float64)
    assert abs((out * G).sum() - (H0 * dH0).sum()) < 1e-9 * abs((out * G).sum())


def test_csr_view_is_a_stable_row_sort():
    rng = np.random.default_rng(2)
    n, e = 60, 500
    idx, val, _ = oracle.graph2adj_arrays(rng.integers(0, n, (e, 2)), None, n)
    row_ptr, col, pos, perm_T = oracle.csr_from_coo(idx, n)
    assert row_ptr[0] == 0 and row_ptr[-1] == 2 * e and np.all(np.diff(row_ptr) >= 0)
    rows = np.repeat(np.arange(n), np.diff(row_ptr))
    assert np.array_equal(idx[pos, 0], rows) and np.array_equal(idx[pos, 1], col)
    for r in range(n):
        seg = pos[row_ptr[r]:row_ptr[r + 1]]
        assert np.all(np.diff(seg) > 0)  # COO order kept inside each row
    assert np.array_equal(rows[perm_T], col) and np.array_equal(col[perm_T], rows)


def test_add_eye_and_modes():
    idx, val, _ = oracle.graph2adj(kat_graph())
    i2, v2, _ = oracle.get_adjacency(idx, val, 5, "symmetric", "before")
    assert i2.shape[0] == 15 and oracle.column_sums(i2[:10], val, 5).tolist() == [3.5, 3, 2, 4.5, 0]
    assert v2[14] == 1.0  # isolated node: deg 1 -> D = 1 -> value 1
    i3, v3, _ = oracle.get_adjacency(idx, val, 5, "bipartite", "after")
    assert v3[10:].tolist() == [1] * 5 and abs(v3[0] - 2.5 / 3.5) < 1e-7
    with pytest.raises(Exception, match="Invalid matrix normalization"):
        oracle.get_adjacency(idx, val, 5, "laplacian")


def test_c_oracle_matches_numpy_oracle(oracle_c):
    rng = np.random.default_rng(3)
    n, e, F, K = 3000, 40000, 47, 10
    idx, val, _ = oracle.graph2adj_arrays(rng.integers(0, n, (e, 2)), rng.random(e).astype(np.float32) + 0.5, n)
    H0 = rng.standard_normal((n, F)).astype(np.float32)
    nnz = idx.shape[0]
    out = np.empty_like(H0)
    scratch = np.empty(nnz + n + n * F, np.float32)
    P = ctypes.c_void_p
    oracle_c.oracle_appnp_propagate_f32.argtypes = [P, P, ctypes.c_int64, ctypes.c_int64, P, ctypes.c_int64,
                                                    ctypes.c_float, ctypes.c_int, ctypes.c_int, P, P]
    oracle_c.oracle_appnp_propagate_f32(idx.ctypes.data, val.ctypes.data, nnz, n, H0.ctypes.data, F,
                                        ctypes.c_float(0.1), K, 1, scratch.ctypes.data, out.ctypes.data)
    expect = oracle.appnp_propagate(idx, val, n, H0, 0.1, K)[-1]
    oracle.assert_close(out, expect, rtol=1e-6, what="C oracle vs NumPy oracle")
    _, nv, D = oracle.get_adjacency(idx, val, n)
    np.testing.assert_allclose(scratch[:nnz], nv, rtol=2e-7)


def test_c_oracle_multithreaded_csr_step_matches_numpy_oracle(oracle_c):
    """The row-parallel CSR step bench.py reports as multi-threaded CPU context."""
    rng = np.random.default_rng(4)
    n, e, F, a = 2000, 25000, 33, 0.1
    idx, val, _ = oracle.graph2adj_arrays(rng.integers(0, n, (e, 2)), rng.random(e).astype(np.float32) + 0.5, n)
    _, nv, _ = oracle.get_adjacency(idx, val, n)
    row_ptr, col, pos, _ = oracle.csr_from_coo(idx, n)
    csr_val = np.ascontiguousarray(nv[pos])
    H = rng.standard_normal((n, F)).astype(np.float32)
    H0 = rng.standard_normal((n, F)).astype(np.float32)
    out = np.zeros((n, F), np.float32)
    P = ctypes.c_void_p
    oracle_c.oracle_appnp_step_csr_omp_f32.argtypes = [P, P, P, P, P, ctypes.c_int64, ctypes.c_float, ctypes.c_int64,
                                                       ctypes.c_int64, P]
    oracle_c.oracle_appnp_step_csr_omp_f32(row_ptr.ctypes.data, col.ctypes.data, csr_val.ctypes.data, H.ctypes.data,
                                           H0.ctypes.data, F, ctypes.c_float(a), 0, n, out.ctypes.data)
    oracle.assert_close(out, oracle.ppr_iteration(idx, nv, H, H0, a), what="CSR/OpenMP step")


# ------------------------------------------------------------------------------------------
# The full-size checker (oracle/oracle_big.py): its CSR route must be the COO route, bit for bit
# ------------------------------------------------------------------------------------------
def _random_graph(n, e, seed, hub=False):
    rng = np.random.default_rng(seed)
    edges = rng.integers(0, n, size=(e, 2)).astype(np.int64)
    if hub:  # a few rows far longer than the GPU's split threshold
        edges[: e // 4, 0] = rng.integers(0, 3, size=e // 4)
    w = (rng.random(e) + 0.25).astype(np.float32)
    return edges, w


def test_big_oracle_csr_is_the_stable_row_sort_of_the_coo_list():
    import oracle_big
    edges, w = _random_graph(300, 4000, 0, hub=True)
    idx, val, _ = oracle.graph2adj_arrays(edges, w, 300)
    row_ptr, col, coo_pos = oracle_big.csr_from_coo(idx, 300)
    e_rp, e_col, e_pos, _ = oracle.csr_from_coo(idx, 300)
    assert np.array_equal(row_ptr, e_rp) and np.array_equal(col, e_col) and np.array_equal(coo_pos, e_pos)


@pytest.mark.parametrize("F", [1, 7, 40, 100])
def test_big_oracle_csr_route_is_bit_identical_to_the_coo_loop(F):
    """The claim oracle_big.py rests on: a row-wise loop over the stable CSR performs, for every
    output element, the same fp32 operations in the same order as TF-CPU's sequential COO loop."""
    import oracle_big
    L = oracle_big.lib()
    n, K, a = 500, 10, 0.1
    edges, w = _random_graph(n, 6000, 1, hub=True)
    big = oracle_big.BigOracle(edges, w, n, keep_idx=True)
    H0 = np.random.default_rng(2).standard_normal((n, F)).astype(np.float32)
    # normalisation and one SpMM, COO route
    nv = oracle.get_adjacency(big.idx, big.raw, n)[1]
    assert np.array_equal(big.norm_coo, nv)
    P = np.zeros((n, F), np.float32)
    L.oracle_spmm_coo_f32(big.idx.ctypes.data, big.norm_coo.ctypes.data, 0, big.nnz, H0.ctypes.data, F, P.ctypes.data)
    assert np.array_equal(big.spmm(H0), P)
    # K = 10 propagation, COO route (the C restatement of filter.py:17-22 under layered.py:52-55)
    scratch = np.empty(big.nnz + n + n * F, np.float32)
    out = np.empty((n, F), np.float32)
    L.oracle_appnp_propagate_f32(big.idx.ctypes.data, big.raw.ctypes.data, big.nnz, n, H0.ctypes.data, F,
                                 ctypes.c_float(a), K, 0, scratch.ctypes.data, out.ctypes.data)
    assert np.array_equal(big.propagate(H0, a, K), out)
    # and against the NumPy oracle within the fp32 tolerance
    oracle.assert_close(out, oracle.appnp_propagate(big.idx, big.raw, n, H0, a, K)[-1], what="C vs NumPy K=10")
    assert np.array_equal(big.step(H0, H0, a), big.propagate(H0, a, 1))


def test_oracle_variant_layers_against_dense_algebra():
    """GCNII / NGCF / loss restatements against the same formulas written with a dense adjacency."""
    rng = np.random.default_rng(5)
    n, F = 60, 7
    edges, w = _random_graph(n, 400, 3)
    idx, val, _ = oracle.graph2adj_arrays(edges, w, n)
    _, nv, _ = oracle.get_adjacency(idx, val, n, dtype=np.float64)
    A = np.zeros((n, n))
    np.add.at(A, (idx[:, 0], idx[:, 1]), nv)
    X, H0 = rng.standard_normal((n, F)), rng.standard_normal((n, F))
    W = rng.standard_normal((F, F)) * 0.3
    got = oracle.gcnii_layer(idx, nv, X, H0, W, 0.1, 0.5, 2, dtype=np.float64)
    b = np.log1p(0.5 / 3)
    np.testing.assert_allclose(got, np.maximum((0.9 * A @ X + 0.1 * H0) @ ((1 - b) * np.eye(F) + b * W), 0), rtol=1e-12, atol=1e-12)
    bias = rng.standard_normal((1, F)) * 0.1
    got = oracle.gcnii_layer(idx, nv, X, H0, W, 0.1, 0.5, 2, bias=bias, dtype=np.float64)
    np.testing.assert_allclose(got, 2 * (np.maximum((0.9 * A @ X + 0.1 * H0) @ ((1 - b) * np.eye(F) + b * W) + bias, 0) - bias),
                               rtol=1e-12, atol=1e-12)
    _, bv, _ = oracle.get_adjacency(idx, val, n, "bipartite", dtype=np.float64)
    Ab = np.zeros((n, n))
    np.add.at(Ab, (idx[:, 0], idx[:, 1]), bv)
    W1, W2 = rng.standard_normal((F, 4)), rng.standard_normal((F, 4))
    b1, b2 = rng.standard_normal((1, 4)), rng.standard_normal((1, 4))
    got = oracle.ngcf_layer(idx, bv, X, W1, b1, W2, b2, dtype=np.float64)
    agg = Ab @ X
    lr = lambda z: np.where(z > 0, z, 0.2 * z)  # noqa: E731
    ref = lr((X * agg) @ W1 + b1) + lr(agg @ W2 + b2)
    np.testing.assert_allclose(got, ref / np.linalg.norm(ref, axis=1, keepdims=True), rtol=1e-10, atol=1e-12)
    logits = rng.standard_normal((n, 5))
    nodes, labels = np.arange(0, 20), rng.integers(0, 5, 20)
    rows = logits[nodes]
    lse = np.log(np.exp(rows).sum(1))
    np.testing.assert_allclose(oracle.node_classification_loss(logits, nodes, labels, dtype=np.float64),
                               (lse - rows[np.arange(20), labels]).mean(), rtol=1e-12)
