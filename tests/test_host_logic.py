"""CPU tests of the host-side mirror that need no GPU: the layer protocol, variable initialisers,
train/eval mode switching, the training loop on a graph-free architecture, the networkx walk of
graph2indices, the task head, and the synthetic generators."""
import numpy as np
import pytest
import torch

import gnntf
import gnntf_oracle as oracle
import synthetic


def test_graph2indices_follows_networkx_order():
    import networkx as nx
    G = nx.DiGraph()
    for u in ["c", "a", "b", "d", "iso"]:
        G.add_node(u)
    for u, v in [("a", "b"), ("b", "a"), ("c", "d"), ("d", "d"), ("a", "c")]:
        G.add_edge(u, v)
    assert gnntf.graph2indices(G) == [[0, 3], [1, 2], [1, 0], [2, 1], [3, 3]] == oracle.graph2indices(G)
    H = gnntf.create_nx_graph(["x", "y"], [("y", "x")])
    assert list(H) == ["x", "y"] and gnntf.graph2indices(H) == [[1, 0]]


def test_layer_protocol_and_mode_switch():
    arch = gnntf.Layered((10, 6))
    assert arch.is_training() and arch.top_shape() == (10, 6)          # layered.py:9
    d = arch.add(gnntf.Dense(4, activation=torch.relu, dropout=0.5))
    arch.add(gnntf.Dropout(0.5))
    assert arch.top_shape() == (10, 4) and arch.top_layer().output_shape == (10, 4)
    assert [tuple(w.var.shape) for w in arch.vars()] == [(6, 4), (1, 4)] and len(d.vars) == 2
    arch.reset()
    W = arch.vars()[0].numpy()
    assert np.abs(W).max() <= 0.5 + 1e-6 and np.abs(W).max() > 0.2       # 'small' = U(±1/sqrt(4)), variables.py:32-34
    assert np.all(arch.vars()[1].numpy() == 0)
    X = torch.ones((10, 6))
    with arch as variables:
        assert arch.is_training() and len(variables) == 2
    assert not arch.is_training()                                        # layered.py:41-42
    out1, out2 = arch(X), arch(X)
    assert torch.equal(out1, out2) and d.value.shape == (10, 4)          # eval: dropout is the identity
    arch.training_mode(True)
    assert not torch.equal(arch(X), arch(X)) or float(out1.abs().sum()) == 0

    class Bad(gnntf.Layer):
        def __build__(self, architecture):
            return None
    with pytest.raises(Exception, match="output shape"):
        arch.add(Bad())
    with pytest.raises(Exception, match="Invalid normalization type"):
        gnntf.WrappedVariable((2, 2), "nope").reset()


def test_node_classification_head():
    logits = torch.tensor([[2.0, 1.0, 0.0], [0.0, 3.0, 0.0], [0.0, 0.0, 1.0], [5.0, 0.0, 0.0]])
    task = gnntf.NodeClassification([0, 1, 3], np.array([0, 1, 2]))
    assert task.predict(logits).tolist() == [0, 1, 0]
    assert abs(task.evaluate(logits) - 2 / 3) < 1e-12
    lp = torch.log_softmax(logits[[0, 1, 3]], 1)
    expect = -(lp[0, 0] + lp[1, 1] + lp[2, 2]) / 3
    assert abs(float(task.loss(logits)) - float(expect)) < 1e-6
    with pytest.raises(Exception, match="requires node labels"):
        gnntf.NodeClassification([0]).loss(logits)
    assert gnntf.acc(torch.tensor([1, 2, 3]), np.array([1, 0, 3])) == pytest.approx(2 / 3)


def test_training_loop_early_stopping_and_best_weight_restore():
    gnntf.set_seed(0)
    rng = np.random.default_rng(0)
    n, classes = 300, 3
    labels = rng.integers(0, classes, n)
    X = rng.standard_normal((n, 8)).astype(np.float32)
    X[np.arange(n), labels] += 3.0
    arch = gnntf.Trainable(X)
    arch.add(gnntf.Dense(16, activation=torch.relu, dropout=0.2))
    arch.add(gnntf.Dense(classes, regularize=False))
    tr, va = np.arange(0, 200), np.arange(200, 300)
    arch.train(train=gnntf.NodeClassification(tr, labels[tr]), valid=gnntf.NodeClassification(va, labels[va]),
               patience=15, epochs=300)
    assert not arch.is_training()
    assert arch.evaluate(gnntf.NodeClassification(va, labels[va])) > 0.9
    pred = arch.predict(gnntf.NodeClassification(va))
    assert gnntf.acc(pred, labels[va]) > 0.9
    first = arch._fast_predict
    arch.predict(gnntf.NodeClassification(tr))
    assert arch._fast_predict is first                                   # cached until reset(), trainable.py:26-29
    arch.reset()
    assert arch._fast_predict is None


def test_synthetic_shapes():
    n, e, _, _ = synthetic.SHAPES["cora"]
    G = synthetic.citation_graph(n, e, seed=0)
    assert G.number_of_nodes() == n and G.number_of_edges() == e
    idx, val, shape = oracle.graph2adj(G)
    assert idx.shape == (2 * e, 2) and shape == (n, n)
    und = {tuple(sorted(p)) for p in idx.tolist()}
    assert len(und) == e // 2                                            # every entry present twice
    n2, edges = synthetic.shaped_edges("arxiv", seed=0)
    assert edges.shape == (synthetic.SHAPES["arxiv"][1], 2) and int(edges.max()) < n2 and int(edges.min()) >= 0
    assert not bool((edges[:, 0] == edges[:, 1]).any())
    deg = torch.bincount(edges.flatten(), minlength=n2)
    assert 8000 < int(deg.max()) < 20000
    a = synthetic.shaped_edges("arxiv", seed=0)[1]
    assert torch.equal(a, edges)                                         # deterministic
    r = synthetic.shaped_edges("arxiv", seed=0, ordering="random")[1]
    assert (r[:, 0] - r[:, 1]).abs().float().median() > 20 * (edges[:, 0] - edges[:, 1]).abs().float().median()
    nr, er = synthetic.rmat_edges(12, 50000, seed=1)
    assert nr == 4096 and er.shape == (50000, 2) and int(er.max()) < nr


def test_reorder_rank_uniformisation_is_a_permutation():
    from gnntf.reorder import _uniformize
    theta = torch.tensor([3.0, 0.1, 6.0, 2.0, 0.05])
    ranks, order = _uniformize(theta)
    assert order.tolist() == [4, 1, 3, 0, 2]
    assert torch.allclose(ranks, torch.tensor([3, 1, 4, 2, 0], dtype=torch.float32) * (2 * np.pi / 5))


def test_pitch_for_row_pitch_rule():
    """ops.pitch_for: the leading dimension of sharded feature buffers — rounded up to 32 floats only when that
    saves >= 10 % of the 128-byte lines a row gather touches and costs <= 35 % more memory."""
    from gnntf.ops import pitch_for
    assert pitch_for(52) == 64          # 208-byte rows: 2.5 lines on average at a dense pitch, 2 at 256 bytes
    assert pitch_for(100) == 100        # 400-byte rows touch 4 lines either way
    assert pitch_for(48) == 48 and pitch_for(40) == 40 and pitch_for(16) == 16
    assert pitch_for(64) == 64 and pitch_for(128) == 128 and pitch_for(256) == 256
    assert pitch_for(7) == 8 and pitch_for(47) == 48      # class widths: padded to float4 first
    for F in range(1, 300):
        ld = pitch_for(F)
        assert ld >= F and ld % 4 == 0 and ld <= 1.35 * ((F + 3) // 4 * 4) + 1e-9
