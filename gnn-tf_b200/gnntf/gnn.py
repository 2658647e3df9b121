"""Host-side mirror of gnntf's graph layer: the ``GNN`` base class, the APPNP / GCN architectures
and the ``NodeClassification`` task, wired to the native propagation ops.

Reference interfaces mirrored:
  * ``GNN`` / ``get_adjacency``        gnntf/core/gnn/gnn.py:29-50
  * ``PPRIteration`` / ``APPNP``       gnntf/core/gnn/architectures/filter.py:6-35
  * ``GCNLayer`` / ``GCN``             gnntf/core/gnn/architectures/gcn.py:77-113
  * ``NodeClassification``             gnntf/core/gnn/graph_predictor.py:10-31
What changes underneath: ``get_adjacency`` + ``tf.sparse.sparse_dense_matmul`` + the teleport
axpy run as fused sm_100a kernels, the eval-mode normalisation is computed once instead of once
per layer per call, and a run of K ``PPRIteration`` layers executes as one fused K-step op with a
fused backward.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as tfn

from . import ops
from .nn import Dense, Dropout, Layer, Predictor, Trainable, as_tensor, identity, relu
from .sparse import SparseAdjacency


class MaskedAdjacency:
    """What ``Layered.sparse_dropout`` returns in training mode (layered.py:50): the raw adjacency
    plus this call's edge keep-mask (COO order); normalisation folds the mask in."""

    def __init__(self, base: SparseAdjacency, keep, rate):
        self.base, self.keep, self.rate = base, keep, rate
        self.indices, self.dense_shape, self.shape = base.indices, base.dense_shape, base.shape

    @property
    def values(self):
        scale = torch.tensor(1.0 / (1.0 - self.rate), dtype=torch.float32, device=self.base.values.device)
        return torch.where(self.keep, self.base.values * scale, torch.zeros_like(self.base.values))


class GNN(Trainable):
    """gnn.py:29-50."""

    def __init__(self, graph, features, preprocessor: Layer = None):
        super().__init__(features)
        if not isinstance(graph, SparseAdjacency):
            raise Exception("graph must be the adjacency returned by gnntf.graph2adj")
        self.graph = graph
        if preprocessor is not None:
            self.add(preprocessor)

    def get_adjacency(self, graph_dropout=0.5, normalized="symmetric", add_eye="none"):
        graph = self.sparse_dropout(self.graph, graph_dropout)            # gnn.py:37
        if isinstance(graph, MaskedAdjacency):
            return self.graph.normalized(normalized, add_eye, keep_mask=graph.keep, rate=graph.rate)
        return self.graph.normalized(normalized, add_eye)                 # gnn.py:38-50 (cached in eval)

    def __call__(self, features):
        """``Layered.__call__`` (layered.py:52-55) with one addition: a maximal run of consecutive
        ``PPRIteration`` layers that share H0 / a / graph_dropout and use the defaults (no feature
        dropout, identity activation) runs as ONE fused K-step op."""
        layers = self.layers()
        i = 0
        while i < len(layers):
            layer = layers[i]
            j = i
            if isinstance(layer, PPRIteration) and layer.fusable(features):
                while j + 1 < len(layers) and isinstance(layers[j + 1], PPRIteration) and layers[j + 1].same_run(layer):
                    j += 1
            if j > i:
                features = PPRIteration.forward_run(self, layers[i:j + 1], features)
            else:
                features = layer(self, features)
            i = j + 1
        return features


class PPRIteration(Layer):
    """filter.py:6-22 — ``act(dropout((1-a)·Â·H + a·H0))`` with Â re-drawn (edge dropout) per call."""

    def __build__(self, architecture: GNN, H0: Layer, restart_probability: float = 0.1, activation=identity,
                  dropout: float = 0, graph_dropout: float = 0.5, restart_transform=identity):
        self.restart_probability = restart_probability
        self.H0 = H0
        self.dropout = dropout
        self.graph_dropout = graph_dropout
        self.activation = activation
        self.restart_transform = restart_transform
        return architecture.top_shape()

    def _alpha(self):
        return self.restart_transform(self.restart_probability)

    def __forward__(self, architecture: GNN, features):
        self.G = architecture.get_adjacency(self.graph_dropout)             # filter.py:18
        a = self._alpha()
        if isinstance(a, torch.Tensor):  # trainable restart probability: keep it differentiable
            propagated = ops.sparse_dense_matmul(self.G, features)         # filter.py:19
            out = propagated * (1 - a) + self.H0.value * a                 # filter.py:21
            return self.activation(architecture.dropout(out, self.dropout))
        keep = None
        if architecture.is_training() and self.dropout != 0:
            keep = torch.rand(features.shape, device=features.device) >= float(self.dropout)
        fused_relu = self.activation in (relu, torch.relu, tfn.relu)
        out = ops.appnp_step(self.G, features, self.H0.value, a, keep, self.dropout, fused_relu)
        return out if (fused_relu or self.activation is identity) else self.activation(out)

    # -- K-run fusion ----------------------------------------------------------------------
    def fusable(self, features):
        # output_regularize != 0 needs this layer's own .value (Layer.loss): such layers run un-fused
        return (self.dropout == 0 and self.activation is identity and self.restart_transform is identity
                and self.output_regularize == 0
                and not isinstance(self.restart_probability, torch.Tensor)
                and getattr(self.H0, "value", None) is features)

    def same_run(self, first):
        return (self.H0 is first.H0 and self.dropout == 0 and self.activation is identity
                and self.output_regularize == 0
                and self.restart_transform is identity and self.restart_probability == first.restart_probability
                and self.graph_dropout == first.graph_dropout)

    @staticmethod
    def forward_run(architecture: GNN, run, features):
        first = run[0]
        K = len(run)
        if architecture.is_training() and first.graph_dropout != 0:
            adjs = [architecture.get_adjacency(first.graph_dropout) for _ in range(K)]  # one mask per iteration
            for layer, G in zip(run, adjs):
                layer.G = G
        else:
            adjs = architecture.get_adjacency(first.graph_dropout)
            for layer in run:
                layer.G = adjs
        out = ops.appnp_propagate(adjs, features, first.restart_probability, K)
        for layer in run[:-1]:  # intermediate iterates are not materialised by the fused op:
            layer.value = None  # a stale value from an earlier un-fused call must not survive
        run[-1].value = out
        return out


class APPNP(GNN):
    """filter.py:25-35 — Dropout(0.5) → Dense(latent, relu, dropout)… → H0 = Dense(num_classes) → K × PPRIteration."""

    def __init__(self, G, features, num_classes: int, a: float = 0.1, latent_dims=[64], iterations=10,
                 dropout=0.6, graph_dropout=0.5, activation=identity, **kwargs):
        super().__init__(G, features, **kwargs)
        self.add(Dropout(0.5))
        for latent_dim in latent_dims:
            self.add(Dense(latent_dim, activation=relu, dropout=dropout))
        H0 = self.add(Dense(num_classes, regularize=False))
        if a is None:
            raise Exception("APPNP(a=None) is broken in the reference (create_var() without a shape); pass a float")
        for _ in range(iterations):
            self.add(PPRIteration(H0, a, graph_dropout=graph_dropout, activation=activation))


class GCNLayer(Layer):
    """gcn.py:77-89 — aggregate first at the INPUT width, then the dense transform."""

    def __build__(self, gcn, outputs: int, activation=relu, bias: bool = True, dropout: float = 0,
                  graph_dropout: float = 0):
        self.W = gcn.create_var((gcn.top_shape()[1], outputs))
        self.b = gcn.create_var((1, outputs), "zero") if bias else 0
        self.activation = activation
        self.dropout = dropout
        self.graph_dropout = graph_dropout
        return (gcn.top_shape()[0], outputs)

    def __forward__(self, gcn, features):
        adjacency = gcn.get_adjacency(self.graph_dropout)
        if self.W.shape[1] < features.shape[1]:
            # (Â·X)·W == Â·(X·W): when the layer narrows, propagate at the OUTPUT width (PubMed layer 1:
            # SpMM at 64 columns instead of 500).  Same value up to fp32 rounding.
            transformed = ops.sparse_dense_matmul(adjacency, features @ self.W)
        else:
            transformed = ops.sparse_dense_matmul(adjacency, features) @ self.W                  # gcn.py:88
        return gcn.dropout(self.activation(transformed + self.b), self.dropout)                  # gcn.py:89


class GCN(GNN):
    """gcn.py:108-113 — the output layer keeps GCNLayer's default relu (a reference quirk)."""

    def __init__(self, G, features, num_classes, latent_dims=[64], layer_type=GCNLayer, **kwargs):
        super().__init__(G, features, **kwargs)
        for latent_dim in latent_dims:
            self.add(layer_type(latent_dim, graph_dropout=0.5, dropout=0.5))
        self.add(layer_type(num_classes))


class NodeClassification(Predictor):
    """graph_predictor.py:10-31."""

    def __init__(self, nodes, labels=None, loss_transform=None):
        self.nodes = nodes
        self.labels = labels
        self.loss_transform = loss_transform

    def _rows(self, features):
        idx = as_tensor(self.nodes, dtype=torch.long, device=features.device)
        return features.index_select(0, idx)  # tf.nn.embedding_lookup

    def _labels(self, device):
        return as_tensor(self.labels, dtype=torch.long, device=device)

    def predict(self, features):
        return torch.argmax(self._rows(features), dim=1)

    def loss(self, features):
        if self.labels is None:
            raise Exception("Evaluation requires node labels")
        if self.loss_transform is not None:
            features = self.loss_transform(features)
        predictions = tfn.log_softmax(self._rows(features), dim=1)
        return tfn.cross_entropy(predictions, self._labels(features.device))  # from_logits CE on log-probs

    def evaluate(self, features):
        if self.labels is None:
            raise Exception("Evaluation requires node labels")
        predictions = torch.argmax(self._rows(features), dim=1)
        wrong = torch.count_nonzero(predictions - self._labels(features.device)).item()
        return 1 - wrong / predictions.shape[0]
