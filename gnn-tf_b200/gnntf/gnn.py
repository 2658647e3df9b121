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
from .nn import (Concatenate, Dense, Dropout, Layer, Predictor, Trainable, as_tensor, dense_tail, identity, leaky_relu,
                 relu)
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
        keep = architecture.dropout_mask(features.shape, self.dropout, features.device)
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
            # one edge mask per iteration (filter.py:18); the adjacencies themselves are never materialised
            drawn = [architecture.sparse_dropout(architecture.graph, first.graph_dropout) for _ in range(K)]
            for layer, G in zip(run, drawn):
                layer.G = G                      # the masked (un-normalised) adjacency of this iteration
            out = ops.appnp_propagate_masked(architecture.graph, [G.keep for G in drawn], first.graph_dropout, features,
                                             first.restart_probability)
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
        if _cacheable(gcn, adjacency, features):
            transformed = aggregate_cached(gcn, adjacency, features) @ self.W                          # gcn.py:88, Â·X kept
        elif self.W.shape[1] < features.shape[1]:
            # (Â·X)·W == Â·(X·W): when the layer narrows, propagate at the OUTPUT width (PubMed layer 1 in
            # training: SpMM at 64 columns instead of 500).  Same value up to fp32 rounding (SURVEY §8f-1;
            # measured against the oracle's (ÂX)W in tests/test_gpu_fullsize.py).
            transformed = ops.sparse_dense_matmul(adjacency, features @ self.W)
        else:
            transformed = ops.sparse_dense_matmul(adjacency, features) @ self.W                        # gcn.py:88
        return dense_tail(gcn, transformed, self.b, self.activation, self.dropout)                     # gcn.py:89


def _cacheable(gcn, adjacency, features):
    return features is gcn.features and not adjacency.has_mask and not features.requires_grad


def aggregate_cached(gcn, adjacency, features):
    """``Â·X`` (gcn.py:88).  When X is the architecture's constant input matrix and Â carries no edge
    mask (eval mode or graph_dropout = 0) the product is the same in every forward of every epoch: it
    is computed once and kept on the adjacency (SURVEY §8f-1 "cache ÂX across epochs")."""
    if _cacheable(gcn, adjacency, features):
        cache = getattr(adjacency, "_agg_cache", None)
        if cache is None or cache[0] is not features:
            adjacency._agg_cache = (features, ops.sparse_dense_matmul(adjacency, features).detach())
        return adjacency._agg_cache[1]
    return ops.sparse_dense_matmul(adjacency, features)


class GCNSpectralPreservingLayer(GCNLayer):
    """gcn.py:93-105 — ``2·dropout(act((Â·X)·W + b) − b)``."""

    def __forward__(self, gcn, features):
        adjacency = gcn.get_adjacency(self.graph_dropout)
        transformed = aggregate_cached(gcn, adjacency, features) @ self.W + self.b                       # gcn.py:104
        return 2 * gcn.dropout(self.activation(transformed) - self.b, self.dropout)                     # gcn.py:105


class GCN(GNN):
    """gcn.py:108-113 — the output layer keeps GCNLayer's default relu (a reference quirk)."""

    def __init__(self, G, features, num_classes, latent_dims=[64], layer_type=GCNLayer, **kwargs):
        super().__init__(G, features, **kwargs)
        for latent_dim in latent_dims:
            self.add(layer_type(latent_dim, graph_dropout=0.5, dropout=0.5))
        self.add(layer_type(num_classes))


class Structural(Layer):
    """gnn.py:5-26 — trainable per-node embeddings concatenated in front of the input features
    (``GNN(preprocessor=Structural(...))``)."""

    def __build__(self, architecture, dims: int = 16, l2_contraint: bool = False, bipartite: int = 0, **kwargs):
        top_shape = architecture.top_shape()
        self.l2_contraint = l2_contraint
        self.embeddings = architecture.create_var((bipartite, dims), **kwargs)
        self.embeddings2 = architecture.create_var((top_shape[0] - bipartite, dims), **kwargs)
        return top_shape[0], dims + top_shape[1]

    def __forward__(self, architecture, features):
        embeddings = self.embeddings2
        if self.embeddings.shape[0] != 0:
            embeddings = torch.cat([self.embeddings, embeddings], dim=0)
        if self.l2_contraint:
            embeddings = l2_normalize(embeddings)
        if features.shape[0] == 0:
            return embeddings
        return torch.cat([embeddings, features], dim=1)


def l2_normalize(x, eps=1e-12):
    """``tf.math.l2_normalize(x, axis=1)``: x · rsqrt(max(Σx², eps))."""
    return x * torch.rsqrt(torch.clamp((x * x).sum(dim=1, keepdim=True), min=eps))


def log1p(x):
    """``tf.math.log1p`` on a Python float (the GCNII beta transformer default, gcn.py:9)."""
    import math
    return math.log1p(x)


class GCNIILayer(Layer):
    """gcn.py:7-27 — ``dropout(act(((1−a)·Â·X + a·H0) · ((1−b)·I + b·W)))``, b = beta_transformer(l/(k+1)).
    The teleport mix is the fused PPR step kernel (same expression as filter.py:21)."""

    def __build__(self, architecture, H0: Layer, a: float, l: float, k: int = 0, activation=identity,
                  beta_transformer=log1p, dropout: float = 0.5, graph_dropout: float = 0.5, regularization=True):
        dim = architecture.top_shape()[1]
        self.W = architecture.create_var((dim, dim), "zero", regularize=regularization)
        self.a, self.l, self.k = a, l, k
        self.activation = activation
        self.dropout = dropout
        self.graph_dropout = graph_dropout
        self.H0 = H0
        self.beta_transformer = beta_transformer
        return architecture.top_shape()

    def _mix(self, gcn, features):
        b = float(self.beta_transformer(self.l / (self.k + 1)))                                          # gcn.py:23
        tradeoff = ops.appnp_step(gcn.get_adjacency(self.graph_dropout), features, self.H0.value, self.a)  # :24-25
        eye = torch.eye(self.W.shape[1], dtype=self.W.dtype, device=self.W.device)
        return tradeoff @ ((1 - b) * eye + b * self.W)                                                    # :26

    def __forward__(self, gcn, features):
        return dense_tail(gcn, self._mix(gcn, features), None, self.activation, self.dropout)             # gcn.py:27


class GCNIISpectralPreservingLayer(GCNIILayer):
    """gcn.py:30-51 — the same mix with a bias inside the activation and removed after it, doubled."""

    def __build__(self, architecture, H0: Layer, a: float, l: float, k: int = 0, activation=identity,
                  beta_transformer=log1p, dropout: float = 0.5, graph_dropout: float = 0.5, regularization=True):
        shape = super().__build__(architecture, H0, a, l, k, activation, beta_transformer, dropout, graph_dropout,
                                  regularization)
        self.bias = architecture.create_var((1, architecture.top_shape()[1]), "zero")
        return shape

    def __forward__(self, gcn, features):
        activation = self._mix(gcn, features) + self.bias                                                # gcn.py:50
        return 2 * gcn.dropout(self.activation(activation) - self.bias, self.dropout)                     # gcn.py:51


class GCNII(GNN):
    """gcn.py:54-74 — Dropout → Dense(latent, relu)… → ``iterations`` × GCNIILayer(H0) → Dense(num_classes)."""

    def __init__(self, graph, features, num_classes, a: float = 0.1, l: float = 0.5, latent_dims=[64], iterations=64,
                 dropout=0.6, convolution_regularization=True, layer_type=GCNIILayer, **kwargs):
        super().__init__(graph, features, **kwargs)
        self.add(Dropout(dropout))
        for latent_dim in latent_dims:
            self.add(Dense(latent_dim, dropout=0, activation=relu))
        H0 = self.top_layer()
        for iteration in range(iterations):
            self.add(layer_type(H0, a, l, iteration, activation=relu, dropout=dropout, graph_dropout=0,
                                regularization=convolution_regularization))
        self.add(Dense(num_classes, dropout=0, regularize=False))


class NGCFLayer(Layer):
    """gcn.py:116-135 — ``l2_normalize(dropout(act((X ∘ ÂX)·W1 + b1) + act((ÂX)·W2 + b2)))`` with the
    bipartite-normalised adjacency drawn ONCE at build time (:127; with node_dropout ≠ 0 the mask drawn
    then is kept for the layer's life, as in the reference)."""

    def __build__(self, gcn, outputs: int, activation=leaky_relu, bias: bool = True, dropout: float = 0,
                  node_dropout: float = 0, regularize: float = 1):
        fan_in = gcn.top_shape()[0]
        bound = 1. / fan_in ** 0.5
        self.W1 = gcn.create_var((gcn.top_shape()[1], outputs), regularize=regularize, normalization=bound)
        self.W2 = gcn.create_var((gcn.top_shape()[1], outputs), regularize=regularize, normalization=bound)
        self.b1 = gcn.create_var((1, outputs), normalization=bound) if bias else 0
        self.b2 = gcn.create_var((1, outputs), normalization=bound) if bias else 0
        self.activation = activation
        self.dropout = dropout
        self.node_dropout = node_dropout
        self.adjacency = gcn.get_adjacency(self.node_dropout, add_eye="none", normalized="bipartite")    # gcn.py:127
        return (gcn.top_shape()[0], outputs)

    def __forward__(self, gcn, features):
        aggregated = ops.sparse_dense_matmul(self.adjacency, features)                                    # gcn.py:131
        output = dense_tail(gcn, (features * aggregated) @ self.W1, self.b1, self.activation, 0) \
            + dense_tail(gcn, aggregated @ self.W2, self.b2, self.activation, 0)                          # gcn.py:133-134
        return l2_normalize(gcn.dropout(output, self.dropout))                                            # gcn.py:135


class NGCF(GNN):
    """gcn.py:138-153."""

    def __init__(self, graph, features, num_classes: int, latent_dims=None, dropout=0.1, **kwargs):
        super().__init__(graph, features, **kwargs)
        if latent_dims is None:
            latent_dims = [num_classes] * 2
        layers = list()
        for latent_dim in latent_dims:
            layers.append(self.add(NGCFLayer(latent_dim, regularize=0.0, dropout=dropout, output_regularize=1)))
        layers.append(self.add(NGCFLayer(num_classes, regularize=0.0, dropout=dropout, output_regularize=1)))
        self.add(Concatenate(layers))


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
        if features.is_cuda and features.dtype == torch.float32 and features.dim() == 2:
            # embedding_lookup + log_softmax + sparse CE in one native kernel (csrc/dense.cu)
            return ops.node_cross_entropy(features, as_tensor(self.nodes, dtype=torch.long, device=features.device),
                                          self._labels(features.device))
        predictions = tfn.log_softmax(self._rows(features), dim=1)
        return tfn.cross_entropy(predictions, self._labels(features.device))  # from_logits CE on log-probs

    def evaluate(self, features):
        if self.labels is None:
            raise Exception("Evaluation requires node labels")
        predictions = torch.argmax(self._rows(features), dim=1)
        wrong = torch.count_nonzero(predictions - self._labels(features.device)).item()
        return 1 - wrong / predictions.shape[0]
