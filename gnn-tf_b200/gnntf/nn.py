"""Host-side mirror of gnntf's generic NN scaffolding on torch tensors (thin Python, no native
work: these are the callers either side of the propagation path).

Reference interfaces mirrored (same names, argument meaning and error behaviour):
  * ``WrappedVariable`` / ``VariableGenerator``  gnntf/core/nn/variables.py:4-67
  * ``Layered`` / ``Layer``                      gnntf/core/nn/layered.py:5-86
  * ``Dense`` / ``Dropout`` / ``Activation``     gnntf/core/nn/layers.py:125-181
  * ``Predictor`` / ``Trainable``                gnntf/core/nn/trainable.py:5-103
Differences that are deliberate: variables are ``torch`` tensors with ``requires_grad`` instead
of ``tf.Variable``; the optimiser is ``torch.optim.Adam`` with Keras' defaults (eps 1e-7).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as tfn


def default_device():
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


def as_tensor(x, dtype=torch.float32, device=None):
    device = device or default_device()
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=device)


def identity(x):
    return x


def relu(x):
    return torch.relu(x)


def leaky_relu(x):
    """``tf.nn.leaky_relu`` (alpha = 0.2), the NGCFLayer default (gcn.py:117)."""
    return tfn.leaky_relu(x, 0.2)


# ------------------------------------------------------------------------------------------
# variables.py
# ------------------------------------------------------------------------------------------
class WrappedVariable:
    """variables.py:4-45.  ``normalization`` picks the initialiser applied by ``reset()``."""

    def __init__(self, shape, normalization="small", trainable=True, regularize=True, name=None):
        self.var = torch.zeros(tuple(shape), dtype=torch.float32, device=default_device(), requires_grad=bool(trainable))
        self.trainable = trainable
        self.regularize = float(regularize)
        self.name = name
        self.normalization = normalization

    def _uniform(self, bound):
        return (torch.rand(self.var.shape, device=self.var.device) * 2 - 1) * bound

    def reset(self):
        shape, kind = self.var.shape, self.normalization
        if isinstance(kind, float):
            new = self._uniform(kind)
        elif kind == "zero":
            new = torch.zeros(shape)
        elif kind == "eye":
            new = torch.eye(shape[1])
        elif kind == "ones":
            new = torch.ones(shape)
        elif kind == "xavier":
            new = self._uniform(math.sqrt(6.0 / (shape[0] + shape[1])))
        elif kind == "he":
            new = self._uniform(math.sqrt(6.0 / shape[0]))
        elif kind == "bernouli":
            new = (torch.round(torch.rand(shape)) * 2 - 1) / shape[1] ** 0.5
        elif kind == "small":  # the default: U(±1/sqrt(shape[1])), variables.py:32-34
            new = self._uniform(1.0 / (shape[1] ** 0.5))
        else:
            raise Exception("Invalid normalization type")
        self.assign(new)

    def identity(self):
        return self.var.detach().clone()

    def numpy(self):
        return self.var.detach().cpu().numpy()

    def assign(self, value):
        with torch.no_grad():
            self.var.copy_(torch.as_tensor(value).to(self.var.device, self.var.dtype).reshape(self.var.shape))

    def apply_gradient(self, optimizer, gradient):
        if gradient is None:
            return
        self.var.grad = gradient
        optimizer.step()


class VariableGenerator:
    """variables.py:48-67."""

    def __init__(self):
        self.__vars = []
        self.__named = {}

    def vars(self):
        return self.__vars

    def create_var(self, *args, shared_name=None, **kwargs):
        if shared_name is not None and shared_name in self.__named:
            return self.__named[shared_name]
        wrapped = WrappedVariable(*args, **kwargs)
        self.__vars.append(wrapped)
        if shared_name is not None:
            self.__named[shared_name] = wrapped.var
        return wrapped.var

    def reset(self):
        for wrapped in self.__vars:
            wrapped.reset()


# ------------------------------------------------------------------------------------------
# layered.py
# ------------------------------------------------------------------------------------------
class Layered(VariableGenerator):
    """Layer stack with a train/eval switch (layered.py:5-55).  Starts in TRAINING mode
    (layered.py:9) and drops to eval on leaving a ``with arch`` block (:41-42)."""

    def __init__(self, input_shape, layers=()):
        super().__init__()
        self.__layers = []
        self.__training = True
        self.input_shape = tuple(input_shape)
        for layer in layers:
            self.add(layer)

    def layers(self):
        return self.__layers

    def top_shape(self):
        return self.__layers[-1].output_shape if self.__layers else self.input_shape

    def top_layer(self):
        return self.__layers[-1]

    def add(self, layer):
        if layer not in self.__layers:
            layer.__late_init__(self)
        self.__layers.append(layer)
        return layer

    def is_training(self):
        return self.__training

    def training_mode(self, training_mode):
        self.__training = training_mode

    def __enter__(self):
        self.__training = True
        return [w.var for w in self.vars() if w.trainable]

    def __exit__(self, exc_type, exc, tb):
        self.__training = False

    def dropout_mask(self, shape, dropout, device=None):
        """The keep-mask ``tf.nn.dropout`` would draw for a tensor of this shape (layered.py:44-45): None in
        eval mode or for rate 0, else one independent Bernoulli(1-rate) per element.  Every dropout of the
        stack goes through here (tests inject masks by overriding it)."""
        if not self.__training or dropout == 0:
            return None
        return torch.rand(tuple(shape), device=device or default_device()) >= float(dropout)

    def dropout(self, features, dropout=0.5):
        """layered.py:44-45 — ``tf.nn.dropout`` only in training mode and for a non-zero rate."""
        keep = self.dropout_mask(features.shape, dropout, features.device)
        if keep is None:
            return features
        return features * keep.to(features.dtype) * (1.0 / (1.0 - float(dropout)))

    def sparse_dropout(self, G, dropout=0.5):
        """layered.py:47-50 — dropout on the nnz value vector, one draw per COO entry.  Returns the
        un-normalised adjacency with a pending edge mask (``MaskedAdjacency``); in eval mode or
        for rate 0 returns ``G`` itself."""
        if dropout == 0 or not self.__training:
            return G
        from .gnn import MaskedAdjacency
        keep = torch.rand(G.n_graph, device=G.raw_val.device) >= float(dropout)
        return MaskedAdjacency(G, keep, float(dropout))

    def __call__(self, features):
        for layer in self.__layers:
            features = layer(self, features)
        return features


class Layer:
    """Plug-in protocol (layered.py:58-86): ``__build__`` returns the output shape and creates the
    variables, ``__forward__`` computes; ``.value`` caches the last output (used as ``H0.value``)."""

    def __init__(self, *args, output_regularize: float = 0, **kwargs):
        self.__args = args
        self.__kwargs = kwargs
        self.output_regularize = output_regularize

    def __late_init__(self, architecture):
        before = set(architecture.vars())
        self.output_shape = self.__build__(architecture, *self.__args, **self.__kwargs)
        if self.output_shape is None:
            raise Exception("Layer __build__ should return an output shape")
        self.vars = set(architecture.vars()) - before
        self.__args = None
        self.__kwargs = None

    def __build__(self, architecture, *args, **kwargs):
        raise Exception("Layer should implment a __build__ method")

    def __forward__(self, architecture, features):
        raise Exception("Layer should implement a __forward__ method")

    def __call__(self, architecture, features):
        self.value = self.__forward__(architecture, features)
        return self.value

    def loss(self):
        if self.output_regularize == 0:
            return 0
        return self.output_regularize * 0.5 * (self.value ** 2).sum()  # tf.nn.l2_loss


# ------------------------------------------------------------------------------------------
# layers.py (the two layer types on either side of the propagation path)
# ------------------------------------------------------------------------------------------
class Dense(Layer):
    """layers.py:125-136 — ``dropout(activation(X·W + b))``."""

    def __build__(self, architecture, outputs: int = None, activation=identity, bias: bool = True,
                  dropout: float = 0, regularize: bool = True):
        inputs = architecture.top_shape()[1]
        if outputs is None:
            outputs = inputs
        self.W = architecture.create_var((inputs, outputs), regularize=regularize)
        self.b = architecture.create_var((1, outputs), "zero", regularize=regularize) if bias else 0
        self.activation = activation
        self.dropout = dropout
        return (architecture.top_shape()[0], outputs)

    def __forward__(self, architecture, features):
        return dense_tail(architecture, features @ self.W, self.b, self.activation, self.dropout)      # layers.py:136


def dense_tail(architecture, Z, bias, activation, dropout):
    """``dropout(activation(Z + b))`` (layers.py:136, gcn.py:89).  One native kernel for the activations
    that have a code (identity / relu / leaky_relu) on CUDA fp32 inputs; any other callable runs eagerly."""
    name = _activation_name(activation)
    if name is None or not Z.is_cuda or Z.dtype != torch.float32:
        return architecture.dropout(activation(Z + bias), dropout)
    from . import ops
    keep = architecture.dropout_mask(Z.shape, dropout, Z.device)
    return ops.bias_act_dropout(Z, bias if isinstance(bias, torch.Tensor) else None, name, keep, dropout)


def _activation_name(fn):
    if fn is identity:
        return "identity"
    if fn in (relu, torch.relu, tfn.relu):
        return "relu"
    if fn is leaky_relu:  # (torch's own leaky_relu defaults to another slope: it runs eagerly)
        return "leaky_relu"
    return None


class Dropout(Layer):
    """layers.py:175-181."""

    def __build__(self, gcn, rate: float = 0.5):
        self.rate = rate
        return gcn.top_shape()

    def __forward__(self, gcn, features):
        return gcn.dropout(features, self.rate)


class Activation(Layer):
    """layers.py:139-172, the parameter-free variants."""

    def __build__(self, architecture, activation: str = "relu", **kwargs):
        table = {"relu": torch.relu, "linear": identity, "tanh": torch.tanh, "exp": torch.exp,
                 "softmax": lambda x: torch.softmax(x, dim=1)}
        self.activation = table[activation] if isinstance(activation, str) else activation
        return architecture.top_shape()

    def __forward__(self, gcn, features):
        return self.activation(features)


class Concatenate(Layer):
    """layers.py:86-101.  Quirk kept: ``__build__`` reports a column-wise concatenation (:93,96) while
    ``__forward__`` concatenates along axis 0 (:100-101), exactly as the reference does."""

    def __build__(self, architecture, H0):
        self.H0 = H0
        if isinstance(H0, list):
            for H in H0:
                if architecture.top_shape()[0] != H.output_shape[0]:
                    raise Exception("Mismatching first dimension to concatenate between shapes " + str(architecture.top_shape())
                                    + " and " + str(H.output_shape))
            return (architecture.top_shape()[0], architecture.top_shape()[1] + H0[0].output_shape[1])
        if architecture.top_shape()[0] != H0.output_shape[0]:
            raise Exception("Mismatching first dimension to concatenate between shapes " + str(architecture.top_shape())
                            + " and " + str(H0.output_shape))
        return (architecture.top_shape()[0], architecture.top_shape()[1] + H0.output_shape[1])

    def __forward__(self, architecture, features):
        if isinstance(self.H0, list):
            return torch.cat([H.value for H in self.H0], dim=0)
        return torch.cat([features, self.H0.value], dim=0)


# ------------------------------------------------------------------------------------------
# trainable.py
# ------------------------------------------------------------------------------------------
class Predictor:
    def predict(self, features):
        raise Exception("Predictors need to implement a predict method")

    def loss(self, features):
        raise Exception("Predictors need to implement a loss method")

    def evaluate(self, features):
        raise Exception("Predictors need to implement an evaluate method")


class Trainable(Layered):
    """trainable.py:16-103 — full-batch training with Adam, L2 on every regularised variable
    (biases included, layers.py:130), early stopping on the validation loss with in-memory
    best-weights restore, and the ``_fast_predict`` cache."""

    def __init__(self, features):
        self.features = as_tensor(features)
        super().__init__(tuple(self.features.shape))
        self._fast_predict = None

    def reset(self):
        super().reset()
        self._fast_predict = None

    def _cached_forward(self):
        if self._fast_predict is None:
            with torch.no_grad():
                self._fast_predict = self(self.features)
        return self._fast_predict

    def predict(self, predictor):
        return predictor.predict(self._cached_forward())

    def loss(self, predictor):
        return predictor.loss(self._cached_forward())

    def evaluate(self, predictor):
        return predictor.evaluate(self._cached_forward())

    def train(self, train, valid=None, test=None, patience: int = 100, learning_rate: float = 0.01,
              regularization: float = 5.E-4, verbose: bool = False, epochs: int = 2000,
              degradation=lambda epoch: 1, batches: int = 1, optimizer=None):
        self.reset()
        params = [w.var for w in self.vars() if w.trainable]
        if optimizer is None:
            optimizer = torch.optim.Adam(params, lr=learning_rate, eps=1e-7)
        if valid is None:
            valid = train
        best_loss = float("inf")
        best_vars = [w.identity() for w in self.vars()]
        patience_remaining = patience
        for epoch in range(epochs):
            self._fast_predict = None
            epoch_loss = 0.0
            for _ in range(batches):
                with self:
                    optimizer.zero_grad(set_to_none=True)
                    batch_loss = train.loss(self(self.features))
                    for layer in self.layers():
                        if layer.output_regularize != 0:
                            batch_loss = batch_loss + layer.loss()
                    for w in self.vars():
                        if w.regularize != 0:
                            batch_loss = batch_loss + regularization * w.regularize * 0.5 * (w.var ** 2).sum()
                    (batch_loss * degradation(epoch)).backward()
                    optimizer.step()
                    epoch_loss = epoch_loss + batch_loss.detach()   # stays on the device
            with torch.no_grad():
                output = self(self.features)  # eval-mode forward (mode dropped by __exit__)
                # the only host synchronisation of the epoch: the early-stopping decision needs the validation
                # loss on the host (trainable.py:84,96-100); the reference also syncs for the training loss (:80)
                valid_loss = float(valid.loss(output))
            patience_remaining -= 1
            if verbose and valid_loss < best_loss:
                epoch_loss = float(epoch_loss)
                train_acc = float(train.evaluate(output))
                test_acc = float("nan") if test is None else float(test.evaluate(output))
                valid_acc = float(valid.evaluate(output))
                print(f"\rEpoch {epoch}  patience {patience_remaining}  Train loss {epoch_loss:.3f} "
                      f"Validation loss {valid_loss:.3f}  Train {train_acc:.3f} Validation {valid_acc:.3f}  "
                      f"Test {test_acc:.3f}", end="")
            if valid_loss < best_loss:
                best_loss, best_vars = valid_loss, [w.identity() for w in self.vars()]
                patience_remaining = patience
            if patience_remaining == 0:
                break
        for w, best in zip(self.vars(), best_vars):
            w.assign(best)
        if verbose:
            print("\r")
