"""TensorFlow binding of the custom ops in gnntf_ops.cc, with the registered gradients.

SOURCE ONLY here (TensorFlow is not installable in the build image; nothing imports this module).
What a maintainer of the reference changes, besides building ``_gnntf_ops.so`` (tf_op/Makefile):

    # gnntf/core/gnn/architectures/filter.py:19     propagated = tf.sparse.sparse_dense_matmul(self.G, features)
    propagated = gnntf_tf.sparse_dense_matmul(self.G, features)
    # or, for the whole run of K PPRIteration layers (filter.py:34-35):
    H_K = gnntf_tf.appnp_propagate(adjacency, H0, alpha=0.1, iterations=10)

The adjacency tensors (row_ptr, col_idx, val, plan) come from ``gnntf_csr_build`` /
``gnntf_normalize_f32`` / ``gnntf_spmm_plan_*`` — called once per graph through ctypes on
``tf.experimental.dlpack`` pointers, exactly as ``gnntf/_native.py`` does for torch.
"""
import os

import tensorflow as tf

_ops = tf.load_op_library(os.path.join(os.path.dirname(__file__), "_gnntf_ops.so"))


class CsrAdjacency:
    """The five tensors of a normalised adjacency; ``transposed`` carries Âᵀ's values on the same
    structure (equal to ``val`` unless edge dropout or bipartite normalisation was applied)."""

    def __init__(self, row_ptr, col_idx, val, plan_header, plan, val_T=None):
        self.row_ptr, self.col_idx, self.val, self.plan_header, self.plan = row_ptr, col_idx, val, plan_header, plan
        self.val_T = val if val_T is None else val_T

    def args(self, transposed=False):
        return (self.row_ptr, self.col_idx, self.val_T if transposed else self.val, self.plan_header, self.plan)


def sparse_dense_matmul(adj: CsrAdjacency, dense):
    @tf.custom_gradient
    def _op(x):
        y = _ops.gnntf_spmm(*adj.args(), x)

        def grad(dy):  # adjoint_a=True branch of TF's gradient; the dA branch is dropped (A depends on no variable)
            return _ops.gnntf_spmm(*adj.args(transposed=True), dy)
        return y, grad
    return _op(dense)


def appnp_propagate(adj: CsrAdjacency, h0, alpha=0.1, iterations=10):
    @tf.custom_gradient
    def _op(x):
        y = _ops.gnntf_appnp_propagate(*adj.args(), x, alpha=alpha, iterations=iterations)

        def grad(dy):  # one adjacency for all steps: the VJP is the same recursion on the transpose
            return _ops.gnntf_appnp_propagate(*adj.args(transposed=True), dy, alpha=alpha, iterations=iterations)
        return y, grad
    return _op(h0)


# Graph-mode registration (tf.function / SavedModel): same gradients by op name.
@tf.RegisterGradient("GnntfSpmm")
def _spmm_grad(op, dy):
    row_ptr, col_idx, val, header, plan, _ = op.inputs
    # symmetric adjacency (eval mode, graph_dropout = 0): Âᵀ = Â.  Masked adjacencies go through
    # sparse_dense_matmul() above, which knows val_T.
    return [None, None, None, None, None, _ops.gnntf_spmm(row_ptr, col_idx, val, header, plan, dy)]


@tf.RegisterGradient("GnntfAppnpPropagate")
def _propagate_grad(op, dy):
    row_ptr, col_idx, val, header, plan, _ = op.inputs
    return [None, None, None, None, None,
            _ops.gnntf_appnp_propagate(row_ptr, col_idx, val, header, plan, dy, alpha=op.get_attr("alpha"),
                                       iterations=op.get_attr("iterations"))]
