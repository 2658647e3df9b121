"""Differentiable host wrappers of the native propagation ops (the "registered gradient" of the
reference's TF ops, here as ``torch.autograd.Function``s over the C-ABI).

* :func:`sparse_dense_matmul` — ``tf.sparse.sparse_dense_matmul`` (filter.py:19, gcn.py:88) and its
  ``adjoint_a`` gradient (trainable.py:78).  No gradient flows into the adjacency values (they
  depend on no variable; the reference computes that branch and discards it).
* :func:`appnp_step` — one fused ``PPRIteration.__forward__`` (filter.py:17-22).
* :func:`appnp_propagate` — K fused iterations (filter.py:34-35 under layered.py:52-55) with the
  fused backward of SURVEY.md Appendix C.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native as nat
from .sparse import NormalizedAdjacency, SparseAdjacency


def _as_norm(adj):
    if isinstance(adj, SparseAdjacency):  # raw graph2adj output used directly: normalized="none"
        return adj.normalized("none")
    if not isinstance(adj, NormalizedAdjacency):
        raise Exception("expected the adjacency returned by graph2adj / get_adjacency")
    return adj


def _dense(x, n_rows=None):
    if x.dtype != torch.float32 or not x.is_cuda:
        raise Exception("features must be a float32 CUDA tensor")
    if x.dim() != 2:
        raise Exception("features must be a 2-D [nodes, features] tensor")
    if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
        x = x.contiguous()
    return x


def pitch_for(F):
    """Leading dimension (in floats) for an [n, F] feature matrix whose rows are gathered whole.  A gather
    costs L1TEX wavefronts per 128-byte LINE touched: a 208-byte row (F = 52) at a 208-byte pitch touches 2.5
    lines on average, at a 256-byte pitch exactly 2 (measured on a 4x2 shard: pass 1 0.64 -> 0.585 ms).  The
    pitch is rounded up to a multiple of 32 floats when that saves >= 10 % of the lines and costs <= 35 % more
    memory; 400-byte rows (F = 100) touch 4 lines either way and stay dense."""
    F4 = (F + 3) // 4 * 4
    aligned = (F4 + 31) // 32 * 32
    if aligned == F4 or aligned > 1.35 * F4:
        return F4
    row, pitch = F4 * 4, F4 * 4
    offs = {(k * pitch) % 128 for k in range(32)}
    unaligned = sum(-(-(o + row) // 128) for o in offs) / len(offs)
    return aligned if unaligned >= 1.1 * (-(-row // 128)) else F4


def _pad4(x):
    """Zero-pad the feature dimension to a multiple of 4 floats so rows are 16-byte aligned and the
    kernels take the float4 path (class-width matrices: 7, 47 ...).  Padding columns stay zero
    through SpMM / teleport / relu, so slicing the result back is exact."""
    F = x.shape[1]
    Fp = (F + 3) // 4 * 4
    if Fp == F:
        return x, F
    out = torch.zeros((x.shape[0], Fp), dtype=x.dtype, device=x.device)
    out[:, :F] = x
    return out, F


def _perm_in(adj, x):
    """External node order -> the adjacency's internal order (identity unless ``reordered()``)."""
    perm = adj.base.perm if not isinstance(adj, (list, tuple)) else adj[0].base.perm
    return x if perm is None else x.index_select(0, perm)


def _perm_out(adj, x):
    inv = adj.base.inv if not isinstance(adj, (list, tuple)) else adj[0].base.inv
    return x if inv is None else x.index_select(0, inv)


def _ld(x):
    return x.stride(0) if x.shape[0] > 1 else max(x.shape[1], x.stride(0))


def spmm_raw(csr_struct, n_rows, B, out=None):
    """C = A·B on raw tensors (no autograd)."""
    L = nat.lib()
    B = _dense(B)
    F = B.shape[1]
    C = out if out is not None else torch.empty((n_rows, F), dtype=torch.float32, device=B.device)
    nat.check(L.gnntf_spmm_f32(ctypes.byref(csr_struct), nat.ptr(B), _ld(B), nat.ptr(C), _ld(C), F,
                               nat.stream_ptr()), "spmm")
    return C


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, adj, H):
        ctx.adj = adj
        Hp, F = _pad4(_perm_in(adj, H)) if H.shape[1] > 8 else (_perm_in(adj, H), H.shape[1])
        out = _perm_out(adj, spmm_raw(adj.struct(Hp.shape[1]), adj.base.n, Hp))
        return out if out.shape[1] == F else out[:, :F].contiguous()

    @staticmethod
    def backward(ctx, g):
        adj = ctx.adj
        if not ctx.needs_input_grad[1]:
            return None, None  # e.g. the constant feature matrix of the first GCN layer
        g = _perm_in(adj, _dense(g))
        gp, F = _pad4(g) if g.shape[1] > 8 else (g, g.shape[1])
        out = _perm_out(adj, spmm_raw(adj.struct_T(gp.shape[1]), adj.base.n, gp))
        return None, (out if out.shape[1] == F else out[:, :F].contiguous())


def sparse_dense_matmul(adj, H):
    """Drop-in for ``tf.sparse.sparse_dense_matmul(adj, H)``."""
    adj = _as_norm(adj)
    H = _dense(H)
    if H.shape[0] != adj.base.n:
        raise Exception(f"dimension mismatch: adjacency is {adj.dense_shape}, features have {H.shape[0]} rows")
    return _SpMM.apply(adj, H)


def _step_raw(struct, H_in, H0, alpha, feat_keep=None, p_scale=1.0, act=nat.ACT_IDENTITY, out=None):
    L = nat.lib()
    F = H_in.shape[1]
    ld = _ld(H_in)
    assert _ld(H0) == ld
    out = out if out is not None else torch.empty_like(H_in)
    nat.check(L.gnntf_appnp_step_f32(ctypes.byref(struct), nat.ptr(H_in), nat.ptr(H0), nat.ptr(out), ld, F,
                                     float(alpha), nat.ptr(feat_keep), float(p_scale), int(act),
                                     nat.stream_ptr()), "appnp_step")
    return out


class _Step(torch.autograd.Function):
    @staticmethod
    def forward(ctx, adj, H, H0, alpha, feat_keep, p_scale, relu):
        H, H0 = _perm_in(adj, H.contiguous()), _perm_in(adj, H0.contiguous())
        keep_internal = _perm_in(adj, feat_keep) if feat_keep is not None else None
        out = _perm_out(adj, _step_raw(adj.struct(H.shape[1]), H, H0, alpha, keep_internal, p_scale,
                                       nat.ACT_RELU if relu else nat.ACT_IDENTITY))
        ctx.adj, ctx.alpha, ctx.p_scale, ctx.relu = adj, alpha, p_scale, relu
        ctx.save_for_backward(feat_keep if feat_keep is not None else torch.empty(0), out if relu else torch.empty(0))
        return out

    @staticmethod
    def backward(ctx, g):
        keep, out = ctx.saved_tensors
        g = g.contiguous()
        if ctx.relu:
            g = g * (out > 0)
        if keep.numel():
            g = g * keep.to(g.dtype) * ctx.p_scale
        dH = _perm_out(ctx.adj, spmm_raw(ctx.adj.struct_T(g.shape[1]), ctx.adj.base.n,
                                         _perm_in(ctx.adj, g.contiguous())).mul_(1.0 - ctx.alpha)) \
            if ctx.needs_input_grad[1] else None
        dH0 = g * ctx.alpha if ctx.needs_input_grad[2] else None
        return None, dH, dH0, None, None, None, None


def appnp_step(adj, H, H0, alpha, feat_keep=None, p_feat=0.0, relu=False):
    """``act(dropout((1-a)·Â·H + a·H0))`` in one kernel (filter.py:19-22)."""
    adj = _as_norm(adj)
    p_scale = 1.0 / (1.0 - p_feat) if feat_keep is not None else 1.0
    if feat_keep is not None:
        feat_keep = feat_keep.to(torch.uint8).contiguous()
    return _Step.apply(adj, _dense(H), _dense(H0), float(alpha), feat_keep, p_scale, bool(relu))


def _struct_array(structs):
    arr = (nat.CsrStruct * len(structs))()
    for i, s in enumerate(structs):
        arr[i] = s
    return arr


def propagate_raw(adjs, H0, alpha, K, out=None, scratch=None):
    """K fused steps on raw tensors.  ``adjs``: one NormalizedAdjacency (eval: shared by all steps)
    or a list of K (training: one edge mask per step, filter.py:18)."""
    L = nat.lib()
    F, ld = H0.shape[1], _ld(H0)
    out = out if out is not None else torch.empty_like(H0)
    if scratch is None and K > 1:
        scratch = torch.empty_like(H0)
    if isinstance(adjs, (list, tuple)):
        arr = _struct_array([a.struct(F) for a in adjs])
        nat.check(L.gnntf_appnp_propagate_multi_f32(arr, K, nat.ptr(H0), nat.ptr(out), nat.ptr(scratch), ld, F,
                                                    float(alpha), nat.stream_ptr()), "appnp_propagate_multi")
    else:
        s = adjs.struct(F)
        nat.check(L.gnntf_appnp_propagate_f32(ctypes.byref(s), nat.ptr(H0), nat.ptr(out), nat.ptr(scratch), ld, F,
                                              float(alpha), K, nat.stream_ptr()), "appnp_propagate")
    return out


def propagate_cluster_raw(adj, H0, alpha, K, cluster_size=0, threads=0, out=None):
    """The K steps in ONE thread-block-cluster launch with the graph and the features resident in the
    cluster's shared memory (gnntf_appnp_propagate_cluster_f32; Cora / PubMed-sized problems only).  An explicit
    alternative: the general entry keeps the cooperative launch, which measured as fast or faster.
    Returns None when the shape does not qualify."""
    L = nat.lib()
    F, ld = H0.shape[1], _ld(H0)
    out = out if out is not None else torch.empty_like(H0)
    s = adj.struct(F)
    rc = L.gnntf_appnp_propagate_cluster_f32(ctypes.byref(s), nat.ptr(H0), nat.ptr(out), ld, F, float(alpha), int(K),
                                             int(cluster_size), int(threads), nat.stream_ptr())
    if rc == nat.GNNTF_E_SHAPE:
        return None
    nat.check(rc, "appnp_propagate_cluster")
    return out


class _Propagate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, adjs, H0, alpha, K):
        H0p, F = _pad4(_perm_in(adjs, H0.contiguous()))
        ctx.adjs, ctx.alpha, ctx.K = adjs, alpha, K
        out = _perm_out(adjs, propagate_raw(adjs, H0p, alpha, K))
        return out if out.shape[1] == F else out[:, :F].contiguous()

    @staticmethod
    def backward(ctx, g):
        L = nat.lib()
        g, F_orig = _pad4(_perm_in(ctx.adjs, g.contiguous()))
        F, ld, K = g.shape[1], _ld(g), ctx.K
        adjs = ctx.adjs if isinstance(ctx.adjs, (list, tuple)) else [ctx.adjs] * K
        arr = _struct_array([a.struct_T(F) for a in adjs])
        dH0 = torch.empty_like(g)
        scratch = torch.empty((2,) + tuple(g.shape), dtype=g.dtype, device=g.device) if K > 1 else None
        nat.check(L.gnntf_appnp_propagate_bwd_f32(arr, K, nat.ptr(g), nat.ptr(dH0), nat.ptr(scratch), ld, F,
                                                  float(ctx.alpha), nat.stream_ptr()), "appnp_propagate_bwd")
        dH0 = _perm_out(ctx.adjs, dH0)
        return None, (dH0 if F == F_orig else dH0[:, :F_orig].contiguous()), None, None


def appnp_propagate(adjs, H0, alpha=0.1, iterations=10):
    """``H ← (1−a)·Â·H + a·H0`` K times, starting from ``H = H0``; differentiable w.r.t. ``H0``."""
    if isinstance(adjs, (list, tuple)):
        adjs = [_as_norm(a) for a in adjs]
        if len(adjs) != iterations:
            raise Exception("one adjacency per iteration is required")
    else:
        adjs = _as_norm(adjs)
    return _Propagate.apply(adjs, _dense(H0), float(alpha), int(iterations))


class _PropagateMasked(torch.autograd.Function):
    """Training-mode K-step propagation (one edge keep-mask per iteration, filter.py:18 → layered.py:47-50)
    WITHOUT K materialised adjacencies: each iteration normalises into one reused scratch value array,
    runs its fused step, and only its mask (uint8 per COO entry) survives; the backward pass recomputes
    the transposed values of iteration k from mask k into the same scratch.  products shape, K = 10:
    1.9 GB of state instead of 11 GB (SURVEY §7 "recompute from the mask")."""

    @staticmethod
    def forward(ctx, base, masks, rate, H0, alpha):
        L = nat.lib()
        K = len(masks)
        H0p, F = _pad4(H0.contiguous())
        n, nnz, dev = base.n, base.csr.nnz, H0p.device
        f32 = dict(dtype=torch.float32, device=dev)
        deg, dinv, val = torch.empty(n, **f32), torch.empty(n, **f32), torch.empty(nnz, **f32)
        bufs = [torch.empty_like(H0p), torch.empty_like(H0p)]
        src = H0p
        for k in range(K):
            base.normalize_into(masks[k], rate, deg, dinv, val=val)
            dst = bufs[k & 1]
            st = base.csr.struct(val, H0p.shape[1])
            nat.check(L.gnntf_appnp_step_f32(ctypes.byref(st), nat.ptr(src), nat.ptr(H0p), nat.ptr(dst), _ld(H0p), H0p.shape[1],
                                             float(alpha), None, 1.0, nat.ACT_IDENTITY, nat.stream_ptr()), "appnp_step")
            src = dst
        ctx.base, ctx.masks, ctx.rate, ctx.alpha = base, masks, rate, alpha
        out = src if K > 0 else H0p.clone()
        return out if out.shape[1] == F else out[:, :F].contiguous()

    @staticmethod
    def backward(ctx, g):
        L = nat.lib()
        base, masks, K, a = ctx.base, ctx.masks, len(ctx.masks), float(ctx.alpha)
        g, F_orig = _pad4(g.contiguous())
        n, nnz, dev, F = base.n, base.csr.nnz, g.device, g.shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        deg, dinv, valT = torch.empty(n, **f32), torch.empty(n, **f32), torch.empty(nnz, **f32)
        dH0 = g * a if K > 0 else g.clone()          # dH0 += a·g_K
        cur = g
        bufs = [torch.empty_like(g), torch.empty_like(g)]
        for k in range(K - 1, -1, -1):               # g_k = (1-a)·Â_kᵀ·g_{k+1};  dH0 += a·g_k (k > 0) ... + g_0
            base.normalize_into(masks[k], ctx.rate, deg, dinv, val=None, val_T=valT)
            st = base.csr.struct(valT, F)
            dst = bufs[k & 1]
            nat.check(L.gnntf_spmm_f32(ctypes.byref(st), nat.ptr(cur), _ld(cur), nat.ptr(dst), _ld(dst), F, nat.stream_ptr()), "spmm")
            dst.mul_(1.0 - a)
            dH0.add_(dst, alpha=(a if k > 0 else 1.0))
            cur = dst
        return None, None, None, (dH0 if F == F_orig else dH0[:, :F_orig].contiguous()), None


def appnp_propagate_masked(base, keep_masks, rate, H0, alpha=0.1):
    """K = len(keep_masks) fused PPR iterations in TRAINING mode: iteration k uses the adjacency
    ``get_adjacency`` builds from edge keep-mask k (COO order, one Bernoulli per entry).  Memory-light
    form of ``appnp_propagate([adj.normalized(keep_mask=m, rate=rate) for m in keep_masks], ...)`` with
    identical results; undirected adjacencies in their own node order only."""
    if base.directed or base.perm is not None:
        adjs = [base.normalized("symmetric", keep_mask=m, rate=rate) for m in keep_masks]
        return appnp_propagate(adjs, H0, alpha, len(keep_masks))
    masks = [m.to(device=base.raw_val.device, dtype=torch.uint8).contiguous() for m in keep_masks]
    for m in masks:
        if m.numel() != base.n_graph:
            raise Exception(f"edge keep-mask must have one entry per COO entry ({base.n_graph}), got {m.numel()}")
    return _PropagateMasked.apply(base, masks, float(rate), _dense(H0), float(alpha))


def appnp_propagate_host(adj, H0_host, alpha=0.1, iterations=10, out_host=None, bufs=None):
    """End-to-end form with HOST feature buffers (pinned for full speed): H2D copy, K fused steps,
    D2H copy, all on the current stream; returns the host tensor after synchronising."""
    L = nat.lib()
    adj = _as_norm(adj)
    n, F = H0_host.shape
    if adj.base.perm is not None:  # reordered adjacency: permute on the device at both ends
        if out_host is None:
            out_host = torch.empty((n, F), dtype=torch.float32, pin_memory=True)
        H0 = H0_host.to("cuda", non_blocking=True)
        out = _perm_out(adj, propagate_raw(adj, _perm_in(adj, H0), alpha, iterations))
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host
    if bufs is None:
        bufs = [torch.empty((n, F), dtype=torch.float32, device="cuda") for _ in range(3)]
    if out_host is None:
        out_host = torch.empty((n, F), dtype=torch.float32, pin_memory=True)
    s = adj.struct(F)
    nat.check(L.gnntf_appnp_propagate_host_f32(ctypes.byref(s), nat.ptr(H0_host), nat.ptr(out_host),
                                               nat.ptr(bufs[0]), nat.ptr(bufs[1]), nat.ptr(bufs[2]), F, F,
                                               float(alpha), int(iterations), nat.stream_ptr()),
              "appnp_propagate_host")
    torch.cuda.current_stream().synchronize()
    return out_host


def appnp_propagate_host_batched(adj, H0_hosts, alpha=0.1, iterations=10, out_hosts=None, work=None):
    """A sequence of HOST feature matrices (pinned for full speed) through the same adjacency:
    ``[appnp_propagate_host(adj, H) for H in H0_hosts]`` as ONE native call that overlaps the upload of
    matrix b+1 and the read-back of matrix b-1 with the K steps of matrix b (three streams, two device
    slots; gnntf_appnp_propagate_host_batched_f32).  ``out_hosts``: host tensors to fill (entries may
    repeat when only some results are wanted); ``work``: optional device workspace of 5·n·F floats."""
    L = nat.lib()
    adj = _as_norm(adj)
    H0_hosts = list(H0_hosts)
    if not H0_hosts:
        return []
    n, F = H0_hosts[0].shape
    if adj.base.perm is not None:   # reordered adjacency: the permutation runs on the device, call by call
        outs = out_hosts or [None] * len(H0_hosts)
        return [appnp_propagate_host(adj, H, alpha, iterations, out_host=o) for H, o in zip(H0_hosts, outs)]
    for H in H0_hosts:
        if tuple(H.shape) != (n, F) or H.dtype != torch.float32 or H.is_cuda or not H.is_contiguous():
            raise ValueError("appnp_propagate_host_batched: every input must be a contiguous float32 HOST tensor of one shape")
    if out_hosts is None:
        out_hosts = [torch.empty((n, F), dtype=torch.float32, pin_memory=True) for _ in H0_hosts]
    out_hosts = list(out_hosts)
    if len(out_hosts) != len(H0_hosts):
        raise ValueError("appnp_propagate_host_batched: one output per input")
    if work is None:
        work = torch.empty((5, n, F), dtype=torch.float32, device="cuda")
    if work.numel() < 5 * n * F:
        raise ValueError("appnp_propagate_host_batched: workspace smaller than 5*n*F floats")
    ins = (ctypes.c_void_p * len(H0_hosts))(*[H.data_ptr() for H in H0_hosts])
    outs = (ctypes.c_void_p * len(out_hosts))(*[o.data_ptr() for o in out_hosts])
    s = adj.struct(F)
    nat.check(L.gnntf_appnp_propagate_host_batched_f32(ctypes.byref(s), ins, outs, len(H0_hosts), nat.ptr(work), F, F,
                                                       float(alpha), int(iterations), nat.stream_ptr()),
              "appnp_propagate_host_batched")
    torch.cuda.current_stream().synchronize()
    return out_hosts


# ------------------------------------------------------------------------------------------
# Fused element-wise stages either side of the path (csrc/dense.cu)
# ------------------------------------------------------------------------------------------
_ACT_CODES = {"identity": nat.ACT_IDENTITY, "relu": nat.ACT_RELU, "leaky_relu": nat.ACT_LEAKY_RELU}


class _BiasActDropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Z, bias, keep, p_scale, act, slope):
        L = nat.lib()
        Z = _dense(Z)
        n, F = Z.shape
        out = torch.empty((n, F), dtype=torch.float32, device=Z.device)
        b = None if bias is None else bias.reshape(-1).contiguous()
        nat.check(L.gnntf_bias_act_dropout_f32(nat.ptr(Z), _ld(Z), nat.ptr(b), nat.ptr(keep), float(p_scale), int(act),
                                               float(slope), nat.ptr(out), F, n, F, nat.stream_ptr()), "bias_act_dropout")
        ctx.keep, ctx.p_scale, ctx.act, ctx.slope = keep, float(p_scale), int(act), float(slope)
        ctx.has_bias, ctx.bias_shape = bias is not None, (None if bias is None else tuple(bias.shape))
        ctx.save_for_backward(out if act != nat.ACT_IDENTITY else torch.empty(0, device=Z.device))
        return out

    @staticmethod
    def backward(ctx, g):
        L = nat.lib()
        (y,) = ctx.saved_tensors
        g = _dense(g)
        n, F = g.shape
        dZ = torch.empty((n, F), dtype=torch.float32, device=g.device)
        yy = y if y.numel() else None
        nat.check(L.gnntf_bias_act_dropout_bwd_f32(nat.ptr(g), _ld(g), nat.ptr(yy), F, nat.ptr(ctx.keep), ctx.p_scale, ctx.act,
                                                   ctx.slope, nat.ptr(dZ), F, n, F, nat.stream_ptr()), "bias_act_dropout_bwd")
        db = dZ.sum(0).reshape(ctx.bias_shape) if (ctx.has_bias and ctx.needs_input_grad[1]) else None
        return (dZ if ctx.needs_input_grad[0] else None), db, None, None, None, None


def bias_act_dropout(Z, bias=None, activation="identity", keep=None, rate=0.0, slope=0.2):
    """``dropout(activation(Z + bias))`` in one kernel (layers.py:135-136, gcn.py:89).  ``keep``: bool/uint8
    [n,F] mask of the elements that survive (None = no dropout); ``rate`` gives the 1/(1-rate) scale."""
    p_scale = 1.0
    if keep is not None:
        keep = keep.to(torch.uint8).contiguous()
        p_scale = 1.0 / (1.0 - float(rate))
    return _BiasActDropout.apply(Z, bias, keep, p_scale, _ACT_CODES[activation], slope)


class _NodeXent(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, nodes, labels):
        L = nat.lib()
        logits = _dense(logits)
        m, C = int(nodes.numel()), logits.shape[1]
        ws = torch.empty(max(m, 1), dtype=torch.float32, device=logits.device)
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        nat.check(L.gnntf_node_xent_f32(nat.ptr(logits), _ld(logits), nat.ptr(nodes), nat.ptr(labels), m, C, nat.ptr(ws),
                                        nat.ptr(loss), nat.stream_ptr()), "node_xent")
        ctx.save_for_backward(logits, nodes, labels)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        L = nat.lib()
        logits, nodes, labels = ctx.saved_tensors
        m, C = int(nodes.numel()), logits.shape[1]
        d = torch.zeros_like(logits)
        gs = g.reshape(1).to(torch.float32).contiguous()
        nat.check(L.gnntf_node_xent_bwd_f32(nat.ptr(logits), _ld(logits), nat.ptr(nodes), nat.ptr(labels), m, C, nat.ptr(gs),
                                            nat.ptr(d), _ld(d), nat.stream_ptr()), "node_xent_bwd")
        return d, None, None


def node_cross_entropy(logits, nodes, labels):
    """``NodeClassification.loss`` (graph_predictor.py:19-25): gather + log-softmax + sparse CE, mean over
    ``nodes`` — one kernel forward, one backward.  ``nodes``/``labels``: int64 CUDA tensors."""
    C = logits.shape[1]
    if labels.numel() and (int(labels.min()) < 0 or int(labels.max()) >= C):
        raise Exception("labels out of range")
    return _NodeXent.apply(logits, nodes.contiguous(), labels.contiguous())
