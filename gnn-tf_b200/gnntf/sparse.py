"""Device-resident adjacency objects: what ``graph2adj`` returns and what ``GNN.get_adjacency``
returns, mirroring the ``tf.sparse.SparseTensor`` fields the reference touches
(``.indices`` ``.values`` ``.dense_shape`` ``.shape``; graph_manipulation.py:16,31, gnn.py:39)
while carrying the CSR view the sm_100a kernels consume.  All arithmetic happens in
``libgnntf_b200.so``; torch only owns the memory.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native as nat

# rows longer than this many entries are split into pieces of `CHUNK` entries (see spmm.cu);
# 128..512 measured flat, 1024 slower
LONG_THRESHOLD = 256
CHUNK = 256


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("gnntf_b200 runs on a CUDA device (sm_100a); no GPU is visible and there is no CPU fallback")
    nat.lib()


class CsrStructure:
    """Stable-by-row CSR of a COO list + the long-row plan.  Values live elsewhere."""

    def __init__(self, n, row_ptr, col_idx, coo_pos, long_threshold=LONG_THRESHOLD, chunk=CHUNK):
        self.n = int(n)
        self.nnz = int(col_idx.numel())
        self.row_ptr, self.col_idx, self.coo_pos = row_ptr, col_idx, coo_pos
        self.long_threshold, self.chunk = int(long_threshold), int(chunk)
        self.row_map = None
        self._partials = None
        self._build_plan()

    def _build_plan(self):
        L = nat.lib()
        dev = self.row_ptr.device
        counts = torch.zeros(2, dtype=torch.int32, device=dev)
        st = nat.stream_ptr()
        nat.check(L.gnntf_spmm_plan_count(nat.ptr(self.row_ptr), self.n, self.long_threshold, self.chunk,
                                          nat.ptr(counts), st), "spmm_plan_count")
        self.n_long, self.n_chunks = (int(x) for x in counts.tolist())  # build-time sync
        i32 = dict(dtype=torch.int32, device=dev)
        self.long_row = torch.empty(self.n_long, **i32)
        self.long_first_chunk = torch.empty(self.n_long, **i32)
        self.long_n_chunks = torch.empty(self.n_long, **i32)
        self.chunk_row = torch.empty(self.n_chunks, **i32)
        self.chunk_begin = torch.empty(self.n_chunks, **i32)
        if self.n_long > 0:
            nat.check(L.gnntf_spmm_plan_fill(nat.ptr(self.row_ptr), self.n, self.long_threshold, self.chunk,
                                             nat.ptr(counts), nat.ptr(self.long_row),
                                             nat.ptr(self.long_first_chunk), nat.ptr(self.long_n_chunks),
                                             nat.ptr(self.chunk_row), nat.ptr(self.chunk_begin), st),
                      "spmm_plan_fill")

    def partials(self, F):
        need = self.n_chunks * ((F + 3) // 4 * 4)
        if need == 0:
            return None
        if self._partials is None or self._partials.numel() < need:
            self._partials = torch.empty(need, dtype=torch.float32, device=self.row_ptr.device)
        return self._partials

    def struct(self, val, F):
        """``gnntf_csr_t`` for these values at feature width F (keeps nothing alive: the caller
        holds the tensors for the duration of the call)."""
        s = nat.CsrStruct()
        s.n_rows, s.nnz = self.n, self.nnz
        s.row_ptr, s.col_idx, s.val = self.row_ptr.data_ptr(), self.col_idx.data_ptr(), val.data_ptr()
        s.row_map = self.row_map.data_ptr() if self.row_map is not None else None
        s.long_threshold, s.chunk = self.long_threshold, self.chunk
        s.n_long, s.n_chunks = self.n_long, self.n_chunks
        if self.n_long > 0:
            s.long_row = self.long_row.data_ptr()
            s.long_first_chunk = self.long_first_chunk.data_ptr()
            s.long_n_chunks = self.long_n_chunks.data_ptr()
            s.chunk_row = self.chunk_row.data_ptr()
            s.chunk_begin = self.chunk_begin.data_ptr()
            s.partials = self.partials(F).data_ptr()
        return s


def build_csr(edges, weights, n, directed=False, add_eye=False, by_column=False, want_coo=True):
    """Run ``gnntf_csr_build``.  ``edges`` int64 [E,2] CUDA, ``weights`` fp32 [E] CUDA or None."""
    _require_cuda()
    L = nat.lib()
    dev = edges.device
    E = int(edges.shape[0])
    nnz = (E if directed else 2 * E) + (n if add_eye else 0)
    ws_bytes = ctypes.c_size_t(0)
    nat.check(L.gnntf_csr_build_ws_bytes(n, E, int(directed), int(add_eye), ctypes.byref(ws_bytes)), "csr_build_ws")
    ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
    indices = torch.empty((nnz, 2), dtype=torch.int64, device=dev) if want_coo else None
    values = torch.empty(nnz, dtype=torch.float32, device=dev) if want_coo else None
    row_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    col_idx = torch.empty(nnz, dtype=torch.int32, device=dev)
    raw_val = torch.empty(nnz, dtype=torch.float32, device=dev)
    coo_pos = torch.empty(nnz, dtype=torch.int32, device=dev)
    nat.check(L.gnntf_csr_build(nat.ptr(edges), nat.ptr(weights), n, E, int(directed), int(add_eye),
                                int(by_column), nat.ptr(indices), nat.ptr(values), nat.ptr(row_ptr),
                                nat.ptr(col_idx), nat.ptr(raw_val), nat.ptr(coo_pos), nat.ptr(ws),
                                ws_bytes.value, nat.stream_ptr()), "csr_build")
    del ws
    return indices, values, CsrStructure(n, row_ptr, col_idx, coo_pos), raw_val


class SparseAdjacency:
    """Return type of :func:`gnntf.graph2adj` — the un-normalised, symmetrised adjacency.

    ``indices`` / ``values`` / ``dense_shape`` are exactly the SparseTensor the reference builds at
    graph_manipulation.py:31 (same order, duplicates kept).  ``csr`` is the derived kernel view.
    """

    def __init__(self, edges, weights, n, directed=False, _source=None, _order=None):
        _require_cuda()
        self.n = int(n)
        self.directed = bool(directed)
        self.edges = edges.contiguous()
        self.weights = None if weights is None else weights.contiguous()
        # a reordered view (see reordered()): internal CSR over relabelled nodes, externally visible
        # COO fields are the source's, untouched
        self.source = _source
        self.perm = _order                       # int64 [n]: internal position -> external node id
        self.inv = None
        if _order is not None:
            self.inv = torch.empty_like(_order)
            self.inv[_order] = torch.arange(self.n, dtype=_order.dtype, device=_order.device)
        coo = _source is None
        self.indices, self.values, self.csr, self.raw_val = build_csr(self.edges, self.weights, self.n, directed,
                                                                      want_coo=coo)
        if not coo:
            self.indices, self.values = _source.indices, _source.values
        self.n_graph = self.csr.nnz
        self.dense_shape = (self.n, self.n)
        self.shape = self.dense_shape
        self._eval_cache = {}
        self._with_eye = None
        self._csc = None

    def reordered(self, power_iters=100, sweeps=30, seed=0, order=None):
        """Same graph with a locality-restoring INTERNAL node order (gnntf/reorder.py).  Everything a
        user can see keeps the reference's numbering: ``indices``/``values`` are this object's, features
        go in and results come out in external node order (the ops permute at the boundary).  COO
        storage order — hence edge-dropout masks and the accumulation order inside every row — is
        unchanged, so results are identical to the un-reordered adjacency."""
        if self.perm is not None:
            return self
        if order is None:
            from .reorder import arrangement_order
            order = arrangement_order(self, power_iters, sweeps, seed)
        order = order.to(device=self.edges.device, dtype=torch.int64).contiguous()
        inv = torch.empty_like(order)
        inv[order] = torch.arange(self.n, dtype=torch.int64, device=order.device)
        return SparseAdjacency(inv[self.edges], self.weights, self.n, self.directed, _source=self, _order=order)

    # -- structure variants --------------------------------------------------------------
    def with_eye(self):
        """CSR of the list with ``tf.sparse.eye`` appended (gnn.py:39,49), built on first use."""
        if self._with_eye is None:
            idx, val, csr, raw = build_csr(self.edges, self.weights, self.n, self.directed, add_eye=True)
            if self.source is not None:  # externally visible indices stay in external numbering
                idx = torch.cat([self.source.indices, torch.arange(self.n, device=idx.device).repeat(2, 1).T.contiguous()])
            self._with_eye = (idx, val, csr, raw)
        return self._with_eye

    def csc(self):
        """Transposed structure (CSR sorted by column) for the directed backward pass."""
        if self._csc is None:
            _, _, csr_t, _ = build_csr(self.edges, self.weights, self.n, self.directed, by_column=True, want_coo=False)
            self._csc = csr_t
        return self._csc

    # -- GNN.get_adjacency ---------------------------------------------------------------
    def normalized(self, normalized="symmetric", add_eye="none", keep_mask=None, rate=0.0):
        """``sparse_dropout`` (layered.py:47-50) with an explicit COO-order keep-mask, then
        ``get_adjacency`` (gnn.py:38-50).  Eval-mode results (no mask) are cached: the reference
        recomputes the identical matrix in every layer of every forward."""
        if normalized not in nat.NORM:
            raise Exception("Invalid matrix normalization")  # gnn.py:46-47
        if add_eye not in nat.EYE:
            add_eye = "none"  # the reference silently ignores unknown add_eye strings (gnn.py:38,48)
        key = (normalized, add_eye)
        if keep_mask is None and key in self._eval_cache:
            return self._eval_cache[key]
        if add_eye == "none":
            indices, csr, raw_val = self.indices, self.csr, self.raw_val
        else:
            indices, _, csr, raw_val = self.with_eye()
        out = NormalizedAdjacency(self, indices, csr, raw_val, normalized, add_eye, keep_mask, rate)
        if keep_mask is None:
            self._eval_cache[key] = out
        return out

    def normalize_into(self, keep_mask, rate, deg, dinv, val=None, val_T=None, normalized="symmetric"):
        """``sparse_dropout`` + ``get_adjacency`` (add_eye = "none") written into CALLER-owned buffers:
        ``deg``/``dinv`` fp32 [n], ``val`` and/or ``val_T`` fp32 [nnz] (CSR order; pass None to skip one).
        The training-mode K-step op keeps ONE scratch value array for all K iterations and recomputes the
        (transposed) values from the saved mask in the backward pass, instead of materialising K×2 arrays
        of nnz floats (products shape, K=10: 1.9 GB instead of 11 GB)."""
        L = nat.lib()
        csr = self.csr
        if self.directed and val_T is not None:
            raise Exception("normalize_into: transposed values of a directed adjacency need the by-column CSR")
        scale = 1.0 if keep_mask is None else 1.0 / (1.0 - float(rate))
        nat.check(L.gnntf_normalize_f32(nat.ptr(csr.row_ptr), nat.ptr(csr.col_idx), nat.ptr(self.raw_val),
                                        nat.ptr(csr.coo_pos), self.n, csr.nnz, self.n_graph, int(self.directed),
                                        nat.ptr(keep_mask), scale, nat.NORM[normalized], nat.EYE["none"],
                                        nat.ptr(deg), nat.ptr(dinv), nat.ptr(val), nat.ptr(val_T), None, nat.stream_ptr()),
                  "normalize")

    def __repr__(self):
        return f"SparseAdjacency(shape={self.dense_shape}, nnz={self.csr.nnz}, directed={self.directed})"


class NormalizedAdjacency:
    """Return type of ``GNN.get_adjacency`` (gnn.py:36-50): same index list, normalised values."""

    def __init__(self, base, indices, csr, raw_val, mode, eye_mode, keep_mask, rate):
        self.base, self.indices, self.csr = base, indices, csr
        self.dense_shape = self.shape = base.dense_shape
        self.mode, self.eye_mode = mode, eye_mode
        n, nnz, dev = base.n, csr.nnz, raw_val.device
        self.has_mask = keep_mask is not None
        self._scale = 1.0
        if self.has_mask:
            keep_mask = keep_mask.to(device=dev, dtype=torch.uint8).contiguous()
            if keep_mask.numel() != base.n_graph:
                raise Exception(f"edge keep-mask must have one entry per COO entry ({base.n_graph}), got {keep_mask.numel()}")
            self._scale = 1.0 / (1.0 - float(rate))
        self._keep, self._raw_val = keep_mask, raw_val
        f32 = dict(dtype=torch.float32, device=dev)
        self.deg = torch.zeros(n, **f32)
        self.dinv = torch.zeros(n, **f32)
        self.val = torch.empty(nnz, **f32)
        want_T = (self.has_mask or mode == "bipartite") and not base.directed  # row-only scaling is not symmetric
        self._val_T = torch.empty(nnz, **f32) if want_T else None
        self._values_coo = None   # COO-order values are produced on first use of .values
        self._run(None)
        self._T = None

    def _run(self, values_coo):
        L = nat.lib()
        base, csr = self.base, self.csr
        nat.check(L.gnntf_normalize_f32(nat.ptr(csr.row_ptr), nat.ptr(csr.col_idx), nat.ptr(self._raw_val),
                                        nat.ptr(csr.coo_pos), base.n, csr.nnz, base.n_graph, int(base.directed),
                                        nat.ptr(self._keep), self._scale, nat.NORM[self.mode], nat.EYE[self.eye_mode],
                                        nat.ptr(self.deg), nat.ptr(self.dinv), nat.ptr(self.val),
                                        nat.ptr(self._val_T), nat.ptr(values_coo), nat.stream_ptr()),
                  "normalize")

    @property
    def values(self):
        """COO-order values: the ``.values`` of the SparseTensor the reference returns.  Computed on
        first use (the kernels only need CSR order; a training forward builds K of these objects)."""
        if self._values_coo is None:
            self._values_coo = torch.empty(self.csr.nnz, dtype=torch.float32, device=self.val.device)
            self._run(self._values_coo)
        return self._values_coo

    def struct(self, F):
        return self.csr.struct(self.val, F)

    def transposed(self):
        """(CsrStructure, values) of Âᵀ.  Undirected: same structure, mirror values (identical to
        Â when no edge mask was drawn).  Directed: the by-column CSR with values gathered through
        its COO positions."""
        if self._T is None:
            if not self.base.directed:
                self._T = (self.csr, self._val_T if self._val_T is not None else self.val)
            else:
                if self.eye_mode != "none":
                    raise Exception("transposed directed adjacency with add_eye is not supported")
                csr_t = self.base.csc()
                self._T = (csr_t, self.values[csr_t.coo_pos.long()].contiguous())
        return self._T

    def struct_T(self, F):
        csr_t, val_t = self.transposed()
        return csr_t.struct(val_t, F)

    def __repr__(self):
        return (f"NormalizedAdjacency(shape={self.dense_shape}, nnz={self.csr.nnz}, normalized={self.mode!r}, "
                f"add_eye={self.eye_mode!r}, masked={self.has_mask})")
