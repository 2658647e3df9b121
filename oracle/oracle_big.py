"""Full-size CPU checker on top of ``oracle_c.c``.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (its ``parity`` block and its CPU
arms) import this; nothing under ``gnn-tf_b200/`` does.  PARITY UNPINNED, like the rest of
``oracle/`` (see ``gnntf_oracle.py``).

It runs the reference's op sequence at the BASELINE sizes (arxiv: 2.3 M entries, products:
123.7 M entries) in seconds:

* ``graph2adj``          graph_manipulation.py:24-31  (append the reversed list, duplicate values)
* ``get_adjacency``      gnn.py:40-42                 (column sums in COO order, sqrt, divide_no_nan,
                                                       row scale then column scale) — C, COO order
* SpMM                   filter.py:19 / gcn.py:88     row-wise over the STABLE-by-row CSR of the COO
                                                       list.  Inside a row the entries keep their COO
                                                       order, so every output element sees exactly the
                                                       fp32 operation sequence of TF-CPU's sequential
                                                       COO loop (``oracle_spmm_coo_f32``); rows run on
                                                       OpenMP threads.  ``tests/test_oracle.py`` checks
                                                       bit-equality of the two on random graphs.
* teleport               filter.py:21                 ``P*(1-a) + H0*a``, unfused
"""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
for _p in (_ROOT, _HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import gnntf_oracle as oracle  # noqa: E402

_LIB = None


def lib():
    """ctypes handle of the C oracle with argument types declared (built on demand)."""
    global _LIB
    if _LIB is None:
        import __graft_entry__ as entry
        L = ctypes.CDLL(entry.build_oracle())
        P, I64, F32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float
        L.oracle_colsum_f32.argtypes = [P, P, I64, I64, P]
        L.oracle_normalize_sym_f32.argtypes = [P, P, I64, I64, P, P]
        L.oracle_spmm_coo_f32.argtypes = [P, P, I64, I64, P, I64, P]
        L.oracle_teleport_f32.argtypes = [P, P, I64, F32, P]
        L.oracle_appnp_propagate_f32.argtypes = [P, P, I64, I64, P, I64, F32, ctypes.c_int, ctypes.c_int, P, P]
        L.oracle_appnp_step_csr_omp_f32.argtypes = [P, P, P, P, P, I64, F32, I64, I64, P]
        L.oracle_csr_from_coo.argtypes = [P, I64, I64, P, P, P]
        L.oracle_gather_f32.argtypes = [P, P, I64, P]
        L.oracle_spmm_csr_omp_f32.argtypes = [P, P, P, P, I64, I64, I64, P]
        L.oracle_appnp_propagate_csr_omp_f32.argtypes = [P, P, P, I64, P, I64, F32, ctypes.c_int, P, P]
        L.oracle_spmm_csr_omp_acc64_f32.argtypes = [P, P, P, P, I64, I64, I64, P]
        L.oracle_appnp_propagate_csr_omp_acc64_f32.argtypes = [P, P, P, I64, P, I64, F32, ctypes.c_int, P, P]
        for name in ("oracle_colsum_f32", "oracle_normalize_sym_f32", "oracle_spmm_coo_f32", "oracle_teleport_f32",
                     "oracle_appnp_propagate_f32", "oracle_appnp_step_csr_omp_f32", "oracle_csr_from_coo",
                     "oracle_gather_f32", "oracle_spmm_csr_omp_f32", "oracle_appnp_propagate_csr_omp_f32",
                     "oracle_spmm_csr_omp_acc64_f32", "oracle_appnp_propagate_csr_omp_acc64_f32"):
            getattr(L, name).restype = None
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data


def csr_from_coo(idx, n):
    """Stable-by-row CSR (row_ptr int64 [n+1], col int32 [nnz], coo_pos int64 [nnz]) in O(nnz)."""
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    nnz = idx.shape[0]
    row_ptr = np.empty(n + 1, np.int64)
    col = np.empty(nnz, np.int32)
    coo_pos = np.empty(nnz, np.int64)
    lib().oracle_csr_from_coo(_p(idx), nnz, n, _p(row_ptr), _p(col), _p(coo_pos))
    return row_ptr, col, coo_pos


class BigOracle:
    """graph2adj -> get_adjacency("symmetric", eval) -> CSR view, for one edge list."""

    def __init__(self, edges, weights, n, directed=False, keep_idx=False):
        L = lib()
        idx, raw, _ = oracle.graph2adj_arrays(np.asarray(edges), weights, n, directed)
        self.n, self.nnz = int(n), int(idx.shape[0])
        self.raw = raw
        self.D = np.empty(self.n, np.float32)
        self.norm_coo = np.empty(self.nnz, np.float32)       # the .values get_adjacency returns (COO order)
        L.oracle_normalize_sym_f32(_p(idx), _p(raw), self.nnz, self.n, _p(self.D), _p(self.norm_coo))
        self.row_ptr, self.col, self.coo_pos = csr_from_coo(idx, self.n)
        self.val = np.empty(self.nnz, np.float32)             # normalised values in CSR order
        L.oracle_gather_f32(_p(self.norm_coo), _p(self.coo_pos), self.nnz, _p(self.val))
        self.idx = idx if keep_idx else None

    def spmm(self, H, val=None, acc64=False):
        """``acc64``: row sums accumulated in double (the error-budget twin used for split rows)."""
        H = np.ascontiguousarray(H, dtype=np.float32)
        out = np.empty_like(H)
        v = self.val if val is None else np.ascontiguousarray(val, dtype=np.float32)
        fn = lib().oracle_spmm_csr_omp_acc64_f32 if acc64 else lib().oracle_spmm_csr_omp_f32
        fn(_p(self.row_ptr), _p(self.col), _p(v), _p(H), H.shape[1], 0, self.n, _p(out))
        return out

    def step(self, H, H0, a, acc64=False):
        H = np.ascontiguousarray(H, dtype=np.float32)
        H0 = np.ascontiguousarray(H0, dtype=np.float32)
        out = self.spmm(H, acc64=acc64)
        lib().oracle_teleport_f32(_p(out), _p(H0), out.size, ctypes.c_float(a), _p(out))
        return out

    def propagate(self, H0, a, K, acc64=False):
        H0 = np.ascontiguousarray(H0, dtype=np.float32)
        out = np.empty_like(H0)
        scratch = np.empty_like(H0)
        fn = lib().oracle_appnp_propagate_csr_omp_acc64_f32 if acc64 else lib().oracle_appnp_propagate_csr_omp_f32
        fn(_p(self.row_ptr), _p(self.col), _p(self.val), self.n, _p(H0), H0.shape[1], ctypes.c_float(a), int(K),
           _p(scratch), _p(out))
        return out


def check_against(big, got, expect32, expect64, split_rows, what, floor_same=None, floor_split=None):
    """The parity rule for results that contain rows the GPU sums in pieces:
    * rows that are NOT split: against the fp32 oracle in the reference's order (``expect32``);
    * split rows (``split_rows`` bool [n]): against the double-accumulating twin (``expect64``) — a
      10^4-term sequential fp32 sum is ~5e-6·‖y‖∞ away from exact arithmetic, which is further
      than the piecewise sum is, so the sequential result cannot be the yardstick for those rows."""
    fs = oracle.FLOOR_REORDERED if floor_same is None else floor_same
    fr = oracle.FLOOR_REORDERED if floor_split is None else floor_split
    keep = ~split_rows
    norm = float(np.max(np.abs(expect32))) if expect32.size else 0.0   # one shared norm for both halves
    oracle.assert_close(got[keep], expect32[keep], what=what + " [un-split rows vs fp32 oracle]", floor=fs, norm=norm)
    if split_rows.any():
        oracle.assert_close(got[split_rows], expect64[split_rows],
                            what=what + " [split rows vs double-accumulating twin]", floor=fr, norm=norm)
