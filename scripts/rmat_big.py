#!/usr/bin/env python
"""R-MAT SpMM at the larger BASELINE sizes (configs[4]): builds the graph on the GPU, frees what it
can, times gnntf_spmm_f32 over F.  One JSON line per (E, F).  usage: rmat_big.py SCALE EDGES [F ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-tf_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gnntf  # noqa: E402
import synthetic  # noqa: E402
from gnntf import ops  # noqa: E402

scale, n_edges = int(sys.argv[1]), int(float(sys.argv[2]))
widths = [int(x) for x in sys.argv[3:]] or [16, 64, 256]
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
n, edges = synthetic.rmat_edges(scale, n_edges, seed=0, device="cuda")
torch.cuda.synchronize()
t0 = time.time()
adj = gnntf.edges2adj(edges, None, n)
torch.cuda.synchronize()
build_ms = (time.time() - t0) * 1e3
del edges
A = adj.normalized("symmetric")
adj.indices = adj.values = None  # the COO view is not needed for the sweep: free 20 bytes per entry
A.indices = None
A._values_coo = None
torch.cuda.empty_cache()
nnz = adj.csr.nnz
deg = adj.csr.row_ptr[1:] - adj.csr.row_ptr[:-1]
for F in widths:
    B = synthetic.features(n, F, 1, "cuda")
    C = torch.empty_like(B)
    s = A.struct(F)
    for _ in range(2):
        ops.spmm_raw(s, n, B, out=C)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.spmm_raw(s, n, B, out=C)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    med = float(np.median(ts))
    gb = (8 * nnz + 4 * (n + 1) + 8 * n * F) / 1e9
    print(json.dumps(dict(config=f"R-MAT SpMM scale {scale}", nodes=n, edges=n_edges, nnz=nnz, F=F, spmm_ms=med,
                          edge_features_per_s=nnz * F / (med * 1e-3), algorithmic_GBs=gb / (med * 1e-3),
                          frac_of_measured_peak=gb / (med * 1e-3) / PEAK, csr_build_ms=build_ms, max_degree=int(deg.max()),
                          long_rows=adj.csr.n_long, pieces=adj.csr.n_chunks,
                          peak_mem_GB=torch.cuda.max_memory_allocated() / 1e9)), flush=True)
    del B, C
    torch.cuda.empty_cache()
