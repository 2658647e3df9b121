#!/usr/bin/env python
"""Per-config measurements beyond the headline bench line (BASELINE.json configs[1], [2], [4]):
GCN 2-layer forward on the PubMed shape, APPNP on the arxiv shape at both widths, node-ordering
sensitivity on the products shape, and the R-MAT SpMM sweep.  Writes one JSON object per line.
usage: python scripts/sweep.py [--rmat-scale 23 --rmat-edges 100000000] > gpurun_out/sweep.jsonl"""
import argparse
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

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


def emit(**kw):
    print(json.dumps(kw), flush=True)


def spmm_bytes(n, nnz, F):
    return 8 * nnz + 4 * (n + 1) + 8 * n * F


def step_bytes(n, nnz, F):
    return 8 * nnz + 4 * (n + 1) + 12 * n * F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rmat-scale", type=int, default=23)
    ap.add_argument("--rmat-edges", type=int, default=100_000_000)
    ap.add_argument("--skip-rmat", action="store_true")
    args = ap.parse_args()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")

    # configs[1]: GCN 2-layer on the PubMed shape (eval forward through the public API)
    n, e, width, classes = synthetic.SHAPES["pubmed"]
    G = synthetic.citation_graph(n, e, seed=0)
    X = synthetic.citation_features(n, width, seed=1)
    gnntf.set_seed(0)
    t0 = time.time()
    adj = gnntf.graph2adj(G)
    torch.cuda.synchronize()
    build_ms = (time.time() - t0) * 1e3
    arch = gnntf.GCN(adj, X, num_classes=classes)
    arch.reset()
    arch.training_mode(False)
    with torch.no_grad():
        med, mn = timed(lambda: arch(arch.features), flush=flush)
        A = adj.normalized("symmetric")
        Xc = arch.features
        m1, _ = timed(lambda: ops.spmm_raw(A.struct(width), n, Xc), flush=flush)
        H = torch.randn((n, 64), device="cuda")
        m2, _ = timed(lambda: ops.spmm_raw(A.struct(64), n, H), flush=flush)
    nnz = adj.csr.nnz
    emit(config="GCN 2-layer PubMed-shaped, eval forward (public API)", nodes=n, nnz=nnz, forward_ms=med, forward_ms_min=mn,
         graph2adj_networkx_plus_build_ms=build_ms,
         spmm_layer1_F500_ms=m1, spmm_layer1_GBs=spmm_bytes(n, nnz, width) / m1 / 1e6,
         spmm_layer2_F64_ms=m2, spmm_layer2_GBs=spmm_bytes(n, nnz, 64) / m2 / 1e6,
         note="L2-resident working set (80 MB / 12 MB): the HBM roofline is not the limiter, launch + latency is")

    # configs[0]/[2]/[3]: APPNP K=10 at the quoted and the class widths, both node orderings
    for name, widths, orderings in (("cora", (7,), ("local",)), ("arxiv", (128, 40), ("local", "random")),
                                    ("products", (100, 47), ("local", "random"))):
        for ordering in orderings:
            if name == "cora":
                n, e, _, _ = synthetic.SHAPES["cora"]
                adj = gnntf.graph2adj(synthetic.citation_graph(n, e, seed=0))
            else:
                n, edges = synthetic.shaped_edges(name, seed=0, ordering=ordering, device="cuda")
                torch.cuda.synchronize()
                t0 = time.time()
                adj = gnntf.edges2adj(edges, None, n)
                torch.cuda.synchronize()
                build_ms = (time.time() - t0) * 1e3
                del edges
            A = adj.normalized("symmetric")
            nnz = adj.csr.nnz
            for F in widths:
                H0, _ = ops._pad4(synthetic.features(n, F, 1, "cuda"))
                out, scratch = torch.empty_like(H0), torch.empty_like(H0)
                small = step_bytes(n, nnz, F) <= 3 * 126e6
                med, mn = timed(lambda: ops.propagate_raw(A, H0, 0.1, 10, out=out, scratch=scratch),
                                reps=10 if name != "products" else 5, flush=flush if small else None)
                gbs = 10 * step_bytes(n, nnz, F) / med / 1e6
                emit(config=f"APPNP K=10 {name}-shaped", ordering=ordering, nodes=n, nnz=nnz, F=F, propagation_ms=med,
                     propagation_ms_min=mn, edge_features_per_s=nnz * F * 10 / (med * 1e-3), algorithmic_GBs=gbs,
                     frac_of_measured_peak=gbs / PEAK, frac_of_8TBs=gbs / 8000.0, csr_build_ms=build_ms,
                     long_rows=adj.csr.n_long, l2_resident=small)
                del H0, out, scratch
            del adj, A
            torch.cuda.empty_cache()

    # configs[4]: R-MAT SpMM sweep
    if not args.skip_rmat:
        n, edges = synthetic.rmat_edges(args.rmat_scale, args.rmat_edges, seed=0, device="cuda")
        torch.cuda.synchronize()
        t0 = time.time()
        adj = gnntf.edges2adj(edges, None, n)
        torch.cuda.synchronize()
        build_ms = (time.time() - t0) * 1e3
        del edges
        A = adj.normalized("symmetric")
        nnz = adj.csr.nnz
        deg = (adj.csr.row_ptr[1:] - adj.csr.row_ptr[:-1])
        for F in (16, 32, 64, 128, 256):
            Bm = synthetic.features(n, F, 1, "cuda")
            C = torch.empty_like(Bm)
            med, mn = timed(lambda: ops.spmm_raw(A.struct(F), n, Bm, out=C), reps=5, warm=2)
            gbs = spmm_bytes(n, nnz, F) / med / 1e6
            emit(config=f"R-MAT SpMM scale {args.rmat_scale}", nodes=n, edges=args.rmat_edges, nnz=nnz, F=F, spmm_ms=med,
                 spmm_ms_min=mn, edge_features_per_s=nnz * F / (med * 1e-3), algorithmic_GBs=gbs,
                 frac_of_measured_peak=gbs / PEAK, frac_of_8TBs=gbs / 8000.0, csr_build_ms=build_ms,
                 max_degree=int(deg.max()), long_rows=adj.csr.n_long, pieces=adj.csr.n_chunks)
            del Bm, C
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
