#!/usr/bin/env python
"""A/B timing of the K=10 propagation for the BASELINE shapes under the current environment
(GNNTF_B200_LIB selects the library build, GNNTF_SPMM_FMA the accumulation form).
usage: [GNNTF_B200_LIB=...] python scripts/ab_kernel.py [tag] [shapes...]  -> one JSON line per config"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-tf_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gnntf  # noqa: E402
import synthetic  # noqa: E402
from gnntf import ops  # noqa: E402


def timed(fn, reps=7, warm=3, flush=None):
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


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "current"
    shapes = sys.argv[2:] or ["arxiv", "products"]
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for name in shapes:
        widths = {"arxiv": [128, 40], "products": [100, 48], "cora": [8], "pubmed": [64, 500], "regular": [100]}[name]
        if name == "regular":
            # products-sized control: every node draws exactly 25 out-edges at log-uniform distance (no hubs,
            # symmetrised degree = 25 + Poisson(25)): separates the cost of the degree distribution from the kernel's
            import math
            n = synthetic.SHAPES["products"][0]
            g = torch.Generator(device="cuda").manual_seed(0)
            u = torch.arange(n, device="cuda").repeat_interleave(25)
            d = torch.exp(torch.rand(u.numel(), generator=g, device="cuda", dtype=torch.float64) * math.log(n // 2)).long().clamp_(1, n // 2 - 1)
            sign = torch.randint(0, 2, (u.numel(),), generator=g, device="cuda") * 2 - 1
            perm = torch.randperm(u.numel(), generator=g, device="cuda")
            edges = torch.stack([u, torch.remainder(u + sign * d, n)], dim=1)[perm].contiguous()
        elif name in synthetic.POWERLAW:
            n, edges = synthetic.shaped_edges(name, seed=0, device="cuda")
        else:
            nn, e, _, _ = synthetic.SHAPES[name]
            G = synthetic.citation_graph(nn, e, seed=0)
            edges = torch.as_tensor(np.asarray(gnntf.graph2indices(G), dtype=np.int64)).cuda()
            n = nn
        adj = gnntf.edges2adj(edges, None, n)
        A = adj.normalized("symmetric")
        for F in widths:
            H0 = synthetic.features(n, F, seed=1, device="cuda")
            out, scratch = torch.empty_like(H0), torch.empty_like(H0)
            small = (8 * adj.csr.nnz + 12 * n * F) < 3 * 126e6
            med, mn = timed(lambda: ops.propagate_raw(A, H0, 0.1, 10, out=out, scratch=scratch), flush=flush if small else None)
            print(json.dumps({"tag": tag, "lib": os.environ.get("GNNTF_B200_LIB", "in-tree"), "fma": os.environ.get("GNNTF_SPMM_FMA", "0"),
                              "shape": name, "F": F, "k10_ms_median": med, "k10_ms_min": mn,
                              "checksum": float(out.double().abs().sum().item())}), flush=True)
        del adj, A


if __name__ == "__main__":
    main()
