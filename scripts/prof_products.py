#!/usr/bin/env python
"""One products-shaped (or --shape arxiv) fused step repeated a few times: the command ncu profiles.
usage: python scripts/prof_products.py [--shape products] [--features 100] [--steps 3]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-tf_b200"))
import torch  # noqa: E402

import gnntf  # noqa: E402
import synthetic  # noqa: E402
from gnntf import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="products")
ap.add_argument("--features", type=int, default=100)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
n, edges = synthetic.shaped_edges(a.shape, seed=0, device="cuda")
adj = gnntf.edges2adj(edges, None, n)
A = adj.normalized("symmetric")
H0 = synthetic.features(n, a.features, seed=1, device="cuda")
out, scratch = torch.empty_like(H0), torch.empty_like(H0)
ops.propagate_raw(A, H0, 0.1, a.steps, out=out, scratch=scratch)
torch.cuda.synchronize()
print("ok", float(out.abs().sum()))
