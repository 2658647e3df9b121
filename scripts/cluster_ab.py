"""A/B of the cluster-resident K-step kernel (csrc/cluster.cu): K=10 propagation time for every cluster size /
CTA size that can hold the shape, next to the default entry and to K separate fused-step launches.
Usage: python scripts/cluster_ab.py > gpurun_out/cluster_ab.jsonl"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-tf_b200"))
import gnntf  # noqa: E402
import synthetic  # noqa: E402
from gnntf import ops  # noqa: E402


def timed(fn, reps=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / reps)
    return best * 1e3  # us


def main():
    K, a = 10, 0.1
    for shape, F in (("cora", 8), ("cora", 64), ("pubmed", 4), ("pubmed", 16), ("pubmed", 64)):
        n, e, _, _ = synthetic.SHAPES[shape]
        adj = gnntf.graph2adj(synthetic.citation_graph(n, e, seed=0))
        A = adj.normalized("symmetric")
        H0 = synthetic.features(n, F, seed=1, device="cuda")
        out, scratch = torch.empty_like(H0), torch.empty_like(H0)
        rec = {"shape": shape, "n": n, "nnz": adj.csr.nnz, "F": F, "K": K}
        rec["default_entry_us"] = timed(lambda: ops.propagate_raw(A, H0, a, K, out=out, scratch=scratch))

        def steps():
            src = H0
            for _ in range(K):
                src = gnntf.appnp_step(A, src, H0, a)      # one fused-step launch (allocates its output)
        rec["k_launches_us"] = timed(steps, reps=50)
        for C in (1, 2, 4, 8, 16):
            for threads in (512, 1024):
                if ops.propagate_cluster_raw(A, H0, a, K, C, threads, out=out) is None:
                    continue
                rec[f"cluster{C}x{threads}_us"] = timed(lambda: ops.propagate_cluster_raw(A, H0, a, K, C, threads, out=out))
        for k in (1, 2, 20):
            if ops.propagate_cluster_raw(A, H0, a, k, 0, 0, out=out) is not None:
                rec[f"auto_K{k}_us"] = timed(lambda: ops.propagate_cluster_raw(A, H0, a, k, 0, 0, out=out))
        print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in rec.items()}), flush=True)


if __name__ == "__main__":
    main()
