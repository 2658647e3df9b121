"""A/B of the feature-row pitch: K=10 propagation on the products / arxiv shapes with the rows of H0, the
ping-pong buffers and the output at several leading dimensions (the kernels take ld >= F).
Usage: python scripts/pitch_ab.py [products|arxiv ...] > gpurun_out/pitch_ab.jsonl"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-tf_b200"))
import gnntf  # noqa: E402
import synthetic  # noqa: E402
from gnntf import ops  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for s, e in ev:
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    t = sorted(s.elapsed_time(e) for s, e in ev)
    return t[len(t) // 2]


def main():
    shapes = sys.argv[1:] or ["products", "arxiv"]
    cases = {"products": [(100, (100, 104, 112, 128)), (48, (48, 64))], "arxiv": [(128, (128,)), (40, (40, 64)), (100, (100, 128))]}
    for shape in shapes:
        n, edges = synthetic.shaped_edges(shape, seed=0, device="cuda")
        adj = gnntf.edges2adj(edges, None, n)
        A = adj.normalized("symmetric")
        del edges
        for F, lds in cases[shape]:
            H = synthetic.features(n, F, seed=1, device="cuda")
            ref = None
            for ld in lds:
                bufs = [torch.zeros((n, ld), dtype=torch.float32, device="cuda") for _ in range(3)]
                H0, out, scratch = (b[:, :F] for b in bufs)
                H0.copy_(H)
                ms = timed(lambda: ops.propagate_raw(A, H0, 0.1, 10, out=out, scratch=scratch))
                same = None
                if ref is None:
                    ref = out.clone()
                else:
                    same = bool(torch.equal(ref, out))
                print(json.dumps({"shape": shape, "F": F, "ld": ld, "k10_ms": round(ms, 3), "bit_equal_to_first": same}), flush=True)
                del bufs, H0, out, scratch


if __name__ == "__main__":
    main()
