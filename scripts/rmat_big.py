#!/usr/bin/env python
"""R-MAT SpMM sweep (BASELINE configs[4]) on 1 GPU or, under torchrun, row-sharded over N GPUs.

  python scripts/rmat_big.py SCALE EDGES [F ...]
  python -m torch.distributed.run --nproc-per-node N ... scripts/rmat_big.py SCALE EDGES [F ...]

Builds the graph on the GPU (every rank builds the same graph from the same seed and keeps its row
range), frees what it can, and times the SpMM over the widths F.  Sharded runs use
``ShardedPropagator.spmm`` (one halo exchange riding in the owned-column launch + the halo-column
pass); the time is the max over ranks.  One JSON line per (E, F) on rank 0, with a ``check`` field:
max relative error of 2,000 sampled rows against a float64 evaluation from the full CSR."""
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

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    from gnntf import dist as gdist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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
max_degree = int(deg.max())
del deg


def sample_check(got_rows, rows, B_full_fn):
    """max |x - y| / ||y||_inf over sampled rows, y evaluated in float64 from the full CSR on the device."""
    rp, col, val = adj.csr.row_ptr, adj.csr.col_idx, A.val
    worst, norm = 0.0, 0.0
    for r, g in zip(rows.tolist(), got_rows):
        s, e = int(rp[r]), int(rp[r + 1])
        y = (val[s:e, None].double() * B_full_fn(col[s:e].long()).double()).sum(0)
        worst = max(worst, float((g.double() - y).abs().max()))
        norm = max(norm, float(y.abs().max()))
    return worst / max(norm, 1e-30)


def feature_rows(F, idx):
    """Rows `idx` of the deterministic dense operand (generated per row so that no rank needs all of it)."""
    g = (idx.double()[:, None] * 0.6180339887 + torch.arange(F, device=idx.device).double()[None, :] * 0.7548776662)
    return (torch.frac(g) * 2 - 1).float()


for F in widths:
    if world == 1:
        B = feature_rows(F, torch.arange(n, device="cuda"))
        C = torch.empty_like(B)
        s = A.struct(F)
        run = lambda: ops.spmm_raw(s, n, B, out=C)  # noqa: E731
        lo, hi = 0, n
    else:
        prop = gdist.ShardedPropagator(adj, A, F, rank, world, push=True)
        lo, hi = prop.lo, prop.hi
        B = feature_rows(F, torch.arange(lo, hi, device="cuda"))
        run = lambda: prop.spmm(B)  # noqa: E731
    for _ in range(2):
        out = run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ts = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = run()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    med = torch.tensor([float(np.median(ts))], device="cuda", dtype=torch.float64)
    rows = torch.randint(lo, hi, (250,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(7 + rank))
    err = torch.tensor([sample_check([out[r - lo] for r in rows.tolist()], rows, lambda idx: feature_rows(F, idx))],
                       device="cuda", dtype=torch.float64)
    halo = torch.tensor([float(prop.n_halo) if world > 1 else 0.0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(med, op=dist.ReduceOp.MAX)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        dist.all_reduce(halo, op=dist.ReduceOp.MAX)
    med = float(med.item())
    gb = (8 * nnz + 4 * (n + 1) + 8 * n * F) / 1e9
    if rank == 0:
        print(json.dumps(dict(config=f"R-MAT SpMM scale {scale}", n_gpus=world, nodes=n, edges=n_edges, nnz=nnz, F=F, spmm_ms=med,
                              edge_features_per_s=nnz * F / (med * 1e-3), algorithmic_GBs=gb / (med * 1e-3),
                              frac_of_measured_peak=gb / (med * 1e-3) / (PEAK * world), csr_build_ms=build_ms,
                              max_degree=max_degree, long_rows=adj.csr.n_long, pieces=adj.csr.n_chunks,
                              exchange=("none" if world == 1 else ("copy-engine all-gather" if prop.copy else "push kernel")),
                              max_halo_rows=int(halo.item()), check_max_rel_err=float(err.item()),
                              peak_mem_GB=torch.cuda.max_memory_allocated() / 1e9)), flush=True)
    if world > 1:
        prop.close()
        del prop
    del B, out
    torch.cuda.empty_cache()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
