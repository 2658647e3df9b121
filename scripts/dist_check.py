#!/usr/bin/env python
"""Multi-GPU correctness check (run under torchrun): the sharded propagation (any ROWSxCOLS grid,
copy-engine all-gather, peer-memory push kernel or NCCL all-to-all) against the single-GPU propagation of the same graph, which
every rank computes locally.  Prints PASS/FAIL per rank; exit code 1 on mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-tf_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gnntf  # noqa: E402
import synthetic  # noqa: E402
from gnntf import dist as gdist  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
grid_arg = sys.argv[1] if len(sys.argv) > 1 else ""
mode = sys.argv[2] if len(sys.argv) > 2 else "auto"      # copy (copy-engine all-gather) | push (fused push kernel) | nccl
push = mode != "nccl"
F = 100
n, edges = synthetic.shaped_edges("arxiv", seed=0, device="cuda")
adj = gnntf.edges2adj(edges, None, n)
A = adj.normalized("symmetric")
R, C = (int(x) for x in grid_arg.split("x")) if grid_arg else gdist.choose_grid(world, F)
grid = gdist.Grid2D(rank, world, R, C)
c0, c1 = gdist.column_range(F, C, grid.c)
prop = gdist.ShardedPropagator(adj, A, c1 - c0, grid.r, R, group=grid.row_group, push=push, copy={"copy": True, "auto": "auto"}.get(mode, False))
ok, worst = True, 0.0
# repeated calls exercise buffer reuse across propagations; odd K with a DIFFERENT H0 per call is the
# write-after-read hazard of the peer-memory push: a fast rank must not overwrite halo rows
# a slow peer is still reading from the previous call
for call, K in enumerate([10, 1, 3, 1, 10, 3, 10, 1, 3]):   # 2nd use of a K captures a CUDA graph, 3rd replays it
    H0 = synthetic.features(n, F, 1 + call, "cuda")
    expect = gnntf.appnp_propagate(A, H0, 0.1, K)
    if call % 2 == rank % 2:
        torch.cuda._sleep(int(2e7))          # skew the ranks: some arrive late at every other call
    got = prop.propagate(H0[prop.lo:prop.hi, c0:c1].contiguous(), 0.1, K)
    ref = expect[prop.lo:prop.hi, c0:c1]
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    worst = max(worst, err)
    ok = ok and err < 1e-5
# sharded plain SpMM (BASELINE config 5 mode)
H = synthetic.features(n, F, 99, "cuda")
ref = gnntf.sparse_dense_matmul(A, H)[prop.lo:prop.hi, c0:c1]
got = prop.spmm(H[prop.lo:prop.hi, c0:c1].contiguous())
err = (got - ref).abs().max().item() / ref.abs().max().item()
worst = max(worst, err)
ok = ok and err < 1e-5
torch.cuda.synchronize()
graphs = sum(isinstance(v, tuple) for v in prop._graphs.values())
print(f"rank {rank} grid {R}x{C} push={prop.push} copy={prop.copy} graphs={graphs} rows {prop.lo}:{prop.hi} cols {c0}:{c1} halo {prop.n_halo} max rel err {worst:.2e} {'PASS' if ok else 'FAIL'}", flush=True)
flag = torch.tensor([0.0 if ok else 1.0], device="cuda")
dist.all_reduce(flag)
prop.close()
dist.destroy_process_group()
sys.exit(1 if flag.item() > 0 else 0)
