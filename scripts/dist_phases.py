#!/usr/bin/env python
"""Per-phase timing of one sharded propagation step, phases serialised so nothing overlaps: fused
peer-memory push (+ completion barrier), NCCL pack + all-to-all for comparison, owned-column pass,
halo-column pass, and the real (overlapped) step.  Run under torchrun.  Diagnostic only.
usage: dist_phases.py [local|random] [ROWSxCOLS]"""
import ctypes
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
os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ordering = sys.argv[1] if len(sys.argv) > 1 else "local"
F = 100
R, C = (int(x) for x in sys.argv[2].split("x")) if len(sys.argv) > 2 else gdist.choose_grid(world, F)
n, edges = synthetic.shaped_edges("products", seed=0, ordering=ordering, device="cuda")
adj = gnntf.edges2adj(edges, None, n)
A = adj.normalized("symmetric")
del edges
grid = gdist.Grid2D(rank, world, R, C)
c0, c1 = gdist.column_range(F, C, grid.c)
Fc = c1 - c0


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


prop = gdist.ShardedPropagator(adj, A, Fc, grid.r, R, group=grid.row_group)
H0 = synthetic.features(prop.n_local, Fc, 1, "cuda")
full = t(lambda: prop.propagate(H0, 0.1, 10), reps=3) / 10
part = prop.parts[0]
src, dst = part["buf"]
nat, L, p = prop.nat, prop.nat.lib(), prop.plan
res = dict(step_ms=full, n_local=prop.n_local, n_halo=prop.n_halo, send_MB=p.send_idx.numel() * Fc * 4 / 1e6,
           recv_MB=prop.n_halo * Fc * 4 / 1e6, owned_nnz=prop.owned.nnz, halo_nnz=prop.halo_part.nnz, halo_rows=prop.halo_part.n)
if R > 1:
    def push():
        nat.check(L.gnntf_halo_push_f32(nat.ptr(src), Fc, nat.ptr(p.send_idx), nat.ptr(prop._send_off),
                                        nat.ptr(prop._peer_ptrs[0][0]), nat.ptr(prop._peer_row0), p.world,
                                        int(p.send_idx.numel()), prop._rotate, Fc, Fc, nat.stream_ptr()))

    def push_barrier():
        push()
        dist.all_reduce(prop._flag, group=grid.row_group)

    send = prop._send_buffer(part)

    def pack_a2a():
        nat.check(L.gnntf_halo_pack_f32(nat.ptr(src), Fc, nat.ptr(p.send_idx), send.shape[0], nat.ptr(send),
                                        Fc, Fc, nat.stream_ptr()))
        gdist.exchange_halo(p, send, src[prop.n_local:], grid.row_group, async_op=False)
    res.update(push=t(push), push_plus_barrier=t(push_barrier), barrier=t(lambda: dist.all_reduce(prop._flag, group=grid.row_group)),
               nccl_pack_a2a=t(pack_a2a))


def pass1():
    s = prop.owned.struct(prop.owned_val, Fc)
    nat.check(L.gnntf_appnp_step_f32(ctypes.byref(s), nat.ptr(src), nat.ptr(part["H0"]), nat.ptr(dst), Fc, Fc, 0.1, None, 1.0, 0,
                                     nat.stream_ptr()))


def pass2():
    if prop.halo_part.n == 0:
        return
    s = prop.halo_part.struct(prop.halo_val, Fc)
    nat.check(L.gnntf_spmm_acc_f32(ctypes.byref(s), nat.ptr(src), Fc, nat.ptr(dst), Fc, Fc, 0.9, nat.stream_ptr()))


res.update(pass1_owned=t(pass1), pass2_halo=t(pass2))
if rank == 0:
    print(f"world={world} grid={R}x{C} ordering={ordering} cols={Fc}", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in res.items()}, flush=True)
dist.barrier()
dist.destroy_process_group()
