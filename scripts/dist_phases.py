#!/usr/bin/env python
"""Per-phase timing of one sharded propagation step (pack / all-to-all / interior / boundary),
each phase synchronised so nothing overlaps — run under torchrun.  Diagnostic only."""
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
ordering = sys.argv[1] if len(sys.argv) > 1 else "local"
n, edges = synthetic.shaped_edges("products", seed=0, ordering=ordering, device="cuda")
adj = gnntf.edges2adj(edges, None, n)
A = adj.normalized("symmetric")
del edges


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


for halves in (1, 2):
    prop = gdist.ShardedPropagator(adj, A, 100, rank, world, halves=halves)
    H0 = synthetic.features(prop.n_local, 100, 1, "cuda")
    full = t(lambda: prop.propagate(H0, 0.1, 10), reps=3) / 10
    out = []
    for part in prop.parts:
        src, dst = part["buf"]
        nat, L = prop.nat, prop.nat.lib()
        F = part["F"]

        def pack():
            nat.check(L.gnntf_halo_pack_f32(nat.ptr(src), F, nat.ptr(prop.plan.send_idx), part["send"].shape[0],
                                            nat.ptr(part["send"]), F, F, nat.stream_ptr()))

        def a2a():
            gdist.exchange_halo(prop.plan, part["send"], src[prop.n_local:], None, async_op=False)

        def pass1():
            import ctypes
            s = prop.owned.struct(prop.owned_val, F)
            nat.check(L.gnntf_appnp_step_f32(ctypes.byref(s), nat.ptr(src), nat.ptr(part["H0"]), nat.ptr(dst), F, F, 0.1,
                                             None, 1.0, 0, nat.stream_ptr()))

        def pass2():
            import ctypes
            s = prop.halo_part.struct(prop.halo_val, F)
            nat.check(L.gnntf_spmm_acc_f32(ctypes.byref(s), nat.ptr(src), F, nat.ptr(dst), F, F, 0.9, nat.stream_ptr()))
        out.append(dict(F=F, pack=t(pack), a2a=t(a2a), pass1_owned=t(pass1), pass2_halo=t(pass2),
                        send_MB=part["send"].numel() * 4 / 1e6, owned_nnz=prop.owned.nnz, halo_nnz=prop.halo_part.nnz))
    if rank == 0:
        print(f"world={world} ordering={ordering} halves={halves} step_ms={full:.3f} n_halo={prop.n_halo} halo_rows={prop.halo_part.n}")
        for o in out:
            print("   ", {k: round(v, 3) for k, v in o.items()})
    del prop
    torch.cuda.empty_cache()
dist.destroy_process_group()
