#!/usr/bin/env python
"""Per-phase kernel times of ONE rank's sharded step, measured on a single GPU: the shard plans of all
ranks are built in this process (no communication), rank 0's propagator is wired to local peer buffers,
and its owned-column pass, halo-column pass and push kernel are timed separately.
usage: python scripts/shard_phases_1gpu.py [world ...]   (products shape; F=100, and F=52 for the 4x2 grid)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnn-tf_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gnntf  # noqa: E402
import synthetic  # noqa: E402
from gnntf import dist as gdist  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def main():
    cases = [(int(w), 100) for w in (sys.argv[1:] or ["2", "4"])] + [(4, 52), (4, 64), (4, 48)]
    n, edges = synthetic.shaped_edges("products", seed=0, device="cuda")
    adj = gnntf.edges2adj(edges, None, n)
    A = adj.normalized("symmetric")
    csr = A.csr
    for world, F in cases:
        bounds = gdist.partition_bounds(csr.row_ptr, world)
        first = [gdist.build_shard_plan(csr.row_ptr, csr.col_idx, A.val, r, world, peer_wants=lambda d: torch.empty(0)) for r in range(world)]
        r = 1 if world > 2 else 0   # an inner rank has halo on both sides
        plan = gdist.build_shard_plan(csr.row_ptr, csr.col_idx, A.val, r, world,
                                      peer_wants=lambda d, r=r: gdist.wanted_rows(first[d].halo_cols, bounds, r))
        halos = [p.n_halo for p in first]
        del first
        prop = gdist.ShardedPropagator(adj, A, F, r, world, plan=plan, push=True, peers="local")
        # aim every peer pointer at one local dummy halo buffer big enough for the largest destination
        dummy = torch.empty((max(halos) + n + 8, F), device="cuda")
        everyone = [dict(recv_counts=[0] * world, n_local=0) for _ in range(world)]
        prop._build_tables(everyone, lambda q, pi, bi: dummy.data_ptr(), lambda q: prop._flags.ptr)
        part = prop.parts[0]
        src, dst = part["buf"]
        src.normal_()
        part["H0"].normal_()
        t1 = timed(lambda: prop._pass1(part, src, dst, 0.1))
        t2 = timed(lambda: prop._pass2(part, src, dst, 0.1))
        tp = timed(lambda: prop._push(part, src, 1))
        tf = timed(lambda: prop._step_push(part, src, dst, 0.1, 1))     # push + pass 1 in ONE launch
        print(json.dumps({"world_rows": world, "F": F, "rank": r, "n_local": prop.n_local, "n_halo": prop.n_halo,
                          "owned_entries": prop.owned.nnz, "halo_entries": prop.halo_part.nnz, "halo_rows": prop.halo_part.n,
                          "n_send": int(plan.send_idx.numel()), "pass1_ms": t1, "pass2_ms": t2, "push_local_ms": tp, "fused_push_pass1_ms": tf,
                          "push_GBps_local": plan.send_idx.numel() * F * 4 / tp / 1e6}), flush=True)
        prop.close()
        del prop, dummy


if __name__ == "__main__":
    main()
