"""world_size-2 (and 3) gloo tests of the row-sharded path's host logic on CPU: nnz-balanced
contiguous partition, column remapping to [owned | halo], the per-peer send lists agreed through
all-to-all, the per-step halo exchange, and the interior/boundary row split.  The per-shard SpMM is
played by the CPU oracle here (tests may use it as the checker); on a GPU box the same plan drives
the native kernels (tests/test_gpu_parity.py::test_sharded_propagator_single_process)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, edges, F, K, alpha, out_dir):
    for p in (os.path.join(ROOT, "gnn-tf_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gnntf_oracle as oracle
        from gnntf import dist as gdist
        idx, val, _ = oracle.graph2adj_arrays(edges, None, n)
        _, nv, _ = oracle.get_adjacency(idx, val, n)
        row_ptr, col_idx, coo_pos, _ = oracle.csr_from_coo(idx, n)
        plan = gdist.build_shard_plan(torch.from_numpy(row_ptr), torch.from_numpy(col_idx.astype(np.int32)),
                                      torch.from_numpy(nv[coo_pos]), rank, world)
        # structural checks
        assert plan.bounds[0] == 0 and plan.bounds[-1] == n and plan.lo == plan.bounds[rank]
        assert int(plan.interior_rows.numel() + plan.boundary_rows.numel()) == plan.n_local
        assert sum(plan.recv_counts) == plan.n_halo and plan.recv_counts[rank] == 0
        assert int(plan.send_idx.numel()) == sum(plan.send_counts)
        if plan.n_halo:
            hc = plan.halo_cols.numpy()
            assert np.all(np.diff(hc) > 0) and np.all((hc < plan.lo) | (hc >= plan.hi))
        if plan.send_idx.numel():
            assert int(plan.send_idx.min()) >= 0 and int(plan.send_idx.max()) < plan.n_local
        rp, col = plan.row_ptr.numpy(), plan.col_idx.numpy()
        for r in plan.interior_rows.tolist():
            assert np.all(col[rp[r]:rp[r + 1]] < plan.n_local)
        for r in plan.boundary_rows.tolist():
            assert np.any(col[rp[r]:rp[r + 1]] >= plan.n_local)
        # emulate the propagation: per step pack -> all-to-all -> interior step -> boundary step
        H0_full = np.random.default_rng(5).standard_normal((n, F)).astype(np.float32)
        H0 = H0_full[plan.lo:plan.hi]
        local_rows = np.repeat(np.arange(plan.n_local), np.diff(rp))
        local_idx = np.stack([local_rows, col.astype(np.int64)], 1)
        H = H0.copy()
        for _ in range(K):
            ext = np.zeros((plan.n_local + plan.n_halo, F), np.float32)
            ext[:plan.n_local] = H
            send = torch.from_numpy(H[plan.send_idx.long().numpy()].copy())
            halo = torch.empty((plan.n_halo, F), dtype=torch.float32)
            gdist.exchange_halo(plan, send, halo)
            ext[plan.n_local:] = halo.numpy()
            P = oracle.spmm_coo(local_idx, plan.val.numpy(), ext, n_rows=plan.n_local)
            H = P * np.float32(1 - alpha) + H0 * np.float32(alpha)
        expect = oracle.appnp_propagate(idx, val, n, H0_full, alpha, K)[-1][plan.lo:plan.hi]
        oracle.assert_close(H, expect, what=f"rank {rank} shard")
        # sub_csr of the boundary rows reproduces those rows
        srp, scol, sval = gdist.sub_csr(plan.row_ptr, plan.col_idx, plan.val, plan.boundary_rows)
        for i, r in enumerate(plan.boundary_rows.tolist()[:20]):
            assert np.array_equal(scol[srp[i]:srp[i + 1]].numpy(), col[rp[r]:rp[r + 1]])
        # all-gather layout of the copy-engine exchange: columns back in global numbering, every rank holds all rows,
        # per step an all-gather of the row blocks, then the owned-column pass and the halo-column pass
        (o_rp, o_col, o_val), (h_rp, h_col, h_val), h_rows = gdist.split_by_column(plan.row_ptr, plan.col_idx, plan.val, plan.n_local)
        og, hg = gdist.global_columns(plan, o_col, h_col)
        assert og.numel() + hg.numel() == col.size
        if og.numel():
            assert int(og.min()) >= plan.lo and int(og.max()) < plan.hi
        if hg.numel():
            assert bool(((hg < plan.lo) | (hg >= plan.hi)).all())
        full_cols = torch.from_numpy(col_idx[row_ptr[plan.lo]:row_ptr[plan.hi]].astype(np.int64))
        merged = torch.where(plan.col_idx.long() < plan.n_local, plan.col_idx.long() + plan.lo,
                             plan.halo_cols[(plan.col_idx.long() - plan.n_local).clamp(min=0)] if plan.n_halo else plan.col_idx.long())
        assert torch.equal(merged, full_cols)
        o_rows = np.repeat(np.arange(plan.n_local), np.diff(o_rp.numpy()))
        h_rows_e = np.repeat(h_rows.numpy(), np.diff(h_rp.numpy()))
        sizes = [plan.bounds[r + 1] - plan.bounds[r] for r in range(world)]
        H = H0.copy()
        for _ in range(K):
            blocks = [torch.empty((max(sizes), F), dtype=torch.float32) for _ in sizes]   # gloo gathers equal shapes
            mine = torch.zeros((max(sizes), F), dtype=torch.float32)
            mine[:plan.n_local] = torch.from_numpy(H)
            dist.all_gather(blocks, mine)
            full = torch.cat([b[:sz] for b, sz in zip(blocks, sizes)]).numpy()
            P = oracle.spmm_coo(np.stack([o_rows, og.numpy().astype(np.int64)], 1), o_val.numpy(), full, n_rows=plan.n_local)
            if hg.numel():
                P = P + oracle.spmm_coo(np.stack([h_rows_e.astype(np.int64), hg.numpy().astype(np.int64)], 1), h_val.numpy(), full,
                                        n_rows=plan.n_local)
            H = P * np.float32(1 - alpha) + H0 * np.float32(alpha)
        oracle.assert_close(H, expect, what=f"rank {rank} shard, all-gather layout", floor=oracle.FLOOR_REORDERED)
        np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([plan.n_local, plan.n_halo]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_plan_and_exchange_gloo(world, tmp_path):
    rng = np.random.default_rng(0)
    n, e, F, K = 400, 3000, 6, 4
    edges = rng.integers(0, n, (e, 2))
    edges[:200, 0] = 7  # a hub row so the nnz-balanced split differs from an equal-rows split
    mp.spawn(_worker, args=(world, _free_port(), n, edges, F, K, 0.1, str(tmp_path)), nprocs=world, join=True)
    sizes = [np.load(tmp_path / f"ok_{r}.npy") for r in range(world)]
    assert sum(int(s[0]) for s in sizes) == n


def test_partition_is_nnz_balanced():
    sys.path.insert(0, os.path.join(ROOT, "gnn-tf_b200"))
    from gnntf import dist as gdist
    deg = torch.tensor([1000] + [1] * 999)
    row_ptr = torch.zeros(1001, dtype=torch.int64)
    torch.cumsum(deg, 0, out=row_ptr[1:])
    b = gdist.partition_bounds(row_ptr, 4)
    assert b[0] == 0 and b[-1] == 1000 and all(b[i] <= b[i + 1] for i in range(4))
    loads = [int(row_ptr[b[i + 1]] - row_ptr[b[i]]) for i in range(4)]
    assert max(loads) <= 1000 + 10 and sum(loads) == 1999       # the hub row is a shard of its own
    assert gdist.partition_bounds(row_ptr, 1) == [0, 1000]


def test_split_by_column_preserves_entries_and_order():
    sys.path.insert(0, os.path.join(ROOT, "gnn-tf_b200"))
    from gnntf import dist as gdist
    rng = np.random.default_rng(4)
    n_local, n_halo = 50, 30
    deg = rng.integers(0, 12, n_local)
    deg[3] = 0
    row_ptr = torch.zeros(n_local + 1, dtype=torch.int32)
    row_ptr[1:] = torch.from_numpy(np.cumsum(deg)).to(torch.int32)
    nnz = int(row_ptr[-1])
    col = torch.from_numpy(rng.integers(0, n_local + n_halo, nnz)).to(torch.int32)
    val = torch.from_numpy(rng.random(nnz).astype(np.float32))
    (o_rp, o_col, o_val), (h_rp, h_col, h_val), rows = gdist.split_by_column(row_ptr, col, val, n_local)
    assert int(o_rp[-1]) + int(h_rp[-1]) == nnz and o_rp.numel() == n_local + 1
    assert bool((o_col < n_local).all()) and bool((h_col >= n_local).all())
    hmap = {int(r): i for i, r in enumerate(rows.tolist())}
    for r in range(n_local):
        s, e = int(row_ptr[r]), int(row_ptr[r + 1])
        c, v = col[s:e], val[s:e]
        own = c < n_local
        assert torch.equal(o_col[int(o_rp[r]):int(o_rp[r + 1])], c[own]) and torch.equal(o_val[int(o_rp[r]):int(o_rp[r + 1])], v[own])
        if (~own).any():
            i = hmap[r]
            assert torch.equal(h_col[int(h_rp[i]):int(h_rp[i + 1])], c[~own]) and torch.equal(h_val[int(h_rp[i]):int(h_rp[i + 1])], v[~own])
        else:
            assert r not in hmap


def test_grid_and_column_helpers():
    sys.path.insert(0, os.path.join(ROOT, "gnn-tf_b200"))
    from gnntf import dist as gdist
    assert gdist.choose_grid(1, 100) == (1, 1) and gdist.choose_grid(2, 100) == (2, 1)
    assert gdist.choose_grid(4, 100) == (4, 1) and gdist.choose_grid(8, 100) == (4, 2)
    assert gdist.choose_grid(8, 47) == (8, 1)            # too narrow to split the columns
    for F, C in ((100, 2), (100, 4), (47, 2), (128, 8), (7, 2)):
        ranges = [gdist.column_range(F, C, c) for c in range(C)]
        assert ranges[0][0] == 0 and ranges[-1][1] == F
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))          # contiguous cover
        assert all((lo % 4 == 0) for lo, _ in ranges)                          # 16-byte aligned starts


def _grid_worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "gnn-tf_b200"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gnntf import dist as gdist
        grid = gdist.Grid2D(rank, world, 2, 2)
        assert (grid.r, grid.c) == (rank // 2, rank % 2)
        t = torch.tensor([float(rank)])
        dist.all_reduce(t, group=grid.row_group)             # sums over the ranks of my column group
        assert t.item() == float(grid.c + (2 + grid.c))      # ranks c and 2+c
        np.save(os.path.join(out_dir, f"grid_{rank}.npy"), np.array([grid.r, grid.c]))
    finally:
        dist.destroy_process_group()


def test_grid2d_process_groups_gloo(tmp_path):
    mp.spawn(_grid_worker, args=(4, _free_port(), str(tmp_path)), nprocs=4, join=True)
    assert sorted(tuple(np.load(tmp_path / f"grid_{r}.npy")) for r in range(4)) == [(0, 0), (0, 1), (1, 0), (1, 1)]
