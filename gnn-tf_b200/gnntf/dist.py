"""Row-sharded multi-GPU APPNP propagation (one process per GPU, ``torch.distributed``).

The reference is single-process (no collectives anywhere); this module is the
B200 scale-out of its propagation loop (filter.py:17-22 under layered.py:52-55), following the
north_star: a contiguous node-range row split (balanced by nnz, not by rows — power-law graphs),
and per propagation step an exchange of exactly the halo feature rows each rank references,
overlapped with the SpMM of the rows that need no halo.

  rank r owns rows [lo_r, hi_r);  its CSR addresses  H_ext = [ owned rows | halo rows ]
  step:  pack rows peers need (native kernel)  ->  all-to-all (NCCL over NVLink)      [comm]
         fused APPNP step on INTERIOR rows (all columns owned)                        [compute, overlaps]
         wait for the halo  ->  fused APPNP step on BOUNDARY rows

:func:`build_shard_plan` is pure index logic on torch tensors (device-agnostic, covered by
world-size-2 gloo tests on CPU); :class:`ShardedPropagator` binds it to the native kernels.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch
import torch.distributed as dist


def partition_bounds(row_ptr, world):
    """Contiguous node ranges with (nearly) equal nnz: bounds[k] = first row whose row_ptr reaches
    k/world of nnz.  Identical on every rank (pure function of row_ptr)."""
    n = row_ptr.numel() - 1
    nnz = int(row_ptr[-1].item())
    targets = torch.tensor([(nnz * k) // world for k in range(1, world)], dtype=row_ptr.dtype, device=row_ptr.device)
    inner = torch.searchsorted(row_ptr[:-1].contiguous(), targets, right=False) if world > 1 else targets
    bounds = [0] + [int(min(max(int(x), 0), n)) for x in inner.tolist()] + [n]
    for k in range(1, len(bounds)):
        bounds[k] = max(bounds[k], bounds[k - 1])
    return bounds


@dataclass
class ShardPlan:
    rank: int
    world: int
    bounds: list            # world+1 row boundaries
    lo: int
    hi: int
    n_local: int
    n_halo: int
    row_ptr: torch.Tensor   # int32 [n_local+1], local
    col_idx: torch.Tensor   # int32 [nnz_local], localized: < n_local owned, else n_local + halo slot
    val: torch.Tensor       # fp32  [nnz_local]
    halo_cols: torch.Tensor  # int64 [n_halo] global ids, sorted (hence grouped by owner rank)
    recv_counts: list       # rows received from each rank
    send_counts: list       # rows sent to each rank
    send_idx: torch.Tensor  # int32 [sum(send_counts)] local row ids, grouped by destination rank
    interior_rows: torch.Tensor  # int32 local row ids with no halo column
    boundary_rows: torch.Tensor  # int32 local row ids with at least one halo column


def build_shard_plan(row_ptr, col_idx, val, rank, world, group=None, peer_wants=None):
    """Slice the (full, normalised) CSR to this rank's rows, remap its columns to
    [owned | halo] numbering, and agree with every peer on which rows to send.
    ``peer_wants`` (single-process emulation only): callable(dst_rank) -> the local row ids rank
    ``dst_rank`` wants from this rank, replacing the all-to-all of the request lists."""
    dev = row_ptr.device
    bounds = partition_bounds(row_ptr, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    n_local = hi - lo
    s, e = int(row_ptr[lo].item()), int(row_ptr[hi].item())
    rp = (row_ptr[lo:hi + 1] - row_ptr[lo]).to(torch.int32).contiguous()
    col = col_idx[s:e].to(torch.int64)
    v = val[s:e].contiguous()
    remote = (col < lo) | (col >= hi)
    halo_cols = torch.unique(col[remote], sorted=True)
    n_halo = int(halo_cols.numel())
    local_col = torch.where(remote, n_local + torch.searchsorted(halo_cols, col), col - lo).to(torch.int32).contiguous()
    # owner of each halo column -> how many rows each peer sends us
    b = torch.tensor(bounds, dtype=torch.int64, device=dev)
    owner = torch.searchsorted(b, halo_cols, right=True) - 1
    recv_counts = torch.bincount(owner, minlength=world).tolist() if n_halo else [0] * world
    # tell every owner which of its rows we need (ids relative to the owner's lo)
    want = (halo_cols - b[owner]).to(torch.int32).contiguous() if n_halo else torch.empty(0, dtype=torch.int32, device=dev)
    if world > 1 and peer_wants is not None:
        lists = [peer_wants(d).to(device=dev, dtype=torch.int32) if d != rank else want[:0] for d in range(world)]
        send_counts = [int(x.numel()) for x in lists]
        send_idx = torch.cat(lists).contiguous()
    elif world > 1:
        rc = torch.tensor(recv_counts, dtype=torch.int64, device=dev)
        sc = torch.empty_like(rc)
        dist.all_to_all_single(sc, rc, group=group)
        send_counts = sc.tolist()
        send_idx = torch.empty(int(sum(send_counts)), dtype=torch.int32, device=dev)
        dist.all_to_all_single(send_idx, want, output_split_sizes=send_counts, input_split_sizes=recv_counts, group=group)
    else:
        send_counts, send_idx = [0], torch.empty(0, dtype=torch.int32, device=dev)
    # interior / boundary rows
    deg = (rp[1:] - rp[:-1]).to(torch.int64)
    csum = torch.zeros(local_col.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(remote.to(torch.int64), 0, out=csum[1:])
    n_remote_in_row = csum[rp[1:].long()] - csum[rp[:-1].long()]
    is_boundary = n_remote_in_row > 0
    rows = torch.arange(n_local, dtype=torch.int32, device=dev)
    del deg
    return ShardPlan(rank, world, bounds, lo, hi, n_local, n_halo, rp, local_col, v, halo_cols, recv_counts,
                     send_counts, send_idx, rows[~is_boundary].contiguous(), rows[is_boundary].contiguous())


def wanted_rows(halo_cols, bounds, owner_rank):
    """Local row ids (relative to the owner's lo) that a rank with halo ``halo_cols`` needs from
    ``owner_rank`` — what the request all-to-all delivers to the owner."""
    lo, hi = bounds[owner_rank], bounds[owner_rank + 1]
    sel = halo_cols[(halo_cols >= lo) & (halo_cols < hi)]
    return (sel - lo).to(torch.int32)


def sub_csr(row_ptr, col_idx, val, rows):
    """Compact CSR of a row subset (``rows`` int32 ascending): (row_ptr, col_idx, val)."""
    dev = row_ptr.device
    r = rows.long()
    start, end = row_ptr[r].long(), row_ptr[r + 1].long()
    deg = end - start
    new_rp = torch.zeros(r.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=new_rp[1:])
    total = int(new_rp[-1].item())
    # entry j of the compact CSR comes from slot start[row_of(j)] + (j - new_rp[row_of(j)])
    row_of = torch.repeat_interleave(torch.arange(r.numel(), device=dev), deg, output_size=total)
    src = start[row_of] + (torch.arange(total, device=dev) - new_rp[row_of])
    return new_rp.to(torch.int32).contiguous(), col_idx[src].contiguous(), val[src].contiguous()


def exchange_halo(plan: ShardPlan, send_buf, halo_out, group=None, async_op=False):
    """All-to-all of the packed rows: ``send_buf`` [sum(send_counts), F] grouped by destination,
    ``halo_out`` [n_halo, F] grouped by source (= sorted halo order)."""
    if plan.world == 1:
        return None
    return dist.all_to_all_single(halo_out, send_buf, output_split_sizes=plan.recv_counts,
                                  input_split_sizes=plan.send_counts, group=group, async_op=async_op)


class ShardedPropagator:
    """APPNP K-step propagation of one shard on one GPU (see module docstring)."""

    def __init__(self, adj, A, F, rank, world, group=None, plan=None, exchange=None):
        from . import _native as nat
        from .sparse import CsrStructure
        self.nat = nat
        self.group, self.F = group, int(F)
        self._exchange = exchange  # test hook: single-process emulation of the all-to-all
        csr = A.csr
        self.plan = p = plan if plan is not None else build_shard_plan(csr.row_ptr, csr.col_idx, A.val, rank, world, group)
        self.lo, self.hi, self.n_local, self.n_halo = p.lo, p.hi, p.n_local, p.n_halo
        self.nnz_local = int(p.col_idx.numel())
        dev = p.row_ptr.device

        def make(rows):
            rp, col, val = sub_csr(p.row_ptr, p.col_idx, p.val, rows)
            st = CsrStructure(int(rows.numel()), rp, col, None)
            st.row_map = rows.contiguous()
            return st, val
        self.interior, self.interior_val = make(p.interior_rows)
        self.boundary, self.boundary_val = make(p.boundary_rows)
        n_ext = self.n_local + self.n_halo
        self.buf = [torch.zeros((n_ext, self.F), dtype=torch.float32, device=dev) for _ in range(2)]
        self.send_buf = torch.empty((int(sum(p.send_counts)), self.F), dtype=torch.float32, device=dev)
        self.H0 = torch.empty((self.n_local, self.F), dtype=torch.float32, device=dev)

    def launches_per_propagation(self, K):
        per_step = 1  # pack
        for st in (self.interior, self.boundary):
            if st.n > 0:
                per_step += 3 if st.n_long > 0 else 1
        return K * per_step

    def _step(self, src, dst, alpha):
        nat, L, p = self.nat, self.nat.lib(), self.plan
        F, st = self.F, self.nat.stream_ptr()
        work = None
        if p.world > 1:
            if self.send_buf.shape[0] > 0:
                nat.check(L.gnntf_halo_pack_f32(nat.ptr(src), F, nat.ptr(p.send_idx), self.send_buf.shape[0],
                                                nat.ptr(self.send_buf), F, F, st), "halo_pack")
            if self._exchange is not None:
                self._exchange(self, self.send_buf, src[self.n_local:])
            else:
                work = exchange_halo(p, self.send_buf, src[self.n_local:], self.group, async_op=True)
        for structure, val, wait in ((self.interior, self.interior_val, False), (self.boundary, self.boundary_val, True)):
            if wait and work is not None:
                work.wait()  # the compute stream waits for the halo rows
            if structure.n == 0:
                continue
            s = structure.struct(val, F)
            nat.check(L.gnntf_appnp_step_f32(ctypes.byref(s), nat.ptr(src), nat.ptr(self.H0), nat.ptr(dst), F, F,
                                             float(alpha), None, 1.0, nat.ACT_IDENTITY, st), "appnp_step")
        if work is not None and self.boundary.n == 0:
            work.wait()

    def propagate(self, H0_local, alpha=0.1, iterations=10):
        """K fused steps on this shard; returns this rank's rows of H_K ([n_local, F] view)."""
        self.H0.copy_(H0_local)
        src, dst = self.buf
        src[:self.n_local].copy_(self.H0)
        for _ in range(iterations):
            self._step(src, dst, alpha)
            src, dst = dst, src
        return src[:self.n_local]

    def propagate_host_timed(self, alpha, iterations, reps=3):
        """End-to-end: pinned host H0 shard -> device, K steps, result shard -> host."""
        import time
        host_in = torch.randn((self.n_local, self.F), dtype=torch.float32).pin_memory()
        host_out = torch.empty_like(host_in).pin_memory()
        dev_in = torch.empty((self.n_local, self.F), dtype=torch.float32, device=self.H0.device)

        def once():
            dev_in.copy_(host_in, non_blocking=True)
            out = self.propagate(dev_in, alpha, iterations)
            host_out.copy_(out, non_blocking=True)
            torch.cuda.synchronize()
        once()
        if self.plan.world > 1:
            dist.barrier(group=self.group)
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        if self.plan.world > 1:
            dist.barrier(group=self.group)
        sec = (time.perf_counter() - t0) / reps
        t = torch.tensor([sec], dtype=torch.float64, device=self.H0.device)
        nbytes = torch.tensor([float(host_in.numel() * 4)], dtype=torch.float64, device=self.H0.device)
        if self.plan.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            dist.all_reduce(nbytes, op=dist.ReduceOp.SUM, group=self.group)
        return {"seconds": float(t.item()), "h2d": int(nbytes.item()), "d2h": int(nbytes.item())}
