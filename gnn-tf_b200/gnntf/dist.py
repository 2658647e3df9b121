"""Row-sharded multi-GPU APPNP propagation (one process per GPU, ``torch.distributed``).

The reference is single-process (no collectives anywhere); this module is the
B200 scale-out of its propagation loop (filter.py:17-22 under layered.py:52-55), following the
north_star: a contiguous node-range row split (balanced by nnz, not by rows — power-law graphs),
and per propagation step an exchange of exactly the halo feature rows each rank references,
overlapped with the SpMM of the rows that need no halo.

  rank r owns rows [lo_r, hi_r);  its CSR addresses  H_ext = [ owned rows | halo rows ]
  step:  pack rows peers need (native kernel)  ->  all-to-all (NCCL over NVLink)      [comm]
         fused APPNP step of ALL rows over their OWNED columns (the bulk of the entries;
         needs no halo)                                                               [compute, overlaps]
         wait for the halo  ->  accumulate the HALO-column entries of the boundary rows
(the shard's CSR is split by column, not by row: on power-law graphs almost every row touches
some remote column, so a row split leaves nothing to overlap with — measured 6 % interior rows
on the products shape at 2 GPUs).

:func:`build_shard_plan` is pure index logic on torch tensors (device-agnostic, covered by
world-size-2 gloo tests on CPU); :class:`ShardedPropagator` binds it to the native kernels.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch
import torch.distributed as dist


def choose_grid(world, F):
    """(row_groups, column_groups) for ``world`` GPUs.  Feature columns propagate independently
    through H <- (1-a)·Â·H + a·H0, so splitting them needs NO communication and shrinks every
    rank's halo bytes by the column factor; but narrower feature rows gather less efficiently and
    every column group re-reads the whole CSR of its row group.  Measured on the products shape,
    F=100: rows only up to 4 GPUs (2x1 33.1 ms vs 1x2 35.6; 4x1 20.2 vs 2x2 21.7),
    two column groups at 8 (4x2 16.4 ms vs 8x1 17.2)."""
    cols = 2 if (world >= 8 and world % 2 == 0 and F >= 96) else 1
    return world // cols, cols


def column_range(F, col_groups, c):
    """Columns [c0, c1) of column group c: equal shares rounded up to a multiple of 4 floats."""
    share = ((F + col_groups - 1) // col_groups + 3) // 4 * 4
    c0 = min(F, c * share)
    return c0, min(F, c0 + share)


class Grid2D:
    """rank -> (row group r of R, column group c of C) with one process group per column group
    (its R ranks exchange halo rows among themselves; column groups never talk)."""

    def __init__(self, rank, world, rows, cols):
        assert rows * cols == world
        self.rank, self.world, self.R, self.C = rank, world, rows, cols
        self.r, self.c = rank // cols, rank % cols
        self.row_group = None
        if world > 1:
            for c in range(cols):  # every rank creates every group, in the same order
                g = dist.new_group(ranks=[r * cols + c for r in range(rows)])
                if c == self.c:
                    self.row_group = g


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


def split_by_column(row_ptr, col_idx, val, n_local):
    """Split a shard-local CSR into the entries over OWNED columns (all rows, identity row map) and
    the entries over HALO columns (compact CSR over the rows that have any, plus those row ids).
    Entry order inside a row is preserved in both parts."""
    dev = row_ptr.device
    n = row_ptr.numel() - 1
    is_halo = col_idx >= n_local
    csum = torch.zeros(col_idx.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(is_halo.to(torch.int64), 0, out=csum[1:])
    rp = row_ptr.long()
    halo_deg = csum[rp[1:]] - csum[rp[:-1]]
    deg = rp[1:] - rp[:-1]
    own_rp = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg - halo_deg, 0, out=own_rp[1:])
    own = (own_rp.to(torch.int32).contiguous(), col_idx[~is_halo].contiguous(), val[~is_halo].contiguous())
    rows = torch.nonzero(halo_deg > 0).flatten().to(torch.int32).contiguous()
    halo_rp = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(halo_deg[rows.long()], 0, out=halo_rp[1:])
    halo = (halo_rp.to(torch.int32).contiguous(), col_idx[is_halo].contiguous(), val[is_halo].contiguous())
    return own, halo, rows


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


class _SharedMatrix:
    """fp32 [rows, cols] device matrix allocated through gnntf_ipc_alloc so that peers can map it
    (CUDA IPC) and store halo rows into it directly; exposed to torch without a copy."""

    def __init__(self, rows, cols, device):
        from . import _native as nat
        self.nat, self.rows, self.cols = nat, int(rows), int(cols)
        ptr = ctypes.c_void_p()
        self.handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(device):
            nat.check(nat.lib().gnntf_ipc_alloc(max(16, self.rows * self.cols * 4), ctypes.byref(ptr), self.handle), "ipc_alloc")
        self.ptr = ptr.value
        self.__cuda_array_interface__ = {"shape": (self.rows, self.cols), "typestr": "<f4", "data": (self.ptr, False),
                                         "version": 2, "strides": None}
        self.tensor = torch.as_tensor(self, device=device)
        self.tensor.zero_()

    def handle_bytes(self):
        return bytes(self.handle)

    def __del__(self):
        try:
            self.nat.lib().gnntf_ipc_free(ctypes.c_void_p(self.ptr))
        except Exception:
            pass


class _EventWork:
    """``work.wait()`` for the single-process emulation hook: the compute stream waits for an event."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class ShardedPropagator:
    """APPNP K-step propagation of one shard on one GPU (see module docstring).

    ``push`` (default): halo rows travel by the fused pack+send kernel over NVLink peer memory
    (``gnntf_halo_push_f32``); if the peers' buffers cannot be mapped every rank falls back to the
    NCCL all-to-all.  ``halves=2`` runs two independent chains over the two halves of the feature
    columns, software-pipelined so one half's exchange is in flight while the other half computes —
    kept for experiments, measured slower than the default."""

    def __init__(self, adj, A, F, rank, world, group=None, plan=None, exchange=None, halves=None, push=True):
        from . import _native as nat
        from .sparse import CsrStructure
        self.nat = nat
        self.group, self.F = group, int(F)
        self._exchange = exchange  # test hook: single-process emulation of the all-to-all
        self.comm_stream = torch.cuda.Stream(priority=-1) if (world > 1 and torch.cuda.is_available()) else None
        csr = A.csr
        self.plan = p = plan if plan is not None else build_shard_plan(csr.row_ptr, csr.col_idx, A.val, rank, world, group)
        self.lo, self.hi, self.n_local, self.n_halo = p.lo, p.hi, p.n_local, p.n_halo
        self.nnz_local = int(p.col_idx.numel())
        dev = p.row_ptr.device

        (o_rp, o_col, o_val), (h_rp, h_col, h_val), h_rows = split_by_column(p.row_ptr, p.col_idx, p.val, self.n_local)
        self.owned, self.owned_val = CsrStructure(self.n_local, o_rp, o_col, None), o_val
        self.halo_part, self.halo_val = CsrStructure(int(h_rows.numel()), h_rp, h_col, None), h_val
        self.halo_part.row_map = h_rows
        self.interior, self.boundary = self.owned, self.halo_part  # (names kept for reports: pass 1 / pass 2)
        if halves is None:
            halves = 1  # column-half pipelining measured slower than the column split at N=2
        first = ((self.F + halves - 1) // halves + 3) // 4 * 4 if halves > 1 else self.F
        widths = [first, self.F - first] if halves > 1 and self.F - first > 0 else [self.F]
        n_ext, n_send = self.n_local + self.n_halo, int(sum(p.send_counts))
        self.push = bool(push) and p.world > 1 and exchange is None
        self.parts, col0, self._shared = [], 0, []
        for w in widths:
            if self.push:
                mats = [_SharedMatrix(n_ext, w, dev) for _ in range(2)]
                self._shared.append(mats)
                bufs = [m.tensor for m in mats]
            else:
                bufs = [torch.zeros((n_ext, w), dtype=torch.float32, device=dev) for _ in range(2)]
            self.parts.append(dict(F=w, col0=col0, buf=bufs, send=None, n_send=n_send,   # send buffer: NCCL path only
                                   H0=torch.empty((self.n_local, w), dtype=torch.float32, device=dev), work=None))
            col0 += w
        if self.push:
            # every rank of the row group must take the same path: agree on whether IPC mapping worked
            try:
                self._map_peers()
                ok = 1.0
            except Exception as err:  # e.g. peers without CUDA IPC / peer access
                ok, self._map_error = 0.0, err
            flag = torch.tensor([ok], dtype=torch.float32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if flag.item() < 1.0:
                self.push = False  # the shared buffers are ordinary device memory for the NCCL path
        if not self.push:
            for part in self.parts:
                self._send_buffer(part)
        # single-part aliases (tests and the emulation hook address them directly)
        self.buf, self.send_buf, self.H0 = self.parts[0]["buf"], self.parts[0]["send"], self.parts[0]["H0"]

    def _send_buffer(self, part):
        if part["send"] is None:
            part["send"] = torch.empty((part["n_send"], part["F"]), dtype=torch.float32, device=part["H0"].device)
        return part["send"]

    def _map_peers(self):
        """Exchange IPC handles and halo layouts inside the row group and build, per part and per
        ping-pong buffer, the device tables gnntf_halo_push_f32 needs."""
        nat, L, p = self.nat, self.nat.lib(), self.plan
        dev = p.row_ptr.device
        mine = dict(handles=[[m.handle_bytes() for m in mats] for mats in self._shared],
                    recv_counts=list(p.recv_counts), n_local=self.n_local)
        everyone = [None] * p.world
        dist.all_gather_object(everyone, mine, group=self.group)
        send_off = [0]
        for c in p.send_counts:
            send_off.append(send_off[-1] + int(c))
        self._send_off = torch.tensor(send_off, dtype=torch.int64, device=dev)
        nxt = send_off[(p.rank + 1) % p.world]
        self._rotate = int(nxt) if nxt < send_off[-1] else 0   # start with the next rank's rows
        # my rows land in peer q's halo region after the rows of lower-ranked owners
        row0 = [int(everyone[q]["n_local"]) + int(sum(everyone[q]["recv_counts"][:p.rank])) for q in range(p.world)]
        self._peer_row0 = torch.tensor(row0, dtype=torch.int64, device=dev)
        self._peer_ptrs, self._opened = [], []
        for pi in range(len(self.parts)):
            per_buf = []
            for bi in range(2):
                ptrs = []
                for q in range(p.world):
                    if q == p.rank or p.send_counts[q] == 0:
                        ptrs.append(0)
                        continue
                    h = (ctypes.c_ubyte * 64).from_buffer_copy(everyone[q]["handles"][pi][bi])
                    out = ctypes.c_void_p()
                    nat.check(L.gnntf_ipc_open(h, ctypes.byref(out)), "ipc_open")
                    self._opened.append(out.value)
                    ptrs.append(out.value)
                per_buf.append(torch.tensor(ptrs, dtype=torch.int64, device=dev))
            self._peer_ptrs.append(per_buf)
        self._flag = torch.zeros(1, dtype=torch.float32, device=dev)

    def launches_per_propagation(self, K):
        per_step = 1 if self.plan.world > 1 else 0  # pack
        for st in (self.owned, self.halo_part):
            if st.n > 0 and st.nnz > 0 or st is self.owned:
                per_step += 2 if st.n_long > 0 else 1
        return K * per_step * len(self.parts)

    # -- one half: exchange and compute -----------------------------------------------------
    def _start_exchange(self, part, src):
        """Pack + all-to-all of this part's halo rows on the comm stream, ordered after everything
        already enqueued on the compute stream (the step that produced ``src``)."""
        nat, L, p = self.nat, self.nat.lib(), self.plan
        part["work"] = None
        if p.world == 1:
            return
        F = part["F"]
        ready = torch.cuda.Event()
        ready.record()                                   # src complete on the compute stream
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
            if self.push:
                # fused pack + send: rows go straight into the peers' halo regions of the SAME
                # ping-pong buffer; a one-element all-reduce is the cross-rank completion barrier
                pi = self.parts.index(part)
                bi = 0 if src.data_ptr() == part["buf"][0].data_ptr() else 1
                n_send = int(p.send_idx.numel())
                if n_send > 0:
                    nat.check(L.gnntf_halo_push_f32(nat.ptr(src), F, nat.ptr(p.send_idx), nat.ptr(self._send_off),
                                                    nat.ptr(self._peer_ptrs[pi][bi]), nat.ptr(self._peer_row0), p.world,
                                                    n_send, self._rotate, F, F, nat.stream_ptr()), "halo_push")
                part["work"] = dist.all_reduce(self._flag, group=self.group, async_op=True)
                return
            send = self._send_buffer(part)
            if send.shape[0] > 0:
                nat.check(L.gnntf_halo_pack_f32(nat.ptr(src), F, nat.ptr(p.send_idx), send.shape[0],
                                                nat.ptr(send), F, F, nat.stream_ptr()), "halo_pack")
            if self._exchange is not None:
                self._exchange(self, send, src[self.n_local:])
                done = torch.cuda.Event()
                done.record()
                part["work"] = _EventWork(done)
            else:
                part["work"] = exchange_halo(p, send, src[self.n_local:], self.group, async_op=True)

    def _compute(self, part, src, dst, alpha):
        nat, L = self.nat, self.nat.lib()
        F, st = part["F"], self.nat.stream_ptr()
        # pass 1: every row over its owned columns, full epilogue (teleport term included)
        s1 = self.owned.struct(self.owned_val, F)
        nat.check(L.gnntf_appnp_step_f32(ctypes.byref(s1), nat.ptr(src), nat.ptr(part["H0"]), nat.ptr(dst), F, F,
                                         float(alpha), None, 1.0, nat.ACT_IDENTITY, st), "appnp_step")
        if part["work"] is not None:
            part["work"].wait()  # the compute stream waits for this part's halo rows
            part["work"] = None
        # pass 2: dst[boundary rows] += (1-a) * (entries over halo columns) . H_halo
        if self.halo_part.n > 0:
            s2 = self.halo_part.struct(self.halo_val, F)
            nat.check(L.gnntf_spmm_acc_f32(ctypes.byref(s2), nat.ptr(src), F, nat.ptr(dst), F, F, 1.0 - float(alpha), st),
                      "spmm_acc")

    def _step(self, src, dst, alpha):
        """One un-pipelined step of the first part (kept for the single-process emulation test)."""
        self._start_exchange(self.parts[0], src)
        self._compute(self.parts[0], src, dst, alpha)

    def propagate(self, H0_local, alpha=0.1, iterations=10):
        """K fused steps on this shard; returns this rank's rows of H_K ([n_local, F])."""
        cur = []
        for part in self.parts:
            part["H0"].copy_(H0_local[:, part["col0"]:part["col0"] + part["F"]])
            src, dst = part["buf"]
            src[:self.n_local].copy_(part["H0"])
            cur.append([src, dst])
        if iterations > 0:
            for part, (src, _) in zip(self.parts, cur):
                self._start_exchange(part, src)
        for k in range(iterations):
            for part, pair in zip(self.parts, cur):
                src, dst = pair
                self._compute(part, src, dst, alpha)
                pair[0], pair[1] = dst, src
                if k + 1 < iterations:      # this half's next exchange overlaps the other half's compute
                    self._start_exchange(part, pair[0])
        if len(self.parts) == 1:
            return cur[0][0][:self.n_local]
        return torch.cat([pair[0][:self.n_local] for pair in cur], dim=1)

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
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if multi:
            dist.barrier()                       # timing spans ALL ranks (every column group)
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        if multi:
            dist.barrier()
        sec = (time.perf_counter() - t0) / reps
        t = torch.tensor([sec], dtype=torch.float64, device=self.H0.device)
        nbytes = torch.tensor([float(host_in.numel() * 4)], dtype=torch.float64, device=self.H0.device)
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(nbytes, op=dist.ReduceOp.SUM)
        return {"seconds": float(t.item()), "h2d": int(nbytes.item()), "d2h": int(nbytes.item())}
