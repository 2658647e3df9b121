"""Row-sharded multi-GPU APPNP propagation (one process per GPU, ``torch.distributed``).

The reference is single-process (no collectives anywhere); this module is the
B200 scale-out of its propagation loop (filter.py:17-22 under layered.py:52-55), following the
north_star: a contiguous node-range row split (balanced by nnz, not by rows — power-law graphs),
and per propagation step an exchange of exactly the halo feature rows each rank references,
overlapped with the SpMM of the rows that need no halo.

  rank r owns rows [lo_r, hi_r);  its CSR addresses  H_ext = [ owned rows | halo rows ]
  step:  ONE launch: the leading CTAs send the rows the peers need straight into their halo buffers over
         NVLink peer memory and publish an epoch flag (st.release.sys); the rest of the grid runs the
         fused APPNP step of ALL rows over their OWNED columns (the bulk of the entries; needs no halo)
         flag wait (ld.acquire.sys, one warp)  ->  accumulate the HALO-column entries of the boundary rows
         (NCCL fallback: pack kernel + all-to-all on a second stream)
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


def global_columns(plan, own_col, halo_col):
    """Column ids of the two parts of :func:`split_by_column` in GLOBAL numbering (the all-gather buffer layout of
    the copy-engine exchange: every rank's buffer holds all rows, its own block at rows [lo, hi))."""
    own = (own_col.long() + plan.lo).to(torch.int32).contiguous()
    halo = plan.halo_cols[halo_col.long() - plan.n_local].to(torch.int32).contiguous()
    return own, halo


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


class _SharedBuffer:
    """Device allocation made through gnntf_ipc_alloc so that peers can map it (CUDA IPC) and store
    into it directly; exposed to torch without a copy.  Freed by :meth:`free` (never by ``__del__``: a
    peer may still have it mapped — ShardedPropagator.close() orders the teardown)."""

    def __init__(self, shape, dtype, device):
        from . import _native as nat
        self.nat, self.shape = nat, tuple(int(x) for x in shape)
        itemsize = torch.empty(0, dtype=dtype).element_size()
        count = 1
        for x in self.shape:
            count *= x
        ptr = ctypes.c_void_p()
        self.handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(device):
            nat.check(nat.lib().gnntf_ipc_alloc(max(16, count * itemsize), ctypes.byref(ptr), self.handle), "ipc_alloc")
        self.ptr = ptr.value
        typestr = {torch.float32: "<f4", torch.int32: "<i4"}[dtype]
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": typestr, "data": (self.ptr, False),
                                         "version": 2, "strides": None}
        self.tensor = torch.as_tensor(self, device=device)
        self.tensor.zero_()

    def handle_bytes(self):
        return bytes(self.handle)

    def free(self):
        if self.ptr:
            self.tensor = None
            self.nat.check(self.nat.lib().gnntf_ipc_free(ctypes.c_void_p(self.ptr)), "ipc_free")
            self.ptr = 0


class _EventWork:
    """``work.wait()`` for the single-process emulation hook: the compute stream waits for an event."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class ShardedPropagator:
    """APPNP K-step propagation (and plain SpMM) of one shard on one GPU (see module docstring).

    ``push`` (default): halo rows travel by the fused pack+send kernel over NVLink peer memory, and the
    ranks synchronise through epoch flags in peer memory written by that same kernel
    (``gnntf_halo_push_signal_f32``: st.release.sys after the rows; the consumer acquires them with
    ``gnntf_flags_wait`` in front of the halo-column pass) — no collective and no host round trip
    inside a propagation.  If the peers' buffers cannot be mapped every rank falls back to the NCCL
    all-to-all.  ``peers="local"`` wires several propagators of ONE process to each other instead of
    using CUDA IPC (single-GPU emulation of the ranks for tests: :func:`connect_local`,
    :func:`propagate_lockstep`)."""

    def __init__(self, adj, A, F, rank, world, group=None, plan=None, exchange=None, halves=None, push=True, peers="ipc",
                 copy="auto"):
        from . import _native as nat
        from .sparse import CsrStructure
        self.nat = nat
        self.group, self.F, self.peers_mode = group, int(F), peers
        self._exchange = exchange  # test hook: single-process emulation of the all-to-all
        self.comm_stream = torch.cuda.Stream(priority=-1) if (world > 1 and torch.cuda.is_available() and peers == "ipc") else None
        csr = A.csr
        self.plan = p = plan if plan is not None else build_shard_plan(csr.row_ptr, csr.col_idx, A.val, rank, world, group)
        self.lo, self.hi, self.n_local, self.n_halo = p.lo, p.hi, p.n_local, p.n_halo
        self.nnz_local = int(p.col_idx.numel())
        dev = p.row_ptr.device

        (o_rp, o_col, o_val), (h_rp, h_col, h_val), h_rows = split_by_column(p.row_ptr, p.col_idx, p.val, self.n_local)
        # copy-engine exchange: every rank's buffers hold ALL rows (own block at its global offset), columns stay global
        if copy == "auto":
            # measured on the products shape: the DMA all-gather moves every row (1.3-1.6x the rows the
            # peers reference) at ~410 GB/s when several peers send at once — hidden behind the owned-column pass
            # with one peer (N=2: 31.5 -> 29.3 ms), slower than the push kernel with three (N=8, 4x2: 14.3 vs 10.9 ms)
            copy = p.world == 2
        self.copy = bool(copy) and bool(push) and p.world > 1 and exchange is None
        self.n_total = int(p.bounds[-1])
        self.row0 = p.lo if self.copy else 0          # first owned row inside the feature buffers
        if self.copy:
            o_col, h_col = global_columns(p, o_col, h_col)
        self.owned, self.owned_val = CsrStructure(self.n_local, o_rp, o_col, None), o_val
        self.halo_part, self.halo_val = CsrStructure(int(h_rows.numel()), h_rp, h_col, None), h_val
        self.halo_part.row_map = h_rows
        self.interior, self.boundary = self.owned, self.halo_part  # (names kept for reports: pass 1 / pass 2)
        if halves is None:
            halves = 1  # column-half pipelining measured slower than the column split at N=2
        first = ((self.F + halves - 1) // halves + 3) // 4 * 4 if halves > 1 else self.F
        widths = [first, self.F - first] if halves > 1 and self.F - first > 0 else [self.F]
        n_ext, n_send = (self.n_total if self.copy else self.n_local + self.n_halo), int(sum(p.send_counts))
        if self.copy:
            halves = 1
            widths = [self.F]
        self.push = bool(push) and p.world > 1 and exchange is None
        self.parts, col0, self._shared = [], 0, []
        self._opened, self._flags, self._closed = [], None, False
        self.use_graph, self._graphs, self._graph_error = True, {}, None
        from .ops import pitch_for
        for w in widths:
            if self.push:
                ld = pitch_for(w)   # rows are gathered whole: a 128-byte-aligned pitch when that saves lines (ops.pitch_for)
                mats = [_SharedBuffer((n_ext, ld), torch.float32, dev) for _ in range(2)]
                self._shared.append(mats)
                bufs = [m.tensor for m in mats]
            else:
                ld = w
                bufs = [torch.zeros((n_ext, w), dtype=torch.float32, device=dev) for _ in range(2)]
            self.parts.append(dict(F=w, ld=ld, col0=col0, buf=bufs, send=None, n_send=n_send,   # send buffer: NCCL path only
                                   H0=torch.zeros((self.n_local, ld), dtype=torch.float32, device=dev), work=None))
            col0 += w
        if self.push:
            # flags[0, q]: epoch of the last push received from rank q; flags[1, q]: rank q's end-of-propagation ack
            self._flags = _SharedBuffer((2, p.world), torch.int32, dev)
            self._epoch = torch.zeros(1, dtype=torch.int32, device=dev)
            self._done = torch.zeros(1, dtype=torch.int32, device=dev)
            self._arange = torch.arange(1024, dtype=torch.int32, device=dev)
            self._epoch_vals = torch.zeros(1024, dtype=torch.int32, device=dev)   # epoch_base + d, refreshed per propagation
            self._copy_streams = ([torch.cuda.Stream() for _ in range(p.world - 1)]
                                  if (self.copy and peers == "ipc" and torch.cuda.is_available()) else None)
        if self.push and peers == "ipc":
            # every rank of the row group must take the same path: agree on whether IPC mapping worked
            try:
                self._map_peers()
                ok = 1.0
            except Exception as err:  # e.g. peers without CUDA IPC / peer access
                ok, self._map_error = 0.0, err
            flag = torch.tensor([ok], dtype=torch.float32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if flag.item() < 1.0:
                if self.copy:
                    raise RuntimeError(f"copy-engine exchange needs CUDA IPC peer mappings: {getattr(self, '_map_error', None)}")
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

    # -- peer tables ------------------------------------------------------------------------
    def _layout(self):
        return dict(recv_counts=list(self.plan.recv_counts), n_local=self.n_local)

    def _build_tables(self, everyone, buffer_ptr, flags_ptr):
        """Device tables the push kernel needs.  ``everyone[q]``: layout dict of rank q;
        ``buffer_ptr(q, part, buf)`` / ``flags_ptr(q)``: addresses of rank q's buffers as seen from here."""
        p = self.plan
        dev = p.row_ptr.device
        send_off = [0]
        for c in p.send_counts:
            send_off.append(send_off[-1] + int(c))
        self._send_off = torch.tensor(send_off, dtype=torch.int64, device=dev)
        nxt = send_off[(p.rank + 1) % p.world]
        self._rotate = int(nxt) if nxt < send_off[-1] else 0   # start with the next rank's rows
        # my rows land in peer q's halo region after the rows of lower-ranked owners
        row0 = [int(everyone[q]["n_local"]) + int(sum(everyone[q]["recv_counts"][:p.rank])) for q in range(p.world)]
        self._peer_row0 = torch.tensor(row0, dtype=torch.int64, device=dev)
        self._peer_ptrs, self._peer_buf_addr = [], []
        for pi in range(len(self.parts)):
            per_buf, per_buf_addr = [], []
            for bi in range(2):
                ptrs = [0 if (q == p.rank or (p.send_counts[q] == 0 and not self.copy)) else buffer_ptr(q, pi, bi)
                        for q in range(p.world)]
                per_buf.append(torch.tensor(ptrs, dtype=torch.int64, device=dev))
                per_buf_addr.append(ptrs)
            self._peer_ptrs.append(per_buf)
            self._peer_buf_addr.append(per_buf_addr)
        # every peer gets my completion flags (also peers I send no rows to: they wait on all slots)
        fl = [0 if q == p.rank else flags_ptr(q) for q in range(p.world)]
        self._peer_flag_addr = fl
        self._peer_data_flags = torch.tensor(fl, dtype=torch.int64, device=dev)
        self._peer_ack_flags = torch.tensor([0 if x == 0 else x + 4 * p.world for x in fl], dtype=torch.int64, device=dev)

    def _map_peers(self):
        """Exchange IPC handles and halo layouts inside the row group and map the peers' buffers."""
        nat, L, p = self.nat, self.nat.lib(), self.plan
        mine = dict(self._layout(), handles=[[m.handle_bytes() for m in mats] for mats in self._shared],
                    flags=self._flags.handle_bytes())
        everyone = [None] * p.world
        dist.all_gather_object(everyone, mine, group=self.group)

        def open_handle(raw):
            h = (ctypes.c_ubyte * 64).from_buffer_copy(raw)
            out = ctypes.c_void_p()
            nat.check(L.gnntf_ipc_open(h, ctypes.byref(out)), "ipc_open")
            self._opened.append(out.value)
            return out.value
        self._build_tables(everyone, lambda q, pi, bi: open_handle(everyone[q]["handles"][pi][bi]),
                           lambda q: open_handle(everyone[q]["flags"]))

    def close(self):
        """Tear the peer mappings down in a safe order: drain this GPU, barrier the group (nobody is still
        storing into anybody's buffers), unmap the peers' allocations, barrier again, free our own.  Tensors
        returned by :meth:`propagate` are views of these buffers and are invalid afterwards."""
        if self._closed:
            return
        self._closed = True
        self._graphs = {}
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        multi = self.group is not None or (dist.is_available() and dist.is_initialized() and self.plan.world > 1
                                           and self.comm_stream is not None)
        if multi:
            dist.barrier(group=self.group)
        for ptr in self._opened:
            self.nat.check(self.nat.lib().gnntf_ipc_close(ctypes.c_void_p(ptr)), "ipc_close")
        self._opened = []
        if multi:
            dist.barrier(group=self.group)
        for part in self.parts:
            part["buf"] = None
        self.buf = None
        for mats in self._shared:
            for m in mats:
                m.free()
        self._shared = []
        if self._flags is not None:
            self._flags.free()
            self._flags = None

    def launches_per_propagation(self, K):
        per_step = 1 if self.plan.world > 1 else 0  # flag-wait kernel (the push rides in the pass-1 launch) / pack kernel
        for st in (self.owned, self.halo_part):
            if st.n > 0 and st.nnz > 0 or st is self.owned:
                per_step += 2 if st.n_long > 0 else 1
        return K * per_step * len(self.parts)

    # -- phases of one step -------------------------------------------------------------------
    def _push(self, part, src, delta):
        """Fused pack + send + completion signal (epoch = base + delta) on the current stream."""
        nat, L, p = self.nat, self.nat.lib(), self.plan
        pi = self.parts.index(part)
        bi = 0 if src.data_ptr() == part["buf"][0].data_ptr() else 1
        F, ld = part["F"], part["ld"]
        nat.check(L.gnntf_halo_push_signal_f32(nat.ptr(src), ld, nat.ptr(p.send_idx), nat.ptr(self._send_off),
                                               nat.ptr(self._peer_ptrs[pi][bi]), nat.ptr(self._peer_row0), p.world,
                                               int(p.send_idx.numel()), self._rotate, ld, F, nat.ptr(self._done),
                                               nat.ptr(self._peer_data_flags), p.rank, nat.ptr(self._epoch), int(delta),
                                               nat.stream_ptr()), "halo_push_signal")

    def _step_push(self, part, src, dst, alpha, delta):
        """Pass 1 (every row over its owned columns) with the halo push of ``src`` and its completion signal
        riding in the SAME launch (gnntf_step_push_f32): the leading CTAs of the grid send, the rest compute."""
        nat, L, p = self.nat, self.nat.lib(), self.plan
        pi = self.parts.index(part)
        bi = 0 if src.data_ptr() == part["buf"][0].data_ptr() else 1
        F, ld = part["F"], part["ld"]
        s1 = self.owned.struct(self.owned_val, F)
        nat.check(L.gnntf_step_push_f32(ctypes.byref(s1), nat.ptr(src), nat.ptr(part["H0"]) if alpha is not None else None,
                                        nat.ptr(dst[self.row0:]), ld, F, float(alpha if alpha is not None else 0.0), nat.ptr(p.send_idx),
                                        nat.ptr(self._send_off), nat.ptr(self._peer_ptrs[pi][bi]), nat.ptr(self._peer_row0),
                                        p.world, int(p.send_idx.numel()), self._rotate, nat.ptr(self._done),
                                        nat.ptr(self._peer_data_flags), p.rank, nat.ptr(self._epoch), int(delta),
                                        nat.stream_ptr()), "step_push")

    def _send_block(self, part, src, delta, side_streams=True):
        """Copy-engine exchange of the step that reads ``src``: this rank's block of rows goes to every peer's
        buffer as one DMA copy per peer (each on its own stream, ordered after everything enqueued so far on the
        current stream), followed by a 4-byte DMA of the epoch  base + delta  into the peer's flag slot."""
        nat, L, p = self.nat, self.nat.lib(), self.plan
        pi = self.parts.index(part)
        bi = 0 if src.data_ptr() == part["buf"][0].data_ptr() else 1
        off = self.lo * part["ld"] * 4
        nbytes = self.n_local * part["ld"] * 4
        epoch_value = self._epoch_vals.data_ptr() + 4 * int(delta)
        ready = None
        if side_streams and self._copy_streams is not None:
            ready = torch.cuda.Event()
            ready.record()
        for j in range(1, p.world):
            q = (p.rank + j) % p.world                     # every rank starts with a different destination
            dst = self._peer_buf_addr[pi][bi][q] + off
            flag = self._peer_flag_addr[q] + 4 * p.rank
            if ready is not None:
                stream = self._copy_streams[j - 1]
                stream.wait_event(ready)
                handle = ctypes.c_void_p(stream.cuda_stream)
            else:
                handle = nat.stream_ptr()
            nat.check(L.gnntf_peer_copy_signal(ctypes.c_void_p(dst), ctypes.c_void_p(src.data_ptr() + off), nbytes,
                                               ctypes.c_void_p(flag), ctypes.c_void_p(epoch_value), handle), "peer_copy_signal")

    def _join_copies(self):
        if self._copy_streams is not None:
            cur = torch.cuda.current_stream()
            for stream in self._copy_streams:
                ev = torch.cuda.Event()
                ev.record(stream)
                cur.wait_event(ev)

    def _propagate_copy(self, H0_local, alpha, iterations, spmm_only=False):
        """Copy-engine mode, per step:  [DMA of this rank's block of H_k to every peer ∥ owned-column pass] ->
        flag wait -> halo-column pass.  The buffers hold all rows; this rank's block sits at rows [lo, hi)."""
        part = self.parts[0]
        F = part["F"]
        src, dst = part["buf"]
        own = slice(self.lo, self.hi)
        if spmm_only:
            src[own, :F].copy_(H0_local)
        else:
            part["H0"][:, :F].copy_(H0_local)
            src[own].copy_(part["H0"])
        if iterations >= self._epoch_vals.numel():
            raise ValueError("at most 1023 iterations per propagation in copy-engine mode")
        torch.add(self._arange, self._epoch, out=self._epoch_vals)
        if iterations > 0:
            self._wait(1, 0)      # every peer is done reading its buffers of the previous propagation
        for k in range(iterations):
            self._send_block(part, src, k + 1)
            self._pass1(part, src, dst, alpha)
            self._wait(0, k + 1)
            self._pass2(part, src, dst, alpha)
            src, dst = dst, src
        self._join_copies()
        self._finish(iterations)
        return src[own, :F]

    def _wait(self, row, delta):
        """Current stream waits until flags[row, q] >= base + delta for every peer q."""
        nat, p = self.nat, self.plan
        flags = self._flags.tensor[row]
        nat.check(nat.lib().gnntf_flags_wait(nat.ptr(flags), p.world, p.rank, nat.ptr(self._epoch), int(delta),
                                             nat.stream_ptr()), "flags_wait")

    def _ack(self, delta):
        nat, p = self.nat, self.plan
        nat.check(nat.lib().gnntf_flags_signal(nat.ptr(self._peer_ack_flags), p.world, p.rank, nat.ptr(self._epoch), int(delta),
                                               nat.stream_ptr()), "flags_signal")

    def _start_exchange(self, part, src, delta=0, first=False):
        """This part's halo exchange for the step that reads ``src``, on the comm stream, ordered after
        everything already enqueued on the compute stream (the step that produced ``src``)."""
        nat, L, p = self.nat, self.nat.lib(), self.plan
        part["work"] = None
        if p.world == 1:
            return
        F = part["F"]
        ready = torch.cuda.Event()
        ready.record()                                   # src complete on the compute stream
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ready)
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

    def _pass1(self, part, src, dst, alpha):
        """Every row over its owned columns; alpha=None: plain SpMM, else the full fused PPR epilogue."""
        nat, L = self.nat, self.nat.lib()
        F, ld, st = part["F"], part["ld"], self.nat.stream_ptr()
        s1 = self.owned.struct(self.owned_val, F)
        if alpha is None:
            nat.check(L.gnntf_spmm_f32(ctypes.byref(s1), nat.ptr(src), ld, nat.ptr(dst[self.row0:]), ld, F, st), "spmm")
        else:
            nat.check(L.gnntf_appnp_step_f32(ctypes.byref(s1), nat.ptr(src), nat.ptr(part["H0"]), nat.ptr(dst[self.row0:]), ld, F,
                                             float(alpha), None, 1.0, nat.ACT_IDENTITY, st), "appnp_step")

    def _pass2(self, part, src, dst, alpha):
        """dst[boundary rows] += (1-a) * (entries over halo columns) · H_halo."""
        nat, L = self.nat, self.nat.lib()
        if self.halo_part.n > 0:
            F, ld = part["F"], part["ld"]
            s2 = self.halo_part.struct(self.halo_val, F)
            scale = 1.0 if alpha is None else 1.0 - float(alpha)
            nat.check(L.gnntf_spmm_acc_f32(ctypes.byref(s2), nat.ptr(src), ld, nat.ptr(dst[self.row0:]), ld, F, scale, nat.stream_ptr()),
                      "spmm_acc")

    def _wait_exchange(self, part):
        work = part["work"]
        part["work"] = None
        if work is not None:
            work.wait()              # NCCL work / emulation event

    def _compute(self, part, src, dst, alpha):
        self._pass1(part, src, dst, alpha)
        self._wait_exchange(part)
        self._pass2(part, src, dst, alpha)

    def _step(self, src, dst, alpha):
        """One un-pipelined step of the first part (kept for the single-process emulation test)."""
        self._start_exchange(self.parts[0], src)
        self._compute(self.parts[0], src, dst, alpha)

    def _finish(self, K):
        """End of a propagation in push mode: join the comm stream, tell the peers this rank is done with its
        halo buffers, advance the epoch base (a device scalar: the sequence replays under a CUDA graph)."""
        if not self.push:
            return
        self._ack(K)
        self._epoch.add_(K)

    def propagate(self, H0_local, alpha=0.1, iterations=10):
        """K fused steps on this shard; returns this rank's rows of H_K ([n_local, F]) — a VIEW of an internal
        ping-pong buffer, valid until the next call (clone it to keep it).

        In push mode the whole sequence (copies, pushes, flag waits, both passes of every step, on the two
        streams) is captured into a CUDA graph on the second call with the same (alpha, K) and replayed
        afterwards: no Python and no launch latency between the ~5 launches of a step.
        The epochs the kernels signal and wait for live in a device scalar, so a replay uses fresh ones."""
        key = (float(alpha), int(iterations))
        if self.push and self.use_graph and self.peers_mode == "ipc" and iterations > 0:
            entry = self._graphs.get(key)
            if entry is None:                       # first call: eager (allocates workspaces), remember we saw it
                self._graphs[key] = "warm"
            elif entry == "warm":                   # second call: capture
                try:
                    self._graphs[key] = self._capture(key)
                except Exception as err:            # keep working without the graph
                    self._graphs[key] = "off"
                    self._graph_error = err
                    torch.cuda.synchronize()
            entry = self._graphs[key]
            if isinstance(entry, tuple):
                graph, static_in, out = entry
                static_in.copy_(H0_local)
                graph.replay()
                return out
        return self._propagate_eager(H0_local, alpha, iterations)

    def _capture(self, key):
        alpha, iterations = key
        static_in = torch.empty((self.n_local, self.F), dtype=torch.float32, device=self.H0.device)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self._propagate_eager(static_in, alpha, iterations)
        return graph, static_in, out

    def _propagate_eager(self, H0_local, alpha, iterations):
        if self.copy:
            return self._propagate_copy(H0_local, alpha, iterations)
        if self.push:
            return self._propagate_push(H0_local, alpha, iterations)
        cur = []
        for part in self.parts:
            part["H0"][:, :part["F"]].copy_(H0_local[:, part["col0"]:part["col0"] + part["F"]])
            src, dst = part["buf"]
            src[:self.n_local].copy_(part["H0"])
            cur.append([src, dst])
        if iterations > 0:
            for part, (src, _) in zip(self.parts, cur):
                self._start_exchange(part, src, delta=1, first=True)
        for k in range(iterations):
            for part, pair in zip(self.parts, cur):
                src, dst = pair
                self._compute(part, src, dst, alpha)
                pair[0], pair[1] = dst, src
                if k + 1 < iterations:      # this half's next exchange overlaps the other half's compute
                    self._start_exchange(part, pair[0], delta=k + 2)
        self._finish(iterations)
        if len(self.parts) == 1:
            return cur[0][0][:self.n_local, :self.parts[0]["F"]]
        return torch.cat([pair[0][:self.n_local, :part["F"]] for part, pair in zip(self.parts, cur)], dim=1)

    def _propagate_push(self, H0_local, alpha, iterations, spmm_only=False):
        """Push mode, ONE stream, per step:  [push of H_k ∥ owned-column pass] -> flag wait -> halo-column pass.
        (alpha=None with spmm_only: a single sharded SpMM.)"""
        cur = []
        for part in self.parts:
            if spmm_only:
                src, dst = part["buf"]
                src[:self.n_local, :part["F"]].copy_(H0_local[:, part["col0"]:part["col0"] + part["F"]])
            else:
                part["H0"][:, :part["F"]].copy_(H0_local[:, part["col0"]:part["col0"] + part["F"]])
                src, dst = part["buf"]
                src[:self.n_local].copy_(part["H0"])
            cur.append([src, dst])
        if iterations > 0:
            self._wait(1, 0)      # every peer is done reading its halo buffers of the previous propagation
        for k in range(iterations):
            for part, pair in zip(self.parts, cur):
                src, dst = pair
                self._step_push(part, src, dst, alpha, k + 1)
            for part, pair in zip(self.parts, cur):
                src, dst = pair
                self._wait(0, k + 1)
                self._pass2(part, src, dst, alpha)
                pair[0], pair[1] = dst, src
        self._finish(iterations)
        if len(self.parts) == 1:
            return cur[0][0][:self.n_local, :self.parts[0]["F"]]
        return torch.cat([pair[0][:self.n_local, :part["F"]] for part, pair in zip(self.parts, cur)], dim=1)

    def spmm(self, H_local):
        """One sharded SpMM ``(Â·H)[lo:hi]`` (BASELINE config 5, the R-MAT sweep): a single halo exchange
        overlapped with the owned-column pass."""
        part = self.parts[0]
        assert len(self.parts) == 1
        if self.copy:
            return self._propagate_copy(H_local, None, 1, spmm_only=True)
        if self.push:
            return self._propagate_push(H_local, None, 1, spmm_only=True)
        src, dst = part["buf"]
        src[:self.n_local, :part["F"]].copy_(H_local)
        self._start_exchange(part, src, delta=1, first=True)
        self._compute(part, src, dst, None)
        return dst[:self.n_local, :part["F"]]

    def propagate_host_batched(self, host_ins, host_outs, alpha, iterations, work=None):
        """End-to-end for a sequence of HOST inputs (this rank's rows, pinned): the upload of input b+1 and the
        read-back of result b-1 overlap the K steps of input b (three streams, two device slots; the result is
        staged out of the ping-pong buffer so that the next propagation may overwrite it).  ``host_outs[b]``
        receives the result of ``host_ins[b]``; entries may repeat.  Synchronises before returning."""
        dev = self.H0.device
        if work is None:
            work = [torch.empty((self.n_local, self.F), dtype=torch.float32, device=dev) for _ in range(4)]
        dev_in, stage = work[:2], work[2:4]
        if getattr(self, "_io_streams", None) is None:
            self._io_streams = (torch.cuda.Stream(), torch.cuda.Stream())
        up, down = self._io_streams
        cur = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(cur)
        up.wait_event(start)
        down.wait_event(start)
        ev_comp, ev_out = [None, None], [None, None]
        for b, (h_in, h_out) in enumerate(zip(host_ins, host_outs)):
            s = b & 1
            with torch.cuda.stream(up):
                if ev_comp[s] is not None:
                    up.wait_event(ev_comp[s])              # propagation b-2 has consumed this slot's input
                dev_in[s].copy_(h_in, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(up)
            cur.wait_event(ev_in)
            if ev_out[s] is not None:
                cur.wait_event(ev_out[s])                  # read-back b-2 has drained this slot's staging buffer
            stage[s].copy_(self.propagate(dev_in[s], alpha, iterations))
            ev_comp[s] = torch.cuda.Event()
            ev_comp[s].record(cur)
            with torch.cuda.stream(down):
                down.wait_event(ev_comp[s])
                h_out.copy_(stage[s], non_blocking=True)
                ev_out[s] = torch.cuda.Event()
                ev_out[s].record(down)
        for ev in ev_out:
            if ev is not None:
                cur.wait_event(ev)
        torch.cuda.synchronize()

    def propagate_host_timed(self, H0_host, alpha, iterations, reps=3):
        """End-to-end timing: this shard's rows of H0 in pinned host memory -> device, K steps, result shard ->
        host, for ``reps`` consecutive steps pipelined by :meth:`propagate_host_batched` (every step moves its own
        input and its own result across PCIe inside the timed region).  Also times ONE un-pipelined call.
        Returns the timings and the host result of the last step (for the caller's parity check)."""
        import time
        dev = self.H0.device
        host_in = H0_host if H0_host.is_pinned() else H0_host.pin_memory()
        outs = [torch.empty_like(host_in).pin_memory() for _ in range(2)]
        work = [torch.empty((self.n_local, self.F), dtype=torch.float32, device=dev) for _ in range(4)]
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

        def timed(n):
            if multi:
                dist.barrier()                       # timing spans ALL ranks (every column group)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            self.propagate_host_batched([host_in] * n, [outs[b & 1] for b in range(n)], alpha, iterations, work=work)
            if multi:
                dist.barrier()
            sec = (time.perf_counter() - t0) / n
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            if multi:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        timed(2)                                     # warm-up (CUDA-graph capture of the propagation included)
        single = timed(1)
        sec = timed(max(1, reps))
        nbytes = torch.tensor([float(host_in.numel() * 4)], dtype=torch.float64, device=dev)
        if multi:
            dist.all_reduce(nbytes, op=dist.ReduceOp.SUM)
        return {"seconds": sec, "single_call_seconds": single, "h2d": int(nbytes.item()), "d2h": int(nbytes.item()),
                "host_out": outs[(max(1, reps) - 1) & 1]}


# ----------------------------------------------------------------------------------------------
# Single-process emulation of the ranks on ONE GPU (tests): same kernels, same tables, same flags
# ----------------------------------------------------------------------------------------------
def connect_local(props):
    """Wire ``props[r]`` (ShardedPropagator(..., peers="local") of rank r, all in this process and on one
    device) to each other: the peer tables point straight at the other objects' buffers."""
    everyone = [pr._layout() for pr in props]
    for pr in props:
        pr._build_tables(everyone, lambda q, pi, bi: props[q]._shared[pi][bi].ptr, lambda q: props[q]._flags.ptr)


def propagate_lockstep(props, H0_locals, alpha=0.1, iterations=10, spmm_only=False):
    """The sequence :meth:`ShardedPropagator.propagate` enqueues per rank, interleaved over all emulated
    ranks on ONE stream so that every flag is written before the kernel that waits for it starts
    (kernels that wait on one another must never share a GPU concurrently).  Returns the per-rank results."""
    cur = []
    for pr, H0 in zip(props, H0_locals):
        part = pr.parts[0]
        if not spmm_only:
            part["H0"][:, :part["F"]].copy_(H0)
        src, dst = part["buf"]
        src[pr.row0:pr.row0 + pr.n_local, :part["F"]].copy_(H0)
        if pr.copy:
            torch.add(pr._arange, pr._epoch, out=pr._epoch_vals)
        cur.append([src, dst])
    K = 1 if spmm_only else iterations
    a = None if spmm_only else alpha
    for pr in props:
        pr._wait(1, 0)
    for k in range(K):
        for pr, (src, dst) in zip(props, cur):
            if pr.copy:                                          # DMA of the block + flag, then the owned-column pass
                pr._send_block(pr.parts[0], src, k + 1, side_streams=False)
                pr._pass1(pr.parts[0], src, dst, a)
            else:
                pr._step_push(pr.parts[0], src, dst, a, k + 1)  # push of H_k + owned-column pass, one launch
        for pr, pair in zip(props, cur):
            src, dst = pair
            pr._wait(0, k + 1)
            pr._pass2(pr.parts[0], src, dst, a)
            pair[0], pair[1] = dst, src
    for pr in props:
        pr._ack(K)
        pr._epoch.add_(K)
    return [pair[0][pr.row0:pr.row0 + pr.n_local, :pr.parts[0]["F"]] for pr, pair in zip(props, cur)]
