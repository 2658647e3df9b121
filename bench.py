#!/usr/bin/env python
"""Benchmark of the hot path: APPNP K=10 propagation (BASELINE.json `metric`).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

One "step" = one full K=10 propagation  H <- (1-a)·Â·H + a·H0  (ten fused launches) over the
named synthetic workload with Â built and normalised beforehand and H0 resident in HBM.
`value` = edge·features processed per second = nnz·F·K_iter·steps / time (whole job, all GPUs);
`ms_per_step` = the metric's other half, the absolute K=10 propagation time.  Prints ONE JSON line
on stdout (rank 0); everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "gnn-tf_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import synthetic  # noqa: E402

METRIC = "appnp_k10_propagation_edge_features_per_s"
UNIT = "edge_features/s"
K_ITER = 10
ALPHA = 0.1


# stdout must carry exactly ONE JSON line: libraries write there too (NCCL prints its version banner
# on stdout), so fd 1 is pointed at stderr for the whole run and the JSON goes to the saved fd.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit_json(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(args, F):
    """ncu-measured DRAM bytes per launch of the dominant kernel for this workload, if one was captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path) or args.scale != 1.0:
        return None
    with open(path) as f:
        entry = json.load(f).get(f"{args.workload}/{args.ordering}/{F}")
    return entry["dram_bytes_per_launch"] if entry else None


def step_bytes(n, nnz, F):
    """Algorithmic bytes of one fused APPNP step: int32 col + fp32 val per entry,
    int32 row_ptr, read H_k, read H0, write H_{k+1}; every array touched once."""
    return 8 * nnz + 4 * (n + 1) + 12 * n * F


# ----------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc, self.path, self.gpu = None, None, gpu_index

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception as e:  # nvidia-smi missing
            log("clock sampler unavailable:", e)
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------
def make_workload(args, device):
    n, e, f_default, classes = synthetic.SHAPES[args.workload]
    F = args.features or (classes if args.workload == "cora" else f_default)
    t0 = time.time()
    if args.workload in synthetic.POWERLAW:
        n, edges = synthetic.shaped_edges(args.workload, seed=0, ordering=args.ordering, device=device, scale=args.scale)
    else:
        G = synthetic.citation_graph(n, e, seed=0)
        import gnntf
        edges = torch.as_tensor(np.asarray(gnntf.graph2indices(G), dtype=np.int64)).to(device)
    log(f"[workload] {args.workload}: n={n} E={edges.shape[0]} F={F} ordering={args.ordering} generated in {time.time() - t0:.1f}s")
    return n, edges, F


def config_dict(args, n, E, nnz, F):
    return {"workload": f"APPNP K={K_ITER} a={ALPHA} on {args.workload}-shaped synthetic graph", "nodes": n, "edges": E,
            "nnz": nnz, "features": F, "iterations": K_ITER, "alpha": ALPHA, "ordering": args.ordering,
            "scale": args.scale, "reorder": bool(getattr(args, "reorder", False)), "normalisation": "symmetric, eval mode (prebuilt)",
            "l2": "no explicit flush: per-step working set (CSR + 3 feature matrices) exceeds the 126 MB L2"
                  if step_bytes(n, nnz, F) > 3 * 126e6 else "working set fits L2: flushed by a 256 MB write between steps"}


# ----------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference's TF-CPU path)
# ----------------------------------------------------------------------------------------------
def big_oracle(n, edges_cpu):
    """The C oracle's view of the workload (oracle/oracle_big.py): graph2adj, get_adjacency, stable CSR."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_big
    t0 = time.time()
    big = oracle_big.BigOracle(edges_cpu, None, n, keep_idx=True)
    log(f"[oracle] graph2adj + get_adjacency + stable CSR on the host: {time.time() - t0:.1f}s")
    return big


def cpu_arm(big, F, seconds_per_step, steps, warmup, multi_thread_context=True):
    """Times the oracle's restatement of filter.py:19-21 (COO-order SpMM loop as in TF-CPU's
    SparseTensorDenseMatMul functor + teleport) on a bounded prefix of the COO list."""
    import oracle_big
    lib = oracle_big.lib()
    n, nnz, idx, norm = big.n, big.nnz, big.idx, big.norm_coo
    H = np.random.default_rng(1).standard_normal((n, F)).astype(np.float32)
    P = np.zeros((n, F), np.float32)
    out = np.empty((n, F), np.float32)
    # calibrate on 2M entries, then size the sample for ~seconds_per_step
    cal = min(nnz, 2_000_000)
    t0 = time.perf_counter()
    lib.oracle_spmm_coo_f32(idx.ctypes.data, norm.ctypes.data, 0, cal, H.ctypes.data, F, P.ctypes.data)
    rate = cal * F / max(time.perf_counter() - t0, 1e-9)
    sample = int(min(nnz, max(cal, rate * seconds_per_step / F)))
    times = []
    for it in range(warmup + steps):
        P[:] = 0
        t0 = time.perf_counter()
        lib.oracle_spmm_coo_f32(idx.ctypes.data, norm.ctypes.data, 0, sample, H.ctypes.data, F, P.ctypes.data)
        frac = sample / nnz
        rows = max(1, int(n * frac))  # the teleport pass scaled to the same fraction of the step
        lib.oracle_teleport_f32(P.ctypes.data, H.ctypes.data, rows * F, ctypes.c_float(ALPHA), out.ctypes.data)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    value = sample * F / t
    omp = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    # context: the same step as a row-parallel CSR SpMM on every host thread (not what TF-CPU does)
    mt = None
    if multi_thread_context:
        rows_cap = int(n * min(1.0, max(0.02, seconds_per_step * value * 4 / (nnz * F))))  # bounded: a few seconds
        args = (big.row_ptr.ctypes.data, big.col.ctypes.data, big.val.ctypes.data, H.ctypes.data, H.ctypes.data, F,
                ctypes.c_float(ALPHA), 0, min(rows_cap, n), out.ctypes.data)
        lib.oracle_appnp_step_csr_omp_f32(*args)
        t0 = time.perf_counter()
        lib.oracle_appnp_step_csr_omp_f32(*args)
        dt = time.perf_counter() - t0
        done = int(big.row_ptr[min(rows_cap, n)])
        mt = {"mt_value": done * F / dt, "mt_cores": omp, "mt_sample": f"rows 0..{min(rows_cap, n)} ({done} entries), 1 timed rep",
              "mt_kind": "row-parallel CSR SpMM + fused teleport on all OpenMP threads (context only: TF-CPU's kernel for this op is a single-threaded COO loop)"}
    desc = (f"{sample} of {nnz} COO entries (prefix, storage order) x F={F}: one PPR iteration's SpMM + teleport, "
            f"{len(times)} timed reps; SpMM loop single-threaded as in TF-CPU, element-wise pass on {omp} OpenMP threads")
    return value, t, sample, desc, omp, mt


def parity_block(big, H0_host, got, split_rows, F):
    """Check the propagation the timed region produced (ALL rows) against the C oracle.
    un-split rows: the reference's order, fp32 (oracle_appnp_propagate_csr_omp_f32, bit-identical to the
    sequential COO loop); rows the GPU sums in pieces: the double-accumulating twin (see oracle_big.check_against).
    max_rel_err = max |x−y| / max(|y|, floor·‖y‖∞) and must stay <= 1e-5."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gnntf_oracle as oracle
    t0 = time.time()
    exp32 = big.propagate(H0_host, ALPHA, K_ITER)
    norm = float(np.abs(exp32).max())
    keep = ~split_rows
    st = oracle.parity_stats(got[keep], exp32[keep], oracle.RTOL, oracle.FLOOR_REORDERED, norm)
    out = {"checked_rows": int(got.shape[0]), "columns": int(F), "oracle": "oracle_c.c (C restatement of filter.py:17-22 over the "
           "stable-by-row CSR, bit-identical to the sequential COO loop), K=10 from the same H0",
           "rtol": oracle.RTOL, "floor": oracle.FLOOR_REORDERED, "norm_inf": norm,
           "unsplit_rows": int(keep.sum()), "unsplit_max_rel_err": st["max_err_over_bound"] * oracle.RTOL,
           "unsplit_max_abs_err_over_norm": st["max_abs_err_over_norm"],
           "unsplit_bit_identical_fraction": float(np.mean(got[keep] == exp32[keep]))}
    worst = st["max_err_over_bound"]
    if split_rows.any():
        exp64 = big.propagate(H0_host, ALPHA, K_ITER, acc64=True)
        s2 = oracle.parity_stats(got[split_rows], exp64[split_rows], oracle.RTOL, oracle.FLOOR_REORDERED, norm)
        s3 = oracle.parity_stats(got[split_rows], exp32[split_rows], oracle.RTOL, 1.0, norm)
        out.update({"split_rows": int(split_rows.sum()), "split_max_rel_err_vs_double_accumulating_twin": s2["max_err_over_bound"] * oracle.RTOL,
                    "split_max_abs_err_over_norm_vs_fp32_sequential": s3["max_abs_err_over_norm"]})
        worst = max(worst, s2["max_err_over_bound"])
    out["max_rel_err"] = worst * oracle.RTOL
    out["ok"] = bool(worst <= 1.0)
    out["wall_s"] = time.time() - t0
    return out


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def dist_setup(n_gpus):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")  # the halo exchange must not queue behind the SpMM grid
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, local, world


def gpu_arm(args):
    import gnntf
    from gnntf import ops
    rank, local, world = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    hbm_peak, peak_src = peaks()
    n, edges, F = make_workload(args, dev)
    E = edges.shape[0]

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    adj = gnntf.edges2adj(edges, None, n)
    if args.reorder:
        adj = adj.reordered()
    A = adj.normalized("symmetric")
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    nnz = adj.csr.nnz
    log(f"[build] graph2adj + CSR + normalise: {build_s * 1e3:.1f} ms (nnz={nnz}, long rows={adj.csr.n_long}, pieces={adj.csr.n_chunks})")
    H0 = synthetic.features(n, F, seed=1, device=dev)
    if F % 4:  # class-width matrices (7, 47): run 16-byte-aligned rows, as gnntf.appnp_propagate does
        H0, _ = ops._pad4(H0)
    F_run = H0.shape[1]

    if world > 1:
        from gnntf import dist as gdist
        R, C = (int(x) for x in args.grid.split("x")) if args.grid else gdist.choose_grid(world, F_run)
        grid = gdist.Grid2D(rank, world, R, C)
        c0, c1 = gdist.column_range(F_run, C, grid.c)
        prop = gdist.ShardedPropagator(adj, A, c1 - c0, grid.r, R, group=grid.row_group, halves=args.halves or None,
                                       push=not args.nccl_exchange, copy=({"auto": "auto", "copy": True, "push": False}[args.exchange] if not args.nccl_exchange else False))
        prop.use_graph = not args.no_graph
        log(f"[rank {rank} = row group {grid.r}/{R}, column group {grid.c}/{C}] rows {prop.lo}:{prop.hi} cols {c0}:{c1} "
            f"nnz {prop.nnz_local} halo rows {prop.n_halo} owned-column entries {prop.owned.nnz} "
            f"halo-column entries {prop.halo_part.nnz} (rows {prop.halo_part.n})")
        H0_local = (H0 if adj.perm is None else H0.index_select(0, adj.perm))[prop.lo:prop.hi, c0:c1].contiguous()
        run = lambda: prop.propagate(H0_local, ALPHA, K_ITER)  # noqa: E731
        launches_per_step = prop.launches_per_propagation(K_ITER)
        how = ("copy-engine DMA of each rank's row block into every peer's buffer over NVLink (all-gather layout), epoch flags "
               "written by DMA and acquired on the device" if prop.copy else
               "fused pack+send+signal kernel over NVLink peer memory, epoch flags acquired on the device" if prop.push
               else "pack kernel + NCCL all-to-all")
        sharding = (f"{R} row groups (contiguous node ranges balanced by nnz; halo rows by {how}, inside a column "
                    f"group) x {C} feature-column groups (no communication)")
        H0_cpu = H0[:, :F].cpu() if (rank == 0 and not args.no_parity) else None   # for the parity check after the timed region
        del H0
    else:
        out = torch.empty_like(H0)
        scratch = torch.empty_like(H0)
        if args.reorder:  # features live in external node order: the permutation in/out is part of the step
            def run():
                with torch.no_grad():
                    return gnntf.appnp_propagate(A, H0, ALPHA, K_ITER)
        else:
            run = lambda: ops.propagate_raw(A, H0, ALPHA, K_ITER, out=out, scratch=scratch)  # noqa: E731
        launches_per_step = K_ITER * (2 if adj.csr.n_long > 0 else 1)  # row kernel (pieces ride in its grid) + long-row reduce
    flush = None
    if step_bytes(n, nnz, F) <= 3 * 126e6:
        flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        run()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e_ in evs:
        if flush is not None:
            flush.fill_(1.0)
        s.record()
        run()
        e_.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    per_step_ms = torch.tensor([s.elapsed_time(e_) for s, e_ in evs], dtype=torch.float64, device=dev)
    total_ms = per_step_ms.sum().reshape(1)
    if world > 1:
        torch.distributed.all_reduce(total_ms, op=torch.distributed.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    ms_per_step = total_ms / args.steps
    value = nnz * F * K_ITER * args.steps / (total_ms * 1e-3)

    result = None
    if rank == 0:
        bstep = step_bytes(n, nnz, F)
        achieved = bstep * K_ITER / (ms_per_step * 1e-3) / 1e9
        result = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args, n, E, nnz, F),   # same keys and values as --impl reference
            "layout": {"padded_features": F_run, "sharding": (sharding if world > 1 else "none")},
            "propagation_ms": {"mean": ms_per_step, "min": float(per_step_ms.min().item()),
                               "median": float(per_step_ms.median().item())},
            "clocks": clocks, "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": measured_traffic(args, F) if world == 1 else None, "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                         "kernel": "fused APPNP step (spmm_rows4_kernel, long-row pieces in its grid, + spmm_long_reduce_kernel)",
                         "algorithmic_bytes_per_launch": bstep,
                         "avg_launch_ms": ms_per_step / K_ITER,
                         "note": "achieved = B_step / (timed region / (steps*K)); B_step = 8*nnz + 4*(N+1) + 12*N*F"},
            "csr_build_ms": build_s * 1e3,
        }
        if world > 1:
            result["layout"]["k_step_launch"] = ("one CUDA graph replay per propagation (both streams captured)"
                                                 if any(isinstance(v, tuple) for v in prop._graphs.values()) else "launched step by step from the host")
            result["roofline"]["note"] += f"; aggregate over {world} GPUs, peak is per-GPU x {world}"
            result["roofline"]["peak"] = hbm_peak * world
            result["roofline"]["frac"] = achieved / (hbm_peak * world)

    # ---- e2e: host buffers in, host buffers out, through the public API ------------------------
    if world == 1:
        H0_host = H0.cpu().pin_memory()
        out_host = torch.empty_like(H0_host).pin_memory()
        bufs = [H0, out, scratch]
        reps = max(1, min(args.steps, 5))
        ops.appnp_propagate_host(A, H0_host, ALPHA, K_ITER, out_host=out_host, bufs=bufs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            ops.appnp_propagate_host(A, H0_host, ALPHA, K_ITER, out_host=out_host, bufs=bufs)
        torch.cuda.synchronize()
        single_s = (time.perf_counter() - t0) / reps
        single_out = out_host.clone()
        # K consecutive steps as ONE batched call: the upload of step i+1 and the read-back of step i-1 overlap the
        # propagation of step i (3 streams, 2 device slots).  Every step still moves its own input and its own
        # result across PCIe inside the timed region.
        nb = max(1, args.steps)
        outs2 = [out_host, torch.empty_like(H0_host).pin_memory()]
        work = torch.empty((5, n, F_run), dtype=torch.float32, device=dev)
        seq_in, seq_out = [H0_host] * nb, [outs2[b % 2] for b in range(nb)]
        ops.appnp_propagate_host_batched(A, seq_in[:2], ALPHA, K_ITER, out_hosts=seq_out[:2], work=work)   # warm-up
        for o in outs2:
            o.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ops.appnp_propagate_host_batched(A, seq_in, ALPHA, K_ITER, out_hosts=seq_out, work=work)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / nb
        same = all(torch.equal(o, single_out) for o in outs2[:min(nb, 2)])
        del work, outs2, single_out
        result["e2e"] = {"value": nnz * F * K_ITER / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * F_run * 4,
                         "d2h_bytes_per_step": n * F_run * 4, "ms_per_step": e2e_s * 1e3,
                         "api": "gnntf.appnp_propagate_host_batched -> gnntf_appnp_propagate_host_batched_f32 (pinned host H0 in, "
                                "host H_K out, every step; the copies of neighbouring steps overlap this step's propagation)",
                         "steps_in_call": nb, "results_equal_single_call": bool(same),
                         "single_call_ms": single_s * 1e3,
                         "single_call_api": "gnntf.appnp_propagate_host -> gnntf_appnp_propagate_host_f32 (H2D, K steps, D2H, one stream)"}
    else:
        # the REAL shard of H0 goes host -> device, the result shard comes back and is compared with the device run
        e2e = prop.propagate_host_timed(H0_local.cpu(), ALPHA, K_ITER, reps=max(1, args.steps))
        same = torch.tensor([1.0 if torch.equal(e2e["host_out"], run().cpu()) else 0.0], device=dev)
        torch.distributed.all_reduce(same, op=torch.distributed.ReduceOp.MIN)
        if rank == 0:
            result["e2e"] = {"value": nnz * F * K_ITER / e2e["seconds"], "unit": UNIT,
                             "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                             "ms_per_step": e2e["seconds"] * 1e3, "single_call_ms": e2e["single_call_seconds"] * 1e3,
                             "api": "gnntf.dist.ShardedPropagator.propagate_host_batched (every rank: pinned host shard in, "
                                    "host result shard out, every step; the copies of neighbouring steps overlap this step's propagation)",
                             "steps_in_call": max(1, args.steps),
                             "host_result_equals_device_run_on_every_rank": bool(same.item() > 0)}

    # ---- parity of the timed propagation against the C oracle (every N), then the CPU baseline --------
    big = None
    if not args.no_parity and not args.reorder:
        # what the timed region produced: this rank's rows (and columns) of H_K, in node order
        with torch.no_grad():
            mine = run()
        torch.cuda.synchronize()
        if world == 1:
            got = mine[:, :F].cpu().numpy()
        else:
            shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
            tag = os.environ.get("MASTER_PORT", "0")
            np.save(os.path.join(shm, f"gnntf_parity_{tag}_{rank}.npy"), mine.cpu().numpy())
            done_file = os.path.join(shm, f"gnntf_parity_{tag}_done")
            if rank == 0 and os.path.exists(done_file):
                os.unlink(done_file)                       # stale marker of an earlier run on this port
            meta = [None] * world
            torch.distributed.all_gather_object(meta, (rank, prop.lo, prop.hi, c0, c1))
            torch.distributed.barrier()
            got = None
            if rank == 0:
                got = np.empty((n, F_run), np.float32)
                for (rk, lo, hi, a0, a1) in meta:
                    path = os.path.join(shm, f"gnntf_parity_{tag}_{rk}.npy")
                    got[lo:hi, a0:a1] = np.load(path)
                    os.unlink(path)
                got = got[:, :F]
        if rank == 0:
            big = big_oracle(n, edges.cpu().numpy())
            deg = np.diff(big.row_ptr)
            from gnntf.sparse import LONG_THRESHOLD
            split_rows = deg > LONG_THRESHOLD
            H0_host = (H0[:, :F].cpu() if world == 1 else H0_cpu).contiguous().numpy()
            result["parity"] = parity_block(big, H0_host, got, split_rows, F)
            log(f"[parity] {json.dumps(result['parity'])}")
            del got, H0_host
            if world > 1:
                open(done_file, "w").close()
        elif world > 1:
            # the oracle runs on rank 0 with every host thread: the other ranks SLEEP until it is done (an NCCL
            # barrier would busy-poll and take the cores away from it)
            t_wait = time.time()
            while not os.path.exists(done_file) and time.time() - t_wait < 900:
                time.sleep(0.2)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t0 = time.time()
        if big is None:
            big = big_oracle(n, edges.cpu().numpy())
        v, t, sample, desc, omp, mt = cpu_arm(big, F, seconds_per_step=6.0, steps=2, warmup=1)
        result["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": omp, "kind": "port", "sample": desc,
                                  "host_cpus": os.cpu_count(), "wall_s": time.time() - t0}
        result["cpu_baseline"].update(mt or {})
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        emit_json(result)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    n, edges, F = make_workload(args, "cpu")
    nnz = 2 * edges.shape[0]
    budget = 150.0 / max(1, args.steps + args.warmup)
    big = big_oracle(n, edges.numpy())
    v, t, sample, desc, omp, mt = cpu_arm(big, F, seconds_per_step=min(20.0, budget), steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": K_ITER * nnz * F / v * 1e3, "ms_per_step_note": "K=10 propagation time extrapolated from the sampled rate",
            "sample_ms": t * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, n, edges.shape[0], nnz, F),
            "cpu_baseline": dict({"value": v, "unit": UNIT, "cores": omp, "kind": "port", "sample": desc,
                                  "host_cpus": os.cpu_count()}, **(mt or {})),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference (TensorFlow) cannot be installed in this image; this is the oracle's C port of its "
                    "CPU path (COO-order single-threaded SpMM loop + teleport), timed on a bounded sample"}
    emit_json(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="products", choices=sorted(synthetic.SHAPES))
    ap.add_argument("--ordering", default="local", choices=["local", "random"])
    ap.add_argument("--features", type=int, default=0, help="feature width (default: the shape's)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (tests only; 1.0 = BASELINE size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run check of the result against the C oracle")
    ap.add_argument("--reorder", action="store_true", help="locality-restoring internal node order (gnntf/reorder.py); its cost is part of the build time")
    ap.add_argument("--grid", default="", help="multi-GPU layout ROWSxCOLS (default: gnntf.dist.choose_grid)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "copy", "push"],
                    help="multi-GPU: copy = every rank's block of rows goes to the peers by copy-engine DMA over NVLink "
                         "(all-gather layout, no SM involved); push = the rows each peer references are sent by the leading "
                         "CTAs of the SpMM launch; auto = copy for row groups of two ranks, push otherwise")
    ap.add_argument("--nccl-exchange", action="store_true", help="multi-GPU: halo rows by NCCL all-to-all instead of the fused peer-memory push")
    ap.add_argument("--halves", type=int, default=0, help="multi-GPU: feature-column chains to pipeline (0 = default)")
    ap.add_argument("--no-graph", action="store_true", help="multi-GPU: launch every step from the host instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
        gpu_arm(args)


if __name__ == "__main__":
    main()
