#!/usr/bin/env python
"""bench.py -- CA-Lanczos s-step blocks/sec on the C3 workload (BASELINE.json configs[2], the configuration the
north_star target is quoted on and the largest that fits one GPU):

    3-D 7-point Laplacian 256^3 (n = 16 777 216, nnz = 117 047 296), s = 8, Newton basis (Leja-ordered Chebyshev
    shifts of the spectral interval), CholQR orthogonalisation, start vector ones(n,1), rows partitioned over the
    N GPUs of one box (strong scaling: the problem is fixed, PA1 ghost zones, one halo exchange per block).

One *step* = one *block* = one outer iteration k>1 of ca_lanczos_basic (ca_lanczos.m:166-225): one matrix powers
kernel (s SpMVs) + projectAndNormalize({Qprev}, V(:,2:s+1), true) including its second pass when the reference's
50 % norm-drop test fires, + the O(s^3) host algebra that extends T.

    python bench.py --gpus N --steps K --warmup W                 (torchrun for N>1, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W  (CPU restatement of the reference, rank 0 only)

Prints ONE JSON line (rank 0).  `value` = blocks/s with everything resident in HBM (CUDA events on libcalz'
stream, max over ranks); `e2e` = the same block through the reference-facing host API (host arrays in pinned
memory in, host arrays out: H2D/D2H inside the timed region); `roofline` = the dominant kernel (the SELL SpMV
of the MPK) against the measured HBM copy peak; `cpu_baseline` = the oracle timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=256, help="grid edge (256 = the BASELINE workload)")
    ap.add_argument("--mz", type=int, default=0, help="z extent (default m): --mz 32 is the per-rank share of the 8-GPU run")
    ap.add_argument("--s", type=int, default=8)
    ap.add_argument("--backend", default="cholqr2", choices=["cholqr", "cholqr2", "tsqr"])
    ap.add_argument("--layout", default="auto", choices=["auto", "selld", "sell", "csr"])
    ap.add_argument("--l2-chunk-mb", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gram", action="store_true")
    ap.add_argument("--no-sell-ref", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-clocks", action="store_true")
    ap.add_argument("--sync-blocks", action="store_true", help="one host synchronisation per block (no pipelining)")
    ap.add_argument("--lag", type=int, default=3, help="projectAndNormalize calls in flight in the pipelined loop")
    ap.add_argument("--e2e-steps", type=int, default=3)
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, no MEASURED_PEAKS.json)"


def workload_name(m, s, backend):
    return "laplace3d_%d^3_7pt_s%d_newton_%s_r=ones" % (m, s, backend)


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region by a separate `nvidia-smi -lms 20` process (the recipe
    of B200_PROFILING.md).  A separate process on purpose: an in-process NVML polling thread competes with the launching
    thread for the GIL and the driver and was measured to perturb the 2-rank run."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0, enabled=True):
        self.rows, self.proc = [], None
        if not enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self, wait_s=3.0):
        """Call right before the timed region.  nvidia-smi needs a while to start: make sure it is already delivering samples."""
        deadline = time.perf_counter() + wait_s
        while self.proc and not self.rows and time.perf_counter() < deadline:
            time.sleep(0.005)
        self.t0 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["not sampled"]}
        t1 = time.perf_counter()
        time.sleep(0.06)                          # let the samples taken inside the region reach the pipe
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        t0 = getattr(self, "t0", 0.0)
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.03] or [r for (_, r) in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_block_seconds(m, mz, s, shifts, backend, repeats=1):
    """One steady-state block of the ORACLE (numpy/scipy restatement of the reference) on an m x m x mz slab of the
    workload's matrix: MPK (matrix_powers_newton.m) + projectAndNormalize({Qprev},V(:,2:s+1),true).  Returns seconds."""
    from ca_lanczos_b200 import gallery
    from oracle import kernels
    A = gallery.laplace3d(m, m, mz)
    n = A.shape[0]
    r = np.ones(n)
    q = r / np.sqrt(r @ r)
    ob = "tsqr" if backend == "tsqr" else "cholqr"
    V = kernels.matrix_powers_newton(A, q, s, shifts, 1)
    Qprev, _, _ = kernels.normalize(V, backend="tsqr")            # first block (untimed set-up)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        V = kernels.matrix_powers_newton(A, Qprev[:, s], s, shifts, 1)
        QZ, RZ = kernels.projectAndNormalize([Qprev], V[:, 1:], True, backend=ob)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        Qprev = np.column_stack([Qprev[:, s], QZ])
    return best, n


def run_reference(args):
    """--impl reference: the reference's own (MATLAB) implementation cannot run here (no Octave/MATLAB in the image), so
    this arm times its CPU restatement (oracle/, kind "port") with all host threads numpy/scipy will use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ca_lanczos_b200 import gallery
    m, s = args.m, args.s
    shifts = gallery.leja_points(0.0, 12.0, s)
    # bounded sample: one block on an m x m x (m/8) slab (1/8 of the rows); blocks/s of the full problem = 1/(8 t)
    div = 8 if m >= 64 else 1
    mz = max(m // div, 2 * s + 2)
    times = []
    for i in range(args.warmup + args.steps):
        if i >= args.warmup + 3 and sum(times) > 150.0:       # keep the arm within a few minutes
            break
        dt, nrows = cpu_block_seconds(m, mz, s, shifts, args.backend)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.median(times))
    scale = (m * m * m) / float(nrows)
    value = 1.0 / (t * scale)
    cores = os.cpu_count()
    sample = "median of %d timed blocks on a %dx%dx%d slab (1/%.0f of the rows), scaled linearly to %d^3; scipy CSR mat-vec is " \
             "single-threaded, BLAS/LAPACK use up to %d threads" % (len(times), m, m, mz, scale, m, cores)
    line = {"impl": "reference", "metric": "ca_lanczos_s_step_blocks_per_sec", "value": value, "unit": "blocks/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(m, s, args.backend), "n": m ** 3, "nnz": 7 * m ** 3 - 6 * m * m, "s": s, "basis": "newton",
                       "orth": args.backend, "layout": "scipy-csr", "partition": "host"},
            "cpu_baseline": {"value": value, "unit": "blocks/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference is MATLAB-only and cannot be executed in this image: oracle/ (numpy/scipy restatement) timed instead"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from ca_lanczos_b200 import api, gallery
    from ca_lanczos_b200.engine import BlockEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE=%d but --gpus %d" % (world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sampler = ClockSampler(local, enabled=not args.no_clocks) if rank == 0 else None      # started early: nvidia-smi is slow to come up
    ctx = api.Context(local)
    if world > 1:
        ids = [api.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.init_comm(world, rank, ids[0])
    ctx.set_option("mpk_l2_chunk_bytes", args.l2_chunk_mb << 20)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    m, s = args.m, args.s
    mz = args.mz or m
    n = m * m * mz
    plane = m * m
    nnz = 7 * n - 2 * plane - 4 * m * mz
    lo, hi = (rank * n) // world, ((rank + 1) * n) // world
    have_lo, have_hi = max(0, lo - s * plane), min(n, hi + s * plane)
    t0 = time.time()
    A = gallery.laplace3d(m, m, mz, row_lo=have_lo, row_hi=have_hi)
    dm = api.DeviceMatrix(A, s_max=s, layout=args.layout, ctx=ctx, n_glob=n, row_begin=have_lo)
    del A
    setup_s = time.time() - t0
    shifts = gallery.leja_points(0.0, 12.0, s)
    K, W = args.steps, args.warmup
    eng = BlockEngine(dm, s, K + W + 10, "newton", shifts, args.backend)
    n_own = dm.n
    q0 = np.full(n_own, 1.0 / np.sqrt(n))                       # r = ones(n,1), normalised (ca_lanczos.m:55)
    eng.first_block(q0)
    for _ in range(W):
        eng.next_block()
    ctx.sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- timed region: exactly K blocks, CUDA events on libcalz' stream, per-phase events for the roofline
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    eng.phase_events = None
    if sampler:
        sampler.mark()                        # waits (before the barrier) until nvidia-smi delivers samples
    barrier()
    if sampler:
        sampler.mark(0.0)
    ctx.launch_count(reset=True)
    host_enqueue_ms = None
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record(stream)
    if args.sync_blocks:
        for i in range(K):
            eng.next_block(events=(ev[i], stream))
    else:
        eng.host_enqueue_s = 0.0
        eng.run_blocks(K, lag=args.lag)          # results of up to `lag` blocks in flight; returns when T holds all K blocks
        host_enqueue_ms = 1e3 * eng.host_enqueue_s / K
    e_stop.record(stream)
    e_stop.synchronize()
    barrier()
    launches = ctx.launch_count()
    clocks = sampler.stop() if sampler else None
    ms_total = e_start.elapsed_time(e_stop)
    if not args.sync_blocks:
        # per-phase split from a second, synchronous pass over a few blocks (not part of `value`)
        kk = min(K, 5)
        eng.next_block()
        barrier()
        for i in range(kk):
            eng.next_block(events=(ev[i], stream))
        ctx.sync()
    else:
        kk = K
    ms_mpk = float(np.mean([ev[i][0].elapsed_time(ev[i][1]) for i in range(kk)]))
    ms_orth = float(np.mean([ev[i][1].elapsed_time(ev[i][2]) for i in range(kk)]))
    if world > 1:
        t = torch.tensor([ms_total, ms_mpk, ms_orth], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_mpk, ms_orth = [float(x) for x in t.tolist()]
    value = K / (ms_total * 1e-3)

    # sanity of the timed run itself: the Ritz values of T must sit inside the spectrum (0,12) and the top one
    # must approach lambda_max = 12 - O(1/m^2); the pass-2 branch must have been exercised like in the reference
    ritz_max = float(np.max(np.linalg.eigvals(eng.T_matrix()).real))
    second_frac = float(np.mean(eng.second)) if eng.second else 0.0

    # ---- Gram X'X (the dense contraction of CholQR) against the fp64 tensor pipe, measured live (1 GPU only; not part of `value`)
    gram = None
    if world == 1 and not args.no_gram:
        import ctypes as C
        from ca_lanczos_b200 import _lib
        tf = C.c_double()
        _lib.check(ctx.lib.calz_dmma_peak(ctx.h, C.byref(tf)), ctx.h)
        Cd = torch.zeros(s * s, dtype=torch.float64, device=dev)
        xp = eng._qcol(1)
        torch.cuda.synchronize(dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(8):
            if rep == 3:
                g0.record(stream)
            _lib.check(ctx.lib.calz_gram(ctx.h, n_own, s, xp, eng.ld, s, xp, eng.ld, Cd.data_ptr()), ctx.h)
        g1.record(stream)
        g1.synchronize()
        gms = g0.elapsed_time(g1) / 5
        gflop = 2.0 * n_own * s * s
        gram = {"kernel": "k_tsmm_tn (Gram X'X of an n x %d block, mma.sync m8n8k4 f64)" % s, "bound": "tensor",
                "achieved": gflop / (gms * 1e-3) / 1e12, "peak": tf.value, "unit": "TFLOP/s",
                "frac": gflop / (gms * 1e-3) / 1e12 / tf.value, "peak_source": "measured live (calz_dmma_peak: register-resident DMMA chains)",
                "launch_ms": gms, "hbm_frac": 8.0 * n_own * s / (gms * 1e-3) / 1e9 / measured_peak()[0],
                "note": "intensity c/4 = %.2f flop/B keeps the Gram HBM-bound on B200 (SURVEY 8d): hbm_frac is the binding fraction" % (s / 4.0)}

    # ---- the same MPK on the UNCOMPRESSED SELL layout (A really streamed from HBM): the conventional roofline number, measured live
    sell_ref = None
    if world == 1 and not args.no_sell_ref and dm.layout == "selld":
        import ctypes as C
        from ca_lanczos_b200 import _lib
        dm2 = api.DeviceMatrix(gallery.laplace3d(m, m, mz), s_max=s, layout="sell", ctx=ctx)
        re = np.ascontiguousarray(shifts, dtype=np.float64)
        Vp, ldp = C.c_void_p(), C.c_int64()
        qp = eng._qcol(1)

        def mpk2():
            _lib.check(ctx.lib.calz_mpk_inplace(dm2.h, C.c_void_p(qp), s, re.ctypes.data_as(_lib.c_dp), None, 1, 0, C.byref(Vp),
                                                C.byref(ldp)), ctx.h)
        for _ in range(3):
            mpk2()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(5):
            mpk2()
        s1.record(stream)
        s1.synchronize()
        sms = s0.elapsed_time(s1) / (5 * s)
        sbytes = 12 * nnz + 4 * (n + 1) + 16 * n
        sell_traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                sell_traffic = json.load(f).get("k_spmv_sell_dram_bytes_per_launch")
        except Exception:
            pass
        sell_ref = {"kernel": "k_spmv_sell (one SpMV step of the MPK, plain SELL-32: 12 B per non-zero streamed)", "bound": "hbm",
                    "achieved": sbytes / (sms * 1e-3) / 1e9, "peak": measured_peak()[0], "unit": "GB/s",
                    "frac": sbytes / (sms * 1e-3) / 1e9 / measured_peak()[0], "traffic": sell_traffic, "launch_ms": sms,
                    "algorithmic_bytes_per_launch": sbytes}
        dm2.close()

    # ---- e2e: the same block through the reference-facing host API (host arrays in, host arrays out)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, ctx, dm, eng, shifts, world, dev, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    n_loc, nnz_loc = dm.info("n_loc"), dm.info("nnz_loc")
    # algorithmic bytes of ONE SpMV launch on this rank: 12*nnz + 4(n+1) + 16 n over the owned rows
    own_nnz = nnz if world == 1 else int(round(nnz * n_own / n))
    spmv_bytes = 12 * own_nnz + 4 * (n_own + 1) + 16 * n_own
    launch_ms = ms_mpk / s
    achieved = spmv_bytes / (launch_ms * 1e-3) / 1e9
    traffic = None
    kname = {"selld": "k_spmv_selld", "sell": "k_spmv_sell", "csr": "k_spmv_csr"}.get(dm.layout, "k_spmv")
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(kname + "_dram_bytes_per_launch")
        if traffic is not None and world > 1:
            traffic = traffic * n_own / n               # the capture is of the 1-GPU launch
    except Exception:
        pass
    note = None
    if dm.layout == "selld":
        note = ("dictionary-coded SELL: A is stored losslessly as 1 code byte per non-zero (7 distinct (offset,value) pairs), so the DRAM "
                "traffic of a launch is ~0.38 GB while the ALGORITHMIC bytes (CSR, 12 B/nnz, SURVEY 8d) stay 1.74 GB: frac > 1 means bytes "
                "not moved, not work not done (bit-identical to the plain SELL kernel, tests/test_gpu_mpk.py); run with --layout sell for "
                "the uncompressed kernel (0.89 of the copy peak)")
    line = {
        "metric": "ca_lanczos_s_step_blocks_per_sec", "value": value, "unit": "blocks/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(m, s, args.backend), "n": n, "nnz": nnz, "s": s, "basis": "newton", "orth": args.backend,
                   "layout": dm.layout, "partition": "rows/%d" % world, "l2": "inputs larger than L2 (A %.2f GB, basis %.2f GB per rank)" %
                   (12 * nnz_loc / 1e9, 8 * n_loc * (s + 1) / 1e9), "l2_chunk_mb": args.l2_chunk_mb, "setup_s": round(setup_s, 1)},
        "phases_ms": {"mpk": ms_mpk, "project_and_normalize": ms_orth, "mpk_share": ms_mpk / (ms_mpk + ms_orth),
                      "host_enqueue_ms_per_block": host_enqueue_ms},
        "roofline": {"kernel": kname + " (one SpMV step of the MPK)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": spmv_bytes, "launch_ms": launch_ms,
                     "traffic_frac": (traffic / (launch_ms * 1e-3) / 1e9 / peak) if traffic else None, "note": note},
        "gpu_launches": int(launches), "clocks": clocks,
        "check": {"ritz_max": ritz_max, "lambda_max": float(6.0 - 6.0 * np.cos(m * np.pi / (m + 1))),
                  "second_pass_fraction": second_frac},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if gram is not None:
        line["roofline_gram"] = gram
    if sell_ref is not None:
        line["roofline_sell"] = sell_ref
    if not args.no_cpu and world == 1:      # reported baseline: rank 0 at N=1 only
        div = 8 if m >= 64 else 1
        mz = max(m // div, 2 * s + 2)
        dt, nrows = cpu_block_seconds(m, mz, s, shifts, args.backend)
        scale = n / float(nrows)
        line["cpu_baseline"] = {"value": 1.0 / (dt * scale), "unit": "blocks/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "one block of the oracle (numpy/scipy restatement; the MATLAB reference cannot run here) on a "
                                          "%dx%dx%d slab = 1/%.0f of the rows (%.1f s), scaled linearly; scipy CSR mat-vec single-threaded, "
                                          "BLAS up to %d threads" % (m, m, mz, scale, dt, os.cpu_count())}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, torch, dist, ctx, dm, eng, shifts, world, dev, barrier):
    """The same block through the drop-in host API: V = matrix_powers_newton(A, q, s, lambda, 1);
    [QZ,RZ] = projectAndNormalize({Qprev}, V(:,2:s+1), true) -- host (pinned) arrays in, host arrays out."""
    from ca_lanczos_b200 import api
    s = args.s
    n_own = dm.n

    def pinned(shape):
        t = torch.empty(shape[::-1], dtype=torch.float64, pin_memory=True)       # (cols, rows) row-major == rows x cols col-major
        return t, t.numpy().T

    tq, q = pinned((n_own, 1))
    tQ, Qprev = pinned((n_own, s + 1))
    tV, V = pinned((n_own, s + 1))
    tZ, QZ = pinned((n_own, s))
    k = eng.k
    Qprev[:, :] = eng.Q[(k - 1) * s : k * s + 1, :n_own].T.cpu().numpy()
    q[:, 0] = Qprev[:, s]
    steps = max(1, min(args.e2e_steps, args.steps))
    api.set_qr_backend(args.backend)
    times = []
    for i in range(steps + 1):
        barrier()
        t0 = time.perf_counter()
        api.matrix_powers_newton(dm, q[:, 0], s, shifts, 1, out=V)
        info = {}
        _, RZ = api.projectAndNormalize([Qprev], V[:, 1:], True, ctx=ctx, info=info, out=QZ)
        barrier()
        dt = time.perf_counter() - t0
        if i > 0:
            times.append(dt)
        Qprev[:, 0] = Qprev[:, s]
        Qprev[:, 1:] = QZ
        q[:, 0] = Qprev[:, s]
    t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    h2d = 8 * n_own * (1 + (s + 1) + s) * world
    d2h = 8 * n_own * ((s + 1) + s) * world
    return {"value": 1.0 / dt, "unit": "blocks/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": steps,
            "api": "matrix_powers_newton + projectAndNormalize (host arrays, pinned), timed with perf_counter around barrier+sync"}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
