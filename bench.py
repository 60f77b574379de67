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
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="library option (calz_set_option), repeatable")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed small-grid comparison with the CPU oracle")
    ap.add_argument("--sync-blocks", action="store_true", help="one host synchronisation per block (no pipelining)")
    ap.add_argument("--lag", type=int, default=3, help="projectAndNormalize calls in flight in the pipelined loop")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--config", default="c3", choices=["c3", "c4", "c5"],
                    help="c3 (default): the BASELINE workload the metric is quoted on; c4: restarted_ca_lanczos on the power-law SPD matrix "
                         "(s=6, TSQR); c5: TSQR vs CholQR sweep on n x (s+1) blocks -- c4/c5 are extra lines kept under profiles/")
    ap.add_argument("--size", dest="n", type=int, default=0,
                    help="c4: matrix order (default 2e7); c5: total rows (default 1e8)  [not --n: torchrun prefix-matches its own options]")
    ap.add_argument("--orth", default="full", choices=["local", "full", "periodic", "selective"],
                    help="c4: orthogonalisation of restarted_ca_lanczos ('local' is the reference default; on this matrix it loses "
                         "orthogonality within the first cycle in the reference as well -- see DESIGN.md)")
    ap.add_argument("--shifts", default="reference", choices=["reference", "chebyshev"],
                    help="Newton shifts: the reference's recipe (2s-step Lanczos -> eig -> Leja, ca_lanczos.m:66-72) or Chebyshev-Leja points")
    ap.add_argument("--cpu-slab", action="store_true", help="CPU arm on a 1/8 slab scaled by 8 (round-1 behaviour) instead of the full problem")
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
class CpuBlocks:
    """The reference's steady-state block on the host cores, as the reference computes it: the ORACLE (numpy/scipy restatement,
    oracle/kernels.py) with the Householder `tsqr` that normalize.m:14 calls -- MPK (matrix_powers_newton.m) +
    projectAndNormalize({Qprev},V(:,2:s+1),true).  Set-up (untimed): the matrix and the first block.  ``step()`` runs one block
    and returns its wall time."""

    def __init__(self, m, mz, s, shifts=None):
        from ca_lanczos_b200 import gallery
        from oracle import drivers, kernels
        self.kernels, self.s = kernels, s
        self.A = gallery.laplace3d(m, m, mz)
        self.n = self.A.shape[0]
        r = np.ones(self.n)
        q = r / np.sqrt(r @ r)
        if shifts is None:      # the reference's own shifts: 2s-step 'full' Lanczos -> eig -> Leja (ca_lanczos.m:66-72)
            shifts = np.diag(drivers.basis_matrix(self.A, q, s, "newton", "full"))[:s].copy()
        self.shifts = np.asarray(shifts, dtype=np.float64)
        V = kernels.matrix_powers_newton(self.A, q, s, self.shifts, 1)
        self.Qprev, _, _ = kernels.normalize(V, backend="tsqr")

    def step(self):
        k, s = self.kernels, self.s
        t0 = time.perf_counter()
        V = k.matrix_powers_newton(self.A, self.Qprev[:, s], s, self.shifts, 1)
        QZ, RZ = k.projectAndNormalize([self.Qprev], V[:, 1:], True, backend="tsqr")
        dt = time.perf_counter() - t0
        self.Qprev = np.column_stack([self.Qprev[:, s], QZ])
        return dt


def cpu_sample(args, shifts=None, max_steps=1, budget_s=150.0, warmup=0):
    """Times up to ``max_steps`` blocks of the CPU arm (at least one), stopping once ``budget_s`` of timed work is spent.
    Default: the FULL problem (no extrapolation); --cpu-slab: a 1/8 slab scaled linearly (round-1 behaviour)."""
    m, s = args.m, args.s
    full_mz = args.mz or m
    mz = full_mz
    if args.cpu_slab and m >= 64:
        mz = max(full_mz // 8, 2 * s + 2)
    cb = CpuBlocks(m, mz, s, shifts)
    times = []
    for i in range(warmup + max_steps):
        dt = cb.step()
        if i >= warmup:
            times.append(dt)
        if sum(times) > budget_s or (i < warmup and dt * (warmup + max_steps) > 2 * budget_s and i + 1 >= 1 and warmup > 1):
            if i < warmup:
                warmup = i + 1          # a block takes many seconds on the host: one warm-up block is all the budget allows
                continue
            break
    t = float(np.median(times))
    scale = (m * m * full_mz) / float(cb.n)
    cores = os.cpu_count()
    what = "the full %dx%dx%d problem" % (m, m, mz) if scale == 1.0 else "a %dx%dx%d slab = 1/%.0f of the rows, scaled linearly" % (m, m, mz, scale)
    sample = ("median of %d timed block(s) of the oracle (numpy/scipy restatement; the MATLAB reference cannot run here) on %s, %.1f s per "
              "block; Householder tsqr as normalize.m:14; scipy CSR mat-vec single-threaded, BLAS/LAPACK up to %d threads"
              % (len(times), what, t * scale, cores))
    return 1.0 / (t * scale), len(times), cores, sample


def run_reference(args):
    """--impl reference: the reference's own (MATLAB) implementation cannot run here (no Octave/MATLAB in the image), so
    this arm times its CPU restatement (oracle/, kind "port") with all host threads numpy/scipy will use, on the SAME
    configuration as the product arm (full 256^3 problem, the reference's shift recipe, Householder QR)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m, s = args.m, args.s
    mz = args.mz or m
    value, nt, cores, sample = cpu_sample(args, None, max_steps=max(1, args.steps), budget_s=120.0, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": "ca_lanczos_s_step_blocks_per_sec", "value": value, "unit": "blocks/s",
            "n_gpus": args.gpus, "steps": nt, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(m, s, args.backend), "n": m * m * mz, "nnz": 7 * m * m * mz - 2 * m * m - 4 * m * mz, "s": s,
                       "basis": "newton", "orth": "tsqr (Householder, normalize.m:14)", "layout": "scipy-csr", "partition": "host"},
            "cpu_baseline": {"value": value, "unit": "blocks/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference is MATLAB-only and cannot be executed in this image: oracle/ (numpy/scipy restatement) timed instead; a host "
                    "block takes many seconds, so the number of timed steps is bounded by a 120 s budget"}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- parity pre-check
def parity_check(args, torch, dist, ctx, world, rank, dev):
    """UNTIMED checker run before the timed region, on every N: a small grid of the same operator
    (24 x 24 x 16N, s as the bench, 4 blocks) through the same N-rank pipeline, compared with the 1-way ORACLE (checker only):
    row bounds / ghost index sets / exchange lists bit-exact, MPK on the owned rows, T entries, basis vectors, Ritz values."""
    from ca_lanczos_b200 import api, gallery
    from ca_lanczos_b200.engine import BlockEngine
    from oracle import drivers, kernels, partition
    m, s, blocks = 24, args.s, 4
    mz = 16 * world
    n, plane = m * m * mz, m * m
    lo, hi = (rank * n) // world, ((rank + 1) * n) // world
    have_lo, have_hi = max(0, lo - (s + 1) * plane), min(n, hi + (s + 1) * plane)
    A = gallery.laplace3d(m, m, mz)
    dm = api.DeviceMatrix(gallery.laplace3d(m, m, mz, row_lo=have_lo, row_hi=have_hi), s_max=s, layout=args.layout, ctx=ctx, n_glob=n,
                          row_begin=have_lo)
    L = dm.info("halo_level") if world > 1 else s
    b = partition.row_bounds(n, world)
    exact = (dm.info("row_lo"), dm.info("row_hi")) == (int(b[rank]), int(b[rank + 1]))
    if world > 1:
        exact &= L == partition.choose_halo_level(A, world, s)
        exact &= bool(np.array_equal(dm.ghost_indices(), partition.ghost_indices(A, lo, hi, L)))
        lists = partition.exchange_lists(A, world, L)
        for q in range(world):
            exact &= bool(np.array_equal(dm.recv_list(q), lists[rank][q])) and bool(np.array_equal(dm.send_list(q), lists[q][rank]))
    r = np.cos(0.61 * np.arange(n) ** 1.5) + 0.3 * np.sin(1.7 * np.arange(n))
    io = {}
    To, Qo = drivers.ca_lanczos(A, r, s, s * blocks, "newton", "local", info=io)      # the reference's own shift recipe inside
    shifts = np.diag(io["Bk"])[:s].copy()
    eng = BlockEngine(dm, s, blocks + 1, "newton", shifts, args.backend)
    eng.first_block((r / np.sqrt(r @ r))[lo:hi])
    eng.run_blocks(blocks - 1)
    T, Ql = eng.T_matrix(), eng.Q_host()
    v = r / np.sqrt(r @ r)
    V = api.matrix_powers_newton(dm, v[lo:hi], s, shifts, 1)
    Vo = kernels.matrix_powers_newton(A, v, s, shifts, 1)
    ro = np.sort(np.linalg.eig(To)[0].real)[::-1]; rg = np.sort(np.linalg.eig(T)[0].real)[::-1]
    errs = [float(np.abs(T - To).max() / np.abs(To).max()), float(np.max(np.linalg.norm(Ql - Qo[lo:hi, : Ql.shape[1]], axis=0))),
            float(np.max(np.linalg.norm(V - Vo[lo:hi], axis=0) / np.linalg.norm(Vo[lo:hi], axis=0))),
            float(np.max(np.abs(rg[:4] - ro[:4]) / np.abs(ro[:4]))), 0.0 if exact else 1.0,
            0.0 if eng.second == [i["second_pass"] for i in io["pan"]] else 1.0]
    dm.close()
    t = torch.tensor(errs, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e = [float(x) for x in t.tolist()]
    return {"grid": "%dx%dx%d" % (m, m, mz), "blocks": blocks, "halo_level": int(L), "T_err": e[0], "Q_err": e[1], "mpk_err": e[2], "ritz_err": e[3],
            "ghost_sets_exact": e[4] == 0.0, "second_pass_pattern_matches": e[5] == 0.0,
            "pass": bool(e[0] <= 1e-10 and e[1] <= 1e-10 and e[2] <= 1e-13 and e[3] <= 1e-8 and e[4] == 0.0 and e[5] == 0.0),
            "what": "N-rank device pipeline vs the 1-way CPU oracle (oracle/drivers.py, oracle/partition.py), untimed; tolerances: T 1e-10 "
                    "relative, basis vectors 1e-10, MPK 1e-13, Ritz values 1e-8, index sets bit-exact"}


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
    for kv in args.opt:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    parity = None if args.no_parity else parity_check(args, torch, dist, ctx, world, rank, dev)

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
    K, W = args.steps, args.warmup
    n_own = dm.n
    q0 = np.full(n_own, 1.0 / np.sqrt(n))                       # r = ones(n,1), normalised (ca_lanczos.m:55)
    if args.shifts == "reference":
        # the reference's shift recipe on the device (ca_lanczos.m:66-72): 2s-step 'full' Lanczos -> eig -> modified Leja order
        from ca_lanczos_b200 import solver
        qd = torch.as_tensor(q0, device=dev)
        torch.cuda.synchronize(dev)
        shifts = solver.newton_shifts(dm, qd.data_ptr(), s, "full")[:s]
        del qd
    else:
        shifts = gallery.leja_points(0.0, 12.0, s)
    eng = BlockEngine(dm, s, K + W + 10, "newton", shifts, args.backend)
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
    # loss of orthogonality of the whole timed run, on the device: compute_orth_err (ca_lanczos.m:99-107) and ||I - Q'Q||_F
    from ca_lanczos_b200 import solver as _solver
    orth_errs = _solver.engine_orth_errors(eng)

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
    # ---- roofline of the dominant kernel: one SpMV launch of the MPK on this rank.
    #  ALGORITHMIC bytes (CSR, 12 B per non-zero + 4(n+1) + 16 n over the owned rows) are what an uncompressed matrix
    #  has to stream; the dictionary-coded layout MOVES far fewer (1 code byte per padded non-zero + slice pointers + x + y over the
    #  local rows, ghosts included).  `achieved`/`frac` are on the bytes the kernel moves, so that frac is a real fraction of the
    #  copy peak; the compression gain is reported separately as `algorithmic_speedup`.
    own_nnz = nnz if world == 1 else int(round(nnz * n_own / n))
    alg_bytes = 12 * own_nnz + 4 * (n_own + 1) + 16 * n_own
    patterns = dm.layout == "selld" and dm.info("pattern_cover_pct") >= 50 and os.environ.get("CALZ_OPTS", "").find("mpk_patterns=0") < 0
    if patterns:      # slice-pattern kernel: one pattern byte per slice, code bytes only for the slices without a pattern, x and y
        cover = dm.info("pattern_cover_pct") / 100.0
        moved = int((1.0 - cover) * dm.info("sell_padded_nnz")) + n_loc // 32 + 16 * n_loc
    elif dm.layout == "selld":
        moved = dm.info("sell_padded_nnz") + 4 * (n_loc // 32 + 2) + 16 * n_loc
    elif dm.layout == "sell":
        moved = 12 * dm.info("sell_padded_nnz") + 4 * (n_loc // 32 + 1) + 16 * n_loc
    else:
        moved = 12 * nnz_loc + 4 * (n_loc + 1) + 16 * n_loc
    launch_ms = ms_mpk / s
    achieved = moved / (launch_ms * 1e-3) / 1e9
    traffic = None
    kname = "k_spmv_selp" if patterns else {"selld": "k_spmv_selld", "sell": "k_spmv_sell", "csr": "k_spmv_csr"}.get(dm.layout, "k_spmv")
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get("%s_dram_bytes_per_launch_n%d" % (kname, world), tj.get(kname + "_dram_bytes_per_launch") if world == 1 else None)
    except Exception:
        pass
    note = ("bytes the kernel moves per launch (model: codes/values + indices + slice pointers + x + y over the local rows) / live CUDA-event "
            "time of the MPK / s; `traffic` = ncu dram bytes per launch of the same kernel at this N (profiles/ncu_traffic.json, null if not "
            "captured at this N)")
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
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "moved_bytes_per_launch": moved,
                     "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_speedup": alg_bytes / moved,
                     "algorithmic_gbs": alg_bytes / (launch_ms * 1e-3) / 1e9, "launch_ms": launch_ms,
                     "traffic_frac": (traffic / (launch_ms * 1e-3) / 1e9 / peak) if traffic else None, "note": note},
        "gpu_launches": int(launches), "clocks": clocks,
        "check": {"ritz_max": ritz_max, "lambda_max": float(6.0 - 6.0 * np.cos(m * np.pi / (m + 1))),
                  "second_pass_fraction": second_frac, "shifts": [float(x) for x in shifts], "shift_recipe": args.shifts,
                  "orth_err_lastblock": orth_errs[0], "orth_err_fro": orth_errs[1], "parity": parity,
                  "T_err": parity["T_err"] if parity else None, "Q_err": parity["Q_err"] if parity else None,
                  "ghost_sets_exact": parity["ghost_sets_exact"] if parity else None},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if gram is not None:
        line["roofline_gram"] = gram
    if sell_ref is not None:
        line["roofline_sell"] = sell_ref
    if not args.no_cpu and world == 1:      # reported baseline: rank 0 at N=1 only, ONE block of the full problem (no extrapolation)
        v, nt, cores, sample = cpu_sample(args, shifts, max_steps=1, budget_s=60.0, warmup=0)
        line["cpu_baseline"] = {"value": v, "unit": "blocks/s", "cores": cores, "kind": "port", "sample": sample}
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
    # ---- PCIe floor of this call sequence: the same bytes with plain pinned copies, one direction at a time (the calls are
    #      synchronous and each output depends on all of its inputs, so the two directions cannot overlap inside a call)
    big = torch.empty((s + 1, n_own), dtype=torch.float64, device=dev)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(2):
        big.copy_(tV, non_blocking=True); torch.cuda.synchronize(dev)
        tV.copy_(big, non_blocking=True); torch.cuda.synchronize(dev)
    pc = (time.perf_counter() - t0) / 4
    h2d_gbs = 8.0 * n_own * (s + 1) / pc / 1e9
    floor_s = (h2d / world + d2h / world) / (h2d_gbs * 1e9)
    del big
    # ---- the same block through the HANDLE mode of the call surface (api.DeviceBlock == mex/calz_vec.m): same public functions,
    #      device blocks in and out, only the small R factors come back to the host
    hv = None
    try:
        Qb = api.DeviceBlock(n_own, s + 1, ctx)
        Qb[:, 0:s + 1] = Qprev
        Vb = api.DeviceBlock(n_own, s + 1, ctx)
        Zb = api.DeviceBlock(n_own, s, ctx)
        ht = []
        for i in range(steps + 2):
            barrier()
            t0 = time.perf_counter()
            api.matrix_powers_newton(dm, Qb[:, s:s + 1], s, shifts, 1, out=Vb)
            _, RZh = api.projectAndNormalize([Qb[:, 0:s + 1]], Vb[:, 1:s + 1], True, ctx=ctx, out=Zb)
            barrier()
            if i > 1:
                ht.append(time.perf_counter() - t0)
            Qb[:, 0:1] = Qb[:, s:s + 1]
            Qb[:, 1:s + 1] = Zb
        th = torch.tensor([float(np.mean(ht))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(th, op=dist.ReduceOp.MAX)
        hv = {"value": 1.0 / float(th.item()), "unit": "blocks/s", "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": int(8 * ((s + 1) * s + s * s)), "steps": steps,
              "api": "the same two public calls with api.DeviceBlock arguments (handle mode, mex/calz_vec.m): blocks stay in HBM, "
                     "the coefficient blocks come back to the host every call; synchronous, one call at a time"}
    except Exception as e:                       # the value-semantics number above must not depend on this extra
        hv = {"error": str(e)}
    return {"value": 1.0 / dt, "unit": "blocks/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": steps,
            "api": "matrix_powers_newton + projectAndNormalize (host arrays, pinned), timed with perf_counter around barrier+sync",
            "pcie_gbs_measured": h2d_gbs, "pcie_floor_blocks_per_s": 1.0 / floor_s, "frac_of_pcie_floor": floor_s / dt,
            "handles": hv}


def _dist_setup(args):
    import torch
    import torch.distributed as dist
    from ca_lanczos_b200 import api
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = api.Context(local)
    if world > 1:
        ids = [api.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.init_comm(world, rank, ids[0])
    return torch, dist, ctx, world, rank, local, dev


# ----------------------------------------------------------------------------------------------------- C4
def run_c4(args):
    """BASELINE configs[3]: restarted_ca_lanczos (restarted_ca_lanczos.m defaults: n_wanted = 10, tol = 1e-8, max_lanczos = 60; --orth
    full|local) on the synthetic power-law SPD matrix (n = 2e7, ~4e8 non-zeros, Jacobi-scaled: gallery.powerlaw_spd_rows), s = 6
    Newton basis, TSQR.  Every rank generates and uploads only its own rows; the level-1 closure of a row block of this graph is
    (almost) every row, so the MPK exchanges the current basis vector before EVERY step (halo level 1) instead of once per block.
    One line: blocks/s over the whole solve (all restarts), plus the solve's own diagnostics and an untimed small-n check of the
    converged eigenvalues against LAPACK."""
    torch, dist, ctx, world, rank, local, dev = _dist_setup(args)
    from ca_lanczos_b200 import api, gallery, restart
    n = args.n or 20_000_000
    s = 6 if args.s == 8 else args.s
    backend = "tsqr" if args.backend == "cholqr2" else args.backend
    orth = args.orth

    # ---- untimed checker: the same N-rank code path at n = 3000, converged eigenvalues against the dense LAPACK spectrum
    def small_check():
        ns = 3000
        l0, h0 = (rank * ns) // world, ((rank + 1) * ns) // world
        As = gallery.powerlaw_spd_rows(ns, 20.0, seed=0, row_lo=l0, row_hi=h0, jacobi=True)
        dms = api.DeviceMatrix(As, s_max=s, layout=args.layout, ctx=ctx, n_glob=ns, row_begin=l0)
        o = restart.DeviceOps(dms, backend=backend, colsum=np.asarray(abs(As).sum(axis=1)).ravel())
        eigs, Qc, nres, rn, oe, order = restart.restarted_ca_lanczos(o, o.from_host(np.ones(h0 - l0)), 60, 10, s, "newton", "full", 1e-8)
        out = {"n": ns, "restarts": int(nres), "nconv": int(len(eigs)), "orth_err_fro": float(oe[-1]) if len(oe) else None,
               "max_rel_residual": float(np.max(rn[-1][:len(eigs)])) if len(eigs) else None}
        if rank == 0:
            ev = np.linalg.eigvalsh(gallery.powerlaw_spd_rows(ns, 20.0, seed=0, jacobi=True).toarray())
            out["max_rel_err_vs_lapack"] = float(max(np.abs(ev - x).min() / abs(x) for x in eigs)) if len(eigs) else None
            out["pass"] = bool(len(eigs) >= 10 and out["max_rel_err_vs_lapack"] <= 1e-8)
        dms.close()
        return out
    check = None if args.no_parity else small_check()
    if world > 1:
        dist.barrier()                    # rank 0 spent ~30 s in LAPACK: re-align the ranks before the next device collective

    lo, hi = (rank * n) // world, ((rank + 1) * n) // world
    t0 = time.time()
    A = gallery.powerlaw_spd_rows(n, 20.0, seed=0, row_lo=lo, row_hi=hi, jacobi=True)
    gen_s = time.time() - t0
    if world > 1:
        dist.barrier()                    # generation time differs from rank to rank
    nnz_own = int(A.nnz)
    colsum = np.asarray(abs(A).sum(axis=1)).ravel()          # symmetric: column sums of |A| restricted to the owned columns
    maxrow = int(np.diff(A.indptr).max())
    t0 = time.time()
    dm = api.DeviceMatrix(A, s_max=s, layout=args.layout, ctx=ctx, n_glob=n, row_begin=lo)
    del A
    upload_s = time.time() - t0
    ops = restart.DeviceOps(dm, backend=backend, colsum=colsum)
    r = ops.from_host(np.ones(hi - lo))                         # r = ones(n,1) as test_restart_diagonal_matrices.m:16

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    # warm-up: one restart cycle
    restart.restarted_ca_lanczos(ops, r, 60, 10, s, "newton", orth, 1e-8, max_restarts=1, want_orth_err=False)
    barrier()
    ctx.launch_count(reset=True)
    sampler = ClockSampler(local, enabled=not args.no_clocks) if rank == 0 else None
    if sampler:
        sampler.mark()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    eigs, Qc, nres, rnorms, orth_err, order = restart.restarted_ca_lanczos(ops, r, 60, 10, s, "newton", orth, 1e-8)
    e1.record(stream)
    e1.synchronize()
    barrier()
    wall = time.perf_counter() - t0
    launches = ctx.launch_count()
    clocks = sampler.stop() if sampler else None
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, wall, float(nnz_own), float(maxrow), float(dm.info("n_long_rows"))], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    else:
        tmax = tsum = t
    ms, wall = float(tmax[0]), float(tmax[1])
    nnz = int(tsum[2])
    iters = 60 // s
    blocks = nres * (iters + 1)
    # a steady-state block alone (MPK + projectAndNormalize({Qprev, Q_conv})), CUDA events, for the per-phase split
    if rank == 0:
        last = rnorms[-1] if len(rnorms) else np.zeros(1)
        line = {"metric": "ca_lanczos_s_step_blocks_per_sec", "value": blocks / (ms * 1e-3), "unit": "blocks/s", "n_gpus": world, "steps": blocks,
                "warmup": iters + 1, "ms_per_step": ms / blocks, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "restarted_ca_lanczos_powerlaw_spd_jacobi_n%d_s%d_newton_%s_%s" % (n, s, orth, backend), "n": n, "nnz": nnz, "s": s,
                           "basis": "newton", "orth": orth + ", " + backend, "layout": dm.layout, "partition": "rows/%d" % world,
                           "halo_level": dm.info("halo_level") if world > 1 else s, "max_lanczos": 60, "n_wanted": 10, "tol": 1e-8,
                           "max_row_nnz": int(tmax[3]), "long_rows_max_per_rank": int(tmax[4]),
                           "nnz_imbalance": float(tmax[2]) * world / max(nnz, 1), "gen_s": round(gen_s, 1), "upload_s": round(upload_s, 1)},
                "solve": {"seconds_device": ms * 1e-3, "seconds_wall": wall, "restarts": int(nres), "blocks": int(blocks),
                          "eigs": [float(x) for x in eigs], "max_rel_residual": float(np.max(last[:len(eigs)])) if len(eigs) else None,
                          "orth_err_fro_last": float(orth_err[-1]) if len(orth_err) else None,
                          "note": "wall time includes the host algebra (eig of T, Leja) and one synchronisation per projectAndNormalize; "
                                  "the O(n) work of a restart (Ritz vectors, residuals, ||I-Q'Q||_F) is inside the timed region"},
                "check": check, "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------- C5
def run_c5(args):
    """BASELINE configs[4]: TSQR (tsqr.m) vs CholQR (cholqr.m) on n x (s+1) blocks, s = 4..16, rows split over the N ranks; input from
    the counter-based generator (gallery.tall_skinny, column j scaled by 2^-j), generated on the device."""
    torch, dist, ctx, world, rank, local, dev = _dist_setup(args)
    import ctypes as C
    from ca_lanczos_b200 import _lib, gallery, solver
    n = args.n or 100_000_000
    lo, hi = (rank * n) // world, ((rank + 1) * n) // world
    nl = hi - lo
    ld = (nl + 31) // 32 * 32
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    peak, peak_src = measured_peak()
    reps = max(2, min(args.steps, 5))
    sweep = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for s in (4, 6, 8, 10, 12, 14, 16):
        c = s + 1
        X = gallery.tall_skinny_device(nl, c, dev, ld, row0=lo)
        Q = torch.empty((c, ld), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        ent = {"s": s, "c": c}
        Rs = {}
        for backend, per in (("tsqr", 16), ("cholqr", 24)):
            Rh = np.zeros((c, c), order="F")
            info = C.c_int()

            def fn():
                if backend == "tsqr":
                    _lib.check(ctx.lib.calz_tsqr(ctx.h, nl, c, X.data_ptr(), ld, Q.data_ptr(), ld, Rh.ctypes.data_as(_lib.c_dp)), ctx.h)
                else:
                    _lib.check(ctx.lib.calz_cholqr(ctx.h, nl, c, X.data_ptr(), ld, Q.data_ptr(), ld, Rh.ctypes.data_as(_lib.c_dp),
                                                   C.byref(info)), ctx.h)
            fn(); fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
            e1.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            oe = solver.orth_error(ctx, nl, [(Q.data_ptr(), ld, c)], "fro")
            Rs[backend] = Rh.copy()
            gbs = per * n * c / (ms * 1e-3) / 1e9
            ent[backend] = {"ms": ms, "compulsory_bytes": per * n * c, "gbs": gbs, "frac_of_copy_peak_per_gpu": gbs / world / peak,
                            "orth_err_fro": oe}
        ent["R_rel_diff"] = float(np.linalg.norm(Rs["tsqr"] - Rs["cholqr"]) / np.linalg.norm(Rs["tsqr"]))
        ent["tsqr_over_cholqr_time"] = ent["tsqr"]["ms"] / ent["cholqr"]["ms"]
        sweep.append(ent)
        del X, Q
    if rank == 0:
        mid = [e for e in sweep if e["s"] == 8][0]
        line = {"metric": "tall_skinny_qr_compulsory_gbs", "value": mid["tsqr"]["gbs"], "unit": "GB/s", "n_gpus": world, "steps": reps, "warmup": 2,
                "ms_per_step": mid["tsqr"]["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": "tsqr_vs_cholqr_sweep_n%d_x_(s+1)_s4..16" % n, "n": n, "partition": "rows/%d" % world,
                                                "input": "splitmix64 counter generator U(-1,1), column j scaled by 2^-j", "peak": peak,
                                                "peak_source": peak_src},
                "sweep": sweep, "note": "value = TSQR at s = 8 (16*n*c compulsory bytes / time, all GPUs); CholQR counts 24*n*c"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c4":
        run_c4(args)
    elif args.config == "c5":
        run_c5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
