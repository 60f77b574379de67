"""Per-kernel timings of the hot path on one B200 (CUDA events on libcalz' stream, L2 flushed by size:
every operand is far larger than the 126 MB L2).  Prints achieved GB/s against the ALGORITHMIC bytes of
SURVEY.md §8(d).   python tools/microbench.py [--m 256] [--s 8] [--reps 10]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from ca_lanczos_b200 import _lib, api, gallery


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=256)
    ap.add_argument("--s", type=int, default=8)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--layouts", default="selld,sell,csr")
    ap.add_argument("--chunks", default="0")
    ap.add_argument("--skip-orth", action="store_true")
    ap.add_argument("--skip-mpk", action="store_true")
    ap.add_argument("--ldpad", type=int, default=0)
    args = ap.parse_args()
    m, s = args.m, args.s
    ctx = api.default_context()
    lib = ctx.lib
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    peak = 6458.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass

    def timed(fn, reps=args.reps, warm=2):
        for _ in range(warm):
            fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    t0 = time.time()
    A = gallery.laplace3d(m)
    n, nnz = A.shape[0], A.nnz
    print("matrix %d^3: n=%d nnz=%d built in %.1fs" % (m, n, nnz, time.time() - t0), flush=True)
    spmv_bytes = 12 * nnz + 4 * (n + 1) + 16 * n
    lam = gallery.leja_points(0.0, 12.0, s)
    re = np.ascontiguousarray(lam)
    v = torch.full((n,), 1.0 / np.sqrt(n), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    out = {}
    for layout in ([] if args.skip_mpk else args.layouts.split(",")):
        t0 = time.time()
        dm = api.DeviceMatrix(A, s_max=s, layout=layout)
        print("layout %s: upload %.1fs sell_padded=%d lanes=%d" % (layout, time.time() - t0, dm.info("sell_padded_nnz"), dm.info("csr_lanes")), flush=True)
        Vp, ld = C.c_void_p(), C.c_int64()

        def mpk():
            _lib.check(lib.calz_mpk_inplace(dm.h, C.c_void_p(v.data_ptr()), s, re.ctypes.data_as(_lib.c_dp), None, 1, 0,
                                            C.byref(Vp), C.byref(ld)), ctx.h)
        for chunk in [int(c) for c in args.chunks.split(",")]:
            ctx.set_option("mpk_l2_chunk_bytes", chunk << 20)
            ms = timed(mpk)
            print("  MPK s=%d %-4s chunk=%3dMB: %8.3f ms  (%.3f ms/SpMV)  %7.1f GB/s algorithmic = %.3f of measured copy peak" %
                  (s, layout, chunk, ms, ms / s, s * spmv_bytes / ms / 1e6, s * spmv_bytes / ms / 1e6 / peak), flush=True)
            out["mpk_%s_%d" % (layout, chunk)] = ms
        ctx.set_option("mpk_l2_chunk_bytes", 0)
        dm.close()
    if args.skip_orth:
        return
    # ---- orthogonalisation kernels on n x c blocks
    ld = (n + 31) // 32 * 32 + args.ldpad
    print("orth blocks: n=%d ld=%d (pad %d)" % (n, ld, args.ldpad))
    Q = torch.randn((s + 1, ld), dtype=torch.float64, device=dev)
    X = torch.randn((s, ld), dtype=torch.float64, device=dev)
    Y = torch.empty((s + 1, ld), dtype=torch.float64, device=dev)
    Cd = torch.zeros((64 * 64,), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    tf = C.c_double()
    _lib.check(lib.calz_dmma_peak(ctx.h, C.byref(tf)), ctx.h)
    print("  fp64 DMMA peak (register-resident m8n8k4 chains): %.1f TFLOP/s" % tf.value, flush=True)
    cases = [("gram c=%d" % s, lambda: lib.calz_gram(ctx.h, n, s, X.data_ptr(), ld, s, X.data_ptr(), ld, Cd.data_ptr()), 8 * n * s),
             ("gram c=%d" % (s + 1), lambda: lib.calz_gram(ctx.h, n, s + 1, Q.data_ptr(), ld, s + 1, Q.data_ptr(), ld, Cd.data_ptr()), 8 * n * (s + 1)),
             ("coeff Q'X M=%d c=%d" % (s + 1, s), lambda: lib.calz_gram(ctx.h, n, s + 1, Q.data_ptr(), ld, s, X.data_ptr(), ld, Cd.data_ptr()), 8 * n * (2 * s + 1))]
    for name, fn, nbytes in cases:
        ms = timed(fn)
        mm, cc = (s, s) if name.startswith("gram c=%d" % s) and "c=%d" % (s + 1) not in name else ((s + 1, s + 1) if name.startswith("gram") else (s + 1, s))
        print("  %-24s %8.3f ms  %7.1f GB/s = %.3f of HBM peak; %.2f TFLOP/s = %.3f of the DMMA peak" %
              (name, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak, 2.0 * n * mm * cc / ms / 1e9, 2.0 * n * mm * cc / ms / 1e9 / tf.value), flush=True)
    R = np.zeros((s + 1, s + 1), order="F")
    info = C.c_int()
    for backend, name, c, nbytes in [("cholqr", "cholqr c=%d" % s, s, 24 * n * s), ("tsqr", "tsqr c=%d" % s, s, 16 * n * s),
                                     ("cholqr", "cholqr c=%d" % (s + 1), s + 1, 24 * n * (s + 1)), ("tsqr", "tsqr c=%d" % (s + 1), s + 1, 16 * n * (s + 1))]:
        src = X if c == s else Q
        rank = C.c_int()
        fn = lambda: _lib.check(lib.calz_normalize(ctx.h, n, c, src.data_ptr(), ld, _lib.QR[backend], 1e-8, Y.data_ptr(), ld,
                                                   R.ctypes.data_as(_lib.c_dp), C.byref(rank)), ctx.h)
        ms = timed(fn, reps=max(3, args.reps // 2))
        print("  %-24s %8.3f ms  %7.1f GB/s = %.3f of peak (compulsory bytes)" % (name, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak), flush=True)
    # full projectAndNormalize with a block that triggers pass 2: X = Q*ones + small
    Qo = torch.linalg.qr(Q[:, :n].T.contiguous())[0]            # n x (s+1) orthonormal (torch, setup only)
    Q[:, :n] = Qo.T
    X[:, :n] = (Qo @ torch.ones((s + 1, s), dtype=torch.float64, device=dev) + (1e-3 / np.sqrt(n)) * torch.randn((n, s), dtype=torch.float64, device=dev)).T
    del Qo
    torch.cuda.synchronize()
    qblk = (C.c_void_p * 1)(Q.data_ptr()); lds = (C.c_int64 * 1)(ld); mc = (C.c_int * 1)(s + 1)
    R1 = np.zeros((s + 1, s), order="F"); Rl = np.zeros((s, s), order="F")
    rp = (_lib.c_dp * 1)(R1.ctypes.data_as(_lib.c_dp))
    second, rank = C.c_int(), C.c_int()
    M, c = s + 1, s
    for backend in ("cholqr", "cholqr2", "tsqr"):
        fn = lambda: _lib.check(lib.calz_project_and_normalize(ctx.h, n, 1, qblk, lds, mc, s, X.data_ptr(), ld, 1, _lib.QR[backend],
                                                               Y.data_ptr(), ld, rp, Rl.ctypes.data_as(_lib.c_dp), C.byref(second),
                                                               C.byref(rank)), ctx.h)
        l0 = ctx.launch_count(True)
        ms = timed(fn, reps=max(3, args.reps // 2))
        nl = ctx.launch_count() / (max(3, args.reps // 2) + 2)
        alg = 8 * n * c + 2 * (8 * n * (2 * M + 3 * c) + 24 * n * c)       # norms + 2 x (project + CholQR)
        print("  pAN %-8s second=%d: %8.3f ms  %7.1f GB/s algorithmic (SURVEY 2-pass bytes) = %.3f of peak, %.0f launches" %
              (backend, second.value, ms, alg / ms / 1e6, alg / ms / 1e6 / peak, nl), flush=True)


if __name__ == "__main__":
    main()
