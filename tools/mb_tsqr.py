"""TSQR vs CholQR on one B200: n x c blocks (C5-style input: counter-based U(-1,1), column j scaled by 2^-j), CUDA events on
libcalz' stream, fraction of the measured copy peak on the COMPULSORY bytes (TSQR 16nc, CholQR 24nc).
    python tools/mb_tsqr.py [--n 16777216] [--cs 5,8,9,13,17] [--reps 5]"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from ca_lanczos_b200 import _lib, api, gallery


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16777216)
    ap.add_argument("--cs", default="5,8,9,13,17")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    ctx = api.default_context()
    lib = ctx.lib
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    peak = 6458.7
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    n = args.n
    ld = (n + 31) // 32 * 32
    # generator check against the numpy one
    Xs = gallery.tall_skinny_device(1000, 5, dev, 1024)[:, :1000].T.cpu().numpy()
    assert np.array_equal(Xs, gallery.tall_skinny(1000, 5)), "device generator differs from gallery.tall_skinny"
    for c in [int(x) for x in args.cs.split(",")]:
        X = gallery.tall_skinny_device(n, c, dev, ld)
        Q = torch.empty((c, ld), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        R = {}
        for backend, nbytes in (("tsqr", 16 * n * c), ("cholqr", 24 * n * c)):
            Rh = np.zeros((c, c), order="F")
            rank = C.c_int()

            def fn():
                _lib.check(lib.calz_normalize(ctx.h, n, c, X.data_ptr(), ld, _lib.QR[backend], 1e-8, Q.data_ptr(), ld,
                                              Rh.ctypes.data_as(_lib.c_dp), C.byref(rank)), ctx.h)
            fn(); fn()
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.reps):
                fn()
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            from ca_lanczos_b200 import solver
            oe = solver.orth_error(ctx, n, [(Q.data_ptr(), ld, c)], "fro")
            R[backend] = Rh.copy()
            print("  n=%d c=%2d %-7s %8.3f ms  %7.1f GB/s on the compulsory bytes = %.3f of the copy peak   ||I-Q'Q||_F = %.2e" %
                  (n, c, backend, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak, oe), flush=True)
        print("  n=%d c=%2d |R_tsqr - R_cholqr|/|R| = %.2e" % (n, c, np.linalg.norm(R["tsqr"] - R["cholqr"]) / np.linalg.norm(R["tsqr"])), flush=True)
        del X, Q


if __name__ == "__main__":
    main()
