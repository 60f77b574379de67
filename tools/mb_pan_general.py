"""projectAndNormalize / project on block lists the single-block fast path does not take: several blocks ({Qprev, Q_conv} of
restarted_ca_lanczos.m:324-327) or one wide block ('full' re-orthogonalisation, ca_lanczos.m:193-197).  Times the tile-panel
path against the legacy kernels (option tile_panels = 1 / 0) at n = 2^24 and prints the bytes each would move at best."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np, torch
from ca_lanczos_b200 import _lib, api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
ctx = api.default_context(); lib = ctx.lib
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
ld = n + 288
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6458.7
MAXQ = 104
Q = torch.empty((MAXQ, ld), dtype=torch.float64, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
# orthonormal-ish columns: random +-1/sqrt(n) entries are orthogonal to ~1/sqrt(n)
for j in range(MAXQ):
    Q[j].copy_((torch.randint(0, 2, (ld,), device=dev, generator=g, dtype=torch.int8).to(torch.float64) * 2 - 1) / np.sqrt(n))
X = torch.empty((16, ld), dtype=torch.float64, device=dev)
Y = torch.empty((16, ld), dtype=torch.float64, device=dev)
torch.cuda.synchronize()


def timed(fn, reps=6):
    for _ in range(2):
        fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream); e1.synchronize()
    return e0.elapsed_time(e1) / reps


def fill_x(c, second):
    X[:c].normal_(generator=g)
    if second:   # large component in span(Q): the norm-drop test fires
        X[:c] += 100.0 * np.sqrt(n) * Q[:c]
    torch.cuda.synchronize()


for blocks, c, backend, second in [([9], 8, "cholqr2", False), ([17], 8, "cholqr2", False), ([25], 8, "cholqr2", False), ([49], 8, "cholqr2", False),
                                   ([97], 8, "cholqr2", False), ([49], 8, "cholqr2", True), ([7, 10], 6, "tsqr", False), ([7, 10], 6, "cholqr2", False),
                                   ([9, 12], 8, "cholqr2", True), ([33], 16, "cholqr2", False)]:
    nb = len(blocks)
    offs = np.cumsum([0] + blocks)
    qb = (C.c_void_p * nb)(*[Q.data_ptr() + 8 * ld * int(o) for o in offs[:-1]])
    lds = (C.c_int64 * nb)(*([ld] * nb)); mc = (C.c_int * nb)(*blocks)
    Rs = [np.zeros((m, c), order="F") for m in blocks]; Rl = np.zeros((c, c), order="F")
    rp = (_lib.c_dp * nb)(*[r.ctypes.data_as(_lib.c_dp) for r in Rs])
    sec, rk = C.c_int(), C.c_int()
    fill_x(c, second)
    M = sum(blocks)
    res = {}
    for opt in (1, 0):
        ctx.set_option("tile_panels", opt)
        fn = lambda: _lib.check(lib.calz_project_and_normalize(ctx.h, n, nb, qb, lds, mc, c, X.data_ptr(), ld, 1, _lib.QR[backend],
                                                               Y.data_ptr(), ld, rp, Rl.ctypes.data_as(_lib.c_dp), C.byref(sec), C.byref(rk)), ctx.h)
        res[opt] = (timed(fn), sec.value, Rl.copy(), [r.copy() for r in Rs])
    ctx.set_option("tile_panels", 1)
    # bytes if every operand crossed HBM once per phase: per sweep coefficients (M + c) + update (M + 2c); CholQR adds read + write of c
    sweeps = 2 if res[1][1] else 1
    nbytes = 8 * n * (sweeps * (2 * M + 3 * c) + 2 * c)
    dR = max(float(np.abs(res[1][2] - res[0][2]).max()), max(float(np.abs(a - b).max()) for a, b in zip(res[1][3], res[0][3])))
    scale = max(1.0, float(np.abs(res[0][2]).max()))
    print("pAN blocks=%-8s c=%2d %-7s second=%d: panels %7.3f ms (%.2f of HBM on the one-crossing bytes)  legacy %7.3f ms (%.2f)  x%.2f   |dR|/|R| = %.1e" %
          (blocks, c, backend, res[1][1], res[1][0], nbytes / res[1][0] / 1e6 / peak, res[0][0], nbytes / res[0][0] / 1e6 / peak,
           res[0][0] / res[1][0], dR / scale), flush=True)
