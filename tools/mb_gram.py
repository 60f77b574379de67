"""tsmm_tn sweep: time calz_gram for several (m, c, same) shapes at n = 2^24."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ca_lanczos_b200 import _lib, api
n = 1 << 24
ctx = api.default_context(); lib = ctx.lib
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
ld = n + 288
Q = torch.randn((64, ld), dtype=torch.float64, device=dev)
X = torch.randn((32, ld), dtype=torch.float64, device=dev)
Cd = torch.zeros((64 * 64,), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
def timed(fn, reps=10):
    for _ in range(2): fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); e1.synchronize()
    return e0.elapsed_time(e1) / reps
gm = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx.set_option("grid_mult", gm)
for m, c, same in [(8, 8, True), (8, 8, False), (9, 8, False), (16, 8, False), (16, 16, True), (9, 9, True), (24, 8, False), (32, 8, False), (64, 8, False), (17, 17, True), (1, 8, False), (8, 1, False)]:
    A = X if same else Q
    ms = timed(lambda: lib.calz_gram(ctx.h, n, m, A.data_ptr(), ld, c, X.data_ptr(), ld, Cd.data_ptr()))
    nbytes = 8 * n * (c if same else m + c)
    print("grid_mult=%d m=%2d c=%2d same=%d: %7.3f ms %7.1f GB/s  %5.2f TFLOP/s(useful)" % (gm, m, c, same, ms, nbytes / ms / 1e6, 2.0 * n * m * c / ms / 1e9), flush=True)
