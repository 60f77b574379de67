"""Single-GPU diagnostic: block pipeline on the (scaled) power-law matrix, all back ends, fused and unfused last pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ca_lanczos_b200 import api, gallery
from ca_lanczos_b200.engine import BlockEngine
from oracle import drivers

A = gallery.powerlaw_spd(20000, 8.0, seed=0)
n = A.shape[0]
r = np.cos(0.61 * np.arange(n) ** 1.5) + 0.3 * np.sin(1.7 * np.arange(n))
s, blocks = 4, 4
io = {}
To, Qo = drivers.ca_lanczos(A, r, s, s * blocks, "newton", "local", info=io)
shifts = np.diag(io["Bk"])[:s].copy()
ctx = api.default_context()
q0 = r / np.sqrt(r @ r)
for layout in ("auto", "csr"):
    dm = api.DeviceMatrix(A, s_max=s, layout=layout, ctx=ctx)
    for backend, fused in (("cholqr2", 1), ("cholqr2", 0), ("cholqr", 1), ("tsqr", 1)):
        ctx.set_option("pan_fused_solve", fused)
        try:
            eng = BlockEngine(dm, s, blocks + 1, "newton", shifts, backend)
            eng.first_block(q0)
            for _ in range(blocks - 1):
                eng.next_block()
            T = eng.T_matrix()
            Q = eng.Q_host()
            print("layout=%s(%s) %s fused=%d: T err %.2e, Q err %.2e, second %s" % (layout, dm.layout, backend, fused,
                  np.abs(T - To).max() / np.abs(To).max(), np.max(np.linalg.norm(Q - Qo[:, :Q.shape[1]], axis=0)), eng.second), flush=True)
        except Exception as e:
            print("layout=%s(%s) %s fused=%d: FAILED %s" % (layout, dm.layout, backend, fused, e), flush=True)
        ctx.set_option("pan_fused_solve", 1)
    dm.close()
