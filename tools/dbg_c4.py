import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ca_lanczos_b200 import api, gallery, restart
from oracle import kernels, drivers
for n in (20000, 2000000):
    A = gallery.powerlaw_spd_rows(n, 20.0, seed=0)
    x = np.cos(0.1 * np.arange(n)) + 1.5
    y = A @ x
    for layout in ("csr", "sell", "auto"):
        dm = api.DeviceMatrix(A, s_max=6, layout=layout)
        yd = api.SpMV(dm, x)
        print(n, layout, dm.layout, "long", dm.info("n_long_rows"), "lanes", dm.info("csr_lanes"), "spmv err", np.abs(yd - y).max() / np.abs(y).max(), flush=True)
        dm.close()
n = 20000
A = gallery.powerlaw_spd_rows(n, 20.0, seed=0)
eo = drivers.restarted_ca_lanczos(A, np.ones(n), 60, 10, 6, "newton", "local", 1e-8)
print("oracle eigs", eo[0][:4], "restarts", eo[2])
eg = restart.device_restarted_ca_lanczos(A, np.ones(n), 60, 10, 6, "newton", "local", 1e-8, backend="tsqr")
print("device eigs", eg[0][:4], "restarts", eg[2], "rn", eg[3][-1].max(), "oe", eg[4][-1])
eg = restart.device_restarted_ca_lanczos(A, np.ones(n), 60, 10, 6, "newton", "local", 1e-8, backend="cholqr2")
print("device(cholqr2) eigs", eg[0][:4], "restarts", eg[2], "rn", eg[3][-1].max(), "oe", eg[4][-1])
