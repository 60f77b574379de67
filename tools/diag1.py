import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ca_lanczos_b200 as ck
from ca_lanczos_b200 import api, gallery
from ca_lanczos_b200.engine import BlockEngine
from oracle import drivers, kernels
def ritz(T): return np.sort(np.linalg.eig(T)[0].real)[::-1]
# 6: zero column tsqr
X = gallery.tall_skinny(500, 3, seed=2); X[:, 1] = 0.0
Q, R = api.tsqr(X); Qo, Ro = kernels.tsqr(X)
print("tsqr zero col R\n", R, "\nRo\n", Ro)
# 3/4 engine
A = gallery.laplace3d(24, 24, 24); n = A.shape[0]; s, nblk = 8, 6
r = np.ones(n); io = {}
To, Qo = drivers.ca_lanczos(A, r, s, s * nblk, "newton", "local", info=io)
shifts = np.diag(io["Bk"]).copy()
for backend in ("tsqr", "cholqr"):
    ig = {}
    Th, Qh = drivers.ca_lanczos(A, r, s, s * nblk, "newton", "local", K=ck, backend=backend, Bk=io["Bk"], info=ig)
    dm = api.DeviceMatrix(A, s_max=s)
    eng = BlockEngine(dm, s, nblk, "newton", shifts, backend)
    eng.first_block(r / np.sqrt(r @ r))
    for _ in range(nblk - 1): eng.next_block()
    T = eng.T_matrix(); Qe = eng.Q_host()
    sc = np.abs(To).max()
    print(backend, "host-driver vs oracle T %.2e" % (np.abs(Th - To).max() / sc), "engine vs oracle %.2e" % (np.abs(T - To).max() / sc),
          "engine vs host-driver %.2e" % (np.abs(T - Th).max() / sc), "ritz e-o %.2e" % np.max(np.abs(ritz(T)[:4] - ritz(To)[:4]) / ritz(To)[:4]),
          "Q1 block diff %.2e" % np.max(np.linalg.norm(Qe[:, :s + 1] - Qo[:, :s + 1], axis=0)),
          "orth e %.2e o %.2e" % (np.linalg.norm(np.eye(Qe.shape[1]) - Qe.T @ Qe), np.linalg.norm(np.eye(Qo.shape[1]) - Qo.T @ Qo)),
          "second", eng.second, [i["second_pass"] for i in io["pan"]])
# start vector not ones: generic
rng = np.random.default_rng(0); r2 = rng.standard_normal(n)
io = {}
To, Qo = drivers.ca_lanczos(A, r2, s, s * nblk, "newton", "local", info=io)
for backend in ("tsqr", "cholqr"):
    Th, Qh = drivers.ca_lanczos(A, r2, s, s * nblk, "newton", "local", K=ck, backend=backend, Bk=io["Bk"])
    print("random start", backend, "T %.2e" % (np.abs(Th - To).max() / np.abs(To).max()), "Q1 %.2e" % np.max(np.linalg.norm(Qh[:, :s + 1] - Qo[:, :s + 1], axis=0)),
          "asym To %.2e" % (np.abs(To - To.T).max() / np.abs(To).max()))
