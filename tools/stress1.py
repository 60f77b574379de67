import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ca_lanczos_b200 as ck
from ca_lanczos_b200 import api, gallery
from oracle import drivers, kernels
g = np.load("tests/golden/c2_diag_s8_newton.npz")
A = gallery.diag_linspace(int(g["n"]), 100.0); s = int(g["s"])
Bk = np.zeros((s + 1, s)); Bk[np.arange(s), np.arange(s)] = g["shifts"]; Bk[np.arange(1, s + 1), np.arange(s)] = 1
r = np.ones(A.shape[0]); q = r / np.sqrt(r @ r)
ref = None; bad = 0
for it in range(40):
    try:
        V = ck.matrix_powers_newton(A, q, s, np.diag(Bk).copy(), 1)
        Q1, R1, rk = ck.normalize(V, backend="tsqr")
        if ref is None: ref = (V.copy(), Q1.copy(), R1.copy())
        dV = np.abs(V - ref[0]).max(); dQ = np.abs(Q1 - ref[1]).max(); dR = np.abs(R1 - ref[2]).max()
        if dV or dQ or dR or not np.all(np.isfinite(R1)):
            bad += 1; print("iter", it, "NONDETERMINISTIC dV %.3e dQ %.3e dR %.3e" % (dV, dQ, dR), "diagR", np.diag(R1))
        T, Q = drivers.ca_lanczos(A, r, s, int(g["iter"]), "newton", "local", K=ck, backend="tsqr", Bk=Bk)
    except Exception as e:
        bad += 1; print("iter", it, "EXC", repr(e)); traceback.print_exc(limit=3)
print("bad", bad, "of 40")
