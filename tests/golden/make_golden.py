"""Generate the golden fixtures under tests/golden/ FROM THE ORACLE (numpy/scipy restatement).

The reference is MATLAB and cannot run here (no Octave/MATLAB), and it ships no goldens of its own,
so these files pin the ORACLE (regression) and give the GPU tests fixed targets; they are not outputs of the
reference.  tools/dump_goldens.m writes the same quantities from the real reference on a machine with
Octave/MATLAB.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ca_lanczos_b200 import gallery  # noqa: E402
from oracle import drivers, kernels  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def dump(name, A, s, iters, basis, **extra):
    r = np.ones(A.shape[0])
    info = {}
    T, Q = drivers.ca_lanczos(A, r, s, iters, basis, "local", info=info)
    ritz = np.sort(np.linalg.eig(T)[0].real)[::-1]
    q = r / np.sqrt(r @ r)
    shifts = np.diag(info["Bk"]).copy()
    if basis == "newton":
        V = kernels.matrix_powers_newton(A, q, s, shifts, 1)
    else:
        V = np.column_stack([q, kernels.matrix_powers_monomial(A, q, s)])
    rows = np.unique(np.linspace(0, A.shape[0] - 1, 257).astype(np.int64))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), T=T, ritz=ritz, shifts=shifts, rows=rows, V_rows=V[rows],
                        Q_rows=Q[rows][:, : 2 * s + 1], s=s, iter=iters, basis=basis,
                        second=np.array([i["second_pass"] for i in info["pan"]]), **extra)
    print(name, "ritz[:3] =", ritz[:3], "orth =", np.linalg.norm(np.eye(Q.shape[1]) - Q.T @ Q, "fro"))


if __name__ == "__main__":
    dump("c1_poisson_s4_monomial", gallery.poisson2d(100), 4, 60, "monomial", m=100)
    dump("c2_diag_s8_newton", gallery.diag_linspace(20000, 100.0), 8, 64, "newton", n=20000)
