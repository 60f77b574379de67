"""CPU check of the device-resident restarted driver's CONTROL FLOW (ca-lanczos_b200/restart.py): the same generic driver is run
over a numpy implementation of the block operations (built from the oracle's kernels) and compared with the oracle's own
restatement of restarted_ca_lanczos.m.  What stays unverified without a GPU is only the thin DeviceOps adapter."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ca_lanczos_b200 import gallery  # noqa: E402
from oracle import drivers, kernels  # noqa: E402

restart = importlib.import_module("ca_lanczos_b200.restart")


class NumpyOps:
    """Block operations on host arrays; a block is a Fortran-ordered 2-D array (views are numpy slices)."""

    def __init__(self, A, backend="tsqr"):
        self.A, self.n, self.backend = A, A.shape[0], backend

    def block(self, cols):
        return np.zeros((self.n, max(int(cols), 1)), order="F")[:, :cols]

    def view(self, B, j0, j1):
        return B[:, j0:j1]

    def ncols(self, B):
        return B.shape[1]

    def spmv(self, x, y):
        y[:, 0] = kernels.SpMV(self.A, x[:, 0])

    def axpy(self, Q, Cm, X, Y):
        Cm = np.asarray(Cm, dtype=np.float64).reshape(Q.shape[1], -1)
        Y[:, :] = (0.0 if X is None else X) - Q @ Cm

    def gram(self, A, B):
        return A.T @ B

    def nrm2(self, x):
        return float(np.sqrt(x[:, 0] @ x[:, 0]))

    def abs_colsum(self, out):
        out[:, 0] = np.asarray(abs(self.A).sum(axis=0)).ravel()

    def matrix_powers(self, q, s, Bk, basis):
        if basis.lower() == "monomial":
            return np.asfortranarray(np.column_stack([q[:, 0], kernels.matrix_powers_monomial(self.A, q[:, 0], s)]))
        return kernels.matrix_powers_newton(self.A, q[:, 0], s, np.diag(Bk).copy(), 1)

    def normalize(self, V, out):
        Q, R, _ = kernels.normalize(V, backend=self.backend)
        out[:, :] = Q
        return R

    def pan(self, blocks, X, out):
        QZ, RZ = kernels.projectAndNormalize([b if (b is not None and b.shape[1] > 0) else None for b in blocks], np.array(X), True,
                                             backend=self.backend)
        out[:, :] = QZ
        return RZ


def _from_host(ops, v):
    b = ops.block(1)
    b[:, 0] = v
    return b


@pytest.mark.parametrize("orth", ["local", "full"])
@pytest.mark.parametrize("basis", ["newton", "monomial"])
def test_generic_restarted_driver_matches_the_oracle_restatement(orth, basis):
    # test_restart_diagonal_matrices.m:8-36 scaled down
    N, s, nw, mx = 600, 4, 4, 40
    if basis == "monomial":
        s, mx = 3, 30                               # the monomial basis of the reference is only usable for small s
    A = gallery.diag_linspace(N, 1.0e2)
    r = np.ones(N)
    eo, Qo, nro, rno, oeo = drivers.restarted_ca_lanczos(A, r, mx, nw, s, basis, orth, 1e-8, backend="tsqr")
    ops = NumpyOps(A)
    eg, Qg, nrg, rng_, oeg, order = restart.restarted_ca_lanczos(ops, _from_host(ops, r), mx, nw, s, basis, orth, 1e-8)
    assert nrg == nro
    np.testing.assert_allclose(eg, eo, rtol=1e-10)
    np.testing.assert_allclose(eg, np.linspace(1, 100, N)[::-1][:nw], rtol=1e-8)
    Qg = Qg[:, order]
    assert Qg.shape == Qo.shape
    for j in range(Qg.shape[1]):                    # Ritz vectors up to sign
        assert min(np.linalg.norm(Qg[:, j] - Qo[:, j]), np.linalg.norm(Qg[:, j] + Qo[:, j])) < 1e-6
    assert rng_.shape == rno.shape and np.allclose(rng_[-1], rno[-1], rtol=1e-3, atol=1e-10)
    # the orthogonality loss is amplified rounding: same length, same order of magnitude (the shifts already differ in the last digits)
    assert oeg.shape == oeo.shape and np.all(np.abs(np.log10(oeg / oeo)) < 2.0) and oeg[0] == pytest.approx(oeo[0], rel=0.5)


@pytest.mark.parametrize("orth", ["periodic", "selective"])
def test_generic_restarted_driver_periodic_and_selective(orth):
    # lanczos_periodic / lanczos_selective of the restarted driver (restarted_ca_lanczos.m:369-552) on the reference's own
    # diagonal test (test_restart_diagonal_matrices.m:8-36 scaled down): same restarts, same re-orthogonalisation breaks /
    # converged-Ritz counts per cycle, eigenvalues equal to the analytic ones
    N, s, nw, mx = 600, 4, 4, 40
    A = gallery.diag_linspace(N, 1.0e2)
    r = np.ones(N)
    io, ig = {}, {}
    eo, Qo, nro, rno, oeo = drivers.restarted_ca_lanczos(A, r, mx, nw, s, "newton", orth, 1e-8, backend="tsqr", info=io)
    ops = NumpyOps(A)
    eg, Qg, nrg, rng_, oeg, order = restart.restarted_ca_lanczos(ops, _from_host(ops, r), mx, nw, s, "newton", orth, 1e-8, log=ig)
    assert nrg == nro
    assert ig["breaks"] == io["breaks"] and ig["nritz"] == io["nritz"]
    assert any(len(b) for b in io["breaks"])                       # the variant really did something on this problem
    np.testing.assert_allclose(eg, eo, rtol=1e-10)
    np.testing.assert_allclose(eg, np.linspace(1, 100, N)[::-1][:nw], rtol=1e-8)
    Qg = Qg[:, order]
    for j in range(Qg.shape[1]):
        assert min(np.linalg.norm(Qg[:, j] - Qo[:, j]), np.linalg.norm(Qg[:, j] + Qo[:, j])) < 1e-6
    assert oeg.shape == oeo.shape and np.all(np.abs(np.log10(oeg / oeo)) < 2.0)


def test_restarted_driver_rejects_unknown_orth():
    A = gallery.diag_linspace(50, 10.0)
    ops = NumpyOps(A)
    with pytest.raises(ValueError):
        restart.restarted_ca_lanczos(ops, _from_host(ops, np.ones(50)), 12, 2, 3, "newton", "sometimes", 1e-8)
    with pytest.raises(ValueError):
        drivers.restarted_ca_lanczos(A, np.ones(50), 12, 2, 3, "newton", "sometimes", 1e-8)


def test_generic_normest_and_shifts_match_the_oracle():
    A = gallery.poisson2d(20)
    ops = NumpyOps(A)
    assert restart.normest(ops) == pytest.approx(drivers.normest(A), rel=1e-12)
    q = _from_host(ops, np.ones(A.shape[0]) / np.sqrt(A.shape[0]))
    Bk = restart.basis_matrix(ops, q, 4, "newton")
    Bo = drivers.basis_matrix(A, q[:, 0], 4, "newton", "local")
    np.testing.assert_allclose(Bk, Bo, rtol=1e-9, atol=1e-12)
