"""BASELINE-size parity against the oracle, per block and from IDENTICAL inputs: one first block and
one steady-state block of C3 (256^3 7-point Laplacian, n = 16 777 216, s = 8 Newton) and of C2 (diagonal, n = 10^6, s = 8
Newton/Leja) through the C ABI with device-resident data, compared with oracle/kernels.py:

    basis vectors V       2-norm relative per column  < 1e-13   (north_star: 1e-10)
    R / coefficient blocks                  relative  < 1e-10
    QZ                    2-norm per column           < 1e-10

The oracle block at 256^3 costs ~1 minute of host time (scipy CSR mat-vec, LAPACK QR of a 16.7 M x 9 block).
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from ca_lanczos_b200 import _lib, api, gallery  # noqa: E402
from oracle import kernels  # noqa: E402


def relcols(A, B):
    return float(np.max(np.linalg.norm(A - B, axis=0) / np.maximum(np.linalg.norm(B, axis=0), 1e-300)))


def _device_mpk(ctx, dm, q, s, shifts):
    """matrix_powers_newton from the host vector q through the zero-copy device entry point; returns the basis on the host."""
    import torch
    n = dm.n
    ld = (n + 31) // 32 * 32
    dev = torch.device("cuda", ctx.device)
    qd = torch.as_tensor(q, device=dev)
    torch.cuda.synchronize(dev)
    re = np.ascontiguousarray(shifts, dtype=np.float64)
    Vp, ldV = C.c_void_p(), C.c_int64()
    _lib.check(ctx.lib.calz_mpk_inplace(dm.h, C.c_void_p(qd.data_ptr()), s, re.ctypes.data_as(_lib.c_dp), None, 1, 0, C.byref(Vp),
                                        C.byref(ldV)), ctx.h)
    ctx.sync()
    # the basis as a torch view of libcalz' workspace (owned rows, leading dimension ldV)
    V = np.empty((n, s + 1), order="F")
    tmp = torch.empty(n, dtype=torch.float64, device=dev)
    for j in range(s + 1):
        _lib.check(ctx.lib.calz_block_axpy(ctx.h, n, 1, C.c_void_p(Vp.value + 8 * ldV.value * j), ldV.value, 1,
                                           np.array([-1.0]).ctypes.data_as(_lib.c_dp), None, ld, C.c_void_p(tmp.data_ptr()), ld), ctx.h)
        ctx.sync()
        V[:, j] = tmp.cpu().numpy()
    return V


def _device_pan(ctx, n, s, Qprev, X, backend):
    import torch
    ld = (n + 31) // 32 * 32
    dev = torch.device("cuda", ctx.device)
    M = Qprev.shape[1]
    Qd = torch.zeros((M, ld), dtype=torch.float64, device=dev)
    Xd = torch.zeros((s, ld), dtype=torch.float64, device=dev)
    Zd = torch.zeros((s, ld), dtype=torch.float64, device=dev)
    for j in range(M):
        Qd[j, :n] = torch.as_tensor(np.ascontiguousarray(Qprev[:, j]), device=dev)
    for j in range(s):
        Xd[j, :n] = torch.as_tensor(np.ascontiguousarray(X[:, j]), device=dev)
    torch.cuda.synchronize(dev)
    qb = (C.c_void_p * 1)(Qd.data_ptr()); lds = (C.c_int64 * 1)(ld); mc = (C.c_int * 1)(M)
    R1 = np.zeros((M, s), order="F"); Rl = np.zeros((s, s), order="F")
    rp = (_lib.c_dp * 1)(R1.ctypes.data_as(_lib.c_dp))
    sec, rk = C.c_int(), C.c_int()
    _lib.check(ctx.lib.calz_project_and_normalize(ctx.h, n, 1, qb, lds, mc, s, C.c_void_p(Xd.data_ptr()), ld, 1, _lib.QR[backend],
                                                  C.c_void_p(Zd.data_ptr()), ld, rp, Rl.ctypes.data_as(_lib.c_dp), C.byref(sec),
                                                  C.byref(rk)), ctx.h)
    ctx.sync()
    QZ = np.empty((n, s), order="F")
    for j in range(s):
        QZ[:, j] = Zd[j, :n].cpu().numpy()
    return QZ, R1, Rl, bool(sec.value)


def _one_block_parity(A, s, shifts, backends):
    n = A.shape[0]
    ctx = api.default_context()
    dm = api.DeviceMatrix(A, s_max=s, ctx=ctx)
    r = np.ones(n)
    q0 = r / np.sqrt(r @ r)                                               # ca_lanczos.m:55
    # ---- oracle: first block (normalize) and one steady-state block
    V1o = kernels.matrix_powers_newton(A, q0, s, shifts, 1)
    Q1o, R1o, _ = kernels.normalize(V1o, backend="tsqr")
    q1 = np.ascontiguousarray(Q1o[:, s])
    V2o = kernels.matrix_powers_newton(A, q1, s, shifts, 1)
    # ---- device MPK from the same start vectors
    V1 = _device_mpk(ctx, dm, q0, s, shifts)
    assert relcols(V1, V1o) < 1e-13
    V2 = _device_mpk(ctx, dm, q1, s, shifts)
    assert relcols(V2, V2o) < 1e-13
    del V1, V2, V1o
    # ---- device projectAndNormalize from the oracle's Qprev and basis block
    for backend in backends:
        info = {}
        QZo, RZo = kernels.projectAndNormalize([Q1o], V2o[:, 1:], True, backend="tsqr" if backend == "tsqr" else "cholqr", info=info)
        QZ, R1, Rl, second = _device_pan(ctx, n, s, Q1o, V2o[:, 1:], backend)
        assert second == info["second_pass"]
        assert np.linalg.norm(R1 - RZo[0]) <= 1e-10 * np.linalg.norm(RZo[0])
        assert np.linalg.norm(Rl - RZo[1]) <= 1e-10 * np.linalg.norm(RZo[1])
        assert float(np.max(np.linalg.norm(QZ - QZo, axis=0))) < 1e-10
        del QZ, QZo
    dm.close()


def test_c3_full_size_block_against_oracle():
    s = 8
    A = gallery.laplace3d(256)
    assert A.shape[0] == 16_777_216 and A.nnz == 117_047_296
    _one_block_parity(A, s, gallery.leja_points(0.0, 12.0, s), ["cholqr2", "tsqr"])


def test_c2_full_size_block_against_oracle():
    s = 8
    A = gallery.diag_linspace(1_000_000, 100.0)
    lam = np.array([99.5677, 1.4323, 45.7883, 81.1407, 19.8593, 64.4648, 7.5729, 93.4271])
    _one_block_parity(A, s, lam, ["cholqr2", "tsqr"])
