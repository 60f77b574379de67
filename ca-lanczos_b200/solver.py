"""Device-resident drivers on top of the hot path: the CALLERS of the reference, so that
a solve never moves an n-vector over PCIe.

    T, Q = ca_lanczos(A, r, s, iter, basis, orth)          mirrors ca_lanczos.m:24 ('local' and 'full')

Everything O(n) runs in libcalz kernels through the C ABI (SpMV, project, cholqr, block_axpy, the BlockEngine pipeline);
torch only owns the device buffers.  The O(s^2) / O(s^3) host algebra follows the reference: the 2s-step Lanczos that
produces the Newton shifts (ca_lanczos.m:66-72 -> lanczos.m:85-134), the modified Leja ordering with the running
capacity rescale (modified_leja.m:24-196, real shifts), newton_basis_matrix.m:13-60, and the T assembly
(ca_lanczos.m:200-223, in engine.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .api import Context, DeviceMatrix, default_context
from .engine import BlockEngine


# ----------------------------------------------------------------------------------------------- Leja / Newton basis
def leja_order(x):
    """Modified Leja ordering of REAL shifts, as leja.m:28-29 -> real_leja.m -> modified_leja.m computes it for the real
    spectra of symmetric matrices: start from the point of largest modulus, then repeatedly take the point maximising
    prod_j |x - y_j| / capacity, with the running capacity estimate and the rescale/unscale of the points
    (modified_leja.m:95-117,:192).  Repeated shifts are an error there (x(1:n) overruns, real_leja.m:87) and here."""
    x = np.sort(np.asarray(x, dtype=np.float64).ravel())
    n = x.size
    if n == 0:
        return x
    if np.any(np.diff(x) == 0):
        raise IndexError("repeated shifts: the reference's real_leja/modified_leja indexes past the unique values")
    x = x.copy()
    out = [int(np.argmax(np.abs(x)))]
    y = [x[out[0]]]
    rest = [i for i in range(n) if i != out[0]]
    capacity = 1.0
    first = True
    while rest:
        npts = len(out)
        if not first and npts > 1:
            old = capacity
            capacity = 1.0
            for i in out[: npts - 1]:
                capacity *= abs(y[npts - 1] - x[i]) ** (1.0 / npts)
            scale = capacity / old
            x = x / scale
            y = [v / scale for v in y]
        first = False
        best, best_val = None, -1.0
        for j in rest:
            p = 1.0
            for i in out:
                p *= abs(x[j] - x[i]) / capacity
            if p > best_val:
                best, best_val = j, p
        if best_val == 0.0 or np.isinf(best_val):
            raise ValueError("Leja product under/overflowed")
        rest.remove(best)
        out.append(best)
        y.append(x[best])
    return np.asarray(y) * capacity


def newton_basis_matrix(shifts, s):
    """newton_basis_matrix.m:13-60 for real shifts: B(k,k) = shift_k, B(k+1,k) = 1."""
    B = np.zeros((s + 1, s))
    for k in range(s):
        B[k, k] = shifts[k]
        B[k + 1, k] = 1.0
    return B


# ----------------------------------------------------------------------------------------------- device helpers
class _Dev:
    """Thin wrappers: device pointers in, libcalz calls out (all on the context's stream)."""

    def __init__(self, dm: DeviceMatrix):
        import torch
        self.torch, self.dm, self.ctx, self.lib = torch, dm, dm.ctx, dm.ctx.lib
        self.n = dm.n
        self.ld = (self.n + 31) // 32 * 32
        self.dev = torch.device("cuda", self.ctx.device)

    def zeros(self, cols):
        t = self.torch.zeros((cols, self.ld), dtype=self.torch.float64, device=self.dev)
        self.torch.cuda.synchronize(self.dev)
        return t

    def col(self, t, j):
        return t.data_ptr() + 8 * self.ld * int(j)

    def spmv(self, x_ptr, y_ptr):
        check(self.lib.calz_spmv(self.dm.h, C.c_void_p(x_ptr), C.c_void_p(y_ptr)), self.ctx.h)

    def axpy(self, q_ptr, m, coeff, x_ptr, y_ptr):
        """y = x - Q*coeff  (Q n x m at q_ptr, coeff length m; one column)."""
        cf = np.ascontiguousarray(coeff, dtype=np.float64).reshape(-1)
        check(self.lib.calz_block_axpy(self.ctx.h, self.n, int(m), C.c_void_p(q_ptr), self.ld, 1, cf.ctypes.data_as(_lib.c_dp),
                                       C.c_void_p(x_ptr) if x_ptr else None, self.ld, C.c_void_p(y_ptr), self.ld), self.ctx.h)

    def project(self, q_ptr, m, x_ptr):
        """coefficients = Q'x ; x -= Q*coefficients   (project.m:34-35 on one block); returns the coefficients."""
        R = np.zeros((m, 1), order="F")
        qb = (C.c_void_p * 1)(q_ptr); lds = (C.c_int64 * 1)(self.ld); mc = (C.c_int * 1)(int(m))
        rp = (_lib.c_dp * 1)(R.ctypes.data_as(_lib.c_dp))
        check(self.lib.calz_project(self.ctx.h, self.n, 1, qb, lds, mc, 1, C.c_void_p(x_ptr), self.ld, 0, rp), self.ctx.h)
        return R[:, 0]

    def normalize_col(self, x_ptr, q_ptr):
        """q = x / sqrt(x'x); returns the norm  (cholqr.m on a single column == lanczos.m:109-110)."""
        R = np.zeros((1, 1), order="F")
        info = C.c_int()
        check(self.lib.calz_cholqr(self.ctx.h, self.n, 1, C.c_void_p(x_ptr), self.ld, C.c_void_p(q_ptr), self.ld,
                                   R.ctypes.data_as(_lib.c_dp), C.byref(info)), self.ctx.h)
        return float(R[0, 0])


def lanczos_T(dev: _Dev, q0_ptr, maxiter: int, orth: str = "local"):
    """lanczos.m:85-134 (lanczos_basic, 'local' or 'fro'): the tridiagonal T of `maxiter` steps, vectors stay on the GPU."""
    Qs = dev.zeros(maxiter + 2)
    r = dev.col(Qs, maxiter + 1)
    dev.torch.cuda.synchronize(dev.dev)
    check(dev.lib.calz_block_axpy(dev.ctx.h, dev.n, 1, C.c_void_p(q0_ptr), dev.ld, 1, np.array([-1.0]).ctypes.data_as(_lib.c_dp),
                                  None, dev.ld, C.c_void_p(dev.col(Qs, 0)), dev.ld), dev.ctx.h)        # Qs(:,1) = q0
    alpha = np.zeros(maxiter); beta = np.zeros(maxiter)
    for j in range(maxiter):
        qj = dev.col(Qs, j)
        dev.spmv(qj, r)                                                  # r = A*Q(:,j)                      (:103)
        if j > 0:
            dev.axpy(dev.col(Qs, j - 1), 1, [beta[j - 1]], r, r)         # r = r - beta(j-1)*Q(:,j-1)       (:105)
        alpha[j] = dev.project(qj, 1, r)[0]                              # alpha = r'q ; r = r - alpha*q    (:107-108)
        beta[j] = dev.normalize_col(r, dev.col(Qs, j + 1))               # beta = sqrt(r'r); q+ = r/beta    (:109-110)
        if orth in ("full", "fro"):
            dev.project(dev.col(Qs, 0), j + 1, dev.col(Qs, j + 1))       # one CGS sweep, no renormalisation (:62-66)
    return np.diag(alpha) + np.diag(beta[: maxiter - 1], 1) + np.diag(beta[: maxiter - 1], -1)


def newton_shifts(dm: DeviceMatrix, q0_ptr, s: int, orth: str = "full"):
    """ca_lanczos.m:66-71: 2s-step Lanczos, eig, Leja order; returns all 2s ordered shifts (the first s are used)."""
    dev = _Dev(dm)
    T = lanczos_T(dev, q0_ptr, 2 * s, orth)
    return leja_order(np.linalg.eigvalsh(T))


# ----------------------------------------------------------------------------------------------- ca_lanczos
def ca_lanczos(A, r, s, iter, basis="newton", orth="local", backend="cholqr2", ctx: Context | None = None, shifts=None,
               return_engine=False, norm_A=None):
    """ca_lanczos.m:24-86 with ca_lanczos_basic (:150-245), device resident.  ``A``: scipy sparse or DeviceMatrix;
    ``r``: host start vector (owned rows).  Returns (T, Q) like the reference (Q as a host array), or the engine."""
    orth = str(orth).lower()
    if orth not in ("local", "full", "periodic", "selective"):
        raise ValueError("ERROR: Unknown orth type: " + orth)
    if basis.lower() not in ("monomial", "newton"):
        raise ValueError("ERROR: Unknown basis type: " + basis)
    s = int(s)
    t = int(np.ceil(iter / s))                                                        # :52
    dm = A if isinstance(A, DeviceMatrix) else DeviceMatrix(A, s_max=max(s, 1), ctx=ctx or default_context())
    dev = _Dev(dm)
    import torch
    rq = torch.as_tensor(np.ascontiguousarray(r, dtype=np.float64), device=dev.dev)
    buf = dev.zeros(2)
    buf[0, : dev.n] = rq
    torch.cuda.synchronize(dev.dev)
    dev.normalize_col(dev.col(buf, 0), dev.col(buf, 1))                               # q = r/sqrt(r'*r)   (:55)
    q_ptr = dev.col(buf, 1)
    if basis.lower() == "newton" and shifts is None:
        shifts = newton_shifts(dm, q_ptr, s, "full")                                  # :68-70
    eng = BlockEngine(dm, s, t + 1, basis, shifts, backend)
    eng.first_block(None, q_ptr=q_ptr)
    if orth == "local":
        eng.run_blocks(t - 1)
    elif orth == "full":
        for _ in range(t - 1):
            eng.next_block(full_reorth=True)
    else:
        _periodic_selective(eng, t, orth, A_host=None if isinstance(A, DeviceMatrix) else A, norm_A=norm_A)
    if return_engine:
        return eng
    return eng.T_matrix(), eng.Q_host()


# ----------------------------------------------------------------------------------------------- periodic / selective (N3)
_EPS = float(np.finfo(np.float64).eps)


def update_omega(omega_in, alpha, beta, anorm, s):
    """ca_lanczos.m:469-539: the omega recurrence (estimated inner products between Lanczos vectors) extended by s rows;
    ``alpha``/``beta`` = diag(T,0) / diag(T,-1).  Host algebra, O((sk)^2) per block; 1-based work arrays as in the reference."""
    n = len(alpha)
    al = np.concatenate([[0.0], np.asarray(alpha, dtype=np.float64)])
    be = np.concatenate([[0.0], np.asarray(beta, dtype=np.float64)])
    Tr = _EPS * anorm
    if omega_in is None or np.size(omega_in) == 0:
        om = np.zeros((s + 2, s + 2))
        om[1, 1] = 1.0
        om[2, 1] = Tr / be[1]; om[2, 2] = 1.0
        rows = range(2, s + 1)
    else:
        m = omega_in.shape[0] - 1
        om = np.zeros((n + 2, n + 2))
        om[1:m + 2, 1:m + 2] = omega_in
        rows = range(m + 1, m + s + 1)
    for j in rows:
        binv = 1.0 / be[j]
        w = be[2] * om[j, 2] + (al[1] - al[j]) * om[j, 1] - be[j] * om[j - 1, 1]
        om[j + 1, 1] = binv * (w + Tr) if w > 0 else binv * (w - Tr)
        for k in range(2, j):
            w = be[k + 1] * om[j, k + 1] + (al[k] - al[j]) * om[j, k] + be[k] * om[j, k - 1] - be[j] * om[j - 1, k]
            om[j + 1, k] = binv * (w + Tr) if w > 0 else binv * (w - Tr)
        om[j + 1, j] = binv * Tr
        om[j + 1, j + 1] = 1.0
    return om[1:, 1:].copy()


def reset_omega(omega_in, anorm, s):
    """ca_lanczos.m:541-551: after a re-orthogonalisation the last s rows of omega drop back to the round-off level."""
    Tr = _EPS * anorm
    m = omega_in.shape[0] - s - 1
    om = omega_in.copy()
    for j in range(m + 1, m + s + 1):          # 1-based row j+1 -> 0-based row j
        om[j, :j] = Tr
        om[j, j] = 1.0
    return om


def _periodic_selective(eng: BlockEngine, t: int, orth: str, A_host=None, norm_A=None):
    """ca_lanczos_periodic (ca_lanczos.m:362-467) / ca_lanczos_selective (:248-359) around the device-resident block engine:
    the omega recurrence / the Ritz convergence test run on the host from T; the re-orthogonalisations
    (projectAndNormalize against ALL previous vectors, :449; against the converged Ritz vectors, :286), the Ritz vector assembly
    (:336-337) and normest (:372) run on the device."""
    import torch
    from . import restart
    dm, ctx, lib, s = eng.dm, eng.ctx, eng.lib, eng.s
    ops = restart.DeviceOps(dm, backend=eng.backend, A_host=A_host)
    if norm_A is None:
        norm_A = restart.normest(ops)
    eng.nbreaks, eng.breaks, eng.norm_A = 0, [], norm_A
    omega = None
    tmp = torch.zeros((s + 1, eng.ld), dtype=torch.float64, device=eng.Q.device)
    QR, nritz = None, 0
    torch.cuda.synchronize(eng.Q.device)
    for k in range(1, t + 1):
        if k > 1:
            extra = [(QR.data_ptr(), eng.ld, nritz)] if (orth == "selective" and nritz > 0) else []
            eng.next_block(extra_blocks=extra)
        T = eng.T                                                            # (s*k+1) x (s*k)
        if orth == "periodic":
            omega = update_omega(omega, np.diag(T, 0), np.diag(T, -1), norm_A, s)          # :433-436
            err = 0.0
            for i in range(1, s + 1):
                row = omega[(k - 1) * s + i, : (k - 1) * s + i]
                err = max(err, float(np.max(np.abs(row))) if row.size else 0.0)
            if err >= np.sqrt(_EPS):                                         # :447-451
                eng.nbreaks += 1
                eng.breaks.append(k)
                mold = (k - 1) * s
                qb = (C.c_void_p * 1)(eng._qcol(0) if mold > 0 else None)
                lds = (C.c_int64 * 1)(eng.ld)
                mc = (C.c_int * 1)(mold)
                s2, r2 = C.c_int(), C.c_int()
                check(lib.calz_project_and_normalize(ctx.h, eng.n, 1, qb, lds, mc, s + 1, C.c_void_p(eng._qcol(mold)), eng.ld, 1,
                                                     _lib.QR[eng.backend], C.c_void_p(tmp.data_ptr()), eng.ld, None, None,
                                                     C.byref(s2), C.byref(r2)), ctx.h)
                ctx.sync()
                eng.Q[mold:mold + s + 1] = tmp                               # torch's stream; libcalz' stream is idle (synchronised)
                torch.cuda.synchronize(eng.Q.device)
                omega = reset_omega(omega, norm_A, s)
        else:
            m = s * k
            Dp, Vp = np.linalg.eig(T[:m, :m])                                # :322
            Vp = np.real(Vp)
            conv = eng.b[k - 1] * np.abs(Vp[m - 1, :]) < norm_A * np.sqrt(_EPS)
            if int(conv.sum()) > nritz:                                      # :330-341
                eng.nbreaks += 1
                eng.breaks.append(k)
                nritz = int(conv.sum())
                QRn = torch.zeros((nritz, eng.ld), dtype=torch.float64, device=eng.Q.device)
                QRo = torch.zeros((nritz, eng.ld), dtype=torch.float64, device=eng.Q.device)
                torch.cuda.synchronize(eng.Q.device)
                cols = np.flatnonzero(conv)
                for c0 in range(0, nritz, 16):                               # y = Q(:,1:k*s)*Vp(:,i): tall-skinny GEMM, 16 columns a call
                    cc = min(16, nritz - c0)
                    cf = np.asfortranarray(-Vp[:, cols[c0:c0 + cc]])
                    check(lib.calz_block_axpy(ctx.h, eng.n, m, C.c_void_p(eng._qcol(0)), eng.ld, cc, cf.ctypes.data_as(_lib.c_dp), None,
                                              eng.ld, C.c_void_p(QRn.data_ptr() + 8 * eng.ld * c0), eng.ld), ctx.h)
                # QR(:,1:nritz) = normalize(QR(:,1:nritz))  (:340), 32 columns at most per QR call: wider sets are orthonormalised
                # block by block (project against the finished ones, then normalize)
                done = 0
                while done < nritz:
                    cc = min(32, nritz - done)
                    src = QRn.data_ptr() + 8 * eng.ld * done
                    dst = QRo.data_ptr() + 8 * eng.ld * done
                    Rr = np.zeros((cc, cc), order="F")
                    rk = C.c_int()
                    if done == 0:
                        check(lib.calz_normalize(ctx.h, eng.n, cc, C.c_void_p(src), eng.ld, _lib.QR[eng.backend], 1e-8, C.c_void_p(dst),
                                                 eng.ld, Rr.ctypes.data_as(_lib.c_dp), C.byref(rk)), ctx.h)
                    else:
                        qb = (C.c_void_p * 1)(QRo.data_ptr()); lds = (C.c_int64 * 1)(eng.ld); mc = (C.c_int * 1)(done)
                        s2 = C.c_int()
                        check(lib.calz_project_and_normalize(ctx.h, eng.n, 1, qb, lds, mc, cc, C.c_void_p(src), eng.ld, 1,
                                                             _lib.QR[eng.backend], C.c_void_p(dst), eng.ld, None, None, C.byref(s2),
                                                             C.byref(rk)), ctx.h)
                    done += cc
                ctx.sync()
                QR = QRo
    eng.nritz = nritz


# ----------------------------------------------------------------------------------------------- orthogonality (N2)
def orth_error(ctx: Context, n: int, blocks, mode: str = "fro", s: int = 0) -> float:
    """Loss of orthogonality of the device-resident basis [Q_1 ... Q_k] (``blocks``: (device pointer, ld, columns) triples):
    ``mode='fro'``  norm(eye - Q'*Q,'fro') as restarted_ca_lanczos.m:165-168; ``mode='lastblock'`` compute_orth_err(Q,s) of
    ca_lanczos.m:99-107 (max |Q_old' * Q_lastblock|).  One pass of DMMA Gram products (calz_orth_error), all-reduced."""
    nb = len(blocks)
    qb = (C.c_void_p * nb)(*[int(b[0]) for b in blocks])
    lds = (C.c_int64 * nb)(*[int(b[1]) for b in blocks])
    mc = (C.c_int * nb)(*[int(b[2]) for b in blocks])
    err = C.c_double()
    check(ctx.lib.calz_orth_error(ctx.h, int(n), nb, qb, lds, mc, {"fro": 0, "lastblock": 1}[mode], int(s), C.byref(err)), ctx.h)
    return float(err.value)


def engine_orth_errors(eng: BlockEngine):
    """(compute_orth_err of ca_lanczos.m:99-107, ||I - Q'Q||_F) of everything the engine has produced so far."""
    cols = eng.s * eng.k + 1
    blk = [(eng._qcol(0), eng.ld, cols)]
    return (orth_error(eng.ctx, eng.n, blk, "lastblock", eng.s), orth_error(eng.ctx, eng.n, blk, "fro"))


# ----------------------------------------------------------------------------------------------- Ritz pairs (N1, N2)
def ritz_residuals(eng: BlockEngine, nev: int | None = None):
    """compute_ritz_rnorm of ca_lanczos.m:88-97 on the device: Ritz values of T (general eig, T is not exactly
    symmetric), Ritz vectors x = Q*Vp(:,i) (tall-skinny GEMM, N1) and relative residuals ||A x - l x|| / ||l x|| (N2),
    sorted by descending Ritz value.  Returns (values, residuals)."""
    T = eng.T_matrix()
    m = T.shape[0]
    lam, Vp = np.linalg.eig(T)
    lam, Vp = np.real(lam), np.real(Vp)
    order = np.argsort(-lam, kind="stable")
    nev = m if nev is None else min(int(nev), m)
    dev = _Dev(eng.dm)
    buf = dev.zeros(4)
    x, ax, rr, tmp = (dev.col(buf, j) for j in range(4))
    vals, res = [], []
    for i in order[:nev]:
        l = float(lam[i])
        cf = np.ascontiguousarray(-Vp[:, i])
        check(dev.lib.calz_block_axpy(dev.ctx.h, dev.n, m, C.c_void_p(eng._qcol(0)), eng.ld, 1, cf.ctypes.data_as(_lib.c_dp),
                                      None, dev.ld, C.c_void_p(x), dev.ld), dev.ctx.h)          # x = Q*Vp(:,i)
        dev.spmv(x, ax)
        dev.axpy(x, 1, [l], ax, rr)                                                                # A*x - l*x
        nr = dev.normalize_col(rr, tmp)
        nx = dev.normalize_col(x, tmp)
        vals.append(l)
        res.append(nr / (abs(l) * nx))
    return np.asarray(vals), np.asarray(res)
