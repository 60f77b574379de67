"""Device-resident ``restarted_ca_lanczos``: the reference's restarted driver with every O(n) operation
behind a small *block-operations* object, so that the basis, the converged Ritz vectors and the restart vector never leave
the GPU (restarted_ca_lanczos.m:4-202, all four orth modes: lanczos_basic 'local'/'full' :288-367, lanczos_selective :369-463,
lanczos_periodic :465-552; generateStartVector 'largest' :204-248).

    eigs, Qconv, nrestarts, rnorms, orth_err = restarted_ca_lanczos(ops, r, max_lanczos, n_wanted_eigs, s, basis, orth, tol)

``ops`` is anything with the methods of :class:`DeviceOps` below (the tests drive the same control flow with a numpy
implementation on the CPU and compare it with the oracle's restatement of the reference).  The O(m^3) host algebra (eig of
the non-symmetric T, the lock-and-sort of converged pairs, the T assembly of :336-359) follows the reference line by line.

The control flow is covered by CPU tests (tests/test_restart_host.py, numpy block operations vs the oracle);
:class:`DeviceOps` chains C-ABI calls only and is checked end to end on the GPU by
tests/test_gpu_driver.py::test_device_resident_restarted_ca_lanczos and by ``bench.py --config c4``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .engine import _solve_upper_right
from .solver import leja_order, newton_basis_matrix


# ------------------------------------------------------------------------------------------------- host algebra
def _extend_T(T, b, k, s, Bk, Rkk_s, Rk_s):
    """restarted_ca_lanczos.m:336-359 (same as ca_lanczos.m:200-223).  ``k`` is 1-based; b[k-1] holds b(k)."""
    Rkk = np.hstack([np.zeros((s, 1)), Rkk_s[:s, :]])                                  # :337
    e1col = np.zeros((s + 1, 1)); e1col[0, 0] = 1.0
    Rk = np.hstack([e1col, np.vstack([Rkk_s[s:s + 1, :s], Rk_s])])                     # :338
    zk, rho, rho_t, bk = Rk[:s, s], Rk[s, s], Rk[s - 1, s - 1], Bk[s, s - 1]            # :339-342
    e1 = np.zeros(s); e1[0] = 1.0
    es = np.zeros(s); es[s - 1] = 1.0
    Rss = Rk[:s, :s]
    Tk = (_solve_upper_right(Rss @ Bk[:s, :], Rss) + (bk / rho_t) * np.outer(zk, es)
          - _solve_upper_right((b[k - 2] * np.outer(e1, es)) @ Rkk[:, :s], Rss))       # :343-347
    b[k - 1] = bk * (rho / rho_t)                                                      # :350
    m = s * (k - 1)
    Tn = np.zeros((m + s + 1, m + s), order="F")                                       # :353-359
    Tn[:m, :m] = T[:m, :m]
    Tn[m - 1, m] = b[k - 2]
    Tn[m, m - 1] = b[k - 2]
    Tn[m:m + s, m:m + s] = Tk
    Tn[m + s, m + s - 1] = b[k - 1]
    return Tn


def _lanczos_T(ops, q, maxiter):
    """lanczos.m:85-134 ('local'): the tridiagonal T of ``maxiter`` steps started from the unit vector ``q``."""
    Qs = ops.block(maxiter + 1)
    rr = ops.block(1)
    ops.axpy(q, np.array([[-1.0]]), None, ops.view(Qs, 0, 1))                          # Q(:,1) = q
    alpha = np.zeros(maxiter); beta = np.zeros(maxiter)
    for j in range(maxiter):
        qj = ops.view(Qs, j, j + 1)
        ops.spmv(qj, rr)                                                               # r = A*Q(:,j)                 (:103)
        if j > 0:
            ops.axpy(ops.view(Qs, j - 1, j), np.array([[beta[j - 1]]]), rr, rr)        # r -= beta(j-1)*Q(:,j-1)       (:105)
        alpha[j] = float(ops.gram(qj, rr)[0, 0])                                       # alpha = r'*q                  (:107)
        ops.axpy(qj, np.array([[alpha[j]]]), rr, rr)                                   # r -= alpha*q                  (:108)
        beta[j] = ops.nrm2(rr)                                                         # beta = sqrt(r'*r)             (:109)
        ops.axpy(rr, np.array([[-1.0 / beta[j]]]), None, ops.view(Qs, j + 1, j + 2))   # Q(:,j+1) = r/beta             (:110)
    return np.diag(alpha) + np.diag(beta[:maxiter - 1], 1) + np.diag(beta[:maxiter - 1], -1)


def basis_matrix(ops, q, s, basis):
    """restarted_ca_lanczos.m:60-71: Bk ((s+1) x s); Newton shifts = Leja-ordered Ritz values of a 2s-step 'local' Lanczos."""
    if basis.lower() == "monomial":
        return np.asfortranarray(np.eye(s + 1)[:, 1:s + 1])
    if basis.lower() != "newton":
        raise ValueError("ERROR: Unknown basis type: " + basis)
    T = _lanczos_T(ops, q, 2 * s)
    return newton_basis_matrix(leja_order(np.linalg.eigvalsh(T)), s)


def normest(ops, tol=1.0e-6):
    """MATLAB normest (restarted_ca_lanczos.m:35) for a SYMMETRIC matrix: power iteration on A'A = A*A from the column sums."""
    x, xn, sx = ops.block(1), ops.block(1), ops.block(1)
    ops.abs_colsum(x)
    e = ops.nrm2(x)
    if e == 0:
        return e
    ops.axpy(x, np.array([[-1.0 / e]]), None, xn)                                      # xn = x/e
    e0, cnt = 0.0, 0
    while abs(e - e0) > tol * e:
        e0 = e
        ops.spmv(xn, sx)                                                               # Sx = S*x
        ops.spmv(sx, x)                                                                # x  = S'*Sx (S symmetric)
        normx = ops.nrm2(x)
        e = normx / ops.nrm2(sx)
        ops.axpy(x, np.array([[-1.0 / normx]]), None, xn)                              # x = x/normx
        cnt += 1
        if cnt > 100:
            break
    return e


# ------------------------------------------------------------------------------------------------- one restart cycle
def _lanczos_basic(ops, Qc, q, Bk, maxiter, s, basis, orth, Q):
    """restarted_ca_lanczos.m:288-367 -- ``while k <= maxiter`` with k starting at 1 and incremented after the body gives
    maxiter+1 blocks (:301).  ``Q``: block storage n x ((maxiter+1)*s+1); returns T ((s*maxiter+1) x s*maxiter)."""
    nblk = maxiter + 1
    b = np.zeros(nblk + 1)
    T = None
    tmp = ops.block(s + 1)
    for k in range(1, nblk + 1):
        qk = q if k == 1 else ops.view(Q, (k - 1) * s, (k - 1) * s + 1)
        V = ops.matrix_powers(qk, s, Bk, basis)                                        # :305-311
        if k == 1:
            Rk = ops.normalize(V, tmp)                                                 # :313
            ops.pan([Qc], tmp, ops.view(Q, 0, s + 1))                                  # :315 (R discarded)
            T = _solve_upper_right(Rk @ Bk, Rk[:s, :s])                                # :317
            b[0] = T[s, s - 1]
        else:
            Qprev = ops.view(Q, (k - 2) * s, (k - 1) * s + 1)
            X = ops.view(V, 1, s + 1)
            dst = ops.view(Q, (k - 1) * s + 1, k * s + 1)
            if orth == "local":
                R = ops.pan([Qprev, Qc], X, dst)                                       # :324-327
                Rkk_s, Rk_s = R[0], R[2]
            else:
                R = ops.pan([Qprev], X, ops.view(tmp, 0, s))                           # :330
                Rkk_s, Rk_s = R[0], R[1]
                Qold = ops.view(Q, 0, (k - 2) * s) if (k - 2) * s > 0 else None
                ops.pan([Qc, Qold], ops.view(tmp, 0, s), dst)                          # :333
            T = _extend_T(T, b, k, s, Bk, Rkk_s, Rk_s)
    return np.asfortranarray(T[:s * maxiter + 1, :s * maxiter])                        # :365-366


def _restarted_periodic_reorth_needed(omega, k, s):
    """restarted_ca_lanczos.m:535-543: row maximum over the first i entries only, no abs(), threshold sqrt(eps/(k*s)) -- as written
    (the non-restarted driver's test, ca_lanczos.m:437-446, is a different one)."""
    err = 0.0
    for i in range(1, s + 1):
        row_err = float(np.max(omega[(k - 1) * s + i, :i]))
        if row_err > err:
            err = row_err
    return err >= np.sqrt(np.finfo(np.float64).eps / (k * s))


def _lanczos_periodic_selective(ops, Qc, q, Bk, maxiter, s, basis, orth, Q, norm_A, log=None):
    """lanczos_selective (restarted_ca_lanczos.m:369-463) and lanczos_periodic (:465-552) over the block operations: maxiter+1
    blocks (``while k <= maxiter``), every block projected against {Qprev, Q_conv} (:500) or {Qprev, Q_conv, QR(:,1:nritz)}
    (:405); periodic: the omega recurrence on the host (:532-534) and, when it trips, the s+1 newest vectors re-orthogonalised
    against everything before them (:544); selective: eig of the current T, Ritz vectors whose estimated residual fell below
    ||A|| sqrt(eps) are formed, orthonormalised (:444-454) and joined to the projection list."""
    from .solver import reset_omega, update_omega
    nblk = maxiter + 1
    b = np.zeros(nblk + 1)
    T = None
    tmp = ops.block(s + 1)
    omega = None
    nritz, breaks = 0, []
    QRraw = QRo = None
    eps = np.finfo(np.float64).eps
    for k in range(1, nblk + 1):
        qk = q if k == 1 else ops.view(Q, (k - 1) * s, (k - 1) * s + 1)
        V = ops.matrix_powers(qk, s, Bk, basis)
        if k == 1:
            Rk = ops.normalize(V, tmp)                                                 # :394 / :489
            ops.pan([Qc], tmp, ops.view(Q, 0, s + 1))                                  # :396 / :491
            T = _solve_upper_right(Rk @ Bk, Rk[:s, :s])
            b[0] = T[s, s - 1]
        else:
            blocks = [ops.view(Q, (k - 2) * s, (k - 1) * s + 1), Qc]
            if orth == "selective":
                blocks.append(ops.view(QRo, 0, nritz) if nritz > 0 else None)
            R = ops.pan(blocks, ops.view(V, 1, s + 1), ops.view(Q, (k - 1) * s + 1, k * s + 1))
            T = _extend_T(T, b, k, s, Bk, R[0], R[-1])                                 # Rk_{1}, Rk_{3} / Rk_{4}
        if orth == "periodic":
            omega = update_omega(omega, np.diag(T, 0).copy(), np.diag(T, -1).copy(), norm_A, s)
            if _restarted_periodic_reorth_needed(omega, k, s):
                breaks.append(k)
                old = ops.view(Q, 0, (k - 1) * s) if (k - 1) * s > 0 else None
                cur = ops.view(Q, (k - 1) * s, k * s + 1)
                ops.pan([old], cur, tmp)                                               # :544 (QZ may not alias X: through tmp)
                ops.axpy(tmp, -np.eye(s + 1), None, cur)
                omega = reset_omega(omega, norm_A, s)
        else:
            m = s * k
            _, Vp = np.linalg.eig(T[:m, :m])                                           # :436
            Vp = np.real(Vp)
            conv = b[k - 1] * np.abs(Vp[m - 1, :]) < norm_A * np.sqrt(eps)             # :438-443
            if int(conv.sum()) > nritz:                                                # :444-454
                breaks.append(k)
                nritz = int(conv.sum())
                if QRraw is None:
                    QRraw, QRo = ops.block(nblk * s), ops.block(nblk * s)
                cols = np.flatnonzero(conv)
                Qm = ops.view(Q, 0, m)
                for c0 in range(0, nritz, 16):                                         # y = Q(:,1:k*s)*Vp(:,i), 16 columns a call
                    cc = min(16, nritz - c0)
                    ops.axpy(Qm, -Vp[:, cols[c0:c0 + cc]], None, ops.view(QRraw, c0, c0 + cc))
                # QR(:,1:nritz) = normalize(QR(:,1:nritz)) (:453); the device QR takes 32 columns at most, wider sets are
                # orthonormalised block by block (against the finished ones, then normalised): the same Q factor
                done = 0
                while done < nritz:
                    cc = min(16, nritz - done)
                    src, dst = ops.view(QRraw, done, done + cc), ops.view(QRo, done, done + cc)
                    if done == 0:
                        ops.normalize(src, dst)
                    else:
                        ops.pan([ops.view(QRo, 0, done)], src, dst)
                    done += cc
    if log is not None:
        log.setdefault("breaks", []).append(breaks)
        log.setdefault("nritz", []).append(nritz)
    return np.asfortranarray(T[:s * maxiter + 1, :s * maxiter])


def restarted_ca_lanczos(ops, r, max_lanczos, n_wanted_eigs=10, s=6, basis="newton", orth="local", tol=1.0e-8,
                         max_restarts=200, want_orth_err=True, log=None):
    """restarted_ca_lanczos.m:4-202.  ``r``: start vector as an ops block (n x 1).
    Returns (conv_eigs, Qconv block view, num_restarts, rnorms, orth_err, order): conv_eigs sorted descending, the matching
    Ritz vectors are the columns ``order`` of the Qconv block."""
    orth = str(orth).lower()
    if orth not in ("local", "full", "periodic", "selective"):
        raise ValueError("lanczos.m: Invalid option value for orth: " + orth)          # :27-31
    norm_A = normest(ops)
    tol = tol * norm_A                                                                 # :35-36
    iters = max_lanczos // s                                                           # :88
    q = ops.block(1)
    ops.axpy(r, np.array([[-1.0 / ops.nrm2(r)]]), None, q)                             # q = r/norm(r)        (:57)
    Bk = basis_matrix(ops, q, s, basis)
    Qstore = ops.block(max_lanczos + n_wanted_eigs + s * iters)                        # converged Ritz vectors (:73)
    Qnew = ops.block((iters + 1) * s + 1)
    x, ax = ops.block(1), ops.block(1)
    conv_eigs, conv_rnorms, orth_err = [], [], []
    rnorms = np.zeros((max_restarts, n_wanted_eigs))
    num_restarts, nconv, restart = 0, 0, True

    def relres(l, xv):                                                                 # ||A x - l x|| / ||l x||   (:141-150)
        ops.spmv(xv, ax)
        ops.axpy(xv, np.array([[l]]), ax, ax)
        return ops.nrm2(ax) / (abs(l) * ops.nrm2(xv))

    while restart and num_restarts < max_restarts:
        num_restarts += 1
        if iters == 0:
            break                                                                      # reference branch uses undefined variables (:91-95)
        Qc = ops.view(Qstore, 0, nconv) if nconv > 0 else None
        if orth in ("periodic", "selective"):                                          # :101-104
            T = _lanczos_periodic_selective(ops, Qc, q, Bk, iters, s, basis, orth, Qnew, norm_A, log)
        else:
            T = _lanczos_basic(ops, Qc, q, Bk, iters, s, basis, orth, Qnew)
        m = s * iters
        Qm = ops.view(Qnew, 0, m)
        Dp, Vp = np.linalg.eig(T[:m, :m])                                              # non-symmetric T => general solver
        Dp = np.real(Dp).copy(); Vp = np.real(Vp).copy()
        beta = T[m, m - 1]
        ritz_norms = beta * np.abs(Vp[m - 1, :])                                       # :103-105
        k = 0
        for i in range(m):                                                             # :107-118 lock converged pairs in front
            if ritz_norms[i] < tol:
                Dp[[i, k]] = Dp[[k, i]]
                Vp[:, [i, k]] = Vp[:, [k, i]]
                ritz_norms[[i, k]] = ritz_norms[[k, i]]
                k += 1
        for i in range(k):                                                             # :135-139 Ritz vectors of the locked pairs
            ops.axpy(Qm, -Vp[:, i:i + 1], None, ops.view(Qstore, nconv + i, nconv + i + 1))
            conv_eigs.append(Dp[i]); conv_rnorms.append(ritz_norms[i])
        if num_restarts > 1:
            rnorms[num_restarts - 1, :nconv] = rnorms[num_restarts - 2, :nconv]
        for i in range(k):
            if nconv + i < n_wanted_eigs:
                rnorms[num_restarts - 1, nconv + i] = relres(conv_eigs[nconv + i], ops.view(Qstore, nconv + i, nconv + i + 1))
        rest = Dp[k:]
        ix = np.argsort(-rest, kind="stable")
        for i in range(max(0, n_wanted_eigs - nconv - k)):                             # :152-160 unconverged leaders (diagnostic)
            ops.axpy(Qm, -Vp[:, k + ix[i]:k + ix[i] + 1], None, x)
            rnorms[num_restarts - 1, nconv + i + k] = relres(rest[ix[i]], x)
        if want_orth_err:                                                              # :165-168  ||I - [Qc Qnew]'[Qc Qnew]||_F
            parts = ([Qc] if Qc is not None else []) + [Qm]
            if hasattr(ops, "orth_fro"):
                orth_err.append(ops.orth_fro(parts))
                parts = []
            tot = sum(ops.ncols(p) for p in parts)
            G = np.zeros((tot, tot))
            o1 = 0
            for p1 in parts:
                o2 = 0
                for p2 in parts:
                    G[o1:o1 + ops.ncols(p1), o2:o2 + ops.ncols(p2)] = ops.gram(p1, p2)
                    o2 += ops.ncols(p2)
                o1 += ops.ncols(p1)
            if parts:
                orth_err.append(float(np.linalg.norm(np.eye(tot) - G, "fro")))
        nconv += k
        restart = not (len(conv_eigs) >= n_wanted_eigs)                                # :178-181
        if restart:
            l = k                                                                      # generateStartVector 'largest' (:209-218)
            for j in range(k, m):
                if Dp[j] > Dp[l]:
                    l = j
            ops.axpy(Qm, -Vp[:, l:l + 1], None, x)
            ops.axpy(x, np.array([[-1.0 / ops.nrm2(x)]]), None, q)
    conv_eigs = np.asarray(conv_eigs)
    ixs = np.argsort(-conv_eigs, kind="stable")
    keep = n_wanted_eigs if not restart else nconv
    # Q_conv(:, ixs)(:, 1:keep) of the reference = the columns ``order`` of the returned block
    return (conv_eigs[ixs][:keep], (ops.view(Qstore, 0, nconv) if nconv else None), num_restarts, rnorms[:num_restarts],
            np.asarray(orth_err), ixs[:keep])


# ------------------------------------------------------------------------------------------------- device operations
class _Blk:
    """n x cols column-major view: device pointer + leading dimension (the owning tensor is kept alive by ``keep``)."""
    __slots__ = ("ptr", "ld", "cols", "keep")

    def __init__(self, ptr, ld, cols, keep=None):
        self.ptr, self.ld, self.cols, self.keep = int(ptr), int(ld), int(cols), keep


class DeviceOps:
    """The block operations on the GPU: every method is one or a few C-ABI calls on the context's stream."""

    def __init__(self, dm, backend="tsqr", A_host=None, colsum=None):
        import torch
        from . import _lib
        self.torch, self._lib, self.dm, self.ctx, self.lib = torch, _lib, dm, dm.ctx, dm.ctx.lib
        self.n = dm.n
        self.ld = (self.n + 31) // 32 * 32
        self.dev = torch.device("cuda", self.ctx.device)
        self.backend = backend
        # normest starts from the column sums of |A| (owned rows of this rank; A symmetric => the row sums of the owned rows)
        self._colsum = (np.ascontiguousarray(colsum, dtype=np.float64) if colsum is not None else
                        None if A_host is None else np.asarray(abs(A_host).sum(axis=0)).ravel().astype(np.float64))
        self._g = torch.zeros(16 * 16, dtype=torch.float64, device=self.dev)
        torch.cuda.synchronize(self.dev)

    # storage
    def block(self, cols):
        t = self.torch.zeros((max(int(cols), 1), self.ld), dtype=self.torch.float64, device=self.dev)
        self.torch.cuda.synchronize(self.dev)
        return _Blk(t.data_ptr(), self.ld, cols, t)

    def from_host(self, v):
        b = self.block(1)
        b.keep[0, :self.n] = self.torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64), device=self.dev)
        self.torch.cuda.synchronize(self.dev)
        return b

    def to_host(self, B):
        self.ctx.sync()
        out = np.empty((self.n, B.cols), order="F")
        for j in range(B.cols):
            t = self.torch.empty(self.n, dtype=self.torch.float64, device=self.dev)
            self._check(self.lib.calz_block_axpy(self.ctx.h, self.n, 1, C.c_void_p(B.ptr + 8 * B.ld * j), B.ld, 1,
                                                 np.array([-1.0]).ctypes.data_as(self._lib.c_dp), None, self.ld, C.c_void_p(t.data_ptr()), self.ld))
            self.ctx.sync()
            out[:, j] = t.cpu().numpy()
        return out

    def view(self, B, j0, j1):
        return _Blk(B.ptr + 8 * B.ld * int(j0), B.ld, int(j1) - int(j0), B.keep)

    def ncols(self, B):
        return B.cols

    def _check(self, st):
        self._lib.check(st, self.ctx.h)

    # O(n) operations
    def spmv(self, x, y):
        self._check(self.lib.calz_spmv(self.dm.h, C.c_void_p(x.ptr), C.c_void_p(y.ptr)))

    def axpy(self, Q, Cm, X, Y):
        Cm = np.asfortranarray(np.asarray(Cm, dtype=np.float64).reshape(Q.cols, -1))
        c = Cm.shape[1]
        self._check(self.lib.calz_block_axpy(self.ctx.h, self.n, Q.cols, C.c_void_p(Q.ptr), Q.ld, c, Cm.ctypes.data_as(self._lib.c_dp),
                                             C.c_void_p(X.ptr) if X is not None else None, X.ld if X is not None else self.ld,
                                             C.c_void_p(Y.ptr), Y.ld))

    def gram(self, A, B):
        out = np.zeros((A.cols, B.cols))
        for i0 in range(0, A.cols, 8):
            mi = min(8, A.cols - i0)
            for j0 in range(0, B.cols, 8):
                cj = min(8, B.cols - j0)
                self._check(self.lib.calz_gram(self.ctx.h, self.n, mi, C.c_void_p(A.ptr + 8 * A.ld * i0), A.ld, cj,
                                               C.c_void_p(B.ptr + 8 * B.ld * j0), B.ld, C.c_void_p(self._g.data_ptr())))
                self.ctx.sync()
                out[i0:i0 + mi, j0:j0 + cj] = self._g[:mi * cj].cpu().numpy().reshape(cj, mi).T
        return out

    def nrm2(self, x):
        return float(np.sqrt(self.gram(x, x)[0, 0]))

    def orth_fro(self, parts):
        """norm(eye - [parts]'*[parts], 'fro')  (restarted_ca_lanczos.m:165-168) in one device pass"""
        from .solver import orth_error
        return orth_error(self.ctx, self.n, [(p.ptr, p.ld, p.cols) for p in parts], "fro")

    def abs_colsum(self, out):
        if self._colsum is None:
            raise ValueError("DeviceOps: pass A_host (the scipy matrix) to use normest")
        t = self.from_host(self._colsum)
        self.axpy(t, np.array([[-1.0]]), None, out)
        self.ctx.sync()

    # the hot path
    def matrix_powers(self, q, s, Bk, basis):
        mono = basis.lower() == "monomial"
        re = None if mono else np.ascontiguousarray(np.diag(Bk)[:s], dtype=np.float64)
        V, ldV = C.c_void_p(), C.c_int64()
        self._check(self.lib.calz_mpk_inplace(self.dm.h, C.c_void_p(q.ptr), int(s), None if mono else re.ctypes.data_as(self._lib.c_dp),
                                              None, 1, 1 if mono else 0, C.byref(V), C.byref(ldV)))
        return _Blk(V.value, ldV.value, s + 1)

    def normalize(self, V, out):
        c = V.cols
        R = np.zeros((c, c), order="F")
        rank = C.c_int()
        self._check(self.lib.calz_normalize(self.ctx.h, self.n, c, C.c_void_p(V.ptr), V.ld, self._lib.QR[self.backend], 1e-8,
                                            C.c_void_p(out.ptr), out.ld, R.ctypes.data_as(self._lib.c_dp), C.byref(rank)))
        return R

    def pan(self, blocks, X, out):
        nb, c = len(blocks), X.cols
        qb = (C.c_void_p * nb)(*[(b.ptr if b is not None else None) for b in blocks])
        lds = (C.c_int64 * nb)(*[(b.ld if b is not None else self.ld) for b in blocks])
        mc = (C.c_int * nb)(*[(b.cols if b is not None else 0) for b in blocks])
        Rs = [np.zeros((b.cols, c), order="F") if b is not None else None for b in blocks]
        rp = (self._lib.c_dp * nb)(*[(r.ctypes.data_as(self._lib.c_dp) if r is not None else None) for r in Rs])
        Rl = np.zeros((c, c), order="F")
        second, rank = C.c_int(), C.c_int()
        self._check(self.lib.calz_project_and_normalize(self.ctx.h, self.n, nb, qb, lds, mc, c, C.c_void_p(X.ptr), X.ld, 1,
                                                        self._lib.QR[self.backend], C.c_void_p(out.ptr), out.ld, rp,
                                                        Rl.ctypes.data_as(self._lib.c_dp), C.byref(second), C.byref(rank)))
        return Rs + [Rl]


def device_restarted_ca_lanczos(A, r, max_lanczos, n_wanted_eigs=10, s=6, basis="newton", orth="local", tol=1.0e-8,
                                backend="tsqr", ctx=None, max_restarts=200):
    """``[conv_eigs, Q_conv, num_restarts, rnorms, orth_err] = restarted_ca_lanczos(A, r, max_lanczos, n_wanted_eigs, s, basis,
    orth, tol)`` with A a scipy sparse SYMMETRIC matrix and r a host vector; Q_conv comes back as a host array."""
    from .api import DeviceMatrix, default_context
    dm = DeviceMatrix(A, s_max=max(int(s), 1), ctx=ctx or default_context())
    ops = DeviceOps(dm, backend=backend, A_host=A)
    eigs, Qc, nres, rn, oe, order = restarted_ca_lanczos(ops, ops.from_host(r), max_lanczos, n_wanted_eigs, s, basis, orth, tol,
                                                         max_restarts)
    Qh = ops.to_host(Qc)[:, order] if Qc is not None else np.zeros((dm.n, 0))
    return eigs, Qh, nres, rn, oe
