"""CPU restatement of the CALLERS of the hot path (ca_lanczos / restarted_ca_lanczos, 'local' and 'full'
orthogonalisation), needed only to carry parity from basis vectors through to T entries and Ritz values.
TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/kernels.py).

The block kernels are taken from a *kernel provider* ``K`` (default: oracle.kernels) exposing the
reference's call surface -- ``matrix_powers_monomial``, ``matrix_powers_newton``, ``normalize``,
``projectAndNormalize``, ``SpMV``.  The parity tests run the same driver once with the oracle provider and
once with the CUDA drop-in provider and compare T, Q and the Ritz values (this is exactly the seam at
which the reference's MATLAB drivers would pick up same-named MEX files).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

from . import kernels as _oracle_kernels
from .leja import leja, newton_basis_matrix


def _mrdivide_upper(B, R):
    """MATLAB ``B/R`` for square upper-triangular R (back substitution): X R = B."""
    return sla.solve_triangular(R, B.T, trans="T", lower=False).T


def _eyeshvec(n):
    e = np.zeros(n); e[n - 1] = 1.0
    return e


# ----------------------------------------------------------------------------- lanczos.m
def lanczos(A, r, maxiter, orth="local", K=_oracle_kernels):
    """lanczos.m:18-60 + lanczos_basic :85-134 ('local' and 'full' = one CGS sweep, no renormalise
    :62-66,112-114).  Returns (T, Q) like ``[T,Q] = lanczos(...)``."""
    orth = str(orth).lower()
    if orth not in ("local", "full"):
        raise NotImplementedError("lanczos orth=%s is out of scope" % orth)
    q = r / np.linalg.norm(r)
    n = q.shape[0]
    Q = np.zeros((n, maxiter + 1), order="F")
    Q[:, 0] = q
    alpha = np.zeros(maxiter); beta = np.zeros(maxiter)
    for j in range(maxiter):
        rr = K.SpMV(A, Q[:, j])
        if j > 0:
            rr = rr - beta[j - 1] * Q[:, j - 1]
        alpha[j] = rr @ Q[:, j]
        rr = rr - alpha[j] * Q[:, j]
        beta[j] = np.sqrt(rr @ rr)
        Q[:, j + 1] = rr / beta[j]
        if orth == "full":
            Rkk = Q[:, : j + 1].T @ Q[:, j + 1]
            Q[:, j + 1] = Q[:, j + 1] - Q[:, : j + 1] @ Rkk
    T = np.diag(alpha) + np.diag(beta[: maxiter - 1], 1) + np.diag(beta[: maxiter - 1], -1)
    return T, Q


def basis_matrix(A, q, s, basis, shift_orth, K=_oracle_kernels):
    """ca_lanczos.m:61-72 / restarted_ca_lanczos.m:60-71 -- change-of-basis matrix Bk ((s+1) x s)."""
    if basis.lower() == "monomial":
        return np.asfortranarray(np.eye(s + 1)[:, 1 : s + 1])
    if basis.lower() == "newton":
        T, _ = lanczos(A, q, 2 * s, shift_orth, K=K)
        basis_eigs = np.linalg.eigvalsh(T)          # T exactly symmetric => MATLAB eig = symmetric solver, ascending
        basis_shifts, _ = leja(basis_eigs, "nonmodified")
        return newton_basis_matrix(basis_shifts, s, 1)
    raise ValueError("ERROR: Unknown basis type: " + basis)


def _matrix_powers(A, q, s, Bk, basis, K):
    """ca_lanczos.m:110-118."""
    if basis.lower() == "monomial":
        V = np.empty((q.shape[0], s + 1), order="F")
        V[:, 0] = q
        V[:, 1:] = K.matrix_powers_monomial(A, q, s)
        return V
    return K.matrix_powers_newton(A, q, s, np.diag(Bk).copy(), 1)


def _extend_T(T, b, k, s, Bk, Rkk_s, Rk_s):
    """ca_lanczos.m:200-223 (identical in restarted_ca_lanczos.m:336-359).  ``k`` is 1-based; b[k-1] is b(k)."""
    Rkk = np.hstack([np.zeros((s, 1)), Rkk_s[:s, :]])
    e1col = np.zeros((s + 1, 1)); e1col[0, 0] = 1.0
    Rk = np.hstack([e1col, np.vstack([Rkk_s[s : s + 1, :s], Rk_s])])
    zk = Rk[:s, s]
    rho = Rk[s, s]
    rho_t = Rk[s - 1, s - 1]
    bk = Bk[s, s - 1]
    e1 = np.zeros(s); e1[0] = 1.0
    es = _eyeshvec(s)
    Rss = Rk[:s, :s]
    Tk = (_mrdivide_upper(Rss @ Bk[:s, :], Rss)
          + (bk / rho_t) * np.outer(zk, es)
          - _mrdivide_upper(((b[k - 2] * np.outer(e1, es)) @ Rkk[:, :s]), Rss))
    b[k - 1] = bk * (rho / rho_t)
    m = s * (k - 1)
    Tn = np.zeros((m + s + 1, m + s), order="F")
    Tn[:m, :m] = T[:m, :m]
    Tn[m - 1, m] = b[k - 2]                       # T12 = b(k-1) * e_last * e1'
    Tn[m, m - 1] = b[k - 2]                       # T21 = b(k-1) * e1 * e_last'
    Tn[m : m + s, m : m + s] = Tk
    Tn[m + s, m + s - 1] = b[k - 1]               # T32 = b(k) * es'
    return Tn


# ----------------------------------------------------------------------------- ca_lanczos.m:469-551 (periodic orthogonalisation)
EPS = np.finfo(np.float64).eps


def update_omega(omega_in, alpha, beta, anorm, s):
    """ca_lanczos.m:469-539 -- the omega recurrence (estimated loss of orthogonality) extended by s rows.  ``alpha``/``beta``:
    diag(T,0) / diag(T,-1).  Written with 1-based index arrays (index 0 unused) so that every line reads like the reference."""
    n = len(alpha)
    al = np.concatenate([[0.0], np.asarray(alpha, dtype=np.float64)])
    be = np.concatenate([[0.0], np.asarray(beta, dtype=np.float64)])
    Tr = EPS * anorm                                                         # :476 "T = eps*anorm"
    if omega_in is None or np.size(omega_in) == 0:
        om = np.zeros((s + 2, s + 2))                                        # 1-based (s+1) x (s+1)
        om[1, 1] = 1.0; om[1, 2] = 0.0
        om[2, 1] = Tr / be[1]; om[2, 2] = 1.0
        rng = range(2, s + 1)
    else:
        m = omega_in.shape[0] - 1
        om = np.zeros((n + 2, n + 2))
        om[1:m + 2, 1:m + 2] = omega_in
        rng = range(m + 1, m + s + 1)
    for j in rng:
        binv = 1.0 / be[j]
        om[j + 1, 1] = be[2] * om[j, 2] + (al[1] - al[j]) * om[j, 1] - be[j] * om[j - 1, 1]
        om[j + 1, 1] = binv * (om[j + 1, 1] + Tr) if om[j + 1, 1] > 0 else binv * (om[j + 1, 1] - Tr)
        for k in range(2, j):
            om[j + 1, k] = be[k + 1] * om[j, k + 1] + (al[k] - al[j]) * om[j, k] + be[k] * om[j, k - 1] - be[j] * om[j - 1, k]
            om[j + 1, k] = binv * (om[j + 1, k] + Tr) if om[j + 1, k] > 0 else binv * (om[j + 1, k] - Tr)
        om[j + 1, j] = binv * Tr
        om[j + 1, j + 1] = 1.0
    return om[1:, 1:].copy()


def reset_omega(omega_in, anorm, s):
    """ca_lanczos.m:541-551."""
    Tr = EPS * anorm
    m = omega_in.shape[0] - s - 1
    om = np.zeros((omega_in.shape[0] + 1, omega_in.shape[1] + 1))
    om[1:, 1:] = omega_in
    for j in range(m + 1, m + s + 1):
        om[j + 1, 1:j + 1] = Tr
        om[j + 1, j + 1] = 1.0
    return om[1:, 1:].copy()


def periodic_reorth_needed(omega, k, s):
    """ca_lanczos.m:437-446: the largest estimated inner product of the s new vectors with everything before them."""
    err = 0.0
    for i in range(1, s + 1):
        row = omega[(k - 1) * s + i, : (k - 1) * s + i]                      # omega((k-1)*s+i+1, 1:(k-1)*s+i), 0-based
        err = max(err, float(np.max(np.abs(row))) if row.size else 0.0)
    return err >= np.sqrt(EPS), err


# ----------------------------------------------------------------------------- ca_lanczos.m
def ca_lanczos(A, r, s, iter, basis, orth="local", K=_oracle_kernels, backend="tsqr", Bk=None, info=None):
    """ca_lanczos.m:24-86 + ca_lanczos_basic :150-245 ('local' and 'full'), ca_lanczos_selective :248-359,
    ca_lanczos_periodic :362-467.  Returns (T, Q)."""
    orth = str(orth).lower()
    if orth in ("periodic", "selective"):
        return _ca_lanczos_periodic_selective(A, r, s, iter, basis, orth, K, backend, Bk, info)
    if orth not in ("local", "full"):
        raise ValueError("ERROR: Unknown orth type: " + orth)
    t = int(np.ceil(iter / s))
    q = r / np.sqrt(r @ r)
    if Bk is None:
        Bk = basis_matrix(A, q, s, basis, "full", K=K)
    n = q.shape[0]
    Q = np.zeros((n, t * s + 1), order="F")
    Q[:, 0] = q
    b = np.zeros(t + 1)
    T = None
    log = []
    for k in range(1, t + 1):
        if k > 1:
            q = Q[:, (k - 1) * s]
        V = _matrix_powers(A, q, s, Bk, basis, K)
        if k == 1:
            Q[:, : s + 1], Rk, _ = K.normalize(V[:, : s + 1], backend=backend)
            T = _mrdivide_upper(Rk @ Bk, Rk[:s, :s])
            b[0] = T[s, s - 1]
        else:
            inf = {}
            Q_, Rk_ = K.projectAndNormalize([Q[:, (k - 2) * s : (k - 1) * s + 1]], V[:, 1 : s + 1], True,
                                            backend=backend, info=inf)
            log.append(inf)
            Rkk_s, Rk_s = Rk_[0], Rk_[1]
            Q[:, (k - 1) * s + 1 : k * s + 1] = Q_
            if orth == "full":
                Q[:, (k - 1) * s + 1 : k * s + 1], _ = K.projectAndNormalize(
                    [Q[:, : (k - 1) * s + 1]], Q[:, (k - 1) * s + 1 : k * s + 1], True, backend=backend)
            T = _extend_T(T, b, k, s, Bk, Rkk_s, Rk_s)
    if info is not None:
        info["pan"] = log
        info["Bk"] = Bk
    return np.asfortranarray(T[: s * t, : s * t]), Q[:, : s * t]


def _ca_lanczos_periodic_selective(A, r, s, iter, basis, orth, K, backend, Bk, info):
    """ca_lanczos_periodic (ca_lanczos.m:362-467) and ca_lanczos_selective (:248-359).  Note ``iter`` here is what the
    dispatcher passes: t = ceil(iter/s) outer steps (:52,:80,:82)."""
    t = int(np.ceil(iter / s))
    q = r / np.sqrt(r @ r)
    if Bk is None:
        Bk = basis_matrix(A, q, s, basis, "full", K=K)
    n = q.shape[0]
    Q = np.zeros((n, t * s + 1), order="F")
    Q[:, 0] = q
    b = np.zeros(t + 1)
    T = None
    norm_A = normest(A)
    omega = None
    nbreaks, breaks, log = 0, [], []
    QR = np.zeros((n, 0), order="F")                                         # selective: converged Ritz vectors
    nritz = 0
    for k in range(1, t + 1):
        if k > 1:
            q = Q[:, (k - 1) * s]
        V = _matrix_powers(A, q, s, Bk, basis, K)
        if k == 1:
            Q[:, : s + 1], Rk, _ = K.normalize(V[:, : s + 1], backend=backend)
            T = _mrdivide_upper(Rk @ Bk, Rk[:s, :s])
            b[0] = T[s, s - 1]
        else:
            inf = {}
            blocks = [Q[:, (k - 2) * s : (k - 1) * s + 1]]
            if orth == "selective":
                blocks.append(QR[:, :nritz] if nritz > 0 else None)          # :286  (an n x 0 block is an empty cell, project.m:33-38)
            Q_, Rk_ = K.projectAndNormalize(blocks, V[:, 1 : s + 1], True, backend=backend, info=inf)
            log.append(inf)
            Rkk_s, Rk_s = Rk_[0], Rk_[-1]                                    # Rk_{1}, Rk_{2} (periodic) / Rk_{3} (selective)
            Q[:, (k - 1) * s + 1 : k * s + 1] = Q_
            T = _extend_T(T, b, k, s, Bk, Rkk_s, Rk_s)
        if orth == "periodic":
            alpha = np.diag(T, 0).copy()
            beta = np.diag(T, -1).copy()
            omega = update_omega(omega, alpha, beta, norm_A, s)              # :433-436
            need, err = periodic_reorth_needed(omega, k, s)
            if need:                                                         # :447-451
                nbreaks += 1
                breaks.append(k)
                old = Q[:, : (k - 1) * s] if (k - 1) * s > 0 else None
                Q[:, (k - 1) * s : k * s + 1], _ = K.projectAndNormalize([old], Q[:, (k - 1) * s : k * s + 1], True, backend=backend)
                omega = reset_omega(omega, norm_A, s)
        else:
            Dp, Vp = np.linalg.eig(T[: s * k, : s * k])                      # :322
            Dp, Vp = np.real(Dp), np.real(Vp)
            conv = b[k - 1] * np.abs(Vp[s * k - 1, :]) < norm_A * np.sqrt(EPS)
            if int(conv.sum()) > nritz:                                      # :330-341
                nbreaks += 1
                breaks.append(k)
                nritz = int(conv.sum())
                QR = np.asfortranarray(Q[:, : k * s] @ Vp[:, conv])
                QR, _, _ = K.normalize(QR, backend=backend)
    if info is not None:
        info["pan"] = log
        info["Bk"] = Bk
        info["nbreaks"] = nbreaks
        info["breaks"] = breaks
        info["nritz"] = nritz
    return np.asfortranarray(T[: s * t, : s * t]), Q[:, : s * t]


# ----------------------------------------------------------------------------- restarted_ca_lanczos.m
def normest(S, tol=1.0e-6):
    """MATLAB ``normest`` (restarted_ca_lanczos.m:35): power iteration on S'S started from the column sums."""
    x = np.asarray(abs(S).sum(axis=0)).ravel().astype(np.float64)
    e = np.linalg.norm(x)
    if e == 0:
        return e
    x = x / e
    e0 = 0.0
    cnt = 0
    while abs(e - e0) > tol * e:
        e0 = e
        Sx = S @ x
        x = S.T @ Sx
        normx = np.linalg.norm(x)
        e = normx / np.linalg.norm(Sx)
        x = x / normx
        cnt += 1
        if cnt > 100:
            break
    return e


def _restarted_lanczos_basic(A, Q_conv, q, Bk, maxiter, s, basis, orth, K, backend):
    """restarted_ca_lanczos.m:288-367 -- note ``while k <= maxiter`` => maxiter+1 blocks (:301)."""
    n = q.shape[0]
    nblk = maxiter + 1
    Q = np.zeros((n, nblk * s + 1), order="F")
    Q[:, 0] = q
    b = np.zeros(nblk + 1)
    T = None
    Qc = None if (Q_conv is None or np.size(Q_conv) == 0) else Q_conv
    for k in range(1, nblk + 1):
        if k > 1:
            q = Q[:, (k - 1) * s]
        V = _matrix_powers(A, q, s, Bk, basis, K)
        if k == 1:
            Q_, Rk, _ = K.normalize(V[:, : s + 1], backend=backend)
            Q[:, : s + 1], _ = K.projectAndNormalize([Qc], Q_, True, backend=backend)
            T = _mrdivide_upper(Rk @ Bk, Rk[:s, :s])
            b[0] = T[s, s - 1]
        else:
            Qprev = Q[:, (k - 2) * s : (k - 1) * s + 1]
            if orth == "local":
                Q_, Rk_ = K.projectAndNormalize([Qprev, Qc], V[:, 1 : s + 1], True, backend=backend)
                Q[:, (k - 1) * s + 1 : k * s + 1] = Q_
                Rkk_s, Rk_s = Rk_[0], Rk_[2]
            else:
                Q_, Rk_ = K.projectAndNormalize([Qprev], V[:, 1 : s + 1], True, backend=backend)
                Rkk_s, Rk_s = Rk_[0], Rk_[1]
                Qold = Q[:, : (k - 2) * s] if (k - 2) * s > 0 else None
                Q[:, (k - 1) * s + 1 : k * s + 1], _ = K.projectAndNormalize([Qc, Qold], Q_, True, backend=backend)
            T = _extend_T(T, b, k, s, Bk, Rkk_s, Rk_s)
    kk = nblk
    return Q[:, : s * (kk - 1)], np.asfortranarray(T[: s * (kk - 1) + 1, : s * (kk - 1)])


def restarted_periodic_reorth_needed(omega, k, s):
    """restarted_ca_lanczos.m:535-543 -- NOT the test of ca_lanczos.m:437-446: here the row maximum runs over the first i entries
    only, without abs(), and the threshold shrinks with the basis size, sqrt(eps/(k*s)).  Kept as written."""
    err = 0.0
    for i in range(1, s + 1):
        row = omega[(k - 1) * s + i, :i]                                     # omega((k-1)*s+i+1, 1:i)
        row_err = float(np.max(row))
        if row_err > err:
            err = row_err
    return err >= np.sqrt(EPS / (k * s)), err


def _restarted_lanczos_periodic_selective(A, Q_conv, q, Bk, maxiter, s, basis, orth, K, backend, norm_A, info=None):
    """lanczos_selective (restarted_ca_lanczos.m:369-463) and lanczos_periodic (:465-552): ``while k <= maxiter`` => maxiter+1
    blocks; every block is projected against {Qprev, Q_conv} (periodic, :500) or {Qprev, Q_conv, QR(:,1:nritz)} (selective, :405)."""
    n = q.shape[0]
    nblk = maxiter + 1
    Q = np.zeros((n, nblk * s + 1), order="F")
    Q[:, 0] = q
    b = np.zeros(nblk + 1)
    T = None
    Qc = None if (Q_conv is None or np.size(Q_conv) == 0) else Q_conv
    omega = None
    QR = np.zeros((n, 0), order="F")
    nritz = 0
    breaks = []
    for k in range(1, nblk + 1):
        if k > 1:
            q = Q[:, (k - 1) * s]
        V = _matrix_powers(A, q, s, Bk, basis, K)
        if k == 1:
            Q_, Rk, _ = K.normalize(V[:, : s + 1], backend=backend)                      # :394 / :489
            Q[:, : s + 1], _ = K.projectAndNormalize([Qc], Q_, True, backend=backend)    # :396 / :491
            T = _mrdivide_upper(Rk @ Bk, Rk[:s, :s])
            b[0] = T[s, s - 1]
        else:
            blocks = [Q[:, (k - 2) * s : (k - 1) * s + 1], Qc]
            if orth == "selective":
                blocks.append(QR[:, :nritz] if nritz > 0 else None)
            Q_, Rk_ = K.projectAndNormalize(blocks, V[:, 1 : s + 1], True, backend=backend)
            Q[:, (k - 1) * s + 1 : k * s + 1] = Q_
            T = _extend_T(T, b, k, s, Bk, Rk_[0], Rk_[-1])                               # Rk_{1}, Rk_{3} / Rk_{4}
        if orth == "periodic":
            alpha = np.diag(T, 0).copy()
            beta = np.diag(T, -1).copy()
            omega = update_omega(omega, alpha, beta, norm_A, s)                          # :532-534
            need, _ = restarted_periodic_reorth_needed(omega, k, s)
            if need:                                                                     # :543-546
                breaks.append(k)
                old = Q[:, : (k - 1) * s] if (k - 1) * s > 0 else None
                Q[:, (k - 1) * s : k * s + 1], _ = K.projectAndNormalize([old], Q[:, (k - 1) * s : k * s + 1], True, backend=backend)
                omega = reset_omega(omega, norm_A, s)
        else:
            Dp, Vp = np.linalg.eig(T[: s * k, : s * k])                                  # :436
            Vp = np.real(Vp)
            conv = b[k - 1] * np.abs(Vp[s * k - 1, :]) < norm_A * np.sqrt(EPS)           # :438-443
            if int(conv.sum()) > nritz:                                                  # :444-454
                breaks.append(k)
                nritz = int(conv.sum())
                QR = np.asfortranarray(Q[:, : k * s] @ Vp[:, conv])
                QR, _, _ = K.normalize(QR, backend=backend)
    if info is not None:
        info.setdefault("breaks", []).append(breaks)
        info.setdefault("nritz", []).append(nritz)
    kk = nblk
    return Q[:, : s * (kk - 1)], np.asfortranarray(T[: s * (kk - 1) + 1, : s * (kk - 1)])


def restarted_ca_lanczos(A, r, max_lanczos, n_wanted_eigs=10, s=6, basis="newton", orth="local", tol=1.0e-8,
                         K=_oracle_kernels, backend="tsqr", max_restarts=200, info=None):
    """restarted_ca_lanczos.m:4-202 (orth 'local' / 'full' :288-367, 'selective' :369-463, 'periodic' :465-552; restart strategy
    'largest' :52,:204-248).

    Returns (conv_eigs, Q_conv, num_restarts, rnorms, orth_err).
    """
    orth = str(orth).lower()
    if orth not in ("local", "full", "periodic", "selective"):
        raise ValueError("lanczos.m: Invalid option value for orth: " + orth)
    norm_A = normest(A)
    tol = tol * norm_A
    n = r.shape[0]
    q = r / np.linalg.norm(r)
    Bk = basis_matrix(A, q, s, basis, "local", K=K)
    Q = np.zeros((n, max_lanczos + n_wanted_eigs + s * (max_lanczos // s)), order="F")
    Q_conv = None
    conv_eigs, conv_rnorms = [], []
    rnorms = np.zeros((max_restarts, n_wanted_eigs))
    orth_err = []
    num_restarts = 0
    restart = True
    nconv = 0
    while restart and num_restarts < max_restarts:
        num_restarts += 1
        iters = max_lanczos // s
        if iters == 0:
            break                                   # reference branch uses undefined variables (:91-95)
        if orth in ("periodic", "selective"):                                            # :101-104
            Q_new, T = _restarted_lanczos_periodic_selective(A, Q_conv, q, Bk, iters, s, basis, orth, K, backend, norm_A, info)
        else:
            Q_new, T = _restarted_lanczos_basic(A, Q_conv, q, Bk, iters, s, basis,
                                                "local" if orth == "local" else "fro", K, backend)
        m = s * iters
        Dp, Vp = np.linalg.eig(T[:m, :m])           # non-symmetric T => general solver (Appendix B)
        Dp = np.real(Dp).copy(); Vp = np.real(Vp).copy()
        beta = T[m, m - 1]
        ritz_norms = beta * np.abs(Vp[m - 1, :])
        k = 0
        for i in range(m):
            if ritz_norms[i] < tol:
                Dp[[i, k]] = Dp[[k, i]]
                Vp[:, [i, k]] = Vp[:, [k, i]]
                ritz_norms[[i, k]] = ritz_norms[[k, i]]
                k += 1
        for i in range(k):
            Q[:, nconv + i] = Q_new @ Vp[:, i]
            conv_eigs.append(Dp[i]); conv_rnorms.append(ritz_norms[i])
        if num_restarts > 1:
            rnorms[num_restarts - 1, :nconv] = rnorms[num_restarts - 2, :nconv]
        for i in range(k):
            if nconv + i < n_wanted_eigs:
                l = conv_eigs[nconv + i]; x = Q[:, nconv + i]
                rnorms[num_restarts - 1, nconv + i] = np.linalg.norm(A @ x - l * x) / np.linalg.norm(l * x)
        rest = Dp[k:]
        ix = np.argsort(-rest, kind="stable")
        for i in range(max(0, n_wanted_eigs - nconv - k)):
            l = rest[ix[i]]; x = Q_new @ Vp[:, k + ix[i]]
            rnorms[num_restarts - 1, nconv + i + k] = np.linalg.norm(A @ x - l * x) / np.linalg.norm(l * x)
        Qall = Q_new if Q_conv is None else np.hstack([Q_conv, Q_new])
        orth_err.append(np.linalg.norm(np.eye(Qall.shape[1]) - Qall.T @ Qall, "fro"))
        nconv += k
        Q_conv = np.asfortranarray(Q[:, :nconv]) if nconv > 0 else None
        restart = not (len(conv_eigs) >= n_wanted_eigs)
        if restart:
            l = k                                  # generateStartVector 'largest' (:209-218)
            for j in range(k, m):
                if Dp[j] > Dp[l]:
                    l = j
            q = Q_new @ Vp[:, l]
            q = q / np.linalg.norm(q)
    conv_eigs = np.asarray(conv_eigs); conv_rnorms = np.asarray(conv_rnorms)
    ixs = np.argsort(-conv_eigs, kind="stable")
    keep = n_wanted_eigs if not restart else nconv
    conv_eigs = conv_eigs[ixs][:keep]
    Qc = Q_conv[:, ixs][:, :keep] if Q_conv is not None else np.zeros((n, 0))
    return conv_eigs, Qc, num_restarts, rnorms[:num_restarts], np.asarray(orth_err)
