"""CPU restatement of the reference's block kernels (the hot path).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference is 100 % MATLAB, Octave/MATLAB are not in the image, and the reference
ships no golden vectors for this path.  Every function below follows the cited
reference lines operation by operation on numpy/scipy (LAPACK ``geqrf/orgqr``, ``potrf``, ``gesdd``, CSR
mat-vec), which is what MATLAB's built-ins call as well.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.
Arrays are fp64; matrices are column-major (Fortran order) like MATLAB's.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp


# ----------------------------------------------------------------------------- MPK
def SpMV(A, v):
    """SpMV.m:6-8 -- ``Av = A*v``."""
    return A @ v


def matrix_powers_monomial(A, q, s):
    """matrix_powers_monomial.m:6-12 -- V = [A q, A^2 q, ..., A^s q]  (n x s, q itself excluded)."""
    n = q.shape[0]
    V = np.zeros((n, s), order="F", dtype=np.result_type(q.dtype, np.float64))
    V[:, 0] = A @ q
    for i in range(1, s):
        V[:, i] = A @ V[:, i - 1]
    return V


def matrix_powers_newton(A, v, s, lam, modifiedp=0):
    """matrix_powers_newton.m:15-54 -- V(:,1)=v; V(:,k+1) = A V(:,k) - lam(k) V(:,k) (n x (s+1)).

    ``modifiedp=1`` (what the drivers pass, ca_lanczos.m:116): real part of a complex shift only, and
    ``+ imag(lam_k)^2 V(:,k-1)`` on the second member of a conjugate pair (:40-41).  Operation order as
    in the reference: full SpMV first, then the shift AXPY (two roundings per element).
    """
    lam = np.asarray(lam).ravel()
    n = v.shape[0]
    cplx = np.iscomplexobj(v) or (modifiedp == 0 and np.iscomplexobj(lam) and np.any(lam.imag != 0))
    V = np.zeros((n, s + 1), order="F", dtype=np.complex128 if cplx else np.float64)
    V[:, 0] = v
    if modifiedp == 0:
        for k in range(s):
            w = SpMV(A, V[:, k])
            lk = lam[k] if cplx else np.real(lam[k])
            V[:, k + 1] = w - lk * V[:, k]
    else:
        for k in range(s):
            w = SpMV(A, V[:, k])
            im = float(np.imag(lam[k]))
            re = float(np.real(lam[k]))
            if im > 0:
                V[:, k + 1] = w - re * V[:, k]
            elif im < 0:
                if k == 0:
                    raise ValueError("k==1, but shift %e has a negative imaginary part" % re)
                V[:, k + 1] = w - re * V[:, k] + im ** 2 * V[:, k - 1]
            else:
                V[:, k + 1] = w - re * V[:, k]
    return V


# ----------------------------------------------------------------------------- QR backends
def tsqr(A):
    """tsqr.m:7-12 -- ``[Q,R]=qr(A,0)`` then ``d=sign(diag(R)); R=diag(d)*R; Q=Q*diag(d)``.

    ``sign(0)=0`` zeroes the column/row of an exactly zero pivot, as in the reference.
    """
    Q, R = np.linalg.qr(np.asarray(A), mode="reduced")
    d = np.sign(np.diag(R))
    R = d[:, None] * R
    Q = Q * d[None, :]
    return np.asfortranarray(Q), np.asfortranarray(R)


def cholqr(X):
    """cholqr.m:3-8 -- ``G=X'*X; R=chol(G); Q=X/R``.  Raises LinAlgError when G is not PD (MATLAB: error)."""
    X = np.asarray(X)
    G = X.T @ X
    R = np.linalg.cholesky(G).T            # upper factor, G = R'R
    Q = sla.solve_triangular(R, X.T, trans="T", lower=False).T   # Q R = X
    return np.asfortranarray(Q), np.asfortranarray(R)


def _chol_shifted(G, adaptive=True):
    """Upper Cholesky factor; on breakdown retry with G + shift*I (shift = 100*c*eps*max(diag G), x100 per retry)."""
    c = G.shape[0]
    shift, nshift = 0.0, 0
    while True:
        try:
            return np.linalg.cholesky(G + shift * np.eye(c)).T, nshift
        except np.linalg.LinAlgError:
            if not adaptive or nshift >= 6:
                raise
            nshift += 1
            shift = 100.0 * c * 2.220446049250313e-16 * np.max(np.diag(G)) if nshift == 1 else shift * 100.0


def cholqr2(X, inv_thresh=32.0):
    """NOT in the reference: cholqr.m made robust -- the restatement of libcalz' CALZ_QR_CHOLQR2 backend.  One CholQR
    pass (shifted if the Cholesky breaks down); while min_j R_jj/||x_j|| < 1/inv_thresh or a shift was needed, repeat
    CholQR on Q (at most 3 more passes), R = R_k ... R_1.  Checked against tsqr (Householder) in the tests."""
    X = np.asarray(X)

    def one_pass(Y):
        G = Y.T @ Y
        R, nshift = _chol_shifted(G)
        Q = sla.solve_triangular(R, Y.T, trans="T", lower=False).T
        again = nshift > 0 or np.min(np.diag(R) / np.sqrt(np.diag(G))) < 1.0 / inv_thresh
        return Q, R, again

    Q, R, again = one_pass(X)
    for _ in range(3):
        if not again:
            break
        Q, R2, again = one_pass(Q)
        R = R2 @ R
    return np.asfortranarray(Q), np.asfortranarray(R)


def normalize(X, opt="None", tol=1.0e-8, backend="tsqr"):
    """normalize.m:3-36 -- QR (tsqr.m:7 at the :14 seam; ``backend='cholqr'`` selects cholqr.m there),
    ``svd(R)``, numerical rank = #{sigma_i > tol*sigma_1} counted up to the first failure (:17-24).

    The ``'randomizeNullSpace'`` option (:28-31, uses ``rand``) is never selected by any in-scope caller
    and is not restated.
    """
    ncols = X.shape[1]
    Q, R = {"tsqr": tsqr, "cholqr": cholqr, "cholqr2": cholqr2}[backend](X)
    S = np.linalg.svd(R, compute_uv=False)
    abs_tol = tol * S[0]
    rank = ncols
    for i in range(ncols):
        if S[i] <= abs_tol:
            rank = i
            break
    if rank != ncols and str(opt).lower() == "randomizenullspace":
        raise NotImplementedError("randomizeNullSpace (normalize.m:28-31,38-51) is out of scope")
    return Q, R, rank


# ----------------------------------------------------------------------------- block Gram-Schmidt
def _isempty(Qi):
    return Qi is None or np.size(Qi) == 0


def project(Q, X, doreorth=False):
    """project.m:7-58 -- for each non-empty block: ``R{i}=Q{i}'*X; X=X-Q{i}*R{i}`` (sequential over blocks).

    ``doreorth=True`` restates :40-57 including the inverted criterion ``max(0.5*before-after) < 0``
    (only reached from restarted_lanczos.m, out of scope).  Returns (X, list of R blocks); an empty block
    gives ``None`` (MATLAB ``[]``).
    """
    if not isinstance(Q, (list, tuple)):
        raise TypeError("Input Q (arg 1) to project() must be cell (block) array.")
    X = np.array(X, order="F", copy=True)
    if len(Q) == 0:
        return X, []
    R = [None] * len(Q)
    if doreorth:
        normBefore = np.sqrt(np.sum(X * X, axis=0))
    for i, Qi in enumerate(Q):
        if not _isempty(Qi):
            R[i] = Qi.T @ X
            X = X - Qi @ R[i]
    if doreorth:
        normAfter = np.sqrt(np.sum(X * X, axis=0))
        if np.max(0.5 * normBefore - normAfter) < 0:
            for i, Qi in enumerate(Q):
                if not _isempty(Qi):
                    R2 = Qi.T @ X
                    X = X - Qi @ R2
                    R[i] = R[i] + R2
    return np.asfortranarray(X), R


def projectAndNormalize(Q, X, doreorth=True, backend="tsqr", info=None):
    """projectAndNormalize.m:3-90.

    norms before (:17-22) -> project (:25) -> normalize (:26) -> norms after from the columns of R (:45-48)
    -> if ``max(|nb-na|./nb) > .5`` (:52): second pass ON THE UN-NORMALISED Y (:63-65), coefficient sum
    ``RZ{i}=RZ{i}+RY{i}`` (:71-73), ``RZ{end}`` = R of the second normalize.  ``info`` (optional dict)
    receives ``second_pass`` and ``rank`` instead of the reference's ``disp('second')`` (:62).
    """
    tol = 0.5
    X = np.asarray(X)
    ncols = X.shape[1]
    nb = len(Q)
    normsBeforeFirst = np.zeros(ncols)
    if doreorth:
        for i in range(ncols):
            normsBeforeFirst[i] = np.sqrt(np.sum(X[:, i] ** 2))
    Y, RY = project(Q, X, False)
    QY, R_, rank = normalize(Y, backend=backend)
    RY = list(RY) + [R_]
    second = False
    if doreorth:
        normsAfterFirst = np.zeros(ncols)
        for i in range(ncols):
            normsAfterFirst[i] = np.sqrt(np.sum(RY[nb][:, i] ** 2))
        with np.errstate(divide="ignore", invalid="ignore"):
            # MATLAB's max() skips NaN (0/0 for an all-zero column); fmax does the same
            second = bool(np.fmax.reduce(np.abs(normsBeforeFirst - normsAfterFirst) / normsBeforeFirst) > tol)
        if not second:
            QZ, RZ = QY, RY
        else:
            Z, RZ = project(Q, Y, False)
            QZ, R_, rank = normalize(Z, backend=backend)
            RZ = list(RZ) + [R_]
            for i in range(nb):
                if RZ[i] is not None:
                    RZ[i] = RZ[i] + RY[i]
    else:
        QZ, RZ = QY, RY
    if info is not None:
        info["second_pass"] = second
        info["rank"] = rank
    return QZ, RZ
