"""CPU restatement of the reference's Newton-basis set-up (host side, O(s^2)).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (no Octave/MATLAB here, no goldens in the reference).  Follows leja.m:23-31,
real_leja.m:18-87, count_multiplicities.m:5-41, modified_leja.m:24-196, newton_basis_matrix.m:13-60
operation by operation (explicit left-to-right products instead of numpy reductions).
"""
from __future__ import annotations

import numpy as np


def _matlab_sort_key(x):
    """MATLAB sorts/uniques complex numbers by abs, then angle; reals by value."""
    x = np.asarray(x)
    if np.iscomplexobj(x) and np.any(x.imag != 0):
        return np.lexsort((np.angle(x), np.abs(x)))
    return np.argsort(x.real, kind="stable")


def count_multiplicities(x, n):
    """count_multiplicities.m:5-41 -- ``unique(x,'first')``: sorted unique values + multiplicities."""
    x = np.asarray(x).ravel()
    order = _matlab_sort_key(x)
    val = x[order]
    first = np.ones(n, dtype=bool)
    first[1:] = val[1:] != val[:-1]
    y = val[first]
    num_unique = y.shape[0]
    if num_unique == n:
        return y, np.ones(n), num_unique
    ii = np.flatnonzero(first)                       # 0-based index of first occurrence in sorted x
    mults = np.zeros(num_unique)
    for k in range(num_unique - 1):
        mults[k] = ii[k + 1] - ii[k]
    mults[num_unique - 1] = n - ii[num_unique - 1]
    return y, mults, num_unique


def _is_conj_pair(a, b):
    """modified_leja.m:26-39."""
    return (np.real(a) == np.real(b)) and (np.imag(a) == -np.imag(b)) and (np.imag(a) != 0)


def _prod(vals):
    p = 1.0
    for v in vals:
        p = p * v
    return p


def modified_leja(x, n, mults):
    """modified_leja.m:24-196 -- modified Leja ordering with the running capacity rescale.

    ``n`` is passed by real_leja as the ORIGINAL count (real_leja.m:87) while ``x`` holds the unique
    values: with repeated shifts ``x(1:n)`` (:49) indexes past the end and MATLAB raises -- so does this
    restatement (IndexError).  Returns (y, outidx) with 0-based ``outidx``.  The points come back rescaled-then-unscaled by the capacity
    estimate (:113-114,:192), i.e. equal to the inputs only up to a few ulp -- kept on purpose.
    """
    x = np.array(x).ravel().copy()
    mults = np.asarray(mults, dtype=np.float64).ravel()
    nx = x.shape[0]
    if n > nx:
        raise IndexError("Index exceeds matrix dimensions: x(1:n) with n=%d > numel(x)=%d (repeated shifts)" % (n, nx))

    # --- modified_leja_start (:41-78)
    if n < 1:
        return np.array([]), []
    if n == 1:
        y, outidx = [x[0]], [0]
    else:
        j = int(np.argmax(np.abs(x[:n])))
        if np.imag(x[j]) == 0:
            y, outidx = [x[j]], [j]
        elif j > 0 and _is_conj_pair(x[j - 1], x[j]):
            if np.imag(x[j - 1]) < 0:
                raise ValueError("Complex conjugate pair out of order at indices %d and %d" % (j, j + 1))
            y, outidx = [x[j - 1], x[j]], [j - 1, j]
        elif j < n - 1 and _is_conj_pair(x[j], x[j + 1]):
            if np.imag(x[j]) < 0:
                x[j] = np.real(x[j]); x[j + 1] = np.real(x[j + 1])
            y, outidx = [x[j], x[j + 1]], [j, j + 1]
        else:
            raise ValueError("Complex shift, not in a pair")
    inidx = [i for i in range(n) if i not in outidx]

    # --- modified_leja_helper (:80-181), recursion unrolled; first level has nargin<8 => capacity=1
    y = list(y)
    num_points = len(outidx)
    capacity = 1.0
    first_level = True
    while len(inidx) > 0:
        if not first_level and num_points > 1:
            old_capacity = capacity
            y_last = y[num_points - 1]
            prev = outidx[: num_points - 1]
            capacity = _prod([abs(y_last - x[i]) ** (mults[i] * (1.0 / num_points)) for i in prev])
            scale = capacity / old_capacity
            x = x / scale
            y = [yy / scale for yy in y]
        first_level = False
        zprod = []
        for j in inidx:
            zprod.append(_prod([(abs(x[j] - x[i]) / capacity) ** mults[i] for i in outidx]))
        k = int(np.argmax(np.asarray(zprod)))
        max_zprod = zprod[k]
        j = inidx[k]
        if max_zprod == 0:
            raise ValueError("Product to maximize is zero; either there are multiple shifts, or the product underflowed")
        if np.isinf(max_zprod):
            raise ValueError("Product to maximize is Inf; must have overflowed")
        if np.imag(x[j]) == 0:
            take = [j]
        elif j > 0 and _is_conj_pair(x[j - 1], x[j]):
            if np.imag(x[j - 1]) < 0:
                raise ValueError("Complex conjugate pair out of order at indices %d and %d" % (j, j + 1))
            take = [j - 1, j]
        elif j < n - 1 and _is_conj_pair(x[j], x[j + 1]):
            if np.imag(x[j]) < 0:
                raise ValueError("Complex conjugate pair out of order at indices %d and %d" % (j + 1, j + 2))
            take = [j, j + 1]
        else:
            raise ValueError("Complex shift, not in a pair")
        inidx = [i for i in inidx if i not in take]
        outidx = outidx + take
        y = y + [x[t] for t in take]
        num_points += len(take)
    y = np.asarray(y) * capacity
    if np.iscomplexobj(y) and np.all(y.imag == 0):
        y = y.real
    return y, outidx


def real_leja(x):
    """real_leja.m:18-87 -- uniquify, sort by real part, fix conjugate-pair order, modified Leja."""
    x = np.asarray(x).ravel()
    n = x.shape[0]
    y, mults, num_unique = count_multiplicities(x, n)
    order = np.argsort(np.real(y), kind="stable")
    y = np.array(y[order]); mults = np.asarray(mults)[order]
    k = 0
    while k < num_unique - 1:
        if np.imag(y[k]) != 0:
            if np.real(y[k]) == np.real(y[k + 1]) and np.imag(y[k]) == -np.imag(y[k + 1]):
                re, im = np.real(y[k]), abs(np.imag(y[k]))
                y[k] = re + 1j * im
                y[k + 1] = re - 1j * im
                k += 2
            else:
                # the reference only prints here and never advances k (infinite loop): fail loudly
                raise ValueError("Error in real_leja, complex numbers.")
        else:
            k += 1
    return modified_leja(y, n, mults)


def leja(x, which=None):
    """leja.m:23-31.  QUIRK kept: ANY second argument -- the callers pass 'nonmodified'
    (ca_lanczos.m:70, restarted_ca_lanczos.m:69) -- takes the real_leja / MODIFIED branch."""
    if which is None:
        raise NotImplementedError("1-arg leja -> nonmodified_leja.m (only ca_lanczos_prop.m; out of scope)")
    return real_leja(x)


def newton_basis_matrix(lam, s, modifiedp=0):
    """newton_basis_matrix.m:13-60 -- (s+1) x s change-of-basis matrix: B(k,k)=lam_k, B(k+1,k)=1;
    modified: real parts on the diagonal and B(k-1,k) = -imag(lam_k)^2 on the conjugate's column."""
    lam = np.asarray(lam).ravel()
    B = np.zeros((s + 1, s), order="F")
    for k in range(s):
        shift = lam[k]
        if modifiedp == 0:
            B[k, k] = np.real(shift) if np.imag(shift) == 0 else shift
        else:
            if np.imag(shift) > 0:
                if k == s - 1 or lam[k] != np.conj(lam[k + 1]):
                    raise ValueError("Modified Leja ordering broken at k = %d" % (k + 1))
                B[k, k] = np.real(shift)
            elif np.imag(shift) < 0:
                if k == 0 or lam[k - 1] != np.conj(lam[k]):
                    raise ValueError("Modified Leja ordering broken at k = %d" % (k + 1))
                B[k, k] = np.real(shift)
                B[k - 1, k] = -np.imag(shift) ** 2
            else:
                B[k, k] = np.real(shift)
        B[k + 1, k] = 1.0
    return B
