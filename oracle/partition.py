"""CPU oracle for the multi-GPU plan: 1-D block-row partition, level-s ghost zones, P-way MPK.
TEST INFRASTRUCTURE ONLY.

The reference has no partitioning at all (single MATLAB process); the scheme is Hoemmen's PA1, which
ca_lanczos.m:3-5 cites as its source.  This module states the integer objects the CUDA
library must reproduce BIT-EXACTLY -- row bounds, ghost index sets, per-peer exchange lists -- using a
boolean sparse closure in scipy, and a P-way matrix powers kernel that must equal the 1-way oracle on the
owned rows.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import kernels


def row_bounds(n: int, P: int) -> np.ndarray:
    """GPU p owns rows [floor(p*n/P), floor((p+1)*n/P)).  Returns the P+1 bounds (int64)."""
    return np.array([(p * n) // P for p in range(P + 1)], dtype=np.int64)


def level_sets(A: sp.csr_matrix, lo: int, hi: int, s: int):
    """R_0 = owned rows; R_k = R_{k-1} U {j : a_ij != 0, i in R_{k-1}} (pattern graph, stored entries).
    Returns ``level`` (int32, length n; level[j] = smallest k with j in R_k, or -1 beyond level s)."""
    n = A.shape[0]
    pat = sp.csr_matrix((np.ones(A.nnz, dtype=np.int8), A.indices, A.indptr), shape=A.shape)
    level = np.full(n, -1, dtype=np.int32)
    level[lo:hi] = 0
    reach = np.zeros(n, dtype=bool); reach[lo:hi] = True
    frontier = reach.copy()
    for k in range(1, s + 1):
        # columns touched by the rows in the frontier
        nxt = (pat.T @ frontier.astype(np.int32)) > 0
        new = nxt & ~reach
        level[new] = k
        reach |= new
        frontier = new
        if not new.any():
            break
    return level


def ghost_indices(A, lo, hi, s):
    """ghost_s(p) = R_s(p) \\ R_0(p), sorted ascending (int64)."""
    level = level_sets(A, lo, hi, s)
    return np.flatnonzero(level > 0).astype(np.int64)


def exchange_lists(A, P, s):
    """recv[p][q] = ghost_s(p) ∩ owned(q), sorted ascending -- what q sends to p once per block."""
    n = A.shape[0]
    b = row_bounds(n, P)
    out = []
    for p in range(P):
        g = ghost_indices(A, int(b[p]), int(b[p + 1]), s)
        out.append([g[(g >= b[q]) & (g < b[q + 1])] for q in range(P)])
    return out


def local_problem(A, lo, hi, s):
    """Local matrix of rank p: rows/cols restricted to R_s(p), local order = ascending GLOBAL index.
    Returns (A_loc csr, loc2glob int64, level int32 per local row, own_off = local index of row ``lo``)."""
    level = level_sets(A, lo, hi, s)
    loc2glob = np.flatnonzero(level >= 0).astype(np.int64)
    A_loc = A[loc2glob][:, loc2glob].tocsr()
    A_loc.sort_indices()
    own_off = int(np.searchsorted(loc2glob, lo))
    return A_loc, loc2glob, level[loc2glob], own_off


def choose_halo_level(A, P, s):
    """Depth L of the ghost closure = MPK steps per halo exchange, as the library picks it when all rows are supplied:
    per rank the deepest level <= s whose cumulative ghost count does not exceed the owned rows (at least 1), then the
    minimum over the ranks.  L = s is PA1 (one exchange per block, stencils); L = 1 the per-step exchange (power-law)."""
    n = A.shape[0]
    b = row_bounds(n, P)
    L = s
    for p in range(P):
        lo, hi = int(b[p]), int(b[p + 1])
        level = level_sets(A, lo, hi, s)
        best, ghosts = 1, 0
        for k in range(1, s + 1):
            ghosts += int(np.count_nonzero(level == k))
            if ghosts <= hi - lo:
                best = k
        L = min(L, best)
    return max(L, 1)


def mpk_partitioned(A, v, s, lam, P, basis="newton", halo_level=None):
    """P-way redundant-ghost MPK.  ``halo_level`` = L (default s = PA1): every rank gets the current basis vector on
    R_L(p), runs L steps on its local matrix with no further exchange (step j of a group on the rows of level <= L-j),
    keeps the owned rows, and exchanges again.  Returns the assembled n x (s+1) basis (first column v for both bases)."""
    n = A.shape[0]
    L = s if halo_level is None else max(1, min(int(halo_level), s))
    b = row_bounds(n, P)
    V = np.zeros((n, s + 1), order="F")
    V[:, 0] = v
    lam = None if lam is None else np.asarray(lam)
    for k0 in range(0, s, L):                            # steps k0+1 .. k0+g
        g = min(L, s - k0)
        for p in range(P):
            lo, hi = int(b[p]), int(b[p + 1])
            A_loc, l2g, lev, off = local_problem(A, lo, hi, L)
            v_loc = V[l2g, k0]                           # the halo exchange of column k0
            if basis == "newton":
                V_loc = kernels.matrix_powers_newton(A_loc, v_loc, g, lam[k0:k0 + g], 1)
            else:
                V_loc = np.column_stack([v_loc, kernels.matrix_powers_monomial(A_loc, v_loc, g)])
            V[lo:hi, k0 + 1:k0 + g + 1] = V_loc[off:off + (hi - lo), 1:g + 1]
    return V
