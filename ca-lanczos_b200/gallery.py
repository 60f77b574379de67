"""Synthetic input matrices for the CA-Lanczos hot path.

All generators are rng-free or fixed-seed and return ``scipy.sparse.csr_matrix`` with
sorted int32 column indices and fp64 values.  The reference's own scripts build their inputs the same
way (``gallery('poisson',m)`` for C1, ``sparse(diag(linspace(1,10^k,N)))`` in
``test_convergence_diagonal_matrices.m:16-19`` for C2); C3..C5 are the synthetic scale-ups named in
BASELINE.json.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _stencil_csr(dims, diag, row_lo=0, row_hi=None):
    """CSR of the (2*d+1)-point Dirichlet Laplacian stencil on a ``dims`` grid (first dim fastest).

    Built directly in CSR (no COO/kron intermediate) so that 256^3 fits comfortably in host RAM.  With
    ``row_lo/row_hi`` only rows [row_lo,row_hi) are generated (shape (row_hi-row_lo) x n, global columns):
    what one rank of a row-partitioned run needs.
    """
    dims = tuple(int(d) for d in dims)
    ntot = int(np.prod(dims))
    row_hi = ntot if row_hi is None else int(row_hi)
    strides = np.cumprod((1,) + dims[:-1]).astype(np.int64)
    idx = np.arange(int(row_lo), row_hi, dtype=np.int64)
    n = idx.shape[0]
    coords = [(idx // st) % d for st, d in zip(strides, dims)]
    # neighbour offsets in ascending column order: -s_d .. -s_1, 0, +s_1 .. +s_d
    offs, valid = [], []
    for k in reversed(range(len(dims))):
        offs.append(-strides[k]); valid.append(coords[k] > 0)
    offs.append(0); valid.append(np.ones(n, dtype=bool))
    for k in range(len(dims)):
        offs.append(strides[k]); valid.append(coords[k] < dims[k] - 1)
    valid = np.stack(valid, axis=1)                      # n x (2d+1), row-major == CSR order
    counts = valid.sum(axis=1)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    cols = (idx[:, None] + np.asarray(offs, dtype=np.int64)[None, :])[valid]
    vals = np.where(np.asarray(offs) == 0, float(diag), -1.0)
    data = np.broadcast_to(vals[None, :], valid.shape)[valid].astype(np.float64)
    it = np.int32 if indptr[-1] < 2**31 else np.int64
    A = sp.csr_matrix((data, cols.astype(np.int32), indptr.astype(it)), shape=(n, ntot))
    A.has_sorted_indices = True
    return A


def poisson2d(m: int) -> sp.csr_matrix:
    """``gallery('poisson',m)``: kron(I,T)+kron(T,I), T=tridiag(-1,2,-1); n=m^2 (C1: m=100)."""
    return _stencil_csr((m, m), 4.0)


def laplace3d(m: int, my: int | None = None, mz: int | None = None, row_lo: int = 0,
              row_hi: int | None = None) -> sp.csr_matrix:
    """7-point Dirichlet Laplacian on an m x my x mz grid, x fastest; diag 6, off -1 (C3: m=256).
    ``row_lo/row_hi``: generate only that row range (rectangular result, global column indices)."""
    my = m if my is None else my
    mz = m if mz is None else mz
    return _stencil_csr((m, my, mz), 6.0, row_lo, row_hi)


def leja_points(a: float, b: float, s: int) -> np.ndarray:
    """s Chebyshev points of [a,b] in Leja order (max-product greedy): the textbook Newton-basis shifts when
    the spectral interval is known (bench.py uses them for the 7-point Laplacian, spectrum in (0,12))."""
    k = np.arange(s)
    x = 0.5 * (a + b) + 0.5 * (b - a) * np.cos((2 * k + 1) * np.pi / (2 * s))
    out = [int(np.argmax(np.abs(x)))]
    rest = [i for i in range(s) if i != out[0]]
    while rest:
        prod = [np.prod(np.abs(x[i] - x[out])) for i in rest]
        j = rest[int(np.argmax(prod))]
        out.append(j); rest.remove(j)
    return x[out]


def diag_linspace(n: int, K: float = 100.0) -> sp.csr_matrix:
    """``sparse(diag(linspace(1,K,n)))`` (test_convergence_diagonal_matrices.m:16-19; C2: n=1e6, K=100)."""
    d = np.linspace(1.0, float(K), int(n))
    idx = np.arange(n, dtype=np.int32)
    return sp.csr_matrix((d, idx, np.arange(n + 1, dtype=np.int32)), shape=(n, n))


def powerlaw_spd(n: int, avg_deg: float = 20.0, seed: int = 0, alpha: float = 0.8) -> sp.csr_matrix:
    """Symmetric power-law SPD matrix (C4): edges (i,j), i ~ Zipf-weighted rank^-alpha, j uniform,
    symmetrised, values -U(0,1), diagonal = sum|row| + 1 (strictly diagonally dominant => SPD)."""
    rng = np.random.default_rng(seed)
    n = int(n)
    m = int(n * avg_deg / 2)
    w = np.arange(1, n + 1, dtype=np.float64) ** (-alpha)
    cdf = np.cumsum(w); cdf /= cdf[-1]
    i = np.searchsorted(cdf, rng.random(m)).astype(np.int64)
    j = rng.integers(0, n, size=m, dtype=np.int64)
    v = -rng.random(m)
    keep = i != j
    i, j, v = i[keep], j[keep], v[keep]
    B = sp.coo_matrix((v, (i, j)), shape=(n, n)).tocsr()
    B = B + B.T
    B.sum_duplicates()
    d = np.asarray(abs(B).sum(axis=1)).ravel() + 1.0
    A = (B + sp.diags(d)).tocsr()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32)
    if A.indptr[-1] < 2**31:
        A.indptr = A.indptr.astype(np.int32)
    return A


def powerlaw_spd_rows(n: int, avg_deg: float = 20.0, seed: int = 0, alpha: float = 0.8, row_lo: int = 0, row_hi: int | None = None,
                      shuffle: bool = True, chunk: int = 1 << 23, jacobi: bool = False) -> sp.csr_matrix:
    """Rows [row_lo,row_hi) of the C4 matrix at scale (n = 2e7, nnz ~ 4e8), generated WITHOUT the other rows: the edge
    list is produced chunk by chunk from a counter-seeded generator (every rank generates the same chunks and keeps the
    edges that touch its rows), so host memory stays bounded by the rank's own share.

    Edges (i,j): i = pi(floor(n*u^(1/(1-alpha)))) -- the continuous inverse CDF of weights rank^-alpha -- j uniform, values
    -U(0,1) on a 2^-30 grid (sums of duplicates and of whole rows are then exact, i.e. independent of the order in which a
    rank meets the edges: the matrix is exactly symmetric and exactly the same for every partition), symmetrised, duplicates summed, diagonal = sum|row| + 1 (strictly diagonally dominant => SPD).  ``shuffle``
    relabels the vertices by the affine permutation pi(r) = (a*r + b) mod n (a coprime to n): the spectrum is unchanged, but
    the hubs are spread over the row partition instead of all landing on rank 0 (the contiguous-block partition of the north
    star is fixed; a power-law graph in degree order would put ~40 % of the non-zeros on the first of 8 ranks).

    ``jacobi``: return D^-1/2 A D^-1/2 (D = diag A; unit diagonal, spectrum in (0,2), still SPD, same pattern).  The unscaled
    matrix is L + I with L a weighted graph Laplacian: ones(n,1) is an exact eigenvector and one eigenvalue per hub dwarfs the
    rest of the spectrum, on which the s = 6 basis blocks of the reference are numerically rank deficient (oracle: rank 3 of 6
    in the second block) -- the scaled matrix is the one bench.py --config c4 solves."""
    n = int(n)
    row_hi = n if row_hi is None else int(row_hi)
    m = int(n * avg_deg / 2)
    a_mul = 0
    if shuffle:
        a_mul = int(0.6180339887 * n) | 1
        while np.gcd(a_mul, n) != 1:
            a_mul += 2
    b_add = n // 3
    rows, cols, vals = [], [], []
    expo = 1.0 / (1.0 - alpha)
    dsum = np.zeros(n) if jacobi else None          # sum |row| of EVERY row (the scaling needs the diagonal of every column)
    for c0 in range(0, m, chunk):
        cnt = min(chunk, m - c0)
        rng = np.random.default_rng([int(seed), c0 // chunk])
        i = np.minimum((n * rng.random(cnt) ** expo).astype(np.int64), n - 1)
        j = rng.integers(0, n, size=cnt, dtype=np.int64)
        v = -(rng.integers(1, 1 << 30, size=cnt, dtype=np.int64).astype(np.float64) * (1.0 / (1 << 30)))
        if shuffle:
            i = (i * a_mul + b_add) % n
        keep = i != j
        if jacobi:
            dsum += np.bincount(i[keep], weights=-v[keep], minlength=n) + np.bincount(j[keep], weights=-v[keep], minlength=n)
        mi = keep & (i >= row_lo) & (i < row_hi)
        mj = keep & (j >= row_lo) & (j < row_hi)
        rows += [i[mi], j[mj]]; cols += [j[mi], i[mj]]; vals += [v[mi], v[mj]]
    rows = np.concatenate(rows) - row_lo; cols = np.concatenate(cols); vals = np.concatenate(vals)
    B = sp.coo_matrix((vals, (rows, cols)), shape=(row_hi - row_lo, n)).tocsr()      # sums duplicates
    del rows, cols, vals
    d = -np.asarray(B.sum(axis=1)).ravel() + 1.0                                     # all off-diagonal values are negative
    nloc = row_hi - row_lo
    if jacobi:
        di = 1.0 / np.sqrt(dsum + 1.0)
        B.sort_indices()
        rr = np.repeat(np.arange(row_lo, row_hi, dtype=np.int64), np.diff(B.indptr))
        B.data = B.data * (di[rr] * di[B.indices])                                   # symmetric: the two factors commute
        d = np.ones(nloc)
    D = sp.csr_matrix((d, np.arange(row_lo, row_hi, dtype=np.int64), np.arange(nloc + 1, dtype=np.int64)), shape=(nloc, n))
    A = (B + D).tocsr()
    A.sort_indices()
    A.indices = A.indices.astype(np.int32)
    return A


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def tall_skinny(n: int, c: int, seed: int = 0, row0: int = 0) -> np.ndarray:
    """C5 input: n x c (Fortran order), entry (i,j) = U(-1,1) from splitmix64(seed, (row0+i)*c+j),
    column j scaled by 2^-j.  Counter-based, so any row slice can be generated independently."""
    with np.errstate(over="ignore"):
        i = np.arange(row0, row0 + n, dtype=np.uint64)[:, None]
        j = np.arange(c, dtype=np.uint64)[None, :]
        ctr = (i * np.uint64(c) + j) + np.uint64(seed) * np.uint64(0xD1B54A32D192ED03)
        bits = _splitmix64(ctr)
    u = (bits >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    X = (2.0 * u - 1.0) * (2.0 ** -np.arange(c, dtype=np.float64))[None, :]
    return np.asfortranarray(X)


def tall_skinny_device(n, c, dev, ld, row0=0, seed=0):
    """``tall_skinny`` generated ON the device with torch integer ops (same counter-based generator, bit-identical; C5 at
    n = 1e8 would take minutes and 13.6 GB of PCIe from the host): a (c, ld) row-major tensor == n x c column-major, ld >= n."""
    import torch
    X = torch.zeros((c, ld), dtype=torch.float64, device=dev)
    M64 = (1 << 64) - 1

    def s64(v):      # python int (mod 2^64) -> signed int64 value
        v &= M64
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(z, k):   # logical shift right on int64 tensors
        return (z >> k) & ((1 << (64 - k)) - 1)
    chunk = 1 << 24
    for j in range(c):
        for r0 in range(0, n, chunk):
            m = min(chunk, n - r0)
            i = torch.arange(row0 + r0, row0 + r0 + m, dtype=torch.int64, device=dev)
            z = i * c + j + s64(seed * 0xD1B54A32D192ED03) + s64(0x9E3779B97F4A7C15)
            z = (z ^ lsr(z, 30)) * s64(0xBF58476D1CE4E5B9)
            z = (z ^ lsr(z, 27)) * s64(0x94D049BB133111EB)
            z = z ^ lsr(z, 31)
            u = lsr(z, 11).to(torch.float64) * (1.0 / (1 << 53))
            X[j, r0:r0 + m] = (2.0 * u - 1.0) * (2.0 ** -j)
    return X
