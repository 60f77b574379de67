"""Host-side mirror of the reference's call surface for the hot path, backed by libcalz.so (CUDA, sm_100a).

The reference is MATLAB; its drivers reach the block kernels by plain function name.  MATLAB/Octave are not
available here, so the host side is written in Python with the SAME names, argument order, defaults and error
behaviour -- ``matrix_powers_monomial``, ``matrix_powers_newton``, ``SpMV``, ``tsqr``, ``cholqr``, ``normalize``,
``project``, ``projectAndNormalize`` -- and talks to the library exclusively through the C ABI (include/calz.h),
exactly what the MEX gateways in mex/ do.  Host arrays in, fresh host arrays out (MATLAB value semantics);
the sparse matrix is uploaded once and cached across calls (keyed on the host object, like the gateway keys
on ``mxGetPr(A)``).

There is no CPU fallback: every function raises if the CUDA library or a device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import CalzError, check

__all__ = [
    "Context", "DeviceMatrix", "DeviceBlock", "default_context", "set_qr_backend", "get_qr_backend",
    "SpMV", "matrix_powers_monomial", "matrix_powers_newton", "tsqr", "cholqr", "normalize", "project",
    "projectAndNormalize", "CalzError",
]

_QR_BACKEND = "tsqr"          # the reference's normalize.m:14 calls tsqr


def set_qr_backend(name: str):
    """Select the QR used at the normalize.m:14 seam: 'tsqr' (reference default) or 'cholqr' (cholqr.m)."""
    global _QR_BACKEND
    if name not in _lib.QR:
        raise ValueError("backend must be 'tsqr', 'cholqr' or 'cholqr2'")
    _QR_BACKEND = name


def get_qr_backend() -> str:
    return _QR_BACKEND


def _dp(a):
    return a.ctypes.data_as(_lib.c_dp)


def _f64_fortran(a, copy=False):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if copy or not a.flags.f_contiguous:
        a = np.array(a, dtype=np.float64, order="F", copy=True)
    return a


class Context:
    """One per process / per GPU (calz_ctx): device, stream, scratch, optional NCCL communicator."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.calz_init(int(device), C.byref(h)))
        self.h = h
        self.device = int(device)
        self.rank, self.nranks = 0, 1
        for kv in filter(None, os.environ.get("CALZ_OPTS", "").split(",")):      # e.g. CALZ_OPTS=mpk_persistent=0,p2p=0
            k, v = kv.split("=")
            self.set_option(k.strip(), int(v))

    def close(self):
        if getattr(self, "h", None):
            self.lib.calz_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.calz_sync(self.h), self.h)

    @property
    def stream(self) -> int:
        return int(self.lib.calz_get_stream(self.h) or 0)

    def set_stream(self, cuda_stream_ptr: int):
        check(self.lib.calz_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), self.h)

    def set_option(self, key: str, value: int):
        check(self.lib.calz_set_option(self.h, key.encode(), int(value)), self.h)

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.calz_launch_count(self.h, 1 if reset else 0))

    def init_comm(self, nranks: int, rank: int, unique_id: bytes):
        path = _lib.nccl_library_path()
        check(self.lib.calz_comm_init(self.h, nranks, rank, unique_id, path.encode() if path else None), self.h)
        self.rank, self.nranks = rank, nranks

    @staticmethod
    def comm_unique_id() -> bytes:
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        path = _lib.nccl_library_path()
        check(lib.calz_comm_unique_id(buf, path.encode() if path else None))
        return buf.raw


_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class DeviceMatrix:
    """Device-resident sparse matrix (calz_mat): the 'other data structure' SpMV.m:3-5 anticipates."""

    def __init__(self, A, s_max: int = 16, layout: str = "auto", ctx: Context | None = None,
                 n_glob: int | None = None, row_begin: int = 0):
        """``A``: scipy sparse holding rows [row_begin, row_begin+A.shape[0]) of the n_glob x n_glob matrix
        (default: the whole matrix)."""
        self.ctx = ctx or default_context()
        lib = self.ctx.lib
        A = sp.csr_matrix(A)
        if not A.has_sorted_indices:
            A = A.copy()
            A.sort_indices()
        n_glob = A.shape[1] if n_glob is None else int(n_glob)
        rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        colind = np.ascontiguousarray(A.indices, dtype=np.int32)
        val = np.ascontiguousarray(A.data, dtype=np.float64)
        h = C.c_void_p()
        check(lib.calz_mat_create_csr(self.ctx.h, n_glob, int(row_begin), int(row_begin) + A.shape[0],
                                      rowptr.ctypes.data_as(_lib.c_i64p), colind.ctypes.data_as(_lib.c_i32p), _dp(val),
                                      int(s_max), _lib.LAYOUT[layout], C.byref(h)), self.ctx.h)
        self.h = h
        self.s_max = int(s_max)
        self.n = self.info("n_own")

    def info(self, key: str) -> int:
        v = C.c_int64()
        check(self.ctx.lib.calz_mat_info(self.h, key.encode(), C.byref(v)), self.ctx.h)
        return int(v.value)

    @property
    def layout(self) -> str:
        return _lib.LAYOUT_NAME[self.info("layout")]

    def _list(self, fn, *args):
        cnt = C.c_int64()
        check(fn(self.h, *args, None, C.byref(cnt)), self.ctx.h)
        out = np.empty(int(cnt.value), dtype=np.int64)
        check(fn(self.h, *args, out.ctypes.data_as(_lib.c_i64p), C.byref(cnt)), self.ctx.h)
        return out

    def ghost_indices(self):
        return self._list(self.ctx.lib.calz_mat_ghost_indices)

    def recv_list(self, peer: int):
        return self._list(self.ctx.lib.calz_mat_recv_list, int(peer))

    def send_list(self, peer: int):
        return self._list(self.ctx.lib.calz_mat_send_list, int(peer))

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.calz_mat_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceBlock:
    """n x cols fp64 block that lives on the device (calz_vec; the MATLAB side is mex/calz_vec.m): the handle mode of the call
    surface.  Every function of the surface (``SpMV``, ``matrix_powers_newton`` / ``_monomial``, ``tsqr``, ``cholqr``, ``normalize``,
    ``project``, ``projectAndNormalize``) returns DeviceBlocks when it is given DeviceBlocks, so V, Q and QZ never cross PCIe between calls; ``B[:, a:b]`` is a zero-copy column view,
    ``B[:, a:b] = other`` a device-to-device copy (or an upload), ``B.to_host()`` brings the block back."""

    def __init__(self, n=None, cols=None, ctx: Context | None = None, _view=None):
        self.ctx = ctx or default_context()
        if _view is not None:
            self.owner, self.h, self.n, self.col0, self.ncols, self.ld, self._base = _view
            return
        h = C.c_void_p()
        check(self.ctx.lib.calz_vec_create(self.ctx.h, int(n), int(cols), C.byref(h)), self.ctx.h)
        dev, nn, cc, ld = C.c_void_p(), C.c_int64(), C.c_int(), C.c_int64()
        check(self.ctx.lib.calz_vec_info(h, C.byref(dev), C.byref(nn), C.byref(cc), C.byref(ld)), self.ctx.h)
        self.owner, self.h, self.n, self.col0, self.ncols, self.ld, self._base = None, h, int(n), 0, int(cols), int(ld.value), int(dev.value)

    @classmethod
    def from_host(cls, X, ctx: Context | None = None):
        X = _f64_fortran(X)
        b = cls(X.shape[0], X.shape[1], ctx)
        check(b.ctx.lib.calz_vec_upload(b.h, 0, X.shape[1], _dp(X), X.shape[0]), b.ctx.h)
        return b

    @property
    def ptr(self) -> int:
        return self._base + 8 * self.ld * self.col0

    @property
    def shape(self):
        return (self.n, self.ncols)

    def _range(self, key):
        if not (isinstance(key, tuple) and len(key) == 2 and key[0] == slice(None)):
            raise IndexError("DeviceBlock views are B[:, a:b]")
        k = key[1]
        if isinstance(k, int):
            k = slice(k, k + 1)
        a, b, st = k.indices(self.ncols)
        if st != 1 or b <= a:
            raise IndexError("DeviceBlock views are contiguous column ranges")
        return a, b

    def __getitem__(self, key):
        a, b = self._range(key)
        return DeviceBlock(ctx=self.ctx, _view=(self.owner or self, self.h, self.n, self.col0 + a, b - a, self.ld, self._base))

    def __setitem__(self, key, value):
        a, b = self._range(key)
        if isinstance(value, DeviceBlock):
            if value.shape != (self.n, b - a):
                raise ValueError("shape mismatch")
            check(self.ctx.lib.calz_vec_copy(self.h, self.col0 + a, value.h, value.col0, b - a), self.ctx.h)
        else:
            X = _f64_fortran(value)
            if X.shape != (self.n, b - a):
                raise ValueError("shape mismatch")
            check(self.ctx.lib.calz_vec_upload(self.h, self.col0 + a, b - a, _dp(X), self.n), self.ctx.h)

    def to_host(self, out=None):
        X = _out_array(out, (self.n, self.ncols))
        check(self.ctx.lib.calz_vec_download(self.h, self.col0, self.ncols, _dp(X), self.n), self.ctx.h)
        return X

    def __del__(self):
        try:
            if self.owner is None and getattr(self, "h", None) and getattr(self.ctx, "h", None):
                self.ctx.lib.calz_vec_destroy(self.h)
        except Exception:
            pass


# ---- device-matrix cache (the MEX gateway keys on mxGetPr(A)+n+nnz; here on the host object)
_mat_cache: dict = {}


def _device_matrix(A, s_needed: int) -> DeviceMatrix:
    if isinstance(A, DeviceMatrix):
        if s_needed > A.s_max:
            raise ValueError("DeviceMatrix was created with s_max=%d < s=%d" % (A.s_max, s_needed))
        return A
    key = id(A)
    ent = _mat_cache.get(key)
    if ent is not None and ent[0]() is A and ent[1].s_max >= s_needed and ent[2] == (A.shape, A.nnz):
        return ent[1]
    dm = DeviceMatrix(A, s_max=max(16, s_needed))
    try:
        ref = weakref.ref(A, lambda _r, k=key: _mat_cache.pop(k, None))
    except TypeError:
        ref = (lambda a: (lambda: a))(A)
    _mat_cache[key] = (ref, dm, (A.shape, A.nnz))
    return dm


# ------------------------------------------------------------------------------------------ MPK
def SpMV(A, v):
    """SpMV.m:6-8 -- ``Av = A*v``."""
    dm = _device_matrix(A, 1)
    if isinstance(v, DeviceBlock):                                 # handle mode: n x 1 device block in, device block out
        if v.shape != (dm.n, 1):
            raise ValueError("SpMV: dimension mismatch")
        y = DeviceBlock(dm.n, 1, dm.ctx)
        check(dm.ctx.lib.calz_spmv(dm.h, C.c_void_p(v.ptr), C.c_void_p(y.ptr)), dm.ctx.h)
        return y
    v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
    if v.shape[0] != dm.n:
        raise ValueError("SpMV: dimension mismatch")
    y = np.empty(dm.n)
    check(dm.ctx.lib.calz_spmv_host(dm.h, _dp(v), _dp(y)), dm.ctx.h)
    return y


def matrix_powers_monomial(A, q, s):
    """matrix_powers_monomial.m:6-12 -- V = [A q, ..., A^s q], n x s (q itself NOT included)."""
    s = int(s)
    dm = _device_matrix(A, s)
    if isinstance(q, DeviceBlock):                                 # handle mode: device block in, device block out
        if q.shape != (dm.n, 1):
            raise ValueError("matrix_powers_monomial: dimension mismatch")
        V = DeviceBlock(dm.n, s, dm.ctx)
        check(dm.ctx.lib.calz_mpk_monomial(dm.h, C.c_void_p(q.ptr), s, C.c_void_p(V.ptr), V.ld), dm.ctx.h)
        return V
    q = np.ascontiguousarray(np.asarray(q, dtype=np.float64).ravel())
    if q.shape[0] != dm.n:
        raise ValueError("matrix_powers_monomial: dimension mismatch")
    V = np.empty((dm.n, s), order="F")
    check(dm.ctx.lib.calz_mpk_monomial_host(dm.h, _dp(q), s, _dp(V), dm.n), dm.ctx.h)
    return V


def _out_array(out, shape):
    """Optional caller-provided result buffer (e.g. pinned host memory); must be fp64, column-major, right shape."""
    if out is None:
        return np.empty(shape, order="F")
    if out.shape != shape or out.dtype != np.float64 or not out.flags.f_contiguous:
        raise ValueError("out must be a float64 Fortran-ordered array of shape %s" % (shape,))
    return out


def matrix_powers_newton(A, v, s, lam, modifiedp=0, out=None):
    """matrix_powers_newton.m:15-54 -- V(:,1)=v; V(:,k+1)=A*V(:,k)-lam(k)*V(:,k); n x (s+1) (v included).
    ``out`` (extension): write the result into a caller-owned array instead of a fresh one."""
    s = int(s)
    dm = _device_matrix(A, s)
    lam = np.asarray(lam).ravel()
    if lam.shape[0] < s:
        raise ValueError("matrix_powers_newton: need at least s shifts")
    re = np.ascontiguousarray(np.real(lam[:s]), dtype=np.float64)
    im = np.ascontiguousarray(np.imag(lam[:s]), dtype=np.float64) if np.iscomplexobj(lam) else None
    handle = isinstance(v, DeviceBlock)
    if handle:                                                     # handle mode: device block in, device block out
        if v.shape != (dm.n, 1):
            raise ValueError("matrix_powers_newton: dimension mismatch")
        V = out if isinstance(out, DeviceBlock) else DeviceBlock(dm.n, s + 1, dm.ctx)
    else:
        v = np.ascontiguousarray(np.asarray(v, dtype=np.float64).ravel())
        if v.shape[0] != dm.n:
            raise ValueError("matrix_powers_newton: dimension mismatch")
        V = _out_array(out, (dm.n, s + 1))
    try:
        if handle:
            check(dm.ctx.lib.calz_mpk_newton(dm.h, C.c_void_p(v.ptr), s, _dp(re), _dp(im) if im is not None else None,
                                             int(modifiedp), C.c_void_p(V.ptr), V.ld), dm.ctx.h)
        else:
            check(dm.ctx.lib.calz_mpk_newton_host(dm.h, _dp(v), s, _dp(re), _dp(im) if im is not None else None,
                                                  int(modifiedp), _dp(V), dm.n), dm.ctx.h)
    except CalzError as e:
        if e.code == 8:           # matrix_powers_newton.m:36-39 error(...)
            raise ValueError(str(e)) from None
        raise
    return V


# ------------------------------------------------------------------------------------------ QR
def tsqr(A, ctx: Context | None = None):
    """tsqr.m:7-12 -- thin QR with diag(R) >= 0."""
    ctx = ctx or default_context()
    if isinstance(A, DeviceBlock):                                 # handle mode: Q a device block, R on the host
        n, c = A.shape
        Q = DeviceBlock(n, c, A.ctx); R = np.empty((c, c), order="F")
        check(A.ctx.lib.calz_tsqr(A.ctx.h, n, c, C.c_void_p(A.ptr), A.ld, C.c_void_p(Q.ptr), Q.ld, _dp(R)), A.ctx.h)
        return Q, R
    A = _f64_fortran(A)
    n, c = A.shape
    Q = np.empty((n, c), order="F"); R = np.empty((c, c), order="F")
    check(ctx.lib.calz_tsqr_host(ctx.h, n, c, _dp(A), n, _dp(Q), n, _dp(R)), ctx.h)
    return Q, R


def cholqr(X, ctx: Context | None = None):
    """cholqr.m:3-8 -- G=X'X; R=chol(G); Q=X/R.  Raises numpy.linalg.LinAlgError if G is not PD (MATLAB: error)."""
    ctx = ctx or default_context()
    if isinstance(X, DeviceBlock):                                 # handle mode
        n, c = X.shape
        Q = DeviceBlock(n, c, X.ctx); R = np.empty((c, c), order="F")
        info = C.c_int(0)
        st = X.ctx.lib.calz_cholqr(X.ctx.h, n, c, C.c_void_p(X.ptr), X.ld, C.c_void_p(Q.ptr), Q.ld, _dp(R), C.byref(info))
        if st == _lib.ERR_CHOL:
            raise np.linalg.LinAlgError("Matrix must be positive definite (pivot %d)" % info.value)
        check(st, X.ctx.h)
        return Q, R
    X = _f64_fortran(X)
    n, c = X.shape
    Q = np.empty((n, c), order="F"); R = np.empty((c, c), order="F")
    info = C.c_int(0)
    st = ctx.lib.calz_cholqr_host(ctx.h, n, c, _dp(X), n, _dp(Q), n, _dp(R), C.byref(info))
    if st == _lib.ERR_CHOL:
        raise np.linalg.LinAlgError("Matrix must be positive definite (pivot %d)" % info.value)
    check(st, ctx.h)
    return Q, R


def normalize(X, opt="None", tol=1.0e-8, backend: str | None = None, ctx: Context | None = None):
    """normalize.m:3-36 -- [Q,R,rank]; QR by the selected backend, rank from svd(R)."""
    if str(opt).lower() == "randomizenullspace":
        raise NotImplementedError("normalize(...,'randomizeNullSpace') is never requested on the hot path")
    ctx = ctx or default_context()
    rank = C.c_int(0)
    if isinstance(X, DeviceBlock):                                 # handle mode
        n, c = X.shape
        Q = DeviceBlock(n, c, X.ctx); R = np.empty((c, c), order="F")
        st = X.ctx.lib.calz_normalize(X.ctx.h, n, c, C.c_void_p(X.ptr), X.ld, _lib.QR[backend or _QR_BACKEND], float(tol),
                                      C.c_void_p(Q.ptr), Q.ld, _dp(R), C.byref(rank))
        if st == _lib.ERR_CHOL:
            raise np.linalg.LinAlgError("Matrix must be positive definite")
        check(st, X.ctx.h)
        return Q, R, int(rank.value)
    X = _f64_fortran(X)
    n, c = X.shape
    Q = np.empty((n, c), order="F"); R = np.empty((c, c), order="F")
    st = ctx.lib.calz_normalize_host(ctx.h, n, c, _dp(X), n, _lib.QR[backend or _QR_BACKEND], float(tol), _dp(Q), n,
                                     _dp(R), C.byref(rank))
    if st == _lib.ERR_CHOL:
        raise np.linalg.LinAlgError("Matrix must be positive definite")
    check(st, ctx.h)
    return Q, R, int(rank.value)


# ------------------------------------------------------------------------------------------ block Gram-Schmidt
def _cell(Q, n):
    """MATLAB cell array of blocks -> ctypes arrays; empty cells ([] / None) are kept as empty."""
    if not isinstance(Q, (list, tuple)):
        raise TypeError("Input Q (arg 1) must be cell (block) array.")          # project.m:12-15
    nb = len(Q)
    keep = []
    ptrs = (_lib.c_dp * max(nb, 1))()
    lds = (C.c_int64 * max(nb, 1))()
    mc = (C.c_int * max(nb, 1))()
    for i, Qi in enumerate(Q):
        if Qi is None or np.size(Qi) == 0:
            ptrs[i] = None; lds[i] = n; mc[i] = 0
            continue
        Qi = _f64_fortran(Qi)
        if Qi.shape[0] != n:
            raise ValueError("block %d has %d rows, expected %d" % (i, Qi.shape[0], n))
        keep.append(Qi)
        ptrs[i] = _dp(Qi); lds[i] = n; mc[i] = Qi.shape[1]
    return nb, ptrs, lds, mc, keep


def project(Q, X, doreorth=False, ctx: Context | None = None):
    """project.m:7-58 -- returns (X, R) with R a list (cell) of Q{i}'*X blocks; empty cells give None."""
    ctx = ctx or default_context()
    if isinstance(X, (list, tuple)):
        raise TypeError("Input X (arg 2) project() must be a column matrix.")   # project.m:16-19
    if isinstance(X, DeviceBlock):                                 # handle mode: a fresh device block comes back (value semantics)
        if not isinstance(Q, (list, tuple)):
            raise TypeError("Input Q (arg 1) must be cell (block) array.")
        n, c = X.shape
        nb = len(Q)
        blocks = [b if (b is not None and not (isinstance(b, np.ndarray) and b.size == 0)) else None for b in Q]
        if any(b is not None and not isinstance(b, DeviceBlock) for b in blocks):
            raise TypeError("project: host arrays and DeviceBlocks cannot be mixed")
        Y = DeviceBlock(n, c, X.ctx)
        Y[:, 0:c] = X
        if nb == 0:
            return Y, []
        qb = (C.c_void_p * nb)(*[(b.ptr if b is not None else None) for b in blocks])
        lds = (C.c_int64 * nb)(*[(b.ld if b is not None else n) for b in blocks])
        mc = (C.c_int * nb)(*[(b.ncols if b is not None else 0) for b in blocks])
        R = [np.zeros((b.ncols, c), order="F") if b is not None else None for b in blocks]
        rp = (_lib.c_dp * nb)(*[(_dp(r) if r is not None else None) for r in R])
        check(X.ctx.lib.calz_project(X.ctx.h, n, nb, qb, lds, mc, c, C.c_void_p(Y.ptr), Y.ld, 1 if doreorth else 0, rp), X.ctx.h)
        return Y, R
    X = _f64_fortran(X, copy=True)
    n, c = X.shape
    nb, ptrs, lds, mc, keep = _cell(Q, n)
    if nb == 0:
        return X, []
    R = [np.zeros((mc[i], c), order="F") if mc[i] > 0 else None for i in range(nb)]
    rp = (_lib.c_dp * nb)(*[(_dp(r) if r is not None else None) for r in R])
    check(ctx.lib.calz_project_host(ctx.h, n, nb, ptrs, lds, mc, c, _dp(X), n, 1 if doreorth else 0, rp), ctx.h)
    return X, R


def projectAndNormalize(Q, X, doreorth=True, backend: str | None = None, info: dict | None = None,
                        ctx: Context | None = None, out=None):
    """projectAndNormalize.m:3-90 -- returns (QZ, RZ) with RZ a list of len(Q)+1 blocks (last = R of the last
    normalize).  ``info`` receives 'second_pass' (the reference prints 'second', :62) and 'rank'."""
    ctx = ctx or default_context()
    if isinstance(X, DeviceBlock):                                 # handle mode: device blocks in, QZ a device block (or `out`)
        if not isinstance(Q, (list, tuple)):
            raise TypeError("Input Q (arg 1) must be cell (block) array.")
        n, c = X.shape
        nb = len(Q)
        blocks = [b if (b is not None and not (isinstance(b, np.ndarray) and b.size == 0)) else None for b in Q]
        if any(b is not None and not isinstance(b, DeviceBlock) for b in blocks):
            raise TypeError("projectAndNormalize: host arrays and DeviceBlocks cannot be mixed")
        qb = (C.c_void_p * max(nb, 1))(*[(b.ptr if b is not None else None) for b in blocks])
        lds = (C.c_int64 * max(nb, 1))(*[(b.ld if b is not None else n) for b in blocks])
        mc = (C.c_int * max(nb, 1))(*[(b.ncols if b is not None else 0) for b in blocks])
        R = [np.zeros((b.ncols, c), order="F") if b is not None else None for b in blocks]
        rp = (_lib.c_dp * max(nb, 1))(*[(_dp(r) if r is not None else None) for r in R])
        Rlast = np.zeros((c, c), order="F")
        QZ = out if isinstance(out, DeviceBlock) else DeviceBlock(n, c, X.ctx)
        second = C.c_int(0); rank = C.c_int(0)
        st = X.ctx.lib.calz_project_and_normalize(X.ctx.h, n, nb, qb, lds, mc, c, C.c_void_p(X.ptr), X.ld, 1 if doreorth else 0,
                                                  _lib.QR[backend or _QR_BACKEND], C.c_void_p(QZ.ptr), QZ.ld, rp, _dp(Rlast),
                                                  C.byref(second), C.byref(rank))
        if st == _lib.ERR_CHOL:
            raise np.linalg.LinAlgError("Matrix must be positive definite")
        check(st, X.ctx.h)
        if info is not None:
            info["second_pass"] = bool(second.value)
            info["rank"] = int(rank.value)
        return QZ, list(R) + [Rlast]
    X = _f64_fortran(X)
    n, c = X.shape
    nb, ptrs, lds, mc, keep = _cell(Q, n)
    R = [np.zeros((mc[i], c), order="F") if (i < nb and mc[i] > 0) else None for i in range(nb)]
    rp = (_lib.c_dp * max(nb, 1))(*[(_dp(r) if r is not None else None) for r in R]) if nb else (_lib.c_dp * 1)()
    Rlast = np.zeros((c, c), order="F")
    QZ = _out_array(out, (n, c))
    second = C.c_int(0); rank = C.c_int(0)
    st = ctx.lib.calz_project_and_normalize_host(ctx.h, n, nb, ptrs, lds, mc, c, _dp(X), n, 1 if doreorth else 0,
                                                 _lib.QR[backend or _QR_BACKEND], _dp(QZ), n, rp, _dp(Rlast),
                                                 C.byref(second), C.byref(rank))
    if st == _lib.ERR_CHOL:
        raise np.linalg.LinAlgError("Matrix must be positive definite")
    check(st, ctx.h)
    if info is not None:
        info["second_pass"] = bool(second.value)
        info["rank"] = int(rank.value)
    return QZ, list(R) + [Rlast]
