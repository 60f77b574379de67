"""Device-resident CA-Lanczos block pipeline.

One *block* is one outer iteration k>1 of ``ca_lanczos_basic`` (ca_lanczos.m:166-225): take the last vector
of the previous block, run the s-step matrix powers kernel, and ``projectAndNormalize({Qprev}, V(:,2:s+1))``.
Here Q, V and every intermediate stay in HBM (torch owns the Q storage, libcalz the basis workspace); only the
(s+1) x s coefficient blocks come back to the host, where the reference's O(s^3) T-matrix algebra
(ca_lanczos.m:200-223) is replayed in numpy.  With a communicator the rows are partitioned over the ranks
(one process per GPU) and the same code runs on every rank.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .api import Context, DeviceMatrix


def _solve_upper_right(B, R):
    """B / R for square upper-triangular R (MATLAB mrdivide): X R = B by forward substitution over columns."""
    B = np.array(B, dtype=np.float64)
    X = np.zeros_like(B)
    for j in range(R.shape[0]):
        X[:, j] = (B[:, j] - X[:, :j] @ R[:j, j]) / R[j, j]
    return X


class BlockEngine:
    def __init__(self, dm: DeviceMatrix, s: int, max_blocks: int, basis: str = "newton", shifts=None,
                 backend: str = "cholqr"):
        import torch

        self.torch = torch
        self.dm, self.ctx, self.lib = dm, dm.ctx, dm.ctx.lib
        self.s, self.backend = int(s), backend
        self.monomial = basis.lower() == "monomial"
        if not self.monomial:
            shifts = np.asarray(shifts).ravel()[: self.s]
            self.re = np.ascontiguousarray(np.real(shifts), dtype=np.float64)
            self.im = np.ascontiguousarray(np.imag(shifts), dtype=np.float64) if np.iscomplexobj(shifts) else None
            self.Bk = np.zeros((s + 1, s))
            for k in range(s):
                self.Bk[k, k] = self.re[k]
                self.Bk[k + 1, k] = 1.0
                if self.im is not None and self.im[k] < 0:
                    self.Bk[k - 1, k] = -self.im[k] ** 2
        else:
            self.re = self.im = None
            self.Bk = np.eye(s + 1)[:, 1:]
        self.n = dm.n                                            # owned rows of this rank
        self.ld = (self.n + 31) // 32 * 32
        self.max_cols = self.s * int(max_blocks) + 1
        dev = torch.device("cuda", self.ctx.device)
        # column-major n x max_cols with leading dimension ld  ==  row-major (max_cols, ld) torch tensor
        self.Q = torch.zeros((self.max_cols, self.ld), dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)          # libcalz runs on its own stream
        self.k = 0
        self.k_enq = 0
        self.host_enqueue_s = 0.0          # host time spent enqueueing blocks (pipelined mode)
        self.pending = []
        self.T = None
        self.b = np.zeros(int(max_blocks) + 2)
        self.second = []
        self._tmp = None
        # host landing zones for the small results
        self._R1 = np.zeros((s + 1, s), order="F")
        self._Rl = np.zeros((s, s), order="F")
        self._Rfirst = np.zeros((s + 1, s + 1), order="F")

    # ---- pointers
    def _qcol(self, j):
        return self.Q.data_ptr() + 8 * self.ld * int(j)

    def _mpk(self, qptr):
        V = C.c_void_p(); ldV = C.c_int64()
        check(self.lib.calz_mpk_inplace(self.dm.h, C.c_void_p(qptr), self.s,
                                        None if self.re is None else self.re.ctypes.data_as(_lib.c_dp),
                                        None if self.im is None else self.im.ctypes.data_as(_lib.c_dp),
                                        1, 1 if self.monomial else 0, C.byref(V), C.byref(ldV)), self.ctx.h)
        return int(V.value), int(ldV.value)

    # ---- ca_lanczos.m:176-182
    def first_block(self, q0, q_ptr=None):
        """q0: host vector (owned rows), already normalised (ca_lanczos.m:55); or ``q_ptr``: the same on the device."""
        torch = self.torch
        s = self.s
        if q_ptr is None:
            q = torch.as_tensor(np.ascontiguousarray(q0, dtype=np.float64), device=self.Q.device)
            torch.cuda.current_stream().synchronize()
            q_ptr = q.data_ptr()
        V, ldV = self._mpk(q_ptr)
        rank = C.c_int()
        check(self.lib.calz_normalize(self.ctx.h, self.n, s + 1, C.c_void_p(V), ldV, _lib.QR[self.backend], 1e-8,
                                      C.c_void_p(self._qcol(0)), self.ld, self._Rfirst.ctypes.data_as(_lib.c_dp),
                                      C.byref(rank)), self.ctx.h)
        Rk = self._Rfirst
        self.T = _solve_upper_right(Rk @ self.Bk, Rk[:s, :s])
        self.b[0] = self.T[s, s - 1]
        self.k = 1
        return Rk.copy()

    # ---- ca_lanczos.m:184-223 ('local')
    def next_block(self, assemble_T: bool = True, events=None, full_reorth: bool = False, extra_blocks=()):
        """``events``: optional ((e0,e1,e2), stream) -- torch CUDA events recorded on libcalz' stream before the MPK,
        between MPK and projectAndNormalize, and after it (bench.py's per-phase timing).
        ``extra_blocks``: further cells of the projection, (device pointer, ld, columns) each, after {Qprev}: the converged Ritz
        vectors of the selective driver (ca_lanczos.m:286) -- their coefficient blocks are discarded like in the reference."""
        s = self.s
        self.k += 1
        k = self.k
        if k * s + 1 > self.max_cols:
            raise RuntimeError("BlockEngine: Q storage exhausted")
        if events is not None:
            events[0][0].record(events[1])
        V, ldV = self._mpk(self._qcol((k - 1) * s))
        if events is not None:
            events[0][1].record(events[1])
        nb = 1 + len(extra_blocks)
        qblk = (C.c_void_p * nb)(self._qcol((k - 2) * s), *[int(b[0]) for b in extra_blocks])
        lds = (C.c_int64 * nb)(self.ld, *[int(b[1]) for b in extra_blocks])
        mc = (C.c_int * nb)(s + 1, *[int(b[2]) for b in extra_blocks])
        rp = (_lib.c_dp * nb)(self._R1.ctypes.data_as(_lib.c_dp), *[None for _ in extra_blocks])
        second = C.c_int(); rank = C.c_int()
        if full_reorth and self._tmp is None:
            self._tmp = self.torch.zeros((s, self.ld), dtype=self.torch.float64, device=self.Q.device)
            self.torch.cuda.synchronize(self.Q.device)
        dst = self._tmp.data_ptr() if full_reorth else self._qcol((k - 1) * s + 1)
        check(self.lib.calz_project_and_normalize(self.ctx.h, self.n, nb, qblk, lds, mc, s, C.c_void_p(V + 8 * ldV), ldV,
                                                  1, _lib.QR[self.backend], C.c_void_p(dst),
                                                  self.ld, rp, self._Rl.ctypes.data_as(_lib.c_dp), C.byref(second),
                                                  C.byref(rank)), self.ctx.h)
        if full_reorth:
            # 'fro' (ca_lanczos.m:193-197): orthogonalise the new block against ALL previous basis vectors as well;
            # the coefficients of this second call are discarded, T is assembled from the first
            qall = (C.c_void_p * 1)(self._qcol(0)); mall = (C.c_int * 1)((k - 1) * s + 1)
            s2, r2 = C.c_int(), C.c_int()
            check(self.lib.calz_project_and_normalize(self.ctx.h, self.n, 1, qall, lds, mall, s, C.c_void_p(dst), self.ld, 1,
                                                      _lib.QR[self.backend], C.c_void_p(self._qcol((k - 1) * s + 1)), self.ld,
                                                      None, None, C.byref(s2), C.byref(r2)), self.ctx.h)
        if events is not None:
            events[0][2].record(events[1])
        self.second.append(bool(second.value))
        if assemble_T:
            self._extend_T(k, self._R1, self._Rl)
        return bool(second.value)

    # ---- asynchronous pipeline: the GPU never waits for the host's O(s^3) algebra or for result read-back
    def _enqueue_block(self):
        import time as _time
        _t0 = _time.perf_counter()
        s = self.s
        k = self.k_enq + 1
        if k * s + 1 > self.max_cols:
            raise RuntimeError("BlockEngine: Q storage exhausted")
        V, ldV = self._mpk(self._qcol((k - 1) * s))
        qblk = (C.c_void_p * 1)(self._qcol((k - 2) * s))
        lds = (C.c_int64 * 1)(self.ld)
        mc = (C.c_int * 1)(s + 1)
        ticket = C.c_int(-1)
        check(self.lib.calz_project_and_normalize_async(self.ctx.h, self.n, 1, qblk, lds, mc, s, C.c_void_p(V + 8 * ldV), ldV, 1,
                                                        _lib.QR[self.backend], C.c_void_p(self._qcol((k - 1) * s + 1)), self.ld,
                                                        C.byref(ticket)), self.ctx.h)
        self.k_enq = k
        self.pending.append((k, int(ticket.value)))
        self.host_enqueue_s += _time.perf_counter() - _t0

    def _collect_one(self, assemble_T=True):
        k, ticket = self.pending.pop(0)
        rp = (_lib.c_dp * 1)(self._R1.ctypes.data_as(_lib.c_dp))
        second, rank, refine = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.calz_pan_collect(self.ctx.h, ticket, rp, self._Rl.ctypes.data_as(_lib.c_dp), C.byref(second), C.byref(rank),
                                        C.byref(refine)), self.ctx.h)
        if refine.value:
            # CholQR2 asked for a refinement pass on this block: everything enqueued after it used an unrefined Q.
            # Roll back and redo this block synchronously (rare: ill-conditioned basis blocks only).
            for _, t in self.pending:
                self.lib.calz_pan_collect(self.ctx.h, t, None, None, None, None, None)
            self.pending = []
            self.ctx.sync()
            self.k = k - 1
            self.next_block(assemble_T)
            self.k_enq = self.k
            return
        self.k = k
        self.second.append(bool(second.value))
        if assemble_T:
            self._extend_T(k, self._R1, self._Rl)

    def run_blocks(self, nblocks: int, lag: int = 3, assemble_T: bool = True):
        """Advance by ``nblocks`` outer iterations with up to ``lag`` projectAndNormalize calls in flight."""
        self.k_enq = self.k
        self.pending = []
        target = self.k + int(nblocks)
        while self.k < target:
            while self.k_enq < target and len(self.pending) < lag:
                self._enqueue_block()
            self._collect_one(assemble_T)

    def _extend_T(self, k, Rkk_s, Rk_s):
        """ca_lanczos.m:200-223."""
        s, Bk, b = self.s, self.Bk, self.b
        Rkk = np.hstack([np.zeros((s, 1)), Rkk_s[:s, :]])
        first = np.zeros((s + 1, 1)); first[0, 0] = 1.0
        Rk = np.hstack([first, np.vstack([Rkk_s[s : s + 1, :], Rk_s])])
        zk, rho, rho_t, bk = Rk[:s, s], Rk[s, s], Rk[s - 1, s - 1], Bk[s, s - 1]
        Rss = Rk[:s, :s]
        e1 = np.zeros(s); e1[0] = 1.0
        es = np.zeros(s); es[s - 1] = 1.0
        Tk = (_solve_upper_right(Rss @ Bk[:s, :], Rss) + (bk / rho_t) * np.outer(zk, es)
              - _solve_upper_right((b[k - 2] * np.outer(e1, es)) @ Rkk[:, :s], Rss))
        b[k - 1] = bk * (rho / rho_t)
        m = s * (k - 1)
        Tn = np.zeros((m + s + 1, m + s))
        Tn[:m, :m] = self.T[:m, :m]
        Tn[m - 1, m] = b[k - 2]
        Tn[m, m - 1] = b[k - 2]
        Tn[m : m + s, m : m + s] = Tk
        Tn[m + s, m + s - 1] = b[k - 1]
        self.T = Tn

    # ---- results
    def T_matrix(self):
        m = self.s * self.k
        return self.T[:m, :m].copy()

    def Q_host(self, ncols=None):
        ncols = self.s * self.k if ncols is None else ncols
        self.ctx.sync()
        return np.asfortranarray(self.Q[:ncols, : self.n].T.cpu().numpy())
