"""World-size-2 CPU test (gloo) of the multi-GPU plan's HOST logic: the library's row partition and ghost level sets
(host-only C entry points, no GPU) drive a real two-process run -- halo exchange of the start vector over gloo
send/recv, redundant-ghost MPK on the local problem, Gram matrices all-reduced -- and the result must equal the
1-way oracle.  The arithmetic is the oracle's (this is the checker), the partition objects are the library's."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, world, port, s, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ca_lanczos_b200 import _lib, gallery
        from oracle import kernels
        lib = _lib.load()
        A = gallery.laplace3d(9, 7, 20)
        n = A.shape[0]
        # ---- library: bounds + level sets (bit-exact objects)
        b = np.zeros(world + 1, dtype=np.int64)
        assert lib.calz_partition_bounds(n, world, b.ctypes.data_as(_lib.c_i64p)) == 0
        lo, hi = int(b[rank]), int(b[rank + 1])
        rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64); col = np.ascontiguousarray(A.indices, dtype=np.int32)
        lev = np.zeros(n, dtype=np.int32)
        assert lib.calz_level_sets(n, 0, n, rowptr.ctypes.data_as(_lib.c_i64p), col.ctypes.data_as(_lib.c_i32p), lo, hi, s,
                                   lev.ctypes.data_as(_lib.c_i32p)) == 0
        loc2glob = np.flatnonzero(lev >= 0)
        ghosts = np.flatnonzero(lev > 0)
        # ---- exchange lists: what I need from each peer; tell the peers (sizes, then indices)
        need = [ghosts[(ghosts >= b[q]) & (ghosts < b[q + 1])] for q in range(world)]
        peer = 1 - rank
        cnt = torch.tensor([need[peer].size]); other = torch.zeros(1, dtype=torch.long)
        reqs = [dist.isend(cnt, peer), dist.irecv(other, peer)]
        [r.wait() for r in reqs]
        give_idx = torch.zeros(int(other.item()), dtype=torch.long)
        reqs = [dist.isend(torch.from_numpy(need[peer].astype(np.int64)), peer), dist.irecv(give_idx, peer)]
        [r.wait() for r in reqs]
        give_idx = give_idx.numpy()
        assert np.all((give_idx >= lo) & (give_idx < hi)) and np.all(np.diff(give_idx) > 0)
        # ---- the ONE halo exchange of a block: owned slices only, ghosts arrive from the peer
        v_full = np.cos(0.3 * np.arange(n)) + 2.0
        v_own = v_full[lo:hi].copy()
        recv = torch.zeros(need[peer].size, dtype=torch.float64)
        reqs = [dist.isend(torch.from_numpy(v_own[give_idx - lo].copy()), peer), dist.irecv(recv, peer)]
        [r.wait() for r in reqs]
        v_loc = np.empty(loc2glob.size)
        own_pos = np.searchsorted(loc2glob, np.arange(lo, hi))
        v_loc[own_pos] = v_own
        v_loc[np.searchsorted(loc2glob, need[peer])] = recv.numpy()
        # ---- redundant-ghost MPK on the local matrix, no further communication
        A_loc = A[loc2glob][:, loc2glob].tocsr()
        lam = np.array([11.0, 1.0, 6.5, 3.0, 9.0, 4.5])[:s]
        V_loc = kernels.matrix_powers_newton(A_loc, v_loc, s, lam, 1)
        V_own = V_loc[own_pos]
        V_ref = kernels.matrix_powers_newton(A, v_full, s, lam, 1)[lo:hi]
        e_mpk = float(np.max(np.abs(V_own - V_ref)) / np.max(np.abs(V_ref)))
        # ---- CholQR with an all-reduced Gram matrix == 1-way cholqr
        X = V_own[:, 1:]
        G = torch.from_numpy(X.T @ X)
        dist.all_reduce(G)
        R = np.linalg.cholesky(G.numpy()).T
        Xfull = kernels.matrix_powers_newton(A, v_full, s, lam, 1)[:, 1:]
        _, R_ref = kernels.cholqr(Xfull)
        e_R = float(np.linalg.norm(R - R_ref) / np.linalg.norm(R_ref))
        out.put((rank, e_mpk, e_R, int(ghosts.size)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("s", [3, 6])
def test_two_rank_gloo_halo_and_allreduce(s):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + s
    procs = [ctx.Process(target=_worker, args=(r, 2, port, s, out)) for r in range(2)]
    [p.start() for p in procs]
    res = [out.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, e_mpk, e_R, ng in res:
        assert ng == 63 * s                       # s planes of 9*7 rows from the single neighbour
        assert e_mpk < 1e-13, (rank, e_mpk)
        assert e_R < 1e-12, (rank, e_R)


def _worker_levels(rank, world, port, s, L, out):
    """halo level L < s: the current basis column is exchanged every L steps (L = 1: the classic per-step exchange)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ca_lanczos_b200 import _lib, gallery
        from oracle import kernels
        lib = _lib.load()
        A = gallery.powerlaw_spd(600, 6.0, seed=2) if L == 1 else gallery.laplace3d(9, 7, 20)
        n = A.shape[0]
        b = np.zeros(world + 1, dtype=np.int64)
        assert lib.calz_partition_bounds(n, world, b.ctypes.data_as(_lib.c_i64p)) == 0
        lo, hi = int(b[rank]), int(b[rank + 1])
        rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64); col = np.ascontiguousarray(A.indices, dtype=np.int32)
        lev = np.zeros(n, dtype=np.int32)
        # only the rows the level-L closure needs are supplied when L == 1 (the owned rows): the library must not ask for more
        rb, re_ = (lo, hi) if L == 1 else (0, n)
        sub = A[rb:re_]
        rp = np.ascontiguousarray(sub.indptr, dtype=np.int64); ci = np.ascontiguousarray(sub.indices, dtype=np.int32)
        assert lib.calz_level_sets(n, rb, re_, rp.ctypes.data_as(_lib.c_i64p), ci.ctypes.data_as(_lib.c_i32p), lo, hi, L,
                                   lev.ctypes.data_as(_lib.c_i32p)) == 0
        loc2glob = np.flatnonzero(lev >= 0)
        ghosts = np.flatnonzero(lev > 0)
        peer = 1 - rank
        need = ghosts[(ghosts >= b[peer]) & (ghosts < b[peer + 1])]
        cnt = torch.tensor([need.size]); other = torch.zeros(1, dtype=torch.long)
        [r.wait() for r in [dist.isend(cnt, peer), dist.irecv(other, peer)]]
        give_idx = torch.zeros(int(other.item()), dtype=torch.long)
        [r.wait() for r in [dist.isend(torch.from_numpy(need.astype(np.int64)), peer), dist.irecv(give_idx, peer)]]
        give_idx = give_idx.numpy()
        own_pos = np.searchsorted(loc2glob, np.arange(lo, hi))
        ghost_pos = np.searchsorted(loc2glob, need)
        A_loc = A[loc2glob][:, loc2glob].tocsr()
        lam = np.array([11.0, 1.0, 6.5, 3.0, 9.0, 4.5])[:s] * (1.0 if L != 1 else 3.0)
        v_full = np.cos(0.3 * np.arange(n)) + 2.0
        V_own = np.zeros((hi - lo, s + 1))
        V_own[:, 0] = v_full[lo:hi]
        exchanges = 0
        for k0 in range(0, s, L):
            g = min(L, s - k0)
            recv = torch.zeros(need.size, dtype=torch.float64)
            [r.wait() for r in [dist.isend(torch.from_numpy(V_own[give_idx - lo, k0].copy()), peer), dist.irecv(recv, peer)]]
            exchanges += 1
            v_loc = np.zeros(loc2glob.size)
            v_loc[own_pos] = V_own[:, k0]
            v_loc[ghost_pos] = recv.numpy()
            V_loc = kernels.matrix_powers_newton(A_loc, v_loc, g, lam[k0:k0 + g], 1)
            V_own[:, k0 + 1:k0 + g + 1] = V_loc[own_pos, 1:g + 1]
        V_ref = kernels.matrix_powers_newton(A, v_full, s, lam, 1)[lo:hi]
        out.put((rank, float(np.max(np.abs(V_own - V_ref)) / np.max(np.abs(V_ref))), exchanges))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("s,L", [(4, 1), (6, 2), (5, 3)])
def test_two_rank_gloo_matrix_powers_with_an_exchange_every_L_steps(s, L):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29620 + 10 * s + L
    procs = [ctx.Process(target=_worker_levels, args=(r, 2, port, s, L, out)) for r in range(2)]
    [p.start() for p in procs]
    res = [out.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, e_mpk, exchanges in res:
        assert exchanges == -(-s // L)
        assert e_mpk < 1e-13, (rank, e_mpk)
