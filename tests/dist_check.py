"""Multi-GPU parity check, one process per GPU (launched by torchrun):

    torchrun --nproc-per-node P tests/dist_check.py [--grid 48]

Every rank builds its row slab (+ level-s ghost rows) of the 3-D Laplacian, runs the device-resident block pipeline
with a P-rank communicator, and rank 0 compares against the 1-way oracle: ghost index sets / exchange lists bit-exact
(oracle/partition.py), basis vectors, T entries, Ritz values, orthogonality.  Exit code != 0 on any mismatch.
Used by tests/test_gpu_dist.py (needs >= 2 GPUs) and run by hand with `gpurun --gpus 2`.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from ca_lanczos_b200 import api, gallery
from ca_lanczos_b200.engine import BlockEngine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=40)
    ap.add_argument("--gridz", type=int, default=0)
    ap.add_argument("--sstep", dest="s", type=int, default=8)       # not "--s": torchrun prefix-matches its own options
    ap.add_argument("--blocks", type=int, default=5)
    ap.add_argument("--backend", default="cholqr2")
    ap.add_argument("--layout", default="auto")
    ap.add_argument("--matrix", default="lap3d", choices=["lap3d", "powerlaw"])
    ap.add_argument("--halo-level", type=int, default=0, help="force the ghost-closure depth L (MPK steps per exchange); 0 = automatic")
    ap.add_argument("--own-rows-only", action="store_true", help="supply only the owned rows (forces L = 1, what the C4 bench does)")
    ap.add_argument("--full-reorth", action="store_true", help="also run the 'fro' driver (second projection against all earlier vectors)")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = api.Context(local)
    ids = [api.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.init_comm(world, rank, ids[0])
    if args.halo_level:
        ctx.set_option("mpk_halo_level", args.halo_level)

    m, s = args.grid, args.s
    mz = args.gridz or m
    if args.matrix == "lap3d":
        n, plane = m * m * mz, m * m
        lo, hi = (rank * n) // world, ((rank + 1) * n) // world
        # supply exactly the level-(s) closure hull of a stencil: s planes either side (clipped)
        have_lo, have_hi = max(0, lo - s * plane - plane), min(n, hi + s * plane + plane)
        A_rows = gallery.laplace3d(m, m, mz, row_lo=have_lo, row_hi=have_hi)
        A = gallery.laplace3d(m, m, mz)
    else:
        # C4-like (scaled): power-law SPD, the level-s closure of a row block is (almost) every row, send lists are scattered
        A = gallery.powerlaw_spd(20000, 8.0, seed=0)
        n = A.shape[0]
        lo, hi = (rank * n) // world, ((rank + 1) * n) // world
        have_lo, have_hi = 0, n
        A_rows = A
    if args.own_rows_only:
        have_lo, have_hi = lo, hi
        A_rows = A[lo:hi]
    dm = api.DeviceMatrix(A_rows, s_max=s, layout=args.layout, ctx=ctx, n_glob=n, row_begin=have_lo)
    fails = []

    def check(name, ok, detail=""):
        if not ok:
            fails.append("rank %d: %s %s" % (rank, name, detail))

    # ---- integer objects: bit-exact against the oracle
    from oracle import partition
    b = partition.row_bounds(n, world)
    check("row bounds", (dm.info("row_lo"), dm.info("row_hi")) == (int(b[rank]), int(b[rank + 1])))
    # depth of the ghost closure: forced, 1 when only the owned rows were supplied, else the oracle's restatement of the automatic rule
    L = dm.info("halo_level")
    L_want = min(args.halo_level, s) if args.halo_level else (1 if args.own_rows_only else partition.choose_halo_level(A, world, s))
    check("halo level", L == L_want, "%d vs %d" % (L, L_want))
    ghosts = partition.ghost_indices(A, lo, hi, L)
    check("ghost set", np.array_equal(dm.ghost_indices(), ghosts), "%d vs %d" % (dm.ghost_indices().size, ghosts.size))
    lists = partition.exchange_lists(A, world, L)
    for q in range(world):
        check("recv list from %d" % q, np.array_equal(dm.recv_list(q), lists[rank][q]))
        check("send list to %d" % q, np.array_equal(dm.send_list(q), lists[q][rank]))

    # ---- block pipeline vs the 1-way oracle driver
    from oracle import drivers, kernels
    r = np.cos(0.61 * np.arange(n) ** 1.5) + 0.3 * np.sin(1.7 * np.arange(n))
    io = {}
    if args.matrix == "lap3d":
        shifts = gallery.leja_points(0.0, 12.0, s)
        Bk = np.zeros((s + 1, s)); Bk[np.arange(s), np.arange(s)] = shifts; Bk[np.arange(1, s + 1), np.arange(s)] = 1.0
        To, Qo = drivers.ca_lanczos(A, r, s, s * args.blocks, "newton", "local", Bk=Bk, info=io)
        qtol = 1e-10
    else:
        # shifts from the oracle's own 2s-step Lanczos + Leja ordering (ca_lanczos.m:66-71).  'local' orthogonalisation lets
        # ||I-Q'Q|| grow to ~1e-6 on this matrix, so the ACCUMULATED trajectories of two backward-stable implementations drift
        # apart by that much; the basis vectors are therefore compared per block from identical inputs below (1e-10).
        To, Qo = drivers.ca_lanczos(A, r, s, s * args.blocks, "newton", "local", info=io)
        shifts = np.diag(io["Bk"])[:s].copy()
        qtol = None
    q0 = (r / np.sqrt(r @ r))[lo:hi]
    eng = BlockEngine(dm, s, args.blocks + 1, "newton", shifts, args.backend)
    eng.first_block(q0)
    for _ in range(args.blocks - 1):
        eng.next_block()
    T = eng.T_matrix()
    Qloc = eng.Q_host()
    sc = np.abs(To).max()
    check("second-pass pattern", eng.second == [i["second_pass"] for i in io["pan"]], str(eng.second))
    check("T entries 1e-10", np.abs(T - To).max() <= 1e-10 * sc, "%.3e" % (np.abs(T - To).max() / sc))
    ro = np.sort(np.linalg.eig(To)[0].real)[::-1]; rg = np.sort(np.linalg.eig(T)[0].real)[::-1]
    check("ritz 1e-8", np.max(np.abs(rg[:4] - ro[:4]) / np.abs(ro[:4])) < 1e-8)
    err = np.max(np.linalg.norm(Qloc - Qo[lo:hi, : Qloc.shape[1]], axis=0))
    if qtol is not None:
        check("Q trajectory %.0e" % qtol, err < qtol, "%.3e" % err)
    # ---- per block, from IDENTICAL inputs (the oracle's Qprev and its MPK output): basis vectors within 1e-10 (north_star)
    import ctypes as C
    from ca_lanczos_b200 import _lib
    ld = eng.ld
    blk_err = 0.0
    for k in range(2, args.blocks + 1):
        Vo_k = kernels.matrix_powers_newton(A, Qo[:, (k - 1) * s], s, shifts, 1)
        Qp = torch.zeros((s + 1, ld), dtype=torch.float64, device=dev)
        Xd = torch.zeros((s, ld), dtype=torch.float64, device=dev)
        Zd = torch.zeros((s, ld), dtype=torch.float64, device=dev)
        Qp[:, : hi - lo] = torch.as_tensor(np.ascontiguousarray(Qo[lo:hi, (k - 2) * s:(k - 1) * s + 1].T), device=dev)
        Xd[:, : hi - lo] = torch.as_tensor(np.ascontiguousarray(Vo_k[lo:hi, 1:].T), device=dev)
        torch.cuda.synchronize(dev)
        qb = (C.c_void_p * 1)(Qp.data_ptr()); lds = (C.c_int64 * 1)(ld); mc = (C.c_int * 1)(s + 1)
        R1 = np.zeros((s + 1, s), order="F"); Rl = np.zeros((s, s), order="F")
        rp = (_lib.c_dp * 1)(R1.ctypes.data_as(_lib.c_dp))
        sec, rk = C.c_int(), C.c_int()
        _lib.check(ctx.lib.calz_project_and_normalize(ctx.h, hi - lo, 1, qb, lds, mc, s, C.c_void_p(Xd.data_ptr()), ld, 1,
                                                      _lib.QR[args.backend], C.c_void_p(Zd.data_ptr()), ld, rp,
                                                      Rl.ctypes.data_as(_lib.c_dp), C.byref(sec), C.byref(rk)), ctx.h)
        ctx.sync()
        QZ = Zd[:, : hi - lo].T.cpu().numpy()
        w = min(s, Qo.shape[1] - ((k - 1) * s + 1))               # the driver returns Q(:,1:s*t): the last block lacks its last vector
        blk_err = max(blk_err, float(np.max(np.linalg.norm(QZ[:, :w] - Qo[lo:hi, (k - 1) * s + 1:(k - 1) * s + 1 + w], axis=0))))
    check("per-block basis vectors 1e-10", blk_err < 1e-10, "%.3e" % blk_err)
    # ---- a TWO-block projection whose second pass does NOT fire (the predicated-off branch of the P-rank TSQR/CholQR2 plan):
    # X = the (ortho)normal columns of the third oracle block, mixed by a well-conditioned matrix, plus a 5 % component in both Q blocks
    rng = np.random.default_rng(7)
    Q1o, Q2o = Qo[:, : s + 1], Qo[:, s + 1: 2 * s + 1]
    Xo = Qo[:, 2 * s + 1: 3 * s + 1] @ (np.eye(s) + 0.1 * rng.standard_normal((s, s))) \
        + 0.05 * Q1o @ rng.standard_normal((s + 1, s)) + 0.05 * Q2o @ rng.standard_normal((s, s))
    pinfo = {}
    QZo, RZo = kernels.projectAndNormalize([Q1o, Q2o], Xo, True, backend="tsqr", info=pinfo)
    B1 = torch.zeros((s + 1, ld), dtype=torch.float64, device=dev); B2 = torch.zeros((s, ld), dtype=torch.float64, device=dev)
    Xd = torch.zeros((s, ld), dtype=torch.float64, device=dev); Zd = torch.zeros((s, ld), dtype=torch.float64, device=dev)
    B1[:, : hi - lo] = torch.as_tensor(np.ascontiguousarray(Q1o[lo:hi].T), device=dev)
    B2[:, : hi - lo] = torch.as_tensor(np.ascontiguousarray(Q2o[lo:hi].T), device=dev)
    Xd[:, : hi - lo] = torch.as_tensor(np.ascontiguousarray(Xo[lo:hi].T), device=dev)
    torch.cuda.synchronize(dev)
    qb = (C.c_void_p * 2)(B1.data_ptr(), B2.data_ptr()); lds = (C.c_int64 * 2)(ld, ld); mc = (C.c_int * 2)(s + 1, s)
    Ra = np.zeros((s + 1, s), order="F"); Rb = np.zeros((s, s), order="F"); Rl = np.zeros((s, s), order="F")
    rp = (_lib.c_dp * 2)(Ra.ctypes.data_as(_lib.c_dp), Rb.ctypes.data_as(_lib.c_dp))
    sec, rk = C.c_int(), C.c_int()
    _lib.check(ctx.lib.calz_project_and_normalize(ctx.h, hi - lo, 2, qb, lds, mc, s, C.c_void_p(Xd.data_ptr()), ld, 1,
                                                  _lib.QR[args.backend], C.c_void_p(Zd.data_ptr()), ld, rp,
                                                  Rl.ctypes.data_as(_lib.c_dp), C.byref(sec), C.byref(rk)), ctx.h)
    ctx.sync()
    check("two-block pan: second pass off on both", (bool(sec.value), pinfo["second_pass"]) == (False, False), "%d %s" % (sec.value, pinfo["second_pass"]))
    sgn = np.sign(np.diag(Rl)) * np.sign(np.diag(RZo[2]))                  # the reference's R has no sign convention (tsqr.m)
    e2 = max(float(np.abs(Ra - RZo[0]).max()), float(np.abs(Rb - RZo[1]).max()),
             float(np.abs(Rl - sgn[:, None] * RZo[2]).max()),
             float(np.abs(Zd[:, : hi - lo].T.cpu().numpy() - QZo[lo:hi] * sgn[None, :]).max()))
    check("two-block pan vs oracle 1e-11", e2 < 1e-11, "%.3e" % e2)
    if args.full_reorth:
        # 'fro' driver (ca_lanczos.m:193-197) over P ranks: every block is projected a second time against ALL earlier vectors
        To_f, Qo_f = drivers.ca_lanczos(A, r, s, s * args.blocks, "newton", "full", Bk=io["Bk"])
        ef = BlockEngine(dm, s, args.blocks + 1, "newton", shifts, args.backend)
        ef.first_block(q0)
        for _ in range(args.blocks - 1):
            ef.next_block(full_reorth=True)
        Tf = ef.T_matrix(); Qf = ef.Q_host()
        check("'fro' T 1e-10", np.abs(Tf - To_f).max() <= 1e-10 * np.abs(To_f).max(), "%.3e" % (np.abs(Tf - To_f).max() / np.abs(To_f).max()))
        efq = np.max(np.linalg.norm(Qf - Qo_f[lo:hi, : Qf.shape[1]], axis=0))
        check("'fro' Q 1e-10", efq < 1e-10, "%.3e" % efq)
        from ca_lanczos_b200 import solver
        of = solver.engine_orth_errors(ef)[1]
        check("'fro' orthogonality 5e-13", of < 5e-13, "%.3e" % of)
    # MPK alone, through the host flavour with the communicator (owned rows in, owned rows out)
    v = r / np.sqrt(r @ r)
    V = api.matrix_powers_newton(dm, v[lo:hi], s, shifts, 1)
    Vo = kernels.matrix_powers_newton(A, v, s, shifts, 1)
    Vp = partition.mpk_partitioned(A, v, s, shifts, world, "newton", halo_level=L)
    check("P-way oracle MPK == 1-way", np.max(np.abs(Vp - Vo)) <= 1e-13 * np.max(np.abs(Vo)))
    e = np.max(np.linalg.norm(V - Vo[lo:hi], axis=0) / np.linalg.norm(Vo[lo:hi], axis=0))
    check("mpk owned rows 1e-13", e < 1e-13, "%.3e" % e)
    # identical small results on every rank (deterministic reductions)
    t = torch.tensor(T.ravel(), device=dev)
    t0 = t.clone()
    dist.broadcast(t0, src=0)
    check("T identical on all ranks", bool(torch.equal(t, t0)))

    allf = [None] * world
    dist.all_gather_object(allf, fails)
    if rank == 0:
        flat = [f for fs in allf for f in fs]
        print("dist_check P=%d %s n=%d s=%d backend=%s layout=%s: %s" % (world, args.matrix, n, s, args.backend, dm.layout, "OK" if not flat else "FAILED"))
        print("  max |T-To|/max|To| = %.2e, Q trajectory err %.2e, per-block Q err %.2e, mpk err %.2e, ghosts %d, halo level %d" %
              (np.abs(T - To).max() / sc, err, blk_err, e, ghosts.size, L))
        for f in flat:
            print("  " + f)
    dist.destroy_process_group()
    sys.exit(1 if any(allf) and any(len(x) for x in allf) else 0)


if __name__ == "__main__":
    main()
