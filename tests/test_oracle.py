"""CPU tests of the oracle (oracle/): analytic known answers, mathematical identities, committed goldens.

The reference ships no golden vectors for this path ("parity unpinned"); what pins the oracle
is (a) the analytic spectra of the reference's own synthetic tests, (b) identities that any correct
implementation satisfies, (c) fixtures generated from the oracle itself (regression pins).
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from ca_lanczos_b200 import gallery
from oracle import drivers, kernels, leja

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _relcols(A, B):
    return np.max(np.linalg.norm(A - B, axis=0) / np.maximum(np.linalg.norm(B, axis=0), 1e-300))


# ----------------------------------------------------------------------------- MPK
def test_monomial_excludes_start_vector_and_newton_includes_it():
    A = gallery.poisson2d(12)
    q = np.linspace(1, 2, A.shape[0])
    Vm = kernels.matrix_powers_monomial(A, q, 4)
    Vn = kernels.matrix_powers_newton(A, q, 4, np.zeros(4), 1)
    assert Vm.shape == (144, 4) and Vn.shape == (144, 5)       # matrix_powers_monomial.m:7 / _newton.m:22-23
    np.testing.assert_array_equal(Vn[:, 0], q)
    np.testing.assert_allclose(Vn[:, 1:], Vm, rtol=0, atol=0)  # zero shifts == monomial, bit for bit


def test_newton_basis_identity_A_Vs_equals_Vs1_B():
    # newton_basis_matrix.m:3-4:  A V(:,1:s) = V B
    A = gallery.poisson2d(20)
    s = 6
    lam = np.array([7.9, 0.1, 4.0, 6.0, 2.0, 5.0])
    v = np.ones(A.shape[0]) / 20.0
    V = kernels.matrix_powers_newton(A, v, s, lam, 1)
    B = leja.newton_basis_matrix(lam, s, 1)
    assert np.linalg.norm(A @ V[:, :s] - V @ B) <= 1e-13 * np.linalg.norm(V)


def test_newton_modified_complex_pair_term():
    # matrix_powers_newton.m:33-41: real part on both members, + imag^2 * V(:,k-1) on the second
    A = gallery.poisson2d(8)
    v = np.arange(1.0, 65.0)
    lam = np.array([2 + 1j, 2 - 1j, 3.0])
    V = kernels.matrix_powers_newton(A, v, 3, lam, 1)
    np.testing.assert_allclose(V[:, 1], A @ v - 2 * v)
    np.testing.assert_allclose(V[:, 2], A @ V[:, 1] - 2 * V[:, 1] + 1.0 * V[:, 0])
    B = leja.newton_basis_matrix(lam, 3, 1)
    assert np.linalg.norm(A @ V[:, :3] - V @ B) <= 1e-12 * np.linalg.norm(V)
    with pytest.raises(ValueError):          # :36-39
        kernels.matrix_powers_newton(A, v, 2, np.array([2 - 1j, 2 + 1j]), 1)


# ----------------------------------------------------------------------------- QR
@pytest.mark.parametrize("c", [1, 5, 9, 17])
def test_tsqr_cholqr_unique_R(c):
    X = gallery.tall_skinny(4000, c, seed=3)
    Q1, R1 = kernels.tsqr(X)
    Q2, R2 = kernels.cholqr(X)
    assert np.all(np.diag(R1) > 0) and np.all(np.diag(R2) > 0)
    assert np.allclose(np.tril(R1, -1), 0)
    kappa = np.linalg.cond(X)
    assert np.linalg.norm(R1 - R2) <= 50 * kappa ** 2 * 1e-16 * np.linalg.norm(R1)     # QR uniqueness
    assert np.linalg.norm(Q1 @ R1 - X) <= 1e-14 * np.linalg.norm(X)
    assert np.linalg.norm(Q1.T @ Q1 - np.eye(c)) <= 1e-14 * c


def test_tsqr_zero_pivot_sign_is_zero():
    X = np.zeros((10, 2), order="F"); X[:, 0] = 1.0     # second column exactly zero => sign(0)=0 (tsqr.m:9)
    Q, R = kernels.tsqr(X)
    assert R[1, 1] == 0 and np.all(Q[:, 1] == 0)


def test_cholqr_raises_when_not_pd():
    X = np.ones((10, 2), order="F")
    with pytest.raises(np.linalg.LinAlgError):
        kernels.cholqr(X)


def test_normalize_rank():
    X = gallery.tall_skinny(500, 4, seed=1)
    X[:, 3] = X[:, 0] * 2 + 1e-12 * X[:, 1]
    _, _, rank = kernels.normalize(X)
    assert rank == 3
    _, _, rank = kernels.normalize(gallery.tall_skinny(500, 4, seed=1))
    assert rank == 4


# ----------------------------------------------------------------------------- project / projectAndNormalize
def test_project_cells_and_empty_blocks():
    X = gallery.tall_skinny(300, 3, seed=5)
    Q1, _ = kernels.tsqr(gallery.tall_skinny(300, 4, seed=6))
    Y, R = kernels.project([Q1, None, np.zeros((300, 0))], X)
    assert R[1] is None and R[2] is None and R[0].shape == (4, 3)
    np.testing.assert_allclose(Q1.T @ Y, 0, atol=1e-14)
    Y0, R0 = kernels.project([], X)
    assert R0 == [] and np.array_equal(Y0, X)


@pytest.mark.parametrize("backend", ["tsqr", "cholqr"])
def test_pan_reconstruction_identity_and_second_pass(backend):
    # Appendix B: X = sum_i Q{i} RZ{i} + QZ RZ{end} holds whether or not pass 2 fires
    n = 2000
    Qa, _ = kernels.tsqr(gallery.tall_skinny(n, 5, seed=11))
    Qb, _ = kernels.tsqr(kernels.project([Qa], gallery.tall_skinny(n, 3, seed=12))[0])
    X_far = gallery.tall_skinny(n, 4, seed=13)                       # mostly outside span(Q): no pass 2
    X_near = Qa @ np.ones((5, 4)) + 1e-3 * gallery.tall_skinny(n, 4, seed=14)   # >50 % norm drop: pass 2
    for X, want in ((X_far, False), (X_near, True)):
        info = {}
        QZ, RZ = kernels.projectAndNormalize([Qa, Qb], X, True, backend=backend, info=info)
        assert info["second_pass"] is want
        rec = Qa @ RZ[0] + Qb @ RZ[1] + QZ @ RZ[2]
        assert np.linalg.norm(rec - X) <= 1e-13 * np.linalg.norm(X)
        assert np.linalg.norm(QZ.T @ QZ - np.eye(4)) <= 1e-10
        assert np.linalg.norm(Qa.T @ QZ) <= 1e-12


# ----------------------------------------------------------------------------- Leja / Newton basis (host)
def test_leja_quirk_any_second_argument_is_modified_branch():
    x = np.array([1.0, 5.0, 2.0, 9.0, 4.0])
    y, idx = leja.leja(x, "nonmodified")            # leja.m:24-30: takes real_leja
    assert sorted(y.tolist()) == pytest.approx(sorted(x.tolist()), rel=1e-14)
    assert y[0] == pytest.approx(9.0, rel=1e-14)    # max modulus first
    assert y[1] == pytest.approx(1.0, rel=1e-14)    # then the farthest point
    with pytest.raises(NotImplementedError):
        leja.leja(x)


def test_leja_points_drift_by_a_few_ulp_only():
    x = np.linspace(1.3, 97.1, 16)
    y, _ = leja.leja(x, "nonmodified")
    assert np.max(np.abs(np.sort(y) - x) / x) < 1e-14


def test_leja_repeated_shifts_raise_like_reference():
    with pytest.raises(IndexError):                 # real_leja.m:87 passes the original n; modified_leja.m:49 x(1:n) overruns
        leja.leja(np.array([1.0, 2.0, 2.0, 3.0]), "nonmodified")


# ----------------------------------------------------------------------------- drivers: analytic spectra
def test_ca_lanczos_poisson_ritz_values_match_analytic_spectrum():
    # C1: gallery('poisson',100), s=4 monomial, 60 steps, r=ones; eigenvalues 4-2cos(i pi/(m+1))-2cos(j pi/(m+1))
    m = 100
    A = gallery.poisson2d(m)
    r = np.ones(m * m)
    info = {}
    T, Q = drivers.ca_lanczos(A, r, 4, 60, "monomial", "local", info=info)
    ritz = np.sort(np.linalg.eig(T)[0].real)[::-1]
    # survey sanity values (restatement of the reference)
    np.testing.assert_allclose(ritz[:5], [7.985225384387, 7.960350846824, 7.924548934762, 7.877896460350,
                                          7.820514144716], rtol=1e-10)
    k = np.arange(1, m + 1)
    lam1 = 2 - 2 * np.cos(k * np.pi / (m + 1))
    spectrum = np.sort((lam1[:, None] + lam1[None, :]).ravel())
    assert abs(ritz[0] - spectrum[-1]) < 2e-2 and ritz[0] <= spectrum[-1] + 1e-10   # Ritz values lie inside the spectrum
    assert sum(i["second_pass"] for i in info["pan"]) >= 13                       # pass 2 fires ~every block
    assert np.linalg.norm(np.eye(Q.shape[1]) - Q.T @ Q, "fro") < 1e-9


def test_ca_lanczos_diagonal_newton_converges_to_extreme_eigenvalues():
    # test_convergence_diagonal_matrices.m:9-22 style: diag(linspace(1,100,N)), r=ones, Newton basis
    N, s = 500, 8
    A = gallery.diag_linspace(N, 100.0)
    T, Q = drivers.ca_lanczos(A, np.ones(N), s, 160, "newton", "full")
    ritz = np.sort(np.linalg.eig(T)[0].real)
    exact = np.linspace(1, 100, N)
    assert abs(ritz[-1] - exact[-1]) < 1e-8 * 100 and abs(ritz[0] - exact[0]) < 1e-8 * 100
    assert np.linalg.norm(np.eye(Q.shape[1]) - Q.T @ Q, "fro") < 1e-8


def test_restarted_ca_lanczos_diagonal():
    # test_restart_diagonal_matrices.m:8-36 scaled down: wanted = largest eigenvalues of diag(linspace(1,K,N))
    N, K = 400, 1.0e2
    A = gallery.diag_linspace(N, K)
    eigs, Qc, nrest, rnorms, oerr = drivers.restarted_ca_lanczos(A, np.ones(N), 40, 4, 4, "newton", "full", 1e-8)
    exact = np.linspace(1, K, N)[::-1][:4]
    np.testing.assert_allclose(eigs, exact, rtol=1e-8)
    assert Qc.shape == (N, 4) and nrest < 200


# ----------------------------------------------------------------------------- goldens (regression pins)
@pytest.mark.parametrize("name", ["c1_poisson_s4_monomial", "c2_diag_s8_newton"])
def test_oracle_matches_committed_goldens(name):
    path = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated")
    g = np.load(path)
    A = gallery.poisson2d(int(g["m"])) if "poisson" in name else gallery.diag_linspace(int(g["n"]), 100.0)
    r = np.ones(A.shape[0])
    T, Q = drivers.ca_lanczos(A, r, int(g["s"]), int(g["iter"]), str(g["basis"]), "local")
    np.testing.assert_allclose(T, g["T"], rtol=0, atol=1e-9 * np.abs(g["T"]).max())
    ritz = np.sort(np.linalg.eig(T)[0].real)[::-1]
    np.testing.assert_allclose(ritz[:8], g["ritz"][:8], rtol=1e-9)
    q = r / np.sqrt(r @ r)
    if str(g["basis"]) == "newton":
        V = kernels.matrix_powers_newton(A, q, int(g["s"]), g["shifts"], 1)
    else:
        V = np.column_stack([q, kernels.matrix_powers_monomial(A, q, int(g["s"]))])
    rows = g["rows"]
    assert _relcols(V[rows], g["V_rows"]) < 1e-12


def test_leja_shifts_agree_with_the_surveys_independent_restatement_up_to_the_tie():
    # SURVEY.md 8c [scratch]: C2-template (N=500, linspace(1,100), s=8), first 8 Leja-ordered shifts from an INDEPENDENT restatement:
    # 99.5677 1.4323 55.2117 19.8593 81.1407 36.5352 93.4271 7.5729.  The spectrum (and so the 16 Ritz values) is symmetric about
    # 50.5, which makes the third Leja choice an exact tie between x and 101-x (|x-a||x-b| is symmetric); the winner depends on
    # the last bit of pow() and every later choice mirrors it.  Both restatements must agree up to that mirror image.
    A = gallery.diag_linspace(500, 100.0)
    io = {}
    drivers.ca_lanczos(A, np.ones(500), 8, 16, "newton", "local", info=io)
    mine = np.diag(io["Bk"])[:8]
    survey = np.array([99.5677, 1.4323, 55.2117, 19.8593, 81.1407, 36.5352, 93.4271, 7.5729])
    np.testing.assert_allclose(mine[:2], survey[:2], atol=5e-5)
    same = np.allclose(mine[2:], survey[2:], atol=5e-5)
    mirrored = np.allclose(mine[2:], 101.0 - survey[2:], atol=5e-5)
    assert same or mirrored


# ---------------------------------------------------------------------------------------------- round 2 additions
def test_powerlaw_rows_generator_matches_whole_matrix():
    """the per-rank generator of the C4 matrix: any row slice equals the same rows of the whole matrix, symmetric, SPD by dominance"""
    n = 4000
    A = gallery.powerlaw_spd_rows(n, 10.0, seed=3)
    assert abs(A - A.T).max() == 0
    d = A.diagonal()
    assert np.all(d - (abs(A).sum(axis=1).A1 - d) >= 1.0 - 1e-12)
    for lo, hi in ((0, 500), (1234, 3000), (3000, 4000)):
        S = gallery.powerlaw_spd_rows(n, 10.0, seed=3, row_lo=lo, row_hi=hi)
        assert abs(S - A[lo:hi]).max() == 0



def test_powerlaw_rows_generator_jacobi_scaling():
    n = 3000
    A = gallery.powerlaw_spd_rows(n, 10.0, seed=3, jacobi=True)
    A0 = gallery.powerlaw_spd_rows(n, 10.0, seed=3)
    d = 1.0 / np.sqrt(A0.diagonal())
    import scipy.sparse as sp
    assert abs(sp.diags(d) @ A0 @ sp.diags(d) - A).max() < 1e-15
    assert abs(A - A.T).max() == 0 and np.all(A.diagonal() == 1.0)
    S = gallery.powerlaw_spd_rows(n, 10.0, seed=3, row_lo=777, row_hi=1999, jacobi=True)
    assert abs(S - A[777:1999]).max() == 0
    ev = np.linalg.eigvalsh(A.toarray())
    assert ev[0] > 0 and ev[-1] < 2.0                               # SPD, spectrum of a normalised Laplacian + I scaled


def test_periodic_driver_on_the_references_own_diagonal_test():
    """test_convergence_diagonal_matrices.m:9-22: diag(linspace(1,100,500)), r = ones, 480 steps, s = 8 Newton, PERIODIC
    orthogonalisation (ca_lanczos.m:362-467 with update_omega :469-539 / reset_omega :541-551).  The eigenvalues are known
    analytically (the diagonal): the restated driver must converge to them and keep semi-orthogonality."""
    from oracle import drivers
    N = 500
    A = gallery.diag_linspace(N, 100.0)
    io = {}
    T, Q = drivers.ca_lanczos(A, np.ones(N), 8, 480, "newton", "periodic", info=io)
    ev = np.sort(np.linalg.eig(T)[0].real)[::-1]
    exact = np.linspace(1.0, 100.0, N)[::-1]
    np.testing.assert_allclose(ev[:20], exact[:20], rtol=1e-10)
    assert io["nbreaks"] >= 10 and io["breaks"] == sorted(io["breaks"])
    assert np.linalg.norm(np.eye(Q.shape[1]) - Q.T @ Q) < 1e-7      # sqrt(eps)-level semi-orthogonality, by construction
    # 'local' on the same problem loses orthogonality completely: the test would fail without the omega recurrence
    Tl, Ql = drivers.ca_lanczos(A, np.ones(N), 8, 480, "newton", "local")
    assert np.linalg.norm(np.eye(Ql.shape[1]) - Ql.T @ Ql) > 1.0


def test_omega_recurrence_product_restatement_equals_oracle():
    """solver.update_omega / reset_omega (product, host algebra of the device driver) against oracle.drivers (test infrastructure)"""
    from ca_lanczos_b200 import solver
    from oracle import drivers
    rng = np.random.default_rng(0)
    s = 4
    om_o = om_s = None
    for k in range(1, 6):
        alpha = rng.uniform(1, 2, s * k); beta = rng.uniform(0.1, 1, s * k)
        om_o = drivers.update_omega(om_o, alpha, beta, 3.0, s)
        om_s = solver.update_omega(om_s, alpha, beta, 3.0, s)
        np.testing.assert_array_equal(om_o, om_s)
        if k == 3:
            om_o = drivers.reset_omega(om_o, 3.0, s); om_s = solver.reset_omega(om_s, 3.0, s)
            np.testing.assert_array_equal(om_o, om_s)


def test_halo_level_rule_and_multi_exchange_mpk():
    from oracle import partition
    A = gallery.laplace3d(8, 8, 64)
    assert partition.choose_halo_level(A, 8, 8) == 4                # 8 planes per rank: 2 x 4 ghost planes <= 8 owned
    assert partition.choose_halo_level(A, 2, 8) == 8
    B = gallery.powerlaw_spd(3000, 6.0, seed=1)
    assert partition.choose_halo_level(B, 4, 4) == 1                # the level-1 closure is (almost) every row
    v = np.cos(np.arange(B.shape[0]) * 0.37) + 2.0
    lam = np.array([30.0, 2.0, 15.0, 7.0])
    one = partition.mpk_partitioned(B, v, 4, lam, 1)
    for P, L in ((4, 1), (3, 2), (2, 4)):
        many = partition.mpk_partitioned(B, v, 4, lam, P, halo_level=L)
        assert np.max(np.abs(many - one)) <= 1e-13 * np.max(np.abs(one))
