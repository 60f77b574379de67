"""GPU parity tests of the matrix powers kernel (K1) against the oracle, through the C ABI (ctypes).

Tolerance: basis vectors within 1e-10 relative (north_star); the SpMV sums differ from the oracle only by
fma contraction / summation order, so the tests assert a much tighter 1e-13 per column (2-norm relative).
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

from ca_lanczos_b200 import api, gallery  # noqa: E402
from oracle import kernels  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-13


def relcols(A, B):
    return float(np.max(np.linalg.norm(A - B, axis=0) / np.maximum(np.linalg.norm(B, axis=0), 1e-300)))


def _random_sym(n, deg, seed):
    rng = np.random.default_rng(seed)
    B = sp.random(n, n, density=deg / n, random_state=rng, format="csr", data_rvs=lambda k: rng.standard_normal(k))
    A = (B + B.T + sp.diags(rng.standard_normal(n))).tocsr()
    A.sort_indices()
    return A


MATS = {
    "poisson100": lambda: gallery.poisson2d(100),                 # C1
    "diag20000": lambda: gallery.diag_linspace(20000, 100.0),     # C2 (scaled)
    "lap3d": lambda: gallery.laplace3d(40, 24, 33),               # C3 (scaled, non-cubic, n not a multiple of 32)
    "random": lambda: _random_sym(5003, 9, 7),                    # ragged rows incl. empty ones
    "powerlaw": lambda: gallery.powerlaw_spd(20000, 8.0, seed=0),  # C4 (scaled): very long and very short rows
    "tiny": lambda: gallery.poisson2d(3),                         # n=9 < one SELL slice
}


def _arrow(n=9000):
    """hub rows far beyond the long-row threshold (2048 entries): the first two rows/columns are full"""
    i = np.arange(n)
    T = sp.diags([-np.ones(n - 1), 4.0 + 0.001 * i, -np.ones(n - 1)], [-1, 0, 1]).tolil()
    T[0, :] = np.cos(0.01 * i) * 0.01; T[:, 0] = (np.cos(0.01 * i) * 0.01).reshape(-1, 1)
    T[1, 2:] = -0.002; T[2:, 1] = -0.002
    T[0, 0] = 50.0
    A = T.tocsr(); A.sort_indices()
    return A


@pytest.mark.parametrize("layout", ["csr", "sell"])
def test_long_rows_take_the_segmented_kernel(layout):
    A = _arrow()
    n = A.shape[0]
    dm = api.DeviceMatrix(A, s_max=4, layout=layout)
    assert dm.info("n_long_rows") == 2
    assert dm.info("nnz_loc") == A.nnz
    q = np.cos(0.1 * np.arange(n)) + 1.5
    y = A @ q
    np.testing.assert_allclose(api.SpMV(dm, q), y, rtol=0, atol=1e-13 * np.linalg.norm(y, np.inf))
    lam = np.array([40.0, 3.0, 20.0, 8.0])
    V = api.matrix_powers_newton(dm, q / np.linalg.norm(q), 4, lam, 1)
    Vo = kernels.matrix_powers_newton(A, q / np.linalg.norm(q), 4, lam, 1)
    assert relcols(V, Vo) < 1e-12
    dm.close()


@pytest.mark.parametrize("layout", ["csr", "sell"])
@pytest.mark.parametrize("name", list(MATS))
def test_spmv_and_monomial(name, layout):
    A = MATS[name]()
    n = A.shape[0]
    dm = api.DeviceMatrix(A, s_max=8, layout=layout)
    assert dm.layout == layout
    q = np.cos(0.1 * np.arange(n)) + 1.5
    np.testing.assert_allclose(api.SpMV(dm, q), A @ q, rtol=0, atol=TOL * np.linalg.norm(A @ q, np.inf) + 1e-300)
    s = 5
    V = api.matrix_powers_monomial(dm, q / np.linalg.norm(q), s)
    Vo = kernels.matrix_powers_monomial(A, q / np.linalg.norm(q), s)
    assert V.shape == (n, s)                                      # q itself is NOT returned
    assert relcols(V, Vo) < TOL
    dm.close()


@pytest.mark.parametrize("layout", ["csr", "sell", "auto"])
@pytest.mark.parametrize("name", ["poisson100", "diag20000", "lap3d", "random"])
def test_newton_basis(name, layout):
    A = MATS[name]()
    n = A.shape[0]
    s = 8
    lam = np.array([99.6, 1.4, 45.8, 81.1, 19.9, 64.5, 7.6, 93.4]) * (8.0 / 100.0 if name != "diag20000" else 1.0)
    v = np.ones(n) / np.sqrt(n)
    V = api.matrix_powers_newton(A if layout == "auto" else api.DeviceMatrix(A, 8, layout), v, s, lam, 1)
    Vo = kernels.matrix_powers_newton(A, v, s, lam, 1)
    assert V.shape == (n, s + 1)
    np.testing.assert_array_equal(V[:, 0], v)                     # start vector is column 1
    assert relcols(V, Vo) < TOL
    V0 = api.matrix_powers_newton(A, v, s, lam)                   # modifiedp defaults to 0 (:16-18): same for real shifts
    assert relcols(V0, Vo) < TOL


@pytest.mark.parametrize("name", ["poisson100", "lap3d", "tiny"])
def test_dictionary_coded_sell_is_bit_identical(name):
    # constant-coefficient stencils have a handful of distinct (offset, value) pairs: one code byte per non-zero,
    # same summation order => not a single bit may differ from the plain SELL kernel; layout=auto picks it by itself
    A = MATS[name]()
    n = A.shape[0]
    v = np.sin(0.37 * np.arange(n)) + 1.2
    lam = np.array([7.9, 0.1, 4.0, 6.0, 2.0])
    outs = {}
    for layout in ("sell", "selld", "auto"):
        dm = api.DeviceMatrix(A, 6, layout)
        if layout != "sell":
            assert dm.layout == "selld" and 0 < dm.info("dict_size") <= 9
        outs[layout] = (api.matrix_powers_newton(dm, v, 5, lam, 1), api.matrix_powers_monomial(dm, v, 4), api.SpMV(dm, v))
        dm.close()
    for k in ("selld", "auto"):
        for a, b in zip(outs["sell"], outs[k]):
            np.testing.assert_array_equal(a, b)
    assert relcols(outs["selld"][0], kernels.matrix_powers_newton(A, v, 5, lam, 1)) < TOL


def _few_values_ragged(n=5000, seed=3):
    # few distinct (offset, value) pairs but a different sparsity pattern in every row: the codes of a warp diverge
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    offs = np.array([-40, -7, -1, 0, 1, 7, 40, 300])
    vals = np.array([-1.0, 0.5, -2.0, 6.0, -2.0, 0.5, -1.0, 0.25])
    rows, cols, data = [], [], []
    for r in range(n):
        for k in np.flatnonzero(rng.random(offs.size) < 0.6):
            c = r + offs[k]
            if 0 <= c < n:
                rows.append(r); cols.append(c); data.append(vals[k])
    return sp.csr_matrix((data, (rows, cols)), shape=(n, n))


@pytest.mark.parametrize("name", ["lap3d", "ragged"])
def test_dictionary_kernel_variants_are_bit_identical(name):
    # the dictionary lives in the constant bank (warp-uniform codes) or in shared memory (ragged codes); persistent
    # software-pipelined grid or one work item per warp: four kernels, one answer
    A = MATS["lap3d"]() if name == "lap3d" else _few_values_ragged()
    n = A.shape[0]
    v = np.cos(0.11 * np.arange(n)) + 0.3
    lam = np.array([5.0, 0.5, 3.0, 1.0])
    ctx = api.default_context()
    ref_dm = api.DeviceMatrix(A, 4, "sell")
    ref = (api.matrix_powers_newton(ref_dm, v, 4, lam, 1), api.matrix_powers_monomial(ref_dm, v, 3))
    ref_dm.close()
    dm = api.DeviceMatrix(A, 4, "selld")
    uni = dm.info("dict_uniform_pct")
    assert 0 <= uni <= 100 and (name == "lap3d" or uni < 50)     # short grid lines: partly uniform; ragged: not at all
    try:
        for patterns, fused in ((1, 0), (0, 0), (0, 1)):          # slice-pattern kernel / coded kernels / one cooperative launch
            for dmode in (-1, 0, 2):
                for persist in (1, 0):
                    ctx.set_option("mpk_patterns", patterns)
                    ctx.set_option("mpk_fused_steps", fused)
                    ctx.set_option("mpk_dict_mode", dmode)
                    ctx.set_option("mpk_persist", persist)
                    np.testing.assert_array_equal(api.matrix_powers_newton(dm, v, 4, lam, 1), ref[0])
                    np.testing.assert_array_equal(api.matrix_powers_monomial(dm, v, 3), ref[1])
    finally:
        ctx.set_option("mpk_dict_mode", -1)
        ctx.set_option("mpk_persist", 1)
        ctx.set_option("mpk_patterns", 1)
        ctx.set_option("mpk_fused_steps", 0)
        dm.close()


def test_slice_patterns_of_a_stencil():
    """7-point Laplacian 64 x 40 x 33: 2 slices per grid line, the line ends miss one neighbour on one lane (lane masks); the three
    boundary states of y and z give 27 patterns at most and every complete slice has one; pattern 0 is the interior."""
    A = gallery.laplace3d(64, 40, 33)
    n = A.shape[0]
    ctx = api.default_context()
    ref_dm = api.DeviceMatrix(A, 4, "sell")
    v = np.cos(0.11 * np.arange(n)) + 0.3
    lam = np.array([5.0, 0.5, 3.0, 1.0])
    ref = (api.matrix_powers_newton(ref_dm, v, 4, lam, 1), api.matrix_powers_monomial(ref_dm, v, 3))
    ref_dm.close()
    dm = api.DeviceMatrix(A, 4, "selld")
    assert 1 <= dm.info("n_patterns") <= 32
    assert dm.info("pattern_cover_pct") >= 99
    assert dm.info("pattern_pair_phase") in (0, 1)
    try:
        for prefetch in (1, 0, 8):                                 # several L2 prefetch distances
            for phase in (-1, 0, 1):                               # the matrix' own slice pairing, and both forced ones
                ctx.set_option("mpk_prefetch", prefetch)
                ctx.set_option("mpk_pair_phase", phase)
                np.testing.assert_array_equal(api.matrix_powers_newton(dm, v, 4, lam, 1), ref[0])
                np.testing.assert_array_equal(api.matrix_powers_monomial(dm, v, 3), ref[1])
    finally:
        ctx.set_option("mpk_prefetch", 1)
        ctx.set_option("mpk_pair_phase", -1)
        dm.close()


def test_slice_pairing_phase_of_a_256_wide_grid():
    # 8 slices per grid line, the first and the last one touch the x boundary: pairing (7 | 0 of the next line) keeps three items
    # of four on the interior path -- the set-up must pick phase 1; odd slice counts and both phases give the same bits as plain SELL
    A = gallery.laplace3d(256, 6, 5)
    n = A.shape[0]
    v = np.sin(0.07 * np.arange(n)) + 0.2
    lam = np.array([4.0, 1.0, 2.5])
    ctx = api.default_context()
    ref_dm = api.DeviceMatrix(A, 3, "sell")
    ref = api.matrix_powers_newton(ref_dm, v, 3, lam, 1)
    ref_dm.close()
    dm = api.DeviceMatrix(A, 3, "selld")
    try:
        assert dm.info("pattern_pair_phase") == 1
        for phase in (-1, 0, 1):
            ctx.set_option("mpk_pair_phase", phase)
            np.testing.assert_array_equal(api.matrix_powers_newton(dm, v, 3, lam, 1), ref)
    finally:
        ctx.set_option("mpk_pair_phase", -1)
        dm.close()


@pytest.mark.parametrize("name", ["poisson100", "lap3d"])
def test_tma_staged_dictionary_kernel_is_bit_identical(name):
    # opt-in variant: the x segments of a CTA's row block are bulk-copied into shared memory (cp.async.bulk + mbarrier)
    A = MATS[name]()
    n = A.shape[0]
    v = np.cos(0.23 * np.arange(n)) + 0.7
    lam = np.array([7.9, 0.1, 4.0, 6.0])
    ctx = api.default_context()
    dm = api.DeviceMatrix(A, 4, "selld")
    assert dm.info("xs_rows") >= 512 and 1 <= dm.info("xs_groups") <= 3
    V0 = api.matrix_powers_newton(dm, v, 4, lam, 1)
    ctx.set_option("mpk_tma_x", 1)
    try:
        V1 = api.matrix_powers_newton(dm, v, 4, lam, 1)
        M1 = api.matrix_powers_monomial(dm, v, 3)
    finally:
        ctx.set_option("mpk_tma_x", 0)
    np.testing.assert_array_equal(V0, V1)
    np.testing.assert_array_equal(M1, api.matrix_powers_monomial(dm, v, 3))


def test_dictionary_coded_sell_falls_back_when_values_are_not_few():
    A = MATS["diag20000"]()                     # 20000 distinct values
    with pytest.raises(api.CalzError):
        api.DeviceMatrix(A, 4, "selld")
    dm = api.DeviceMatrix(A, 4, "auto")
    assert dm.layout == "sell"
    A = MATS["random"]()
    assert api.DeviceMatrix(A, 4, "auto").layout in ("sell", "csr")


def test_newton_complex_pair_and_errors():
    A = gallery.poisson2d(30)
    v = np.sin(np.arange(900.0)) + 2
    lam = np.array([3 + 0.5j, 3 - 0.5j, 1.0, 6 + 2j, 6 - 2j])
    V = api.matrix_powers_newton(A, v, 5, lam, 1)
    assert relcols(V, kernels.matrix_powers_newton(A, v, 5, lam, 1)) < TOL
    with pytest.raises(ValueError):                               # matrix_powers_newton.m:36-39
        api.matrix_powers_newton(A, v, 2, np.array([3 - 0.5j, 3 + 0.5j]), 1)
    with pytest.raises(api.CalzError):                            # complex vectors are out of scope
        api.matrix_powers_newton(A, v, 2, np.array([3 - 0.5j, 3 + 0.5j]), 0)
    with pytest.raises(ValueError):
        api.matrix_powers_newton(A, v[:-1], 2, lam, 1)


def test_csc_ingest_of_nonsymmetric_matrix():
    # the MEX gateways hand MATLAB's CSC arrays to calz_mat_create_csc64, which transposes on the host
    import ctypes as C
    from ca_lanczos_b200 import _lib
    rng = np.random.default_rng(3)
    A = sp.random(700, 700, density=0.01, random_state=rng, format="csc") + sp.eye(700, format="csc") * 2
    A = A.tocsc(); A.sort_indices()
    ctx = api.default_context()
    jc = A.indptr.astype(np.uint64); ir = A.indices.astype(np.uint64); pr = A.data.astype(np.float64)
    h = C.c_void_p()
    _lib.check(ctx.lib.calz_mat_create_csc64(ctx.h, 700, jc.ctypes.data_as(C.POINTER(C.c_uint64)),
                                             ir.ctypes.data_as(C.POINTER(C.c_uint64)), pr.ctypes.data_as(_lib.c_dp), 4, 0,
                                             C.byref(h)), ctx.h)
    x = rng.standard_normal(700); y = np.empty(700)
    _lib.check(ctx.lib.calz_spmv_host(h, x.ctypes.data_as(_lib.c_dp), y.ctypes.data_as(_lib.c_dp)), ctx.h)
    np.testing.assert_allclose(y, A @ x, rtol=1e-13, atol=1e-13)
    ctx.lib.calz_mat_destroy(h)


@pytest.mark.parametrize("chunk", [1 << 16, 1 << 20])
def test_l2_temporal_blocking_is_bit_identical(chunk):
    # the skewed chunk schedule only reorders launches: results must not change by a single bit
    A = gallery.laplace3d(32, 32, 64)
    v = np.ones(A.shape[0]) / np.sqrt(A.shape[0])
    lam = np.array([11.9, 0.2, 6.0, 9.0, 3.0, 7.5])
    ctx = api.default_context()
    dm = api.DeviceMatrix(A, 6, "sell")
    V0 = api.matrix_powers_newton(dm, v, 6, lam, 1)
    ctx.set_option("mpk_l2_chunk_bytes", chunk)
    try:
        V1 = api.matrix_powers_newton(dm, v, 6, lam, 1)
    finally:
        ctx.set_option("mpk_l2_chunk_bytes", 0)
    np.testing.assert_array_equal(V0, V1)
    assert relcols(V0, kernels.matrix_powers_newton(A, v, 6, lam, 1)) < TOL


@pytest.mark.parametrize("name", ["c1_poisson_s4_monomial", "c2_diag_s8_newton"])
def test_mpk_against_committed_goldens(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    A = gallery.poisson2d(int(g["m"])) if "poisson" in name else gallery.diag_linspace(int(g["n"]), 100.0)
    r = np.ones(A.shape[0]); q = r / np.sqrt(r @ r)
    s = int(g["s"])
    if str(g["basis"]) == "newton":
        V = api.matrix_powers_newton(A, q, s, g["shifts"], 1)
    else:
        V = np.column_stack([q, api.matrix_powers_monomial(A, q, s)])
    assert relcols(V[g["rows"]], g["V_rows"]) < 1e-12


def test_full_size_c2_basis_identity():
    # C2 at BASELINE size (n = 1e6): size-independent property  A V(:,1:s) = V(:,1:s+1) B  (newton_basis_matrix.m:3-4)
    n, s = 1_000_000, 8
    A = gallery.diag_linspace(n, 100.0)
    lam = np.array([99.5677, 1.4323, 45.7883, 81.1407, 19.8593, 64.4648, 7.5729, 93.4271])
    v = np.ones(n) / np.sqrt(n)
    V = api.matrix_powers_newton(A, v, s, lam, 1)
    B = np.zeros((s + 1, s)); B[np.arange(s), np.arange(s)] = lam; B[np.arange(1, s + 1), np.arange(s)] = 1
    assert np.linalg.norm(A @ V[:, :s] - V @ B) <= 1e-13 * np.linalg.norm(V)
    d = np.linspace(1, 100, n)                                    # diagonal A: closed form prod(d - lam_i) * v
    exact = np.prod(d[:, None] - lam[None, :], axis=1) * v        # rows with d ~ lam_i cancel: compare norm-wise
    assert np.linalg.norm(V[:, s] - exact) <= 1e-12 * np.linalg.norm(exact)
