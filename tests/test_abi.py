"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/calz.h declares,
fails loudly without a GPU, and its host-only planning functions (partition, ghost level sets) are bit-exact
against the oracle.  No compute entry point is called here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ca_lanczos_b200 import _lib, gallery
from oracle import partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "calz.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(calz_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "libcalz.so does not export " + s
        assert s in _lib.SIGNATURES, "ctypes signature table misses " + s
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.calz_version() >= 100


def test_mex_gateways_only_use_declared_symbols():
    syms = set(_declared_symbols())
    mexdir = os.path.join(ROOT, "mex")
    used = set()
    for f in os.listdir(mexdir):
        if f.endswith((".cpp", ".h")):
            used |= set(re.findall(r"\b(calz_[a-z0-9_]+)\s*\(", open(os.path.join(mexdir, f)).read()))
    helpers = {u for u in used if u.startswith("calz_mex_")} | {"calz_vec_mex", "calz_vec"}    # gateway-local helpers, the handle gateway and its MATLAB class
    assert used - helpers and (used - helpers) <= syms


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    st = lib.calz_init(0, C.byref(h))
    assert st == 2 and not h.value                               # CALZ_ERR_CUDA
    assert b"no CPU fallback" in lib.calz_last_error(None)
    from ca_lanczos_b200 import api
    with pytest.raises(_lib.CalzError):
        api.cholqr(np.ones((8, 2)))


@pytest.mark.parametrize("n,P", [(10, 1), (10, 3), (16777216, 8), (7, 8), (20000000, 8), (100000001, 7)])
def test_partition_bounds_bit_exact(n, P):
    lib = _lib.load()
    b = np.zeros(P + 1, dtype=np.int64)
    assert lib.calz_partition_bounds(n, P, b.ctypes.data_as(_lib.c_i64p)) == 0
    np.testing.assert_array_equal(b, partition.row_bounds(n, P))
    assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)


def _levels_lib(A, lo, hi, s, row_begin=0, row_end=None):
    lib = _lib.load()
    n = A.shape[0]
    row_end = n if row_end is None else row_end
    sub = A[row_begin:row_end]
    rowptr = np.ascontiguousarray(sub.indptr, dtype=np.int64)
    col = np.ascontiguousarray(sub.indices, dtype=np.int32)
    out = np.zeros(n, dtype=np.int32)
    st = lib.calz_level_sets(n, row_begin, row_end, rowptr.ctypes.data_as(_lib.c_i64p), col.ctypes.data_as(_lib.c_i32p),
                             lo, hi, s, out.ctypes.data_as(_lib.c_i32p))
    return st, out


@pytest.mark.parametrize("name,P,s", [("poisson", 4, 3), ("lap3d", 8, 4), ("lap3d", 3, 8), ("powerlaw", 4, 2), ("diag", 4, 8)])
def test_ghost_level_sets_bit_exact_vs_oracle(name, P, s):
    A = {"poisson": lambda: gallery.poisson2d(30), "lap3d": lambda: gallery.laplace3d(12, 10, 24),
         "powerlaw": lambda: gallery.powerlaw_spd(3000, 6.0, seed=1), "diag": lambda: gallery.diag_linspace(101)}[name]()
    n = A.shape[0]
    b = partition.row_bounds(n, P)
    for p in range(P):
        lo, hi = int(b[p]), int(b[p + 1])
        st, lev = _levels_lib(A, lo, hi, s)
        assert st == 0
        np.testing.assert_array_equal(lev, partition.level_sets(A, lo, hi, s))
        ghosts = np.flatnonzero(lev > 0)
        np.testing.assert_array_equal(ghosts, partition.ghost_indices(A, lo, hi, s))


def test_level_sets_report_missing_rows():
    A = gallery.laplace3d(8, 8, 16)            # planes of 64 rows
    st, _ = _levels_lib(A, 512, 768, 3, row_begin=512 - 64, row_end=768 + 64)   # only one ghost plane supplied, need 2
    assert st == 7                              # CALZ_ERR_CLOSURE
    st, _ = _levels_lib(A, 512, 768, 3, row_begin=512 - 128, row_end=768 + 128)
    assert st == 0


def test_exchange_lists_are_sorted_and_partition_the_ghosts():
    A = gallery.laplace3d(6, 6, 16)
    P, s = 4, 3
    lists = partition.exchange_lists(A, P, s)
    b = partition.row_bounds(A.shape[0], P)
    for p in range(P):
        g = partition.ghost_indices(A, int(b[p]), int(b[p + 1]), s)
        cat = np.concatenate(lists[p])
        np.testing.assert_array_equal(cat, g)
        assert lists[p][p].size == 0
        for q in range(P):
            assert np.all(np.diff(lists[p][q]) > 0)
            # slab partition of a stencil: s planes of 36 rows from each direct neighbour only
            if abs(p - q) == 1:
                assert lists[p][q].size == 36 * s
            elif p != q:
                assert lists[p][q].size == 0


@pytest.mark.parametrize("basis", ["newton", "monomial"])
def test_partitioned_mpk_oracle_equals_one_way(basis):
    A = gallery.laplace3d(7, 5, 24)
    v = np.cos(np.arange(A.shape[0]) * 0.37) + 2.0
    lam = np.array([11.0, 1.0, 6.5, 3.0])
    one = partition.mpk_partitioned(A, v, 4, lam, 1, basis)
    for P in (2, 3, 8):
        many = partition.mpk_partitioned(A, v, 4, lam, P, basis)
        assert np.max(np.abs(many - one)) <= 1e-13 * np.max(np.abs(one))
