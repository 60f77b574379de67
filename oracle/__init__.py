"""CPU oracle for the CA-Lanczos hot path.  TEST INFRASTRUCTURE ONLY -- never shipped, never measured as
the product.

PARITY UNPINNED: the reference (magnusgrandin/ca-lanczos) is 100 % MATLAB; neither Octave nor MATLAB exists
in this image or on the GPU boxes, and the reference holds no golden vectors, KATs or fixtures for this
path.  The oracle is therefore a numpy/scipy restatement that follows the reference's
operation order line by line (citations in every docstring), pinned only by (a) analytic spectra of the
reference's own synthetic tests (test_convergence_diagonal_matrices.m:16-19,
test_restart_diagonal_matrices.m:21-31), (b) mathematical identities (QR uniqueness, the basis identity
A V_s = V_{s+1} B, the reconstruction identity of projectAndNormalize) and (c) fixtures generated FROM
THIS ORACLE by tests/golden/make_golden.py.  tools/dump_goldens.m regenerates (c) from the real reference
on any machine that has Octave/MATLAB.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package.
"""
from . import kernels, leja, drivers, partition  # noqa: F401
