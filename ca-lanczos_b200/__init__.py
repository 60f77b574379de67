"""B200-native CA-Lanczos hot path: s-step matrix powers kernel + block orthogonalisation.

``from ca_lanczos_b200 import matrix_powers_newton, projectAndNormalize, ...`` gives the reference's call
surface (same names and argument order as the MATLAB functions) backed by libcalz.so; ``gallery`` builds the
synthetic inputs of SURVEY.md §8d; ``engine.BlockEngine`` is the device-resident block pipeline (needs torch).
"""
from . import gallery  # noqa: F401
from . import solver  # noqa: F401
from .api import (  # noqa: F401
    CalzError, Context, DeviceMatrix, SpMV, cholqr, default_context, get_qr_backend, matrix_powers_monomial,
    matrix_powers_newton, normalize, project, projectAndNormalize, set_qr_backend, tsqr,
)
