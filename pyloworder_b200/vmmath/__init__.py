"""`pyLOM.vmmath` (alias `pyLOM.math`) entry points of the POD hot path, same names and signatures
as pyLOM/vmmath/__init__.py:10-18; every call lands in hand-written sm_100a CUDA through the C ABI."""
from .maths import matmul, matmulp, vecmat, vector_sum, vector_norm
from .averaging import temporal_mean, subtract_mean, temporal_variance, norm_variance
from .truncation import compute_truncation_residual
from .stats import RMSE, energy
from .svd import (qr, svd, tsqr, tsqr_svd, randomized_qr, randomized_svd, init_qr_streaming, update_qr_streaming,
                  next_power_of_2)
