"""paddle.linalg namespace. Parity: python/paddle/linalg.py."""
from .ops.linalg import (cholesky, cholesky_inverse, cholesky_solve, cond, corrcoef, cov, det, eig, eigh, eigvals, eigvalsh,  # noqa: F401
                         fp8_fp8_half_gemm_fused, householder_product, inv, lstsq, lu, lu_solve, lu_unpack, matrix_exp, matrix_norm,
                         matrix_power, matrix_rank, multi_dot, norm, ormqr, pca_lowrank, pinv, qr, slogdet, solve, svd, svd_lowrank,
                         svdvals, triangular_solve, vecdot, vector_norm)
from .ops.linalg import matmul, cross, dist, bmm, mv, dot  # noqa: F401
from .ops.math import diagonal  # noqa: F401
