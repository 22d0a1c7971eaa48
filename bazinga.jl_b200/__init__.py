"""bazinga_b200 -- B200 (sm_100a) implementation of the PANOC+/ALPS hot path of aldma/Bazinga.jl.

The directory is named ``bazinga.jl_b200`` after the reference; import it as ``bazinga_b200`` (the
shim ``bazinga_b200.py`` at the repository root registers this directory under that name).

Host code here is a thin mirror of the reference's operator/plugin interface over the C ABI in
``include/bazinga_b200.h``; all arithmetic runs in the CUDA kernels under ``csrc/``.
"""
from ._capi import BazingaError, LIB_PATH, STATUS_SYMBOLS, load_library  # noqa: F401
from .oracles import (  # noqa: F401
    ClosedSet, ConstraintBasisPursuit, ConstraintObstacle, ConstraintObstacleRed, FreeSet, IdentityFunction, IndBox, IndFree,
    IndicatorSet, LeastSquares, NonsmoothCostObstacleL1, NonsmoothCostObstacleL2, NonsmoothCostObstacleL2Red, NonsmoothCostObstacleRedL1,
    SmoothCostObstacleL2, NormL0, NormL0Box, NormL1, NormL1Box,
    NormL1Nonneg, NormLpPowerBox, NormLpPowerNonneg, PairSet, Quadratic, SetEOR, SetMPVCA, SetP, SetXOR, SmoothCostObstacleL1, SmoothCostP,
    ConstraintP, Zero, ZeroSet,
)
from .matrixprox import NuclearNorm, Rank, SchattenNormLpPower, check_and_reshape_as_matrix  # noqa: F401
from .results import DATA_COLUMNS, STATS_COLUMNS, demo_rows, demo_stats, write_demo_results  # noqa: F401
from .solver import (  # noqa: F401
    BATCH_KINDS, Context, LBFGS, NoAcceleration, PANOCplus, Problem, SharedMatrixBatch, alps, als, alps_batch, als_batch,
    comm_unique_id, default_context, default_subsolver, row_shard,
)
