"""scimloperators.jl_b200 -- B200-native evaluation backend for the cached in-place `mul!` path of
SciMLOperators.jl (TensorProductOperator and Added/Composed/Scaled operator trees).

The product is `libscimlop_b200.so` (csrc/, C ABI in include/scimlop_b200.h); this package is the host-side
mirror of the reference's operator interface over that ABI.  The directory name contains a dot, so import it
through the `scimlop_b200` shim at the repo root (`import scimlop_b200 as sb`)."""
from . import _lib
from ._lib import Handle, ScimlopError, handle, load, OPT_TAIL_SPLIT, OPT_PEER_PIPELINE, OPT_PEER_CHUNKS, OPT_PEER_STAGE2_SMS, OPT_MODE_PIPELINE
from .array import DeviceArray
from .operators import (AbstractSciMLOperator, AddedOperator, BandedMatrix, BatchedDiagonalOperator, ComposedOperator,
                        DiagonalOperator, IdentityOperator, MatrixOperator, NullOperator, ScalarOperator,
                        ScaledOperator, SparseMatrixCSR, TensorProductOperator, TensorSumOperator, adapt, cache_operator, copy, iscached, kron,
                        kronsum, mul_, update_coefficients_, InvertibleOperator, InvertedOperator, factorize, has_ldiv,
                        inv, invert_matrix, ldiv_, LUFactorization, BlockDiagonalOperator, WOperator)
from .distributed import ColumnShard, allgather_cols, shard_columns
from . import krylov

__all__ = [n for n in dir() if not n.startswith("_")]
