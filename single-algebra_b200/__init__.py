"""single-algebra_b200 — B200-native (sm_100a) implementation of the sparse-PCA hot path of
SingleRust/single-algebra behind the reference's own type surface.  See DESIGN.md."""
from . import _native
from ._native import SalgError
from .api import (Context, CscMatrix, CsrMatrix, DeviceCsr, Direction, MaskedSparsePCA, MaskedSparsePCABuilder,
                  PowerIterationNormalizer, SVDMethod, SparsePCA, SparsePCABuilder, default_context,
                  device_count, op_cholqr2, op_small_svd, op_spmm, op_spmm_bench, op_tall_gram, set_default_context,
                  synth_device)
from . import dist, synth

TRANSFORM_EXACT = _native.TRANSFORM_EXACT
TRANSFORM_REFERENCE_COMPAT = _native.TRANSFORM_REFERENCE_COMPAT

__all__ = [
    "CscMatrix",
    "Context", "CsrMatrix", "DeviceCsr", "Direction", "MaskedSparsePCA", "MaskedSparsePCABuilder",
    "PowerIterationNormalizer", "SVDMethod", "SparsePCA", "SparsePCABuilder", "SalgError", "default_context",
    "device_count", "set_default_context", "synth_device", "op_spmm", "op_cholqr2", "op_small_svd",
    "op_spmm_bench", "op_tall_gram", "dist", "synth", "TRANSFORM_EXACT", "TRANSFORM_REFERENCE_COMPAT",
]
