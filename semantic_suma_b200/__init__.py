"""semantic_suma_b200 -- a Blackwell (sm_100a) projective-ICP + surfel-fusion core behind the operator surface of
PRBonn/semantic_suma (SuMa++): Preprocessing, Frame2Model / LieGaussNewton, SurfelMap, SurfelMapping.

The product is libsuma_b200.so (hand-written CUDA + a C ABI, include/suma_b200.h); this package is the Python mirror of
the reference's classes used by the tests and the benchmark. There is no CPU fallback.
"""
from .api import (Context, Frame, Frame2Model, LieGaussNewton, Params, Preprocessing, SumaError, SurfelMap,  # noqa: F401
                  SurfelMapping, SURFEL_DTYPE, colmajor, default_params, from_colmajor, lib, library_path)
