"""fnm_b200: B200-native (sm_100a) Fourier-layer hot path behind the reference's nn.Module API.

    from fnm_b200.models import FNO2d            # drop-in for models.FNO2d of the reference
    from fnm_b200 import Adam                    # drop-in for util.Adam.Adam

The arithmetic runs in libfnm_b200.so (hand-written CUDA, C ABI in include/fnm_b200.h).
"""
from . import _lib, models, ops, parallel, plan, shared
from ._lib import FnmError, build, version
from .adam import Adam
from .models import FND1d, FND2d, FNF1d, FNF2d, FNN1d, FNN2d, FNO1d, FNO1d2, FNO2d, FNO2d1
from .ops import get_compute_mode, set_compute_mode
from .shared import (MLP, LinearDecoder1d, LinearDecoder2d, LinearFunctionals1d, LinearFunctionals2d,
                     SpectralConv1d, SpectralConv2d)

__all__ = ["Adam", "FnmError", "build", "version", "models", "ops", "parallel", "plan", "shared",
           "FNO1d", "FNO2d", "FNF1d", "FNF2d", "FND1d", "FND2d", "FNN1d", "FNN2d", "FNO1d2", "FNO2d1", "MLP",
           "SpectralConv1d", "SpectralConv2d", "LinearFunctionals1d", "LinearFunctionals2d",
           "LinearDecoder1d", "LinearDecoder2d", "set_compute_mode", "get_compute_mode"]
