"""B200-native analytic NNGP/NTK kernels behind the Neural Tangents API.

Drop-in for the hot path of google/neural-tangents (reference v0.6.6):
`stax` `kernel_fn(x1, x2, get)` for Dense/Conv/Relu/Erf/AvgPool/GlobalAvgPool/
Flatten/FanOut/FanInSum networks and the `nt.batch` Gram tiling.  All arithmetic
runs in hand-written sm_100a CUDA behind the C-ABI of `include/ntk_b200.h`
(`libntk_b200.so`); there is no CPU fallback.

Public names mirror `neural_tangents/__init__.py:21-33`; `predict` holds the closed-form inference that
consumes the Gram matrices (`gp_inference`, `gradient_descent_mse_ensemble`).
"""
from . import predict  # noqa: F401
from . import stax  # noqa: F401
from ._config import config  # noqa: F401
from .batching import batch  # noqa: F401
from .kernel import Kernel  # noqa: F401

__version__ = '0.1.0'
