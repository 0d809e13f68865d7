"""Drop-in for the reference's ``4-bit quantization`` package: ``quant_ultra`` (module factories) and
``quantization`` (the NumPy export helpers, computed on the GPU)."""
from . import quant_ultra, quantization  # noqa: F401
from .quant_ultra import (activation_quantize_fn, batchNorm1d_Q_fn, batchNorm2d_Q_fn, conv2d_Q_fn,  # noqa: F401
                          linear_Q_fn, uniform_quantize, weight_quantize_fn)
