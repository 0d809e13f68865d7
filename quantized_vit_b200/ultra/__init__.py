"""Drop-in for the reference's ``4-bit quantization`` package: ``quant_ultra`` (module factories) and
``quantization`` (the NumPy export helpers, computed on the GPU)."""
from . import quant_ultra, quantization  # noqa: F401
from .quant_ultra import (activation_quantize_fn, conv2d_Q_fn, linear_Q_fn, uniform_quantize,  # noqa: F401
                          weight_quantize_fn)
