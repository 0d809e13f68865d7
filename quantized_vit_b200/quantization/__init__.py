"""Mirror of ``only_train_once/quantization/__init__.py:1-2`` (re-exports the layer and model helpers)."""
from .quant_layers import *  # noqa: F401,F403
from .quant_layers import (LAYER_TO_QUANTLAYER, DGEQuantizer, NanInGradientError, QuantizationMode, QuantizationType,
                           QuantizeConv2d, QuantizeLinear, QuantizeMixin, SymQuantizerLinear, SymQuantizerNonLinear,
                           _get_quantizer, check_nan_flags, initialize_quant_layer)  # noqa: F401
from .quant_model import get_bitwidth_dict, get_quant_param_dict, model_to_quantize_model  # noqa: F401
from .geta_step import GetaQuantParamStepper  # noqa: F401
