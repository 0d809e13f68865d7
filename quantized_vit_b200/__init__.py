"""quantized_vit_b200 - B200-native (sm_100a) forward/backward of the quantized Conv2d/Linear layers of
LongAoTianxia/Quantized_ViT behind the reference's own nn.Module API.

Layout (only what the hot path needs):
  csrc/            hand-written CUDA kernels + the C ABI (include/qvit_b200.h) -> libqvit_b200.so
  _lib.py, ops.py  ctypes binding / tensor-level wrappers (PyTorch = device memory + streams only)
  quantization/    drop-in QuantizeLinear / QuantizeConv2d / quantizer Functions / model_to_quantize_model
  ultra/           drop-in quant_ultra factories + the NumPy-export integer helpers (quantization.py)
  engine/          fused whole-model inference drivers (ViT, UltraNet) built on the same kernels
"""
__version__ = "0.1.0"
