"""Fused whole-model inference drivers built on the hot-path kernels."""
from .vit import ViTInferenceEngine, VIT_CONFIGS  # noqa: F401
from .ultranet import UltraNetEngine  # noqa: F401
