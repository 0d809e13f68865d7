"""Import the UNMODIFIED reference modules from /root/reference (this container only).

TEST INFRASTRUCTURE - never imported by the product package.  Only `oracle/make_golden.py`
and the `-m "not gpu"` cross-check tests use it, and only when /root/reference exists
(it does not exist on the GPU box).

The reference's top-level ``only_train_once/__init__.py`` imports its graph tracer, which needs
``torch.onnx._globals`` (gone in torch 2.11).  The hot-path files do not, so a stub package with
the right ``__path__`` is registered and the sub-packages are imported normally.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("QVIT_REFERENCE_ROOT", "/root/reference")
_GETA_DIR = os.path.join(REFERENCE_ROOT, "QViT_with_GETA")
_ULTRA_DIR = os.path.join(REFERENCE_ROOT, "4-bit quantization")


def use_root(root: str) -> None:
    """Point the loader at another copy of the reference tree (oracle/_ref on the GPU box, see oracle/make_ref.py)."""
    global REFERENCE_ROOT, _GETA_DIR, _ULTRA_DIR
    REFERENCE_ROOT = root
    _GETA_DIR = os.path.join(root, "QViT_with_GETA")
    _ULTRA_DIR = os.path.join(root, "4-bit quantization")


def available() -> bool:
    return os.path.isdir(os.path.join(_GETA_DIR, "only_train_once", "quantization"))


def _stub_only_train_once():
    if "only_train_once" in sys.modules and getattr(sys.modules["only_train_once"], "__qvit_stub__", False):
        return
    sys.dont_write_bytecode = True  # the reference tree is read-only
    pkg = types.ModuleType("only_train_once")
    pkg.__path__ = [os.path.join(_GETA_DIR, "only_train_once")]
    pkg.__qvit_stub__ = True
    sys.modules["only_train_once"] = pkg


def quant_layers():
    """reference `only_train_once/quantization/quant_layers.py` as a module."""
    _stub_only_train_once()
    return importlib.import_module("only_train_once.quantization.quant_layers")


def quant_model():
    _stub_only_train_once()
    return importlib.import_module("only_train_once.quantization.quant_model")


def vit_model():
    sys.dont_write_bytecode = True
    if _GETA_DIR not in sys.path:
        sys.path.insert(0, _GETA_DIR)
    return importlib.import_module("vit_model")


def _ultra_path():
    sys.dont_write_bytecode = True
    if _ULTRA_DIR not in sys.path:
        sys.path.insert(0, _ULTRA_DIR)


def quant_ultra():
    _ultra_path()
    return importlib.import_module("quant_ultra")


def quantization_np():
    _ultra_path()
    return importlib.import_module("quantization")


def mymodel():
    _ultra_path()
    return importlib.import_module("mymodel")


def qnn_mem_process():
    _ultra_path()
    return importlib.import_module("qnn_mem_process")
