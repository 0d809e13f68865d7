"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
the ctypes prototypes cover the header, and error paths that need no GPU behave.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qvit_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qvit_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from quantized_vit_b200 import build
    path = build.build_library()
    assert os.path.exists(path)
    return path


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("qvit_quantize_sym", "qvit_gemm_i8", "qvit_sym_backward", "qvit_im2col_quantize_sym", "qvit_bn_fold",
                 "qvit_ultra_conv_bn_act", "qvit_pack_int4", "qvit_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    handle = ctypes.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, f"declared in include/qvit_b200.h but not exported: {missing}"


def test_ctypes_prototypes_cover_the_header(built_lib):
    from quantized_vit_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    L = _lib.lib()
    assert L.qvit_abi_version() == 1


def test_epilogue_struct_layout_matches_c(built_lib):
    from quantized_vit_b200 import _lib
    # 2*int32, 2 ptr, float(+pad), 3 ptr, int64, 4 ptr  -> 8 + 16 + 8 + 24 + 8 + 32 = 96 bytes on LP64
    assert ctypes.sizeof(_lib.Epilogue) == 96
    assert _lib.Epilogue.scale_const.offset == 24 and _lib.Epilogue.ld_res.offset == 56 and _lib.Epilogue.flags.offset == 88


def test_invalid_arguments_are_rejected_without_a_gpu(built_lib):
    from quantized_vit_b200 import _lib
    L = _lib.lib()
    assert L.qvit_quantize_sym(None, 1, 1, 1, None, None, None, None, 1, None, None) == 1      # QVIT_ERR_INVALID
    assert b"null" in L.qvit_last_error()
    assert L.qvit_pack_int4(1, 3, 1, None) == 1                                                # odd n
    assert L.qvit_bn_fold(1, 1, 1, 1, 1e-5, 7, 4, 1, 1, None) == 1                              # bad mode
    assert L.qvit_gemm_i8(None, 16, 0, None, 16, 4, 4, 16, None, 4, None, 0, None) == 1         # NULL epilogue
    assert L.qvit_ultra_quantize_weight(1, 4, 9, 0, 1, 1, None) == 1                               # w_bit > 8


def test_product_refuses_cpu_tensors():
    import torch
    from quantized_vit_b200.quantization import QuantizeLinear, QuantizationMode
    m = QuantizeLinear(8, 4, quant_mode=QuantizationMode.WEIGHT_AND_ACTIVATION)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 8))


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "quantized_vit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "/root/reference" not in txt, f"{f} reads the reference tree"


def test_engines_refuse_cpu_devices():
    import torch
    from quantized_vit_b200.engine import UltraNetEngine, ViTInferenceEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ViTInferenceEngine({"cls_token": torch.zeros(1, 1, 8)}, depth=1, num_heads=1, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        UltraNetEngine({}, device="cpu")


def test_bench_reference_arm_contract_keys(monkeypatch, capsys):
    """bench.py --impl reference prints ONE JSON line with the contract keys (the CPU timing itself is stubbed here)."""
    import json
    import sys
    import bench
    monkeypatch.setattr(bench, "cpu_arm", lambda sub, repeats, warmup=1: (5.0, 3.2, 8, "port"))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
