"""GPU parity of the DoReFa-style quant_ultra drop-ins (K1'/K2'/K5) against reference-generated goldens
(tests/golden/ultra.npz).  tanh on the GPU differs from the CPU's in the last ulp, so weight codes must match except
at rounding ties (<= 1 in 1000 here, and then by one level)."""
import numpy as np
import pytest
import torch

from oracle import ref_ultra
from tests.conftest import norm_close

pytestmark = pytest.mark.gpu


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def test_kat_weight_codes(golden):
    from quantized_vit_b200.ultra import quantization as qz
    g = golden("ultra")
    w = _t(g["kat.w"]).float().cuda()
    assert qz.weight_quantize_int(w, 4).cpu().tolist() == [-4, 1, -1, 3, 2, 5, -7] == g["kat.int4"].tolist()


@pytest.mark.parametrize("bit", [2, 4, 8])
def test_weight_quantize_fn(golden, bit):
    from quantized_vit_b200.ultra import weight_quantize_fn
    from quantized_vit_b200.ultra import quantization as qz
    g = golden("ultra")
    w = _t(g["w.x"]).cuda()
    with torch.no_grad():
        y = weight_quantize_fn(bit)(w).cpu().numpy()
    ref = g[f"w.bit{bit}.torch"]
    n = 2 ** (bit - 1) - 1
    codes, rcodes = np.round(y * n), np.round(ref * n)
    assert np.abs(codes - rcodes).max() <= 1 and (codes != rcodes).mean() <= 1e-3
    same = codes == rcodes
    assert np.array_equal(y[same], ref[same])                       # values bit-identical wherever the code agrees
    ci = qz.weight_quantize_int(w, bit).cpu().numpy()
    ri = g[f"w.bit{bit}.np_int"]
    assert np.abs(ci - ri).max() <= 1 and (ci != ri).mean() <= 1e-3


def test_weight_quantize_1bit_quirk_is_nan(golden):
    from quantized_vit_b200.ultra import weight_quantize_fn
    g = golden("ultra")
    with torch.no_grad():
        y = weight_quantize_fn(1)(_t(g["w.x"]).cuda())
    assert bool(torch.isnan(y).all()) and bool(np.isnan(g["w.bit1.torch"]).all())


@pytest.mark.parametrize("bit", [2, 4, 8])
def test_activation_quantize_fn(golden, bit):
    from quantized_vit_b200.ultra import activation_quantize_fn
    from quantized_vit_b200 import ops
    g = golden("ultra")
    x = _t(g["a.x"]).cuda()
    with torch.no_grad():
        y = activation_quantize_fn(bit)(x).cpu().numpy()
    assert np.array_equal(y, g[f"a.bit{bit}"])
    codes, _ = ops.ultra_act(x, bit)
    assert torch.equal(codes.cpu().to(torch.int64), ref_ultra.ultra_act_codes(_t(g["a.x"]), bit))


def test_conv_and_linear_q(golden):
    from quantized_vit_b200.ultra import conv2d_Q_fn, linear_Q_fn
    g = golden("ultra")
    conv = conv2d_Q_fn(4)(3, 16, kernel_size=3, stride=1, padding=1, bias=False)
    conv.load_state_dict({"weight": _t(g["conv.w"])})
    with torch.no_grad():
        y = conv.cuda()(_t(g["conv.x"]).cuda()).cpu().numpy()
    ok, err, scale = norm_close(y, g["conv.y"], 1e-3)
    assert ok and err <= 1e-5 * scale
    conv1 = conv2d_Q_fn(4)(16, 12, kernel_size=1, stride=1, padding=0)
    conv1.load_state_dict({"weight": _t(g["conv1.w"]), "bias": _t(g["conv1.b"])})
    with torch.no_grad():
        y = conv1.cuda()(_t(g["conv1.x"]).cuda()).cpu().numpy()
    ok, err, scale = norm_close(y, g["conv1.y"], 1e-3)
    assert ok and err <= 1e-5 * scale
    lin = linear_Q_fn(4)(40, 24)
    lin.load_state_dict({"weight": _t(g["lin.w"]), "bias": _t(g["lin.b"])})
    with torch.no_grad():
        y = lin.cuda()(_t(g["lin.x"]).cuda()).cpu().numpy()
    ok, err, scale = norm_close(y, g["lin.y"], 1e-3)
    assert ok and err <= 1e-5 * scale
    assert sorted(lin.state_dict()) == ["bias", "weight"] and type(conv).__base__ is torch.nn.Conv2d


def test_conv_q_training_path_has_ste_gradients(golden):
    from quantized_vit_b200.ultra import conv2d_Q_fn, activation_quantize_fn
    torch.manual_seed(0)
    conv = conv2d_Q_fn(2)(3, 16, kernel_size=3, padding=1).cuda()
    act = activation_quantize_fn(3)
    x = torch.rand(1, 3, 32, 32, device="cuda")
    out = torch.mean(act(conv(x)))                      # quant_ultra.py:225-242 __main__ demo
    out.backward()
    assert conv.weight.grad is not None and torch.isfinite(conv.weight.grad).all()


def test_bn_fold_and_integer_thresholds(golden):
    from quantized_vit_b200 import ops
    from quantized_vit_b200.ultra import quantization as qz
    g = golden("ultra")
    for pre in ("fold", "fold64"):
        gam, bet, mu, var = (_t(g[f"{pre}.{k}"]).cuda() for k in ("gamma", "beta", "mean", "var"))
        w, b = qz.bn_act_w_bias_float(gam.float(), bet.float(), mu.float(), var.float(), 1e-5)
        assert np.allclose(w.cpu().numpy(), g[f"{pre}.w"], rtol=1e-6) and np.allclose(b.cpu().numpy(), g[f"{pre}.b"], rtol=1e-5, atol=1e-7)
        inc, bias = qz.bn_act_quantize_int(gam, bet, mu, var, 1e-5, w_bit=4, in_bit=4, out_bit=4, l_shift=8)   # fp64 inputs
        assert np.array_equal(inc.cpu().numpy(), g[f"{pre}.inc"]) and np.array_equal(bias.cpu().numpy(), g[f"{pre}.bias"])
    assert g["fold.inc"].tolist() == [9362, 1170] and g["fold.bias"].tolist() == [-245754, -86016]
    # torch-formula fold equals nn.BatchNorm2d eval
    torch.manual_seed(1)
    bn = torch.nn.BatchNorm2d(6).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2)
    s, b = ops.bn_fold(bn.weight.cuda(), bn.bias.cuda(), bn.running_mean.cuda(), bn.running_var.cuda(), bn.eps, mode=0)
    xx = torch.randn(2, 6, 4, 4)
    assert torch.allclose(xx * s.cpu().view(1, -1, 1, 1) + b.cpu().view(1, -1, 1, 1), bn(xx), rtol=1e-5, atol=1e-6)


def test_batchnorm_q_modules_match_reference(golden):
    """QU:94-207 against outputs of the UNMODIFIED reference classes (tests/golden/ultra_bnq.npz; only torch >= 2's
    Python-level `eps <= 0` guard in front of the unchanged ATen batch_norm was bypassed when generating it)."""
    from quantized_vit_b200.ultra import batchNorm1d_Q_fn, batchNorm2d_Q_fn
    g = golden("ultra_bnq")
    for dim, fn in ((2, batchNorm2d_Q_fn), (1, batchNorm1d_Q_fn)):
        x = _t(g[f"bn{dim}d.x"])
        for bits in (2, 4, 8):
            bn = fn(bits)(6).eval()
            with torch.no_grad():
                for k in ("weight", "bias", "running_mean", "running_var"):
                    getattr(bn, k).copy_(_t(g[f"bn{dim}d.{k}"]))
            with torch.no_grad():
                got = bn.cuda()(x.cuda()).cpu().numpy()
            ref = g[f"bn{dim}d.bit{bits}.y"]
            assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), (dim, bits, np.abs(got - ref).max())
            assert sorted(bn.state_dict()) == ["bias", "num_batches_tracked", "running_mean", "running_var", "weight"]


def test_hls_parameter_layout_matches_reference(golden):
    """qvit_pack_hls_weights (w_to_hls_array) and inc_bias_to_hls_array: the FPGA parameter words of the reference exporter
    (qnn_mem_process.py:84-170), bit for bit, for every UltraNet layer shape of ultranet_param_gen.py and a ragged one."""
    from quantized_vit_b200.ultra import quantization as qz
    g = golden("ultra_hls")
    for n in sorted({k.split(".")[0] for k in g.files}):
        o, i, k, simd, pe = (int(v) for v in g[f"{n}.cfg"])
        words = qz.w_to_hls_array(torch.from_numpy(g[f"{n}.codes"]).cuda(), 4, simd, pe).cpu().numpy()
        assert np.array_equal(words.view(np.uint64), g[f"{n}.words"]), n
        hi, hb = qz.inc_bias_to_hls_array(torch.from_numpy(g[f"{n}.inc"]).cuda(), torch.from_numpy(g[f"{n}.bias"]).cuda(), pe)
        assert np.array_equal(hi.cpu().numpy(), g[f"{n}.hls_inc"]) and np.array_equal(hb.cpu().numpy(), g[f"{n}.hls_bias"])
    with pytest.raises(AssertionError, match="out_ch mod pe"):
        qz.w_to_hls_array(torch.zeros(6, 3, 3, 3, dtype=torch.int8).cuda(), 4, 3, 4)
