"""Whole-model GPU parity: the fused engines against logits / feature maps produced by the UNMODIFIED reference
models (tests/golden/vit_*.npz, ultranet.npz - see oracle/make_golden.py).

Tolerance (BASELINE.json north_star): final fp32 logits within 1e-3 relative (norm-wise: max|delta| <= 1e-3 * max|ref|)
with identical top-1.  Activation codes inside the network are compared layer by layer for UltraNet; they may
differ from the reference only where the reference's own fp32 conv output sits within fp32 rounding noise of a
rounding boundary (the reference accumulates code/15 * code/7 products in fp32; we accumulate integers exactly)."""
import numpy as np
import pytest
import torch

from tests import fixtures
from tests.conftest import norm_close

pytestmark = pytest.mark.gpu


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def _vit_case(golden, name):
    g = golden(name)
    img, patch, dim, depth, heads, classes = [int(v) for v in g["cfg"]]
    if "x" in g.files:
        sd = {k[3:]: _t(g[k]) for k in g.files if k.startswith("sd.")}
        x = _t(g["x"])
    else:
        sd = fixtures.vit_state_dict(img, patch, dim, depth, heads, classes, seed=int(g["fill_seed"]))
        for k, v in zip(g["q.names"], g["q.values"]):
            sd[str(k)] = torch.tensor([float(v)], dtype=torch.float32)
        x = fixtures.vit_input(int(g["batch"]), img, seed=1)
        assert abs(float(x.double().sum()) - float(g["x_sum"])) < 1e-6
    return g, sd, x, dict(depth=depth, num_heads=heads, patch_size=patch)


VIT_CASES = ["vit_tiny_w4a4_init", "vit_tiny_w4a4_calib", "vit_tiny_w4a8_calib", "vit_tiny_nl_w8a8_calib",
             "vit_b16_w4a4_init", "vit_b16_w4a4_calib"]


@pytest.mark.parametrize("name", VIT_CASES)
def test_vit_engine_logits_match_reference(golden, name):
    from quantized_vit_b200.engine import ViTInferenceEngine
    g, sd, x, cfg = _vit_case(golden, name)
    eng = ViTInferenceEngine(sd, precision="fp32", **cfg)
    logits = eng(x.cuda()).cpu().numpy()
    ref = g["logits"]
    ok, err, scale = norm_close(logits, ref, 1e-3)
    print(f"{name}: max|delta|/max|ref| = {err / scale:.3e}, min top-1 margin {float(g['margin'].min()):.3e}")
    assert ok, f"{name}: {err / scale:.3e} > 1e-3"
    assert np.array_equal(logits.argmax(-1), g["top1"])
    assert int(eng.flags.item()) == 0


def test_vit_engine_cuda_graph_replay_is_identical(golden):
    from quantized_vit_b200.engine import ViTInferenceEngine
    g, sd, x, cfg = _vit_case(golden, "vit_tiny_w4a4_calib")
    eng = ViTInferenceEngine(sd, **cfg)
    eager = eng(x.cuda()).clone()
    xs, ys, graph = eng.capture(x.shape[0], x.shape[-1])
    xs.copy_(x.cuda())
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(ys, eager)


def test_vit_dropin_modules_equal_engine(golden):
    """The same network assembled from the drop-in QuantizeLinear/QuantizeConv2d modules (module-by-module path, fp32 in
    and out of every layer) agrees with the reference logits as well."""
    from quantized_vit_b200.quantization import QuantizeConv2d, QuantizeLinear, QuantizationMode, QuantizationType
    import torch.nn.functional as F
    g, sd, x, cfg = _vit_case(golden, "vit_tiny_w4a4_calib")
    depth, heads, patch = cfg["depth"], cfg["num_heads"], cfg["patch_size"]
    WA, LIN = QuantizationMode.WEIGHT_AND_ACTIVATION, QuantizationType.SYMMETRIC_LINEAR

    def make(prefix):
        w = sd[f"{prefix}.weight"]
        if w.dim() == 4:
            m = QuantizeConv2d(w.shape[1], w.shape[0], patch, patch, 0, bias=True, quant_type=LIN, quant_mode=WA)
        else:
            m = QuantizeLinear(w.shape[1], w.shape[0], bias=True, quant_type=LIN, quant_mode=WA)
        m.load_state_dict({k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + ".")})
        return m.cuda().eval()
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        h = make("patch_embed.proj")(x.cuda()).flatten(2).transpose(1, 2)
        B, D = h.shape[0], h.shape[-1]
        h = torch.cat((sdc["cls_token"].expand(B, -1, -1), h), 1) + sdc["pos_embed"]
        for i in range(depth):
            p = f"blocks.{i}"
            y = F.layer_norm(h, (D,), sdc[f"{p}.norm1.weight"], sdc[f"{p}.norm1.bias"], 1e-6)
            qkv = make(f"{p}.attn.qkv")(y).reshape(B, -1, 3, heads, D // heads).permute(2, 0, 3, 1, 4)
            a = ((qkv[0] @ qkv[1].transpose(-2, -1)) * (D // heads) ** -0.5).softmax(-1)
            y = (a @ qkv[2]).transpose(1, 2).reshape(B, -1, D)
            h = h + make(f"{p}.attn.proj")(y)
            y = F.layer_norm(h, (D,), sdc[f"{p}.norm2.weight"], sdc[f"{p}.norm2.bias"], 1e-6)
            h = h + make(f"{p}.mlp.fc2")(F.gelu(make(f"{p}.mlp.fc1")(y)))
        h = F.layer_norm(h, (D,), sdc["norm.weight"], sdc["norm.bias"], 1e-6)
        logits = make("head")(h[:, 0]).cpu().numpy()
    ok, err, scale = norm_close(logits, g["logits"], 1e-3)
    assert ok and np.array_equal(logits.argmax(-1), g["top1"]), f"{err / scale:.3e}"


def test_ultranet_engine_matches_reference(golden):
    from quantized_vit_b200.engine import UltraNetEngine
    g = golden("ultranet")
    sd = fixtures.ultranet_state_dict()
    x = fixtures.ultranet_input(1, seed=int(g["x_seed"]))
    assert abs(float(x.double().sum()) - float(g["x_sum"])) < 1e-6
    for input_bits in (8, None):
        eng = UltraNetEngine(sd, input_bits=input_bits)
        taps = []
        feats = eng(x.cuda(), taps=taps).cpu().numpy()
        total = mism = 0
        for i, t in enumerate(taps):
            got = t.cpu().numpy().transpose(0, 3, 1, 2)           # NHWC -> NCHW
            ref = g[f"tap{i}.codes"]
            assert got.shape == ref.shape
            assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1, f"tap {i}: codes differ by more than one level"
            mism += int((got != ref).sum())
            total += ref.size
        print(f"input_bits={input_bits}: {mism} of {total} activation codes differ from the reference by one level")
        assert mism <= 1e-4 * total
        ok, err, scale = norm_close(feats, g["feats"], 1e-3 if mism == 0 else 5e-2)
        assert ok, f"feats: {err / scale:.3e}"


def test_ultranet_layers_teacher_forced_bit_exact(golden):
    """Every fused integer layer fed with the REFERENCE's previous-layer codes reproduces the reference's codes exactly,
    except where the exact rational pre-activation lies within 2e-6 of a rounding boundary."""
    from quantized_vit_b200 import ops
    from quantized_vit_b200.engine import UltraNetEngine
    g = golden("ultranet")
    eng = UltraNetEngine(fixtures.ultranet_state_dict(), input_bits=8)
    for i in range(1, 8):
        L = eng.layers[i]
        prev = _t(g[f"tap{i - 1}.codes"]).permute(0, 2, 3, 1).contiguous().cuda()
        got = ops.ultra_conv_bn_act(prev, L["codes_ohwi"], L["pad"], 1.0 / (15 * 7), L["scale"], L["bias"], 15, L["pool"])
        got = got.cpu().numpy().transpose(0, 3, 1, 2)
        ref = g[f"tap{i}.codes"]
        nbad = int((got != ref).sum())
        assert nbad <= 2, f"layer {i}: {nbad} codes differ"
        assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1


def test_ultranet_cuda_graph(golden):
    from quantized_vit_b200.engine import UltraNetEngine
    eng = UltraNetEngine(fixtures.ultranet_state_dict())
    x = fixtures.ultranet_input(1).cuda()
    eager = eng(x).clone()
    xs, ys, graph = eng.capture(1)
    xs.copy_(x)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(ys, eager)
