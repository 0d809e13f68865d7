"""CPU: the oracle restatement against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py) and the de-facto known-answer vectors of SURVEY.md section 4."""
import numpy as np
import pytest
import torch

from oracle import ref_geta, ref_ultra, ref_models


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def T(a):
    return torch.from_numpy(np.asarray(a))


def _case_params(g, name):
    t = float(g[f"{name}.t"])
    return float(g[f"{name}.d"]), float(g[f"{name}.q_m"]), (None if np.isnan(t) else t)


def test_quantizer_forward_bit_exact(golden):
    g = golden("geta_quantizers")
    for name in g["cases"]:
        d, q_m, t = _case_params(g, name)
        x = T(g[f"{name}.x"])
        y = ref_geta.sym_forward(x, d, q_m, t).numpy()
        ref = g[f"{name}.y"]
        assert np.array_equal(y, ref, equal_nan=True), name
        codes = ref_geta.sym_codes(x, d, q_m, t).numpy()
        fin = np.isfinite(ref)
        dq = np.float32(abs(np.float32(d)))
        assert np.array_equal((codes.astype(np.float32) * dq)[fin], ref[fin]), name


def test_quantizer_backward_matches_reference(golden):
    g = golden("geta_quantizers")
    for name in g["cases"]:
        if f"{name}.g" not in g.files:
            continue
        d, q_m, t = _case_params(g, name)
        r = ref_geta.sym_backward(T(g[f"{name}.x"]), T(g[f"{name}.g"]), d, q_m, t)
        assert np.array_equal(r["grad_x"].numpy(), g[f"{name}.grad_x"]), name
        for k in ("grad_d", "grad_qm", "grad_t"):
            if k in r:
                assert np.array_equal(r[k].numpy(), g[f"{name}.{k}"], equal_nan=True), (name, k)


def test_kat_round_half_even():
    x = torch.tensor([.25, .75, 1.25, 1.75, -.25, -.75])
    # SURVEY section 4 KAT: values [0,1,1,2,-0,-1] with d=0.5  <=> codes [0,2,2,4,0,-2] (ties to even)
    assert ref_geta.sym_codes(x, 0.5, 100.0).tolist() == [0, 2, 2, 4, 0, -2]
    assert ref_geta.sym_forward(x, 0.5, 100.0).tolist() == [0, 1, 1, 2, 0, -1]
    # saturation level is itself rounded half-even
    assert ref_geta.saturation_code(0.1, 0.75) == 8 and ref_geta.saturation_code(0.1, 0.65) in (6, 7)
    big = torch.randn(1 << 16, generator=torch.Generator().manual_seed(3)) * 0.05
    codes = ref_geta.sym_codes(big, 0.2 / 7, 0.2)
    assert set(codes.unique().tolist()) == set(range(-7, 8))


def _layer_q(g, name, which):
    key = f"{name}.sd.d_quant_{which}"
    if key not in g.files:
        return None
    q = {"d": T(g[key]), "q_m": T(g[f"{name}.sd.q_m_{which}"])}
    if f"{name}.sd.t_quant_{which}" in g.files:
        q["t"] = T(g[f"{name}.sd.t_quant_{which}"])
    return q


def test_layers_match_reference(golden):
    g = golden("geta_layers")
    for name in g["cases"]:
        w = T(g[f"{name}.sd.weight"])
        b = T(g[f"{name}.sd.bias"]) if f"{name}.sd.bias" in g.files else None
        x = T(g[f"{name}.x"])
        wq, aq = _layer_q(g, name, "wt"), _layer_q(g, name, "act")
        if name.startswith("lin"):
            r = ref_geta.quantize_linear_forward(x, w, b, wq, aq)
            assert ref_geta.bit_width(wq["d"].item(), wq["q_m"].item(), wq.get("t", torch.ones(1)).item()).__round__() \
                == int(g[f"{name}.weight_bit"])
        else:
            cin, cout, k, s, p, dil, groups = [int(v) for v in g[f"{name}.conv"]]
            r = ref_geta.quantize_conv2d_forward(x, w, b, wq, aq, s, p, dil, groups)
        assert np.array_equal(r["y"].numpy(), g[f"{name}.y"]), name
        if "acc" in r:     # exact integer result reproduces the fp32 reference within fp32 accumulation noise
            scale = (wq["d"].abs() * aq["d"].abs()).double().item()
            y_int = r["acc"].double() * scale
            if b is not None:
                y_int = y_int + (b.double() if name.startswith("lin") else b.double().view(1, -1, 1, 1))
            ref = g[f"{name}.y"].astype(np.float64)
            assert np.abs(y_int.numpy() - ref).max() <= 1e-5 * np.abs(ref).max(), name


def test_init_quant_params(golden):
    g = golden("geta_layers")
    d, q_m = ref_geta.init_quant_params(T(g["lin_w4a4_init.sd.weight"]), 4)
    assert np.array_equal(d.numpy(), g["lin_w4a4_init.sd.d_quant_wt"])
    assert np.array_equal(q_m.numpy(), g["lin_w4a4_init.sd.q_m_wt"])
    assert np.array_equal(d.numpy(), g["lin_w4a4_init.sd.d_quant_act"])      # QL:436-437


def test_ultra_quantizers(golden):
    g = golden("ultra")
    assert ref_ultra.np_weight_quantize_int(g["kat.w"], 4).tolist() == [-4, 1, -1, 3, 2, 5, -7]
    assert np.array_equal(ref_ultra.np_weight_quantize_int(g["kat.w"], 4), g["kat.int4"])
    assert np.array_equal(ref_ultra.np_weight_quantize_float(g["kat.w"], 4), g["kat.float4"])
    w = T(g["w.x"])
    for b in (2, 4, 8):
        assert np.array_equal(ref_ultra.ultra_weight_values(w, b).numpy(), g[f"w.bit{b}.torch"])
        codes = ref_ultra.ultra_weight_codes(w, b).numpy()
        n = np.float32(2 ** (b - 1) - 1)
        assert np.array_equal(codes.astype(np.float32) / n, g[f"w.bit{b}.torch"])
        # the torch (fp32) and NumPy (fp64) reference paths agree on these weights
        if b != 2:          # at 2 bits the torch path is sign() (QU:15-16) and differs from QZ:24-31 by design
            assert np.array_equal(codes, g[f"w.bit{b}.np_int"])
        assert np.array_equal(ref_ultra.np_weight_quantize_int(g["w.x"].astype(np.float64), b), g[f"w.bit{b}.np_int"])
    assert np.array_equal(ref_ultra.ultra_weight_values(w, 1).numpy(), g["w.bit1.torch"], equal_nan=True)  # all-NaN quirk
    a = T(g["a.x"])
    for b in (2, 4, 8):
        assert np.array_equal(ref_ultra.ultra_act_values(a, b).numpy(), g[f"a.bit{b}"])
        n = np.float32(2 ** b - 1)
        assert np.array_equal(ref_ultra.ultra_act_codes(a, b).numpy().astype(np.float32) / n, g[f"a.bit{b}"])


def test_ultra_layers(golden):
    g = golden("ultra")
    y = ref_ultra.conv2d_q_forward(T(g["conv.x"]), T(g["conv.w"]), None, 4, 1, 1)
    assert np.array_equal(y.numpy(), g["conv.y"])
    y = ref_ultra.conv2d_q_forward(T(g["conv1.x"]), T(g["conv1.w"]), T(g["conv1.b"]), 4, 1, 0)
    assert np.array_equal(y.numpy(), g["conv1.y"])
    y = ref_ultra.linear_q_forward(T(g["lin.x"]), T(g["lin.w"]), T(g["lin.b"]), 4)
    assert np.array_equal(y.numpy(), g["lin.y"])


def test_bn_fold_and_pack_kats(golden):
    g = golden("ultra")
    w, b = ref_ultra.np_bn_fold(g["fold.gamma"], g["fold.beta"], g["fold.mean"], g["fold.var"], 1e-5)
    assert np.array_equal(w, g["fold.w"]) and np.array_equal(b, g["fold.b"])
    np.testing.assert_allclose(w, [1.99996, 0.24999875], rtol=1e-6)
    inc, bias = ref_ultra.np_bn_act_quantize_int(g["fold.gamma"], g["fold.beta"], g["fold.mean"], g["fold.var"],
                                                 1e-5, 4, 4, 4, 8)
    assert inc.tolist() == [9362, 1170] and bias.tolist() == [-245754, -86016]
    inc, bias = ref_ultra.np_bn_act_quantize_int(g["fold64.gamma"], g["fold64.beta"], g["fold64.mean"],
                                                 g["fold64.var"], 1e-5, 4, 4, 4, 8)
    assert np.array_equal(inc, g["fold64.inc"]) and np.array_equal(bias, g["fold64.bias"])
    assert ref_ultra.pack_words([1, -1, 7, -7], 4) == 0x97f1 == int(g["pack.word"])
    for row, word in zip(g["pack16.codes"], g["pack16.words"]):
        assert ref_ultra.pack_words(row, 4) == int(word)
        by = ref_ultra.pack_int4_bytes(row)
        assert int.from_bytes(by.tobytes(), "little") == int(word)
        assert np.array_equal(ref_ultra.unpack_int4_bytes(by), row)


def test_hls_parameter_layout(golden):
    """oracle.hls_weight_words / hls_inc_bias against QNNLayerMemProcess.w_to_hls_array / inc_bias_to_hls_array run on
    the reference itself (tests/golden/ultra_hls.npz, oracle/make_golden.py::golden_ultra_hls)."""
    g = golden("ultra_hls")
    names = sorted({k.split(".")[0] for k in g.files})
    assert "ragged" in names and len(names) == 6
    for n in names:
        o, i, k, simd, pe = (int(v) for v in g[f"{n}.cfg"])
        words = ref_ultra.hls_weight_words(g[f"{n}.codes"], 4, simd, pe)
        assert words.shape == g[f"{n}.words"].shape and words.shape[1] == int(g[f"{n}.w_tiles"])
        assert np.array_equal(words, g[f"{n}.words"]), n
        hi, hb = ref_ultra.hls_inc_bias(g[f"{n}.inc"], g[f"{n}.bias"], pe)
        assert np.array_equal(hi, g[f"{n}.hls_inc"]) and np.array_equal(hb, g[f"{n}.hls_bias"])


def test_ultranet_whole_model(golden):
    g = golden("ultranet")
    from tests.fixtures import ultranet_state_dict, ultranet_input
    sd = ultranet_state_dict()
    x = ultranet_input()
    assert abs(x.double().sum().item() - float(g["x_sum"])) < 1e-9
    taps = []
    feats = ref_models.ultranet_features(sd, x, taps=taps)
    assert np.array_equal(feats.numpy(), g["feats"])
    for i in range(8):
        assert np.array_equal(torch.round(taps[i] * 15).numpy().astype(np.uint8), g[f"tap{i}.codes"])
    io = ref_models.yolo_decode(feats, (160, 320))
    np.testing.assert_allclose(io.numpy(), g["io"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["vit_tiny_w4a4_init", "vit_tiny_w4a4_calib", "vit_tiny_w4a8_calib",
                                  "vit_tiny_nl_w8a8_calib"])
def test_vit_tiny_whole_model(golden, name):
    g = golden(name)
    sd = {k[3:]: T(g[k]) for k in g.files if k.startswith("sd.")}
    img, patch, dim, depth, heads, classes = [int(v) for v in g["cfg"]]
    logits = ref_models.vit_forward(sd, T(g["x"]), depth, heads, patch)
    assert np.array_equal(logits.numpy(), g["logits"])


@pytest.mark.parametrize("name", ["vit_b16_w4a4_init", "vit_b16_w4a4_calib", "vit_b16_w8a8_calib", "vit_l16_w4a8_init",
                                  "vit_l16_w4a8_calib"])
def test_vit_b16_whole_model(golden, name):
    """ViT-B/16 at batch 2: weights come from the seeded fill, quantizer params from the fixture
    (they were produced by the reference's own initialize_quant_layer / calibration)."""
    g = golden(name)
    from tests.fixtures import vit_state_dict, vit_input
    img, patch, dim, depth, heads, classes = [int(v) for v in g["cfg"]]
    sd = vit_state_dict(img, patch, dim, depth, heads, classes, seed=int(g["fill_seed"]))
    qref = dict(zip(g["q.names"].tolist(), g["q.values"].tolist()))
    if name == "vit_b16_w4a4_init":   # our own restatement of initialize_quant_layer must give the same params
        for k, v in qref.items():
            if k.endswith("d_quant_wt"):
                d, q_m = ref_geta.init_quant_params(sd[k.replace("d_quant_wt", "weight")], 4)
                assert d.item() == np.float32(v) and q_m.item() == np.float32(qref[k.replace("d_quant", "q_m")])
    for k, v in qref.items():
        sd[k] = torch.tensor([v], dtype=torch.float32)
    x = vit_input(int(g["batch"]), img)
    assert abs(x.double().sum().item() - float(g["x_sum"])) < 1e-6
    logits = ref_models.vit_forward(sd, x, depth, heads, patch)
    assert np.array_equal(logits.numpy(), g["logits"])


def test_batchnorm_q_pinned_to_reference(golden):
    """QU:94-207: outputs of the UNMODIFIED reference classes with only torch >= 2's Python-level
    `eps <= 0` guard bypassed (oracle/make_golden.py::golden_bnq) - bit for bit."""
    g = golden("ultra_bnq")
    for dim, fwd in ((2, ref_ultra.batchnorm2d_q_forward), (1, ref_ultra.batchnorm1d_q_forward)):
        gam, bet, mu, var = (_t(g[f"bn{dim}d.{k}"]) for k in ("weight", "bias", "running_mean", "running_var"))
        x = _t(g[f"bn{dim}d.x"])
        for bits in (2, 4, 8):
            assert np.array_equal(fwd(x, gam, bet, mu, var, 1e-5, bits).numpy(), g[f"bn{dim}d.bit{bits}.y"]), (dim, bits)


@pytest.mark.parametrize("tag", ["lin", "nl"])
def test_qat_autograd_matches_reference(golden, tag):
    """Config 3 parity of the ORACLE: ref_models.vit_forward_autograd (quantizers differentiated by ref_geta.SymQuantFn)
    against the reference's own autograd on the depth-2, D = 768 ViT (tests/golden/qat_vit_d768_*.npz): loss, logits, every
    quantizer-scalar gradient and the digest of every other gradient."""
    import os
    from tests import fixtures
    g = golden(f"qat_vit_d768_{tag}")
    img, patch, dim, depth, heads, classes = [int(v) for v in g["cfg"]]
    torch.set_num_threads(os.cpu_count() or 1)
    sd = fixtures.vit_state_dict(img, patch, dim, depth, heads, classes, seed=int(g["fill_seed"]))
    for k, v in zip(g["q.names"], g["q.values"]):
        sd[str(k)] = torch.tensor([float(v)], dtype=torch.float32)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = fixtures.vit_input(int(g["batch"]), img, seed=1)
    assert abs(float(x.double().sum()) - float(g["x_sum"])) < 1e-6
    logits = ref_models.vit_forward_autograd(params, x, depth, heads, patch)
    loss = torch.nn.functional.cross_entropy(logits, _t(g["labels"]))
    loss.backward()
    assert np.array_equal(logits.detach().numpy(), g["logits"])
    assert abs(loss.item() - float(g["loss"])) <= 1e-6
    worst = 0.0
    for n in [str(s) for s in g["g.names"]]:
        st, smp = ref_models.grad_digest(params[n].grad)
        ref_st, ref_smp = g[f"g.{n}.stats"], g[f"g.{n}.samples"]
        tol = 1e-5 * max(abs(ref_st[1]), 1e-30)
        assert abs(st[0] - ref_st[0]) <= tol and abs(st[1] - ref_st[1]) <= tol and abs(st[2] - ref_st[2]) <= 1e-5 * ref_st[2] + 1e-12, n
        err = np.abs(smp - ref_smp).max() / max(np.abs(ref_smp).max(), 1e-30)
        worst = max(worst, err)
        assert err <= 1e-5, (n, err)
    print(f"qat {tag}: {len(g['g.names'])} gradients, worst sampled deviation {worst:.2e}")
