"""Generate tests/golden/*.npz by IMPORTING AND RUNNING THE UNMODIFIED REFERENCE (build container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

TEST INFRASTRUCTURE.  The fixtures pin the oracle (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_*.py) to outputs of the reference itself; the GPU box has no /root/reference, so the
vectors are committed together with this script.  Everything is seeded; re-running reproduces the
files bit-for-bit on the same torch build (torch 2.11.0+cu128, CPU, fp32).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import _refload as R  # noqa: E402
from oracle import ref_models  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _np(t):
    return t.detach().cpu().numpy()


def edge_inputs(d: float, q_m: float) -> torch.Tensor:
    """Hand-picked edge cases (SURVEY.md appendix B truth table): zeros, denormal, exact ties,
    |x| == q_m, beyond range, infinities."""
    vals = [0.0, -0.0, 1e-45, -1e-45, 0.5 * d, -0.5 * d, 1.5 * d, 2.5 * d, -2.5 * d, 3.5 * d,
            q_m, -q_m, q_m * (1 - 1e-7), q_m * (1 + 1e-7), 2 * q_m, -3 * q_m, float("inf"), float("-inf"),
            6.5 * d, -6.5 * d, 7.5 * d, 0.49999997 * d, 0.50000006 * d]
    return torch.tensor(vals, dtype=torch.float32)


def golden_quantizers():
    ql = R.quant_layers()
    out = {}
    cases = []
    # (name, x, d, q_m, t)
    g = _gen(100)
    x_w = torch.randn(96, 80, generator=g) * 0.02
    qm = x_w.abs().max().item()
    cases.append(("w4", x_w, qm / 7, qm, None))
    cases.append(("w8", x_w, qm / 127, qm, None))
    x_a = torch.randn(4, 33, 80, generator=_gen(101))
    cases.append(("a4", x_a, 2.5 / 7, 2.5, None))
    cases.append(("a8", x_a, 3.0 / 127, 3.0, None))
    cases.append(("a4_learned", x_a, 0.3791, 2.2113, None))          # q_m/d not an integer
    cases.append(("edge4", edge_inputs(0.1, 0.7), 0.1, 0.7, None))
    cases.append(("edge_negqm", edge_inputs(0.1, 0.7), 0.1, -0.7, None))
    cases.append(("edge_negd", edge_inputs(0.1, 0.7), -0.1, 0.7, None))
    cases.append(("nl_t1", x_a, 2.5 / 7, 2.5, 1.0))
    cases.append(("nl_t07", x_a, 0.25, 2.5, 0.7))
    cases.append(("nl_t13", x_a.abs() * 0.5 + 0.01, 0.21, 1.9, 1.3))
    cases.append(("nl_w", x_w, (qm + 1e-6) / 7, qm, 1.0))
    clip = torch.tensor((-2.0, 2.0))
    q_s = torch.tensor(0.0)
    names = []
    for name, x, d, q_m, t in cases:
        names.append(name)
        xd = x.clone().requires_grad_(True)
        dp = torch.nn.Parameter(torch.tensor([d], dtype=torch.float32))
        qp = torch.nn.Parameter(torch.tensor([q_m], dtype=torch.float32))
        gout = torch.randn(x.shape, generator=_gen(200 + len(names)))
        if t is None:
            y = ql.SymQuantizerLinear.apply(xd, dp, qp, clip, q_s)
            tp = None
        else:
            tp = torch.nn.Parameter(torch.tensor([t], dtype=torch.float32))
            y = ql.SymQuantizerNonLinear.apply(xd, dp, qp, tp, clip, q_s)
        finite = bool(torch.isfinite(x).all())
        out[f"{name}.x"] = _np(x)
        out[f"{name}.d"] = np.float32(d)
        out[f"{name}.q_m"] = np.float32(q_m)
        out[f"{name}.t"] = np.float32(np.nan if t is None else t)
        out[f"{name}.y"] = _np(y)
        if finite:
            y.backward(gout)
            out[f"{name}.g"] = _np(gout)
            out[f"{name}.grad_x"] = _np(xd.grad)
            out[f"{name}.grad_d"] = _np(dp.grad)
            out[f"{name}.grad_qm"] = _np(qp.grad)
            if tp is not None:
                out[f"{name}.grad_t"] = _np(tp.grad)
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "geta_quantizers.npz"), **out)


def golden_layers():
    ql = R.quant_layers()
    out = {}
    names = []

    def run_linear(name, in_f, out_f, bias, bits, act_bits, qtype, mode, x, seed):
        torch.manual_seed(seed)
        lin = torch.nn.Linear(in_f, out_f, bias=bias)
        m = ql.QuantizeLinear.from_module(lin, quant_type=qtype, quant_mode=mode, num_bits=bits)
        if act_bits is not None and mode == ql.QuantizationMode.WEIGHT_AND_ACTIVATION:
            amax = x.abs().max().item() * 0.6
            with torch.no_grad():
                m.q_m_act.fill_(amax)
                m.d_quant_act.fill_(amax / (2 ** (act_bits - 1) - 1))
        y = m(x)
        names.append(name)
        out[f"{name}.x"] = _np(x)
        out[f"{name}.y"] = _np(y)
        for k, v in m.state_dict().items():
            out[f"{name}.sd.{k}"] = _np(v)
        out[f"{name}.weight_bit"] = np.int64(m.weight_bit)
        out[f"{name}.activation_bit"] = np.int64(m.activation_bit)

    L, NL = ql.QuantizationType.SYMMETRIC_LINEAR, ql.QuantizationType.SYMMETRIC_NONLINEAR
    WO, WA = ql.QuantizationMode.WEIGHT_ONLY, ql.QuantizationMode.WEIGHT_AND_ACTIVATION
    x1 = torch.randn(3, 17, 64, generator=_gen(300))
    run_linear("lin_w4a4", 64, 48, True, 4, 4, L, WA, x1, 1)
    run_linear("lin_w4a8", 64, 48, True, 4, 8, L, WA, x1, 2)
    run_linear("lin_w4a4_init", 64, 48, True, 4, None, L, WA, x1 * 0.05, 3)      # act params = weight params (QL:436)
    run_linear("lin_w4_wo", 64, 48, False, 4, None, L, WO, x1, 4)
    run_linear("lin_w18_wo", 256, 128, True, 18, None, L, WO, torch.randn(1, 256, generator=_gen(301)), 5)
    run_linear("lin_nl_w8a8", 64, 48, True, 8, 8, NL, WA, x1, 6)
    run_linear("lin_odd_w4a4", 50, 37, True, 4, 4, L, WA, torch.randn(5, 50, generator=_gen(302)), 7)

    def run_conv(name, cin, cout, k, s, p, dil, groups, bias, bits, act_bits, qtype, mode, x, seed):
        torch.manual_seed(seed)
        conv = torch.nn.Conv2d(cin, cout, k, s, p, dil, groups, bias=bias)
        m = ql.QuantizeConv2d.from_module(conv, quant_type=qtype, quant_mode=mode, num_bits=bits)
        if act_bits is not None and mode == WA:
            amax = x.abs().max().item() * 0.6
            with torch.no_grad():
                m.q_m_act.fill_(amax)
                m.d_quant_act.fill_(amax / (2 ** (act_bits - 1) - 1))
        y = m(x)
        names.append(name)
        out[f"{name}.x"] = _np(x)
        out[f"{name}.y"] = _np(y)
        out[f"{name}.conv"] = np.array([cin, cout, k, s, p, dil, groups], dtype=np.int64)
        for kk, v in m.state_dict().items():
            out[f"{name}.sd.{kk}"] = _np(v)

    xc = torch.randn(2, 3, 32, 32, generator=_gen(310))
    run_conv("conv_patch_w4a4", 3, 32, 8, 8, 0, 1, 1, True, 4, 4, L, WA, xc, 11)
    run_conv("conv_3x3_w4a4", 3, 8, 3, 2, 1, 1, 1, False, 4, 4, L, WA, xc, 12)
    run_conv("conv_3x3_w4a8_dil", 4, 8, 3, 1, 2, 2, 1, True, 4, 8, L, WA, torch.randn(1, 4, 13, 11, generator=_gen(311)), 13)
    run_conv("conv_groups_wo", 4, 8, 3, 1, 1, 1, 2, True, 8, None, L, WO, torch.randn(1, 4, 9, 9, generator=_gen(312)), 14)
    # channel counts the implicit-GEMM tensor-core conv serves (C in {16, 32, 64, 128}): plain, strided, dilated
    run_conv("conv_c32_3x3_w4a4", 32, 48, 3, 1, 1, 1, 1, True, 4, 4, L, WA, torch.randn(2, 32, 19, 23, generator=_gen(313)), 15)
    run_conv("conv_c16_s2_w4a8", 16, 40, 3, 2, 1, 1, 1, False, 4, 8, L, WA, torch.randn(3, 16, 17, 12, generator=_gen(314)), 16)
    run_conv("conv_c64_dil2_w8a8", 64, 64, 3, 1, 2, 2, 1, True, 8, 8, L, WA, torch.randn(1, 64, 14, 14, generator=_gen(315)), 17)
    run_conv("conv_c16_1x1_w4a4", 16, 36, 1, 1, 0, 1, 1, True, 4, 4, L, WA, torch.randn(2, 16, 10, 20, generator=_gen(316)), 18)
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "geta_layers.npz"), **out)


def golden_ultra():
    qu, qz, mp = R.quant_ultra(), R.quantization_np(), R.qnn_mem_process()
    out = {}
    kat = np.array([-0.6, 0.1, -0.2, 0.5, 0.3, 0.8, -3.9])
    out["kat.w"] = kat
    out["kat.int4"] = qz.weight_quantize_int(kat, bit=4)
    out["kat.float4"] = qz.weight_quantize_float(kat, bit=4)
    w = torch.randn(16, 3, 3, 3, generator=_gen(400)) * 0.5
    for b in (2, 4, 8):
        out[f"w.bit{b}.torch"] = _np(qu.weight_quantize_fn(b)(w))
        out[f"w.bit{b}.np_int"] = qz.weight_quantize_int(_np(w).astype(np.float64), bit=b)
    out["w.bit1.torch"] = _np(qu.weight_quantize_fn(1)(w))
    out["w.x"] = _np(w)
    a = torch.randn(2, 5, 7, 9, generator=_gen(401)) * 0.7 + 0.3
    out["a.x"] = _np(a)
    for b in (2, 4, 8):
        out[f"a.bit{b}"] = _np(qu.activation_quantize_fn(b)(a))
    # Conv2d_Q / Linear_Q
    torch.manual_seed(5)
    conv = qu.conv2d_Q_fn(4)(3, 16, kernel_size=3, stride=1, padding=1, bias=False)
    xin = torch.rand(2, 3, 20, 24, generator=_gen(402))
    out["conv.w"] = _np(conv.weight)
    out["conv.x"] = _np(xin)
    out["conv.y"] = _np(conv(xin))
    conv1 = qu.conv2d_Q_fn(4)(16, 12, kernel_size=1, stride=1, padding=0)
    x1 = qu.activation_quantize_fn(4)(torch.rand(2, 16, 5, 6, generator=_gen(403)) * 1.3 - 0.1)
    out["conv1.w"], out["conv1.b"], out["conv1.x"], out["conv1.y"] = _np(conv1.weight), _np(conv1.bias), _np(x1), _np(conv1(x1))
    lin = qu.linear_Q_fn(4)(40, 24)
    xl = torch.randn(7, 40, generator=_gen(404))
    out["lin.w"], out["lin.b"], out["lin.x"], out["lin.y"] = _np(lin.weight), _np(lin.bias), _np(xl), _np(lin(xl))
    # BatchNorm2d_Q
    bn = qu.batchNorm2d_Q_fn(4)(6).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(6, generator=_gen(405)) * 1.5)
        bn.bias.copy_(torch.randn(6, generator=_gen(406)) * 0.5)
        bn.running_mean.copy_(torch.randn(6, generator=_gen(407)) * 0.3)
        bn.running_var.copy_(torch.rand(6, generator=_gen(408)) * 2 + 0.2)
    xb = torch.randn(2, 6, 4, 5, generator=_gen(409))
    for k in ("weight", "bias", "running_mean", "running_var"):
        out[f"bnq.{k}"] = _np(getattr(bn, k))
    out["bnq.x"] = _np(xb)
    try:                       # QU:125-130 calls F.batch_norm(eps=0): torch >= 2.x raises ValueError, so the
        out["bnq.y"] = _np(bn(xb))          # reference forward cannot run here -> row 14 is "parity unpinned"
    except ValueError as e:
        out["bnq.error"] = np.array(str(e))
    # NumPy BN fold + integer thresholds (QZ) - silence the reference's prints
    import contextlib, io
    gamma, beta = np.array([1.0, 0.5]), np.array([0.1, -0.2])
    mean, var = np.array([0.3, -0.1]), np.array([0.25, 4.0])
    wf, bf = qz.bn_act_w_bias_float(gamma, beta, mean, var, 1e-5)
    with contextlib.redirect_stdout(io.StringIO()):
        inc, bias = qz.bn_act_quantize_int(gamma, beta, mean, var, 1e-5, w_bit=4, in_bit=4, out_bit=4, l_shift=8)
    out["fold.gamma"], out["fold.beta"], out["fold.mean"], out["fold.var"] = gamma, beta, mean, var
    out["fold.w"], out["fold.b"], out["fold.inc"], out["fold.bias"] = wf, bf, inc, bias
    rng = np.random.RandomState(7)
    g2, b2 = rng.rand(64) * 0.4 + 0.1, rng.randn(64) * 0.3
    m2, v2 = rng.randn(64) * 0.5, rng.rand(64) * 3 + 0.05
    with contextlib.redirect_stdout(io.StringIO()):
        inc2, bias2 = qz.bn_act_quantize_int(g2, b2, m2, v2, 1e-5, w_bit=4, in_bit=4, out_bit=4, l_shift=8)
    wf2, bf2 = qz.bn_act_w_bias_float(g2, b2, m2, v2, 1e-5)
    out["fold64.gamma"], out["fold64.beta"], out["fold64.mean"], out["fold64.var"] = g2, b2, m2, v2
    out["fold64.w"], out["fold64.b"], out["fold64.inc"], out["fold64.bias"] = wf2, bf2, inc2, bias2
    # pack convention
    out["pack.codes"] = np.array([1, -1, 7, -7], dtype=np.int64)
    out["pack.word"] = np.uint64(mp.array_to_string([1, -1, 7, -7], 4))
    codes = rng.randint(-7, 8, size=(5, 16))
    out["pack16.codes"] = codes
    out["pack16.words"] = np.array([mp.array_to_string(list(r), 4) for r in codes], dtype=np.uint64)
    np.savez_compressed(os.path.join(OUT, "ultra.npz"), **out)


HLS_CASES = [  # (name, out_ch, in_ch, kernel, simd, pe): the shapes of ultranet_param_gen.py:14-24 plus a ragged one
    ("conv0", 16, 3, 3, 3, 16), ("conv1", 32, 16, 3, 16, 8), ("conv3", 64, 64, 3, 16, 4), ("conv4", 64, 64, 3, 8, 2),
    ("conv8", 36, 64, 1, 8, 2), ("ragged", 4, 5, 3, 8, 2)]


def golden_ultra_hls():
    """FPGA parameter layout of the reference (qnn_mem_process.py:84-170): int4 weight codes -> [pe][tiles] words of
    simd * w_bit bits, BN thresholds -> [pe][a_tiles].  The methods are called on a QNNLayerMemProcess object created
    without its file-reading constructor (only pe / simd / w_bit are used by them)."""
    mp = R.qnn_mem_process()
    out = {}
    rng = np.random.RandomState(11)
    for name, o, i, k, simd, pe in HLS_CASES:
        w = rng.randint(-7, 8, size=(o, i, k, k)).astype(np.int32)
        proc = object.__new__(mp.QNNLayerMemProcess)
        proc.pe, proc.simd, proc.w_bit = pe, simd, 4
        con_w = w.transpose(0, 2, 3, 1).reshape(o, -1)                    # qnn_mem_process.py:152-154
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            res = proc.w_to_hls_array(con_w)
        out[f"{name}.codes"] = w.astype(np.int8)
        out[f"{name}.cfg"] = np.array([o, i, k, simd, pe], dtype=np.int64)
        out[f"{name}.words"] = np.array(res, dtype=np.uint64)            # simd * 4 <= 64 bits in every case
        out[f"{name}.w_tiles"] = np.int64(proc.w_tiles)
        inc = rng.randint(-2000, 2000, size=o).astype(np.int32)
        bias = rng.randint(-(1 << 20), 1 << 20, size=o).astype(np.int32)
        hi, hb = proc.inc_bias_to_hls_array(inc, bias)
        out[f"{name}.inc"], out[f"{name}.bias"], out[f"{name}.hls_inc"], out[f"{name}.hls_bias"] = inc, bias, hi, hb
    np.savez_compressed(os.path.join(OUT, "ultra_hls.npz"), **out)


GETA_STEP_CASES = [  # name, hyper-parameters, stages of the consecutive steps
    ("sgd_range", dict(variant="sgd", lr=0.1, lr_quant=1e-3, first_momentum=0.0, dampening=0.0, weight_decay=None), ["range"] * 3),
    ("sgdm_wd_descent", dict(variant="sgd", lr=0.05, lr_quant=2e-3, first_momentum=0.9, dampening=0.1, weight_decay=1e-2), ["descent"] * 4),
    ("adam_range", dict(variant="adam", lr=1e-3, lr_quant=1e-3, first_momentum=0.9, second_momentum=0.999, weight_decay=None), ["descent", "range", "range", "range"]),
    ("adamw_all", dict(variant="adamw", lr=3e-3, lr_quant=5e-3, first_momentum=0.9, second_momentum=0.99, weight_decay=0.05), ["descent", "range", "range", "fix", "fix"]),
]


def _geta_fake_params(rng, nonlinear):
    """Six layers: W+A (with and without t), weight-only; values around a 4..8-bit operating point."""
    params = {}
    for li in range(6):
        layer = f"blocks.{li}.fc"
        qm = float(rng.uniform(0.2, 3.0))
        params[f"{layer}.d_quant_wt"] = qm / float(rng.choice([3, 7, 15, 127])) * float(rng.uniform(0.8, 1.2))
        params[f"{layer}.q_m_wt"] = qm * (1 if li % 3 else -1)              # q_m may be negative: |.| is taken
        if nonlinear and li % 2 == 0:
            params[f"{layer}.t_quant_wt"] = float(rng.uniform(0.8, 1.2))
        if li != 5:                                                          # last layer: weight-only
            qa = float(rng.uniform(0.5, 4.0))
            params[f"{layer}.d_quant_act"] = qa / float(rng.choice([7, 15, 127]))
            params[f"{layer}.q_m_act"] = qa
            if nonlinear and li % 2 == 0:
                params[f"{layer}.t_quant_act"] = float(rng.uniform(0.8, 1.2))
    return params


def golden_geta_step():
    """Quantizer-scalar half of GETA.step(), run on the REAL class: the methods are called on a GETA object created without
    its constructor (no model / graph needed for them), with one hand-made param_group."""
    import contextlib, importlib, io
    R._stub_only_train_once()
    geta_mod = importlib.import_module("only_train_once.optimizer.geta")
    out = {}
    rng = np.random.RandomState(21)
    for name, hp, stages in GETA_STEP_CASES:
        for nonlinear in (False, True):
            tag = f"{name}.{'nl' if nonlinear else 'lin'}"
            init = _geta_fake_params(rng, nonlinear)
            names = list(init)
            params = [torch.nn.Parameter(torch.tensor([v], dtype=torch.float32)) for v in init.values()]
            opt = object.__new__(geta_mod.GETA)
            group = dict(p_names=names, params=params, grad_variant=dict(), is_prunable=False, active_redundant_idxes=[])
            group.update(dampening=0.0, second_momentum=0.0, first_momentum=0.0)
            group.update(hp)
            opt.param_groups = [group]
            opt.num_steps, opt.safe_guard = 0, 1e-8
            opt.first_moment_grads, opt.second_moment_grads = dict(), dict()
            opt.min_bit_wt, opt.max_bit_wt, opt.min_bit_act, opt.max_bit_act = 4, 8, 3, 6
            opt.grad_clip_min, opt.grad_clip_max = -1.0, 1.0
            out[f"{tag}.names"] = np.array(names)
            out[f"{tag}.init"] = np.array(list(init.values()), dtype=np.float32)
            bit_dict = None
            grads_all, after_all = [], []
            for si, stage in enumerate(stages):
                g = (rng.randn(len(names)) * rng.choice([0.05, 0.5, 3.0], size=len(names))).astype(np.float32)
                if si == 1:
                    g[3] = np.nan if False else g[3]
                    g[1::5] = 0.0                                            # some exact zeros
                for p, gi in zip(params, g):
                    p.grad = torch.tensor([gi], dtype=torch.float32)
                if si == len(stages) - 1:
                    params[2].grad = None                                    # "if p.grad is None: continue"
                    g[2] = np.nan                                            # marker: no gradient
                opt.grad_clipping_names = None
                for p in params:                                             # GETA.grad_clipping (geta.py:160-165), None-safe
                    if p.grad is not None:
                        p.grad = p.grad.clamp(min=opt.grad_clip_min, max=opt.grad_clip_max)
                opt.num_steps += 1
                opt.compute_grad_variant()
                with contextlib.redirect_stdout(io.StringIO()):
                    if stage == "descent":
                        opt.gradient_descent_step(group)
                    elif stage == "range":
                        opt.partial_projected_gradient_descent_step_range_wt(group)
                        opt.partial_projected_gradient_descent_step_range_act(group)
                    else:
                        if bit_dict is None:
                            bit_dict = opt.get_bitwidth_dict(group)
                            out[f"{tag}.bit_layers"] = np.array(list(bit_dict))
                            out[f"{tag}.bit_wt"] = np.array([bit_dict[k]["weight"] for k in bit_dict], dtype=np.float64)
                            out[f"{tag}.bit_act"] = np.array([bit_dict[k].get("activation", -1) for k in bit_dict], dtype=np.float64)
                        opt.partial_projected_gradient_descent_step_fix(group, bit_dict)
                grads_all.append(g)
                after_all.append(np.array([float(p.data) for p in params], dtype=np.float32))
            out[f"{tag}.grads"] = np.stack(grads_all)
            out[f"{tag}.after"] = np.stack(after_all)
            out[f"{tag}.stages"] = np.array(stages)
    np.savez_compressed(os.path.join(OUT, "geta_step.npz"), **out)


def golden_ultranet():
    mm = R.mymodel()
    torch.manual_seed(0)
    net = mm.UltraNetQua().eval()
    sd = net.state_dict()
    ref_models.fill_state_dict_(sd, seed=11, weight_std=0.3)
    # BN affine in the ranges SURVEY.md section 8d suggests so activations are not all saturated
    for k in list(sd.keys()):
        g = _gen(ref_models._seed_for(k, 12))
        if k.endswith(".weight") and sd[k].dim() == 1:
            sd[k].copy_(torch.rand(sd[k].shape, generator=g) * 0.2 + 0.2)
        if k.endswith(".bias") and sd[k].dim() == 1 and not k.startswith("layers.28"):
            sd[k].copy_(torch.rand(sd[k].shape, generator=g) * 0.4 + 0.3)
    net.load_state_dict(sd)
    x = torch.rand(1, 3, 160, 320, generator=_gen(1))
    x = torch.round(x * 255) / 255                    # 8-bit input grid (ultranet_param_gen.py:15)
    with torch.no_grad():
        feats = net.layers(x)
        io, _ = net(x)
        taps = []
        h = x
        for i, layer in enumerate(net.layers):
            h = layer(h)
            if i in (3, 7, 11, 15, 18, 21, 24, 27):
                taps.append(h)
    out = {"x_seed": np.int64(1), "feats": _np(feats), "io": _np(io),
           "x_sum": np.float64(x.double().sum().item())}
    for i, t in enumerate(taps):
        out[f"tap{i}.codes"] = _np(torch.round(t * 15)).astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "ultranet.npz"), **out)


def _vit_golden(vm, qm, name, cfg, batch, num_bits, qtype, act_bits=None, calibrate=False, seed=0, store_sd=False):
    torch.manual_seed(0)
    model = vm.VisionTransformer(**cfg)
    sd = model.state_dict()
    ref_models.fill_state_dict_(sd, seed=seed, weight_std=0.02)
    model.load_state_dict(sd)
    model = qm.model_to_quantize_model(model, num_bits=num_bits, quant_type=qtype,
                                       quant_mode="weight_and_activation").eval()
    x = torch.randn(batch, 3, cfg["img_size"], cfg["img_size"], generator=_gen(1))
    if calibrate or act_bits is not None:
        # fixture B: q_m_act = 99.9th percentile of |layer input| on this batch
        ql = R.quant_layers()
        hooks, stats = [], {}

        def mk(nm):
            def hook(mod, inp):
                a = inp[0].detach().abs().flatten()
                kth = max(1, int(round(0.999 * a.numel())))
                stats[nm] = a.kthvalue(kth).values.item()
            return hook
        bits = act_bits or num_bits
        for nm, mod in model.named_modules():
            if isinstance(mod, (ql.QuantizeLinear, ql.QuantizeConv2d)):
                hooks.append(mod.register_forward_pre_hook(mk(nm)))
        # calibrate layer by layer in one pass: hooks see inputs produced with already-calibrated predecessors
        # only if applied sequentially, so iterate: run, set, re-run until stable (2 passes suffice for a fixture)
        for _ in range(2):
            with torch.no_grad():
                model(x)
            for nm, mod in model.named_modules():
                if nm in stats:
                    with torch.no_grad():
                        mod.q_m_act.fill_(stats[nm])
                        mod.d_quant_act.fill_(stats[nm] / (2 ** (bits - 1) - 1))
        for h in hooks:
            h.remove()
    with torch.no_grad():
        logits = model(x)
    out = {"logits": _np(logits), "x_sum": np.float64(x.double().sum().item()),
           "cfg": np.array([cfg["img_size"], cfg["patch_size"], cfg["embed_dim"], cfg["depth"], cfg["num_heads"],
                            cfg["num_classes"]], dtype=np.int64),
           "batch": np.int64(batch), "fill_seed": np.int64(seed)}
    qnames, qvals = [], []
    for k, v in model.state_dict().items():
        if any(s in k for s in ("d_quant", "q_m", "t_quant")):
            qnames.append(k)
            qvals.append(v.item())
    out["q.names"] = np.array(qnames)
    out["q.values"] = np.array(qvals, dtype=np.float32)
    top2 = torch.topk(logits, 2, dim=-1).values
    out["top1"] = _np(logits.argmax(-1))
    out["margin"] = _np(top2[:, 0] - top2[:, 1])
    if store_sd:
        for k, v in model.state_dict().items():
            out[f"sd.{k}"] = _np(v)
        out["x"] = _np(x)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, "logits[0,:4]", logits[0, :4].tolist(), "min margin", float(out["margin"].min()))


def golden_vit():
    vm, qm = R.vit_model(), R.quant_model()
    tiny = dict(img_size=32, patch_size=8, embed_dim=64, depth=2, num_heads=4, num_classes=10)
    _vit_golden(vm, qm, "vit_tiny_w4a4_init", tiny, 3, 4, "symmetric+linear", store_sd=True)
    _vit_golden(vm, qm, "vit_tiny_w4a4_calib", tiny, 3, 4, "symmetric+linear", calibrate=True, store_sd=True)
    _vit_golden(vm, qm, "vit_tiny_w4a8_calib", tiny, 3, 4, "symmetric+linear", act_bits=8, store_sd=True)
    _vit_golden(vm, qm, "vit_tiny_nl_w8a8_calib", tiny, 3, 8, "symmetric+nonlinear", calibrate=True, store_sd=True)
    base = dict(img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12, num_classes=1000)
    _vit_golden(vm, qm, "vit_b16_w4a4_init", base, 2, 4, "symmetric+linear")
    _vit_golden(vm, qm, "vit_b16_w4a4_calib", base, 2, 4, "symmetric+linear", calibrate=True)


def golden_vit_large():
    """Config 4: ViT-L/16 (vit_model.py:419-433), num_bits=4 conversion, 8-bit activation codes.
    `init`: q_m_act stays max|W| and d_quant_act = q_m_act / 127 (the survey's literal recipe); `calib`: q_m_act = 99.9th
    percentile of the layer input, d = q_m / 127.  Batch 2; weights are re-created on the GPU box by fill_state_dict_."""
    vm, qm = R.vit_model(), R.quant_model()
    large = dict(img_size=224, patch_size=16, embed_dim=1024, depth=24, num_heads=16, num_classes=1000)
    _vit_golden(vm, qm, "vit_l16_w4a8_calib", large, 2, 4, "symmetric+linear", act_bits=8)
    # literal recipe: convert at 4 bits, then d_quant_act = q_m_act / 127
    torch.manual_seed(0)
    model = vm.vit_large_patch16_224(num_classes=1000)
    sd = model.state_dict()
    ref_models.fill_state_dict_(sd, seed=0, weight_std=0.02)
    model.load_state_dict(sd)
    model = qm.model_to_quantize_model(model, num_bits=4, quant_type="symmetric+linear", quant_mode="weight_and_activation").eval()
    with torch.no_grad():
        for m in model.modules():
            if hasattr(m, "q_m_act"):
                m.d_quant_act.copy_(m.q_m_act / 127)
    x = torch.randn(2, 3, 224, 224, generator=_gen(1))
    with torch.no_grad():
        logits = model(x)
    out = {"logits": _np(logits), "x_sum": np.float64(x.double().sum().item()),
           "cfg": np.array([224, 16, 1024, 24, 16, 1000], dtype=np.int64), "batch": np.int64(2), "fill_seed": np.int64(0)}
    qn = [k for k in model.state_dict() if any(t in k for t in ("d_quant", "q_m", "t_quant"))]
    out["q.names"] = np.array(qn)
    out["q.values"] = np.array([model.state_dict()[k].item() for k in qn], dtype=np.float32)
    top2 = torch.topk(logits, 2, dim=-1).values
    out["top1"], out["margin"] = _np(logits.argmax(-1)), _np(top2[:, 0] - top2[:, 1])
    np.savez_compressed(os.path.join(OUT, "vit_l16_w4a8_init.npz"), **out)
    print("vit_l16_w4a8_init logits[0,:4]", logits[0, :4].tolist())


def golden_vit_b16_w8a8():
    vm, qm = R.vit_model(), R.quant_model()
    base = dict(img_size=224, patch_size=16, embed_dim=768, depth=12, num_heads=12, num_classes=1000)
    _vit_golden(vm, qm, "vit_b16_w8a8_calib", base, 2, 8, "symmetric+linear", calibrate=True)


def golden_qat():
    """Config 3 parity: the REFERENCE's autograd (SymQuantizerLinear / NonLinear.backward through QuantizeLinear /
    QuantizeConv2d inside the reference VisionTransformer) on a depth-2, D = 768 ViT, W&A 4-bit, batch 2, cross-entropy.
    Committed: loss, logits, every quantizer-scalar gradient, and a digest (sum, abs-sum, l2, 256 strided samples) of every
    other parameter's gradient."""
    vm, qm = R.vit_model(), R.quant_model()
    ql = R.quant_layers()
    cfg = dict(img_size=224, patch_size=16, embed_dim=768, depth=2, num_heads=12, num_classes=10)
    for tag, qtype in (("lin", "symmetric+linear"), ("nl", "symmetric+nonlinear")):
        torch.manual_seed(0)
        model = vm.VisionTransformer(**cfg)
        sd = model.state_dict()
        ref_models.fill_state_dict_(sd, seed=5, weight_std=0.02)
        model.load_state_dict(sd)
        model = qm.model_to_quantize_model(model, num_bits=4, quant_type=qtype, quant_mode="weight_and_activation")
        x = torch.randn(2, 3, 224, 224, generator=_gen(1))
        labels = torch.randint(0, 10, (2,), generator=_gen(2))
        # calibrated activation ranges (fixture B recipe), two passes
        stats, hooks = {}, []

        def mk(nm):
            def hook(mod, inp):
                a = inp[0].detach().abs().flatten()
                stats[nm] = a.kthvalue(max(1, int(round(0.999 * a.numel())))).values.item()
            return hook
        for nm, mod in model.named_modules():
            if isinstance(mod, (ql.QuantizeLinear, ql.QuantizeConv2d)):
                hooks.append(mod.register_forward_pre_hook(mk(nm)))
        model.eval()
        for _ in range(2):
            with torch.no_grad():
                model(x)
            for nm, mod in model.named_modules():
                if nm in stats:
                    with torch.no_grad():
                        mod.q_m_act.fill_(stats[nm])
                        mod.d_quant_act.fill_(stats[nm] / 7)
        for h in hooks:
            h.remove()
        if tag == "nl":
            with torch.no_grad():
                for mod in model.modules():
                    if hasattr(mod, "t_quant_act"):
                        mod.t_quant_act.fill_(0.9)
                        mod.t_quant_wt.fill_(1.05)
        model.train()
        logits = model(x)
        loss = torch.nn.functional.cross_entropy(logits, labels)
        loss.backward()
        out = {"cfg": np.array([224, 16, 768, 2, 12, 10], dtype=np.int64), "fill_seed": np.int64(5), "batch": np.int64(2),
               "labels": _np(labels), "loss": np.float64(loss.item()), "logits": _np(logits),
               "x_sum": np.float64(x.double().sum().item())}
        qn = [k for k in model.state_dict() if any(t in k for t in ("d_quant", "q_m", "t_quant"))]
        out["q.names"] = np.array(qn)
        out["q.values"] = np.array([model.state_dict()[k].item() for k in qn], dtype=np.float32)
        names = []
        for n, p in model.named_parameters():
            assert p.grad is not None, n
            names.append(n)
            st, smp = ref_models.grad_digest(p.grad)
            out[f"g.{n}.stats"], out[f"g.{n}.samples"] = st, smp
        out["g.names"] = np.array(names)
        np.savez_compressed(os.path.join(OUT, f"qat_vit_d768_{tag}.npz"), **out)
        print("qat", tag, "loss", loss.item(), "grad d_quant_act blocks.0.attn.qkv",
              dict(model.named_parameters())["blocks.0.attn.qkv.d_quant_act"].grad.item())


def golden_bnq():
    """BatchNorm2d_Q / BatchNorm1d_Q (QU:94-207).  Their forward ends in F.batch_norm(eps = eps * 0); torch >= 2 added a
    PYTHON-level guard (eps <= 0 raises) in front of the unchanged ATen op.  The golden runs the unmodified reference
    classes with that guard bypassed: the module's `F.batch_norm` is pointed at a shim that forwards the very same
    arguments to torch.batch_norm (what F.batch_norm did when the reference was written)."""
    qu = R.quant_ultra()

    class _F:
        def __getattr__(self, k):
            return getattr(torch.nn.functional, k)

        @staticmethod
        def batch_norm(input, running_mean, running_var, weight=None, bias=None, training=False, momentum=0.1, eps=1e-5):
            return torch.batch_norm(input, weight, bias, running_mean, running_var, training, momentum, eps, False)
    saved = qu.F
    qu.F = _F()
    out = {}
    try:
        for dim, fn, shape in ((2, qu.batchNorm2d_Q_fn, (2, 6, 4, 5)), (1, qu.batchNorm1d_Q_fn, (7, 6))):
            for bits in (2, 4, 8):
                bn = fn(bits)(6).eval()
                with torch.no_grad():
                    bn.weight.copy_(torch.rand(6, generator=_gen(405)) * 1.5)
                    bn.bias.copy_(torch.randn(6, generator=_gen(406)) * 0.5)
                    bn.running_mean.copy_(torch.randn(6, generator=_gen(407)) * 0.3)
                    bn.running_var.copy_(torch.rand(6, generator=_gen(408)) * 2 + 0.2)
                xb = torch.randn(shape, generator=_gen(409 + dim))
                with torch.no_grad():
                    y = bn(xb)
                for k in ("weight", "bias", "running_mean", "running_var"):
                    out[f"bn{dim}d.{k}"] = _np(getattr(bn, k))
                out[f"bn{dim}d.x"] = _np(xb)
                out[f"bn{dim}d.bit{bits}.y"] = _np(y)
    finally:
        qu.F = saved
    np.savez_compressed(os.path.join(OUT, "ultra_bnq.npz"), **out)


if __name__ == "__main__":
    if not R.available():
        raise SystemExit("reference tree not found (this script only runs in the build container)")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["quantizers", "layers", "ultra", "ultranet", "vit"]
    for w in which:
        globals()[f"golden_{w}"]()
        print("wrote", w)
