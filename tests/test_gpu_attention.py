"""Tensor-core (exact 3-way bf16 split, tcgen05 kind::f16, fp32 TMEM accumulation) attention core (glue beside the hot path, vit_model.py:141-149): fp32-equivalent accuracy.
Reference = the same op sequence as the reference module in float64 on a deliberately sharp softmax (scores up to
+-70).  Bar: max|delta| <= 6e-6 * max|ref|; measured 3e-6 (the library fp32 kernel scores 1e-6 on the same input, a
single-pass TF32 kernel ~5e-4, bf16 ~4e-3)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, H):
    B, T, C3 = qkv.shape
    hd = C3 // 3 // H
    q, k, v = qkv.double().reshape(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    a = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1)
    return (a @ v).transpose(1, 2).reshape(B, T, H * hd)


@pytest.mark.parametrize("B,T,H", [(2, 197, 12), (1, 50, 3), (3, 128, 2), (1, 208, 1), (2, 129, 4), (1, 1, 1), (5, 17, 16)])
def test_attention_matches_fp64(B, T, H):
    from quantized_vit_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = torch.randn(B, T, 3 * H * 64, generator=g).cuda()
    qkv[..., : H * 64] *= 3.0                        # sharper softmax: large score range
    out = ops.attention_f32(qkv, H)
    ref = _ref(qkv, H)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    plain = torch.nn.functional.scaled_dot_product_attention(
        *(qkv.reshape(B, T, 3, H, 64).permute(2, 0, 3, 1, 4))).transpose(1, 2).reshape(B, T, H * 64)
    err_plain = float((plain.double() - ref).abs().max() / ref.abs().max())
    print(f"B={B} T={T} H={H}: tensor-core err {err:.2e} (library fp32 kernel {err_plain:.2e})")
    assert err <= 6e-6


def test_attention_rejects_unsupported_shapes():
    from quantized_vit_b200 import ops
    with pytest.raises(RuntimeError, match="head_dim == 64"):
        ops.attention_f32(torch.randn(1, 10, 3 * 2 * 32).cuda(), 2)
    with pytest.raises(RuntimeError, match="T <= 208"):
        ops.attention_f32(torch.randn(1, 300, 3 * 64).cuda(), 1)


@pytest.mark.parametrize("B,T,H", [(2, 197, 12), (3, 129, 2), (1, 5, 1)])
@pytest.mark.parametrize("nonlinear", [False, True])
def test_fused_quantizer_equals_separate_quantize(B, T, H, nonlinear):
    """qvit_attention_quantize_sym = qvit_attention_f32 followed by qvit_quantize_sym (the consumer's quantize_act,
    quant_layers.py:356-381), bit for bit, and the optional fp32 context equals the unfused one."""
    from quantized_vit_b200 import ops
    g = torch.Generator().manual_seed(B * 77 + T)
    qkv = torch.randn(B, T, 3 * H * 64, generator=g).cuda()
    d, qm, t = 0.031, 0.217, (0.9 if nonlinear else None)
    ctx = ops.attention_f32(qkv, H)
    want = ops.quantize_sym(ctx.view(B * T, H * 64), d, qm, t, ld_codes=ops.pad16(H * 64))
    codes, ctx2 = ops.attention_quantize_sym(qkv, H, d, qm, t, want_context=True)
    assert torch.equal(ctx2, ctx)
    assert torch.equal(codes, want)
    codes_only, none = ops.attention_quantize_sym(qkv, H, d, qm, t)
    assert none is None and torch.equal(codes_only, want)


# ------------------------------------------------------------------------------------------ two-plane fp16 kernel (pipelined)
def _planes(qkv, H):
    """fp32 qkv [B, T, 3*H*64] -> two-plane fp16 [B*T, 2*3*H*64] with one power of two per part from the data's own range
    (the engine derives it from a static bound instead; any exponent that keeps |x 2^e| < 65504 is exact)."""
    from quantized_vit_b200 import ops
    B, T, C3 = qkv.shape
    D = C3 // 3
    exps = [ops.f16x2_exponent(float(qkv[..., i * D:(i + 1) * D].abs().max())) for i in range(3)]
    col_exp = torch.cat([torch.full((D,), e, dtype=torch.int32) for e in exps]).cuda()
    fl = ops.new_flags(qkv.device)
    planes = ops.split2_f16(qkv.reshape(B * T, C3), col_exp, flags=fl)
    assert int(fl.item()) == 0
    return planes, exps


@pytest.mark.parametrize("B,T,H", [(2, 197, 12), (1, 50, 3), (3, 128, 2), (1, 208, 1), (2, 129, 4), (1, 1, 1), (5, 17, 16), (40, 197, 12)])
def test_attention_f16x2_matches_fp64(B, T, H):
    """qvit_attention_f16x2: hi/lo fp16 planes, three product terms, fp32 accumulation in TMEM, TMA-fed, pipelined over two
    S/P buffers.  Same bar as the 3 x bf16 kernel (6e-6 of max|ref| on a deliberately sharp softmax); (40, 197, 12) gives
    every CTA several (batch, head) units so that all buffer rotations of the pipeline are exercised."""
    from quantized_vit_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = torch.randn(B, T, 3 * H * 64, generator=g).cuda()
    qkv[..., : H * 64] *= 3.0
    planes, exps = _planes(qkv, H)
    sums = planes.float().view(B * T, 2, -1).sum(1)
    scale = torch.cat([torch.full((H * 64,), 2.0 ** e) for e in exps]).cuda()
    assert (sums / scale - qkv.view(B * T, -1)).abs().max() <= 2.0 ** -21 * qkv.abs().max()      # hi + lo carries >= 22 bits
    _, out = ops.attention_f16x2(planes, B, T, H, exps, want_codes=False, want_context=True)
    ref = _ref(qkv, H)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    print(f"B={B} T={T} H={H}: two-plane fp16 tensor-core err {err:.2e}")
    assert err <= 6e-6


@pytest.mark.parametrize("nonlinear", [False, True])
def test_attention_f16x2_fused_quantizer(nonlinear):
    from quantized_vit_b200 import ops
    B, T, H = 3, 197, 4
    qkv = torch.randn(B, T, 3 * H * 64, generator=torch.Generator().manual_seed(9)).cuda()
    planes, exps = _planes(qkv, H)
    d, qm, t = 0.031, 0.217, (0.9 if nonlinear else None)
    codes, ctx = ops.attention_f16x2(planes, B, T, H, exps, d, qm, t, want_context=True)
    want = ops.quantize_sym(ctx.view(B * T, H * 64), d, qm, t, ld_codes=ops.pad16(H * 64))
    assert torch.equal(codes, want)
    codes_only, none = ops.attention_f16x2(planes, B, T, H, exps, d, qm, t)
    assert none is None and torch.equal(codes_only, want)


def _attention_ref64(qkv, H, gout):
    B, T, D3 = qkv.shape
    x = qkv.double().requires_grad_(True)
    q, k, v = x.reshape(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    p = torch.softmax(q @ k.transpose(-2, -1) * 64 ** -0.5, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, T, D3 // 3)
    o.backward(gout.double())
    lse2 = torch.logsumexp(q.detach() @ k.detach().transpose(-2, -1) * 64 ** -0.5, dim=-1) / np.log(2.0)
    return o.detach(), x.grad, lse2


@pytest.mark.parametrize("B,T,H,gscale", [(3, 197, 4, 1.0), (2, 208, 2, 1e-5), (5, 50, 2, 30.0), (1, 129, 12, 1.0), (37, 197, 12, 1e-3)])
def test_attention_train_forward_backward(B, T, H, gscale):
    """csrc/attention_train.cu against float64 autograd of ViTAttention's core (vit_model.py:141-149): forward at the fp32 level,
    backward at the 16-bit operand level (stated tolerance 1e-4 of the largest gradient; gradients of any magnitude)."""
    from quantized_vit_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + T)
    qkv = (torch.randn(B, T, 3 * H * 64, generator=g) * 1.5).cuda()
    qkv[:, :, : H * 64] *= 2.0                                   # peaked rows: |s| up to ~ 20
    gout = (torch.randn(B, T, H * 64, generator=g) * gscale).cuda()
    o_ref, dqkv_ref, lse_ref = _attention_ref64(qkv, H, gout)
    out, lse = ops.attention_train_fwd(qkv, H)
    assert float((out.double() - o_ref).abs().max()) <= 3e-6 * float(o_ref.abs().max())
    assert float((lse[:, :, :T].double() - lse_ref).abs().max()) <= 2e-5
    dqkv = ops.attention_train_bwd(qkv, out, lse, gout, H)
    D = H * 64
    for part, name in enumerate(("dq", "dk", "dv")):
        got, want = dqkv[:, :, part * D:(part + 1) * D].double(), dqkv_ref[:, :, part * D:(part + 1) * D]
        err = float((got - want).abs().max() / want.abs().max())
        assert err <= 1e-4, f"{name}: {err:.2e}"


def test_attention_module_uses_train_kernels():
    """ViTAttention of the drop-in model (engine/vit_module.py) under autograd against the SDPA formulation it replaces."""
    from quantized_vit_b200.engine.vit_module import ViTAttention
    torch.manual_seed(3)
    att = ViTAttention(128, 2).cuda()
    x = torch.randn(4, 197, 128, device="cuda", requires_grad=True)
    y = att(x)
    y.square().sum().backward()
    gx, gw = x.grad.clone(), att.qkv.weight.grad.clone()
    x.grad = None
    att.zero_grad()
    qkv = att.qkv(x).reshape(4, 197, 3, 2, 64).permute(2, 0, 3, 1, 4)
    o = torch.nn.functional.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
    y2 = att.proj(o.transpose(1, 2).reshape(4, 197, 128))
    y2.square().sum().backward()
    assert float((y - y2).detach().abs().max()) <= 1e-5 * float(y2.detach().abs().max())
    assert float((gx - x.grad).abs().max()) <= 2e-4 * float(x.grad.abs().max())
    assert float((gw - att.qkv.weight.grad).abs().max()) <= 2e-4 * float(att.qkv.weight.grad.abs().max())
