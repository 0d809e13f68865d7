"""Config 5: standalone QuantLinear GEMM sweep, M = 197*B, 4-bit weights, 4/8-bit activations.
kernel mode : int8 codes in, weight codes resident -> bf16 out / int8-requantised out   (TOP/s vs int8 tensor peak)
drop-in mode: fp32 in -> quantize kernel -> GEMM -> fp32 out                            (HBM-bound; GB/s too)
Device-timed by replaying a CUDA graph of 10 calls (no host launch latency inside; operands of the small-B cases stay in L2)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops
from tools.quick_bench import timeit

rows = []
for (K, N) in [(768, 2304), (768, 3072), (3072, 768), (1024, 4096)]:
    w = torch.nn.init.trunc_normal_(torch.empty(N, K), std=0.02).cuda()
    qm_w = ops.absmax(w); d_w = qm_w / 7
    wc = ops.quantize_sym(w, d_w, qm_w, None, ld_codes=K)
    bias = torch.zeros(N, device="cuda")
    for B in (1, 4, 16, 64, 256, 1024):
        M = 197 * B
        x = torch.randn(M, K, device="cuda")
        for abits in (4, 8):
            qm_a = torch.tensor([2.5], device="cuda"); d_a = qm_a / (2 ** (abits - 1) - 1)
            ac = ops.quantize_sym(x, d_a, qm_a, None, ld_codes=K)
            hint = (2 ** (abits - 1) - 1) * 7 * K
            r = {"K": K, "N": N, "B": B, "M": M, "a_bits": abits}
            for name, kw in (("bf16", dict(out_kind=ops.QVIT_OUT_BF16)),
                             ("i8", dict(out_kind=ops.QVIT_OUT_I8, next_q=(d_a, qm_a, None)))):
                out = ops.gemm_i8(ac, wc, K, N, scale_a=d_a, scale_w=d_w, bias=bias, acc_abs_max=hint, **kw)
                med, _ = timeit(lambda: ops.gemm_i8(ac, wc, K, N, scale_a=d_a, scale_w=d_w, bias=bias, out=out, acc_abs_max=hint, **kw), iters=10, graph=True)
                r[f"kernel_{name}_us"] = med * 1e3
                r[f"kernel_{name}_tops"] = 2.0 * M * K * N / (med * 1e-3) / 1e12
            def dropin():
                a = ops.quantize_sym(x, d_a, qm_a, None, ld_codes=K)
                return ops.gemm_i8(a, wc, K, N, out_kind=ops.QVIT_OUT_F32, scale_a=d_a, scale_w=d_w, bias=bias, acc_abs_max=hint)
            med, _ = timeit(dropin, iters=10, graph=True)
            r["dropin_f32_us"] = med * 1e3
            r["dropin_tops"] = 2.0 * M * K * N / (med * 1e-3) / 1e12
            r["dropin_hbm_gbs"] = (M * K * 5 + N * K + M * K + M * N * 4) / (med * 1e-3) / 1e9
            rows.append(r)
            print(json.dumps(r), flush=True)
