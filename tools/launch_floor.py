"""Developer check: per-launch floor of the step's kernels inside a CUDA graph (tiny problems: prologue + launch + exit)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops
from tools.quick_bench import timeit
T = lambda v: torch.tensor([v], dtype=torch.float32, device="cuda")
a = torch.randint(-7, 8, (128, 128), dtype=torch.int8, device="cuda")
w = torch.randint(-7, 8, (256, 128), dtype=torch.int8, device="cuda")
d, qm = T(0.3), T(2.1)
out32 = torch.empty(128, 256, device="cuda")
out8 = torch.empty(128, 256, dtype=torch.int8, device="cuda")
x = torch.randn(8, 768, device="cuda"); g = torch.ones(768, device="cuda"); b = torch.zeros(768, device="cuda")
qkv = torch.randn(1, 8, 3 * 64, device="cuda")
cases = [("gemm f32 (1 tile)", lambda: ops.gemm_i8(a, w, 128, 256, out_kind=ops.QVIT_OUT_F32, scale_a=d, scale_w=d, out=out32)),
         ("gemm i8+gelu (1 tile)", lambda: ops.gemm_i8(a, w, 128, 256, out_kind=ops.QVIT_OUT_I8, act=ops.QVIT_ACT_GELU, scale_a=d, scale_w=d, next_q=(d, qm, None), out=out8)),
         ("layernorm_quantize (8 rows)", lambda: ops.layernorm_quantize(x, g, b, 1e-6, d, qm, None)),
         ("attention (1 head, 8 tokens)", lambda: ops.attention_quantize_sym(qkv, 1, d, qm, None)),
         ("torch add (8 x 768)", lambda: torch.add(x, x))]
for name, fn in cases:
    med, best = timeit(fn, iters=50, graph=True)
    print(f"{name:32s} {med * 1e3:7.2f} us per launch (graph of 50, back to back)")
