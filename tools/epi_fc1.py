"""fc1-shaped GEMM (768 -> 3072, GELU + requantize -> int8, M = 50432) a few times: target of ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops
M, K, N = 50432, 768, 3072
a = torch.randint(-7, 8, (M, K), dtype=torch.int8, device="cuda")
w = torch.randint(-7, 8, (N, K), dtype=torch.int8, device="cuda")
bias = torch.randn(N, device="cuda")
T = lambda v: torch.tensor([v], dtype=torch.float32, device="cuda")
out = None
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = ops.gemm_i8(a, w, K, N, out_kind=ops.QVIT_OUT_I8, bias=bias, act=ops.QVIT_ACT_GELU, next_q=(T(0.3), T(2.1), None),
                      scale_a=T(0.1), scale_w=T(0.01), acc_abs_max=49 * K, out=out)
torch.cuda.synchronize()
print("ok", int(out.abs().max()))
