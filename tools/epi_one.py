"""Developer tool: run ONE epilogue case of the tensor-core GEMM a few times (for ncu captures).
usage: epi_one.py <mode> <case> [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops, _lib

mode, case = int(sys.argv[1]), sys.argv[2]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
M, K, N = 50432, int(os.environ.get("K", 768)), int(os.environ.get("N", 3072))
a = torch.randint(-7, 8, (M, K), dtype=torch.int8, device="cuda")
w = torch.randint(-7, 8, (N, K), dtype=torch.int8, device="cuda")
bias = torch.randn(N, device="cuda")
cases = {"none": dict(out_kind=ops.QVIT_OUT_NONE, backend=ops.QVIT_GEMM_TCGEN05),
         "i32": dict(out_kind=ops.QVIT_OUT_I32),
         "f32": dict(out_kind=ops.QVIT_OUT_F32, bias=bias),
         "bf16": dict(out_kind=ops.QVIT_OUT_BF16, bias=bias),
         "i8": dict(out_kind=ops.QVIT_OUT_I8, bias=bias, next_q=(0.3, 2.1, None)),
         "i8+gelu": dict(out_kind=ops.QVIT_OUT_I8, bias=bias, act=ops.QVIT_ACT_GELU, next_q=(0.3, 2.1, None))}
kw = dict(cases[case], scale_a=0.1, scale_w=0.01, acc_abs_max=49 * K)
_lib.lib().qvit_gemm_set_cta_group(mode)
if kw["out_kind"] != ops.QVIT_OUT_NONE:
    kw["out"] = ops.gemm_i8(a, w, K, N, **kw)
for _ in range(iters):
    ops.gemm_i8(a, w, K, N, **kw)
torch.cuda.synchronize()
print("ok", mode, case)
