import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops
from tools.quick_bench import timeit
M, K, N = 50432, 768, 3072
a = torch.randint(-7, 8, (M, K), dtype=torch.int8, device="cuda")
w = torch.randint(-7, 8, (N, K), dtype=torch.int8, device="cuda")
bias = torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda")
cases = [("none", dict(out_kind=ops.QVIT_OUT_NONE, backend=ops.QVIT_GEMM_TCGEN05)),
         ("i32", dict(out_kind=ops.QVIT_OUT_I32)),
         ("f32", dict(out_kind=ops.QVIT_OUT_F32, bias=bias)),
         ("f32+res", dict(out_kind=ops.QVIT_OUT_F32, bias=bias, residual=res)),
         ("f32+gelu", dict(out_kind=ops.QVIT_OUT_F32, bias=bias, act=ops.QVIT_ACT_GELU)),
         ("f32+relu", dict(out_kind=ops.QVIT_OUT_F32, bias=bias, act=ops.QVIT_ACT_RELU)),
         ("bf16", dict(out_kind=ops.QVIT_OUT_BF16, bias=bias)),
         ("bf16+gelu", dict(out_kind=ops.QVIT_OUT_BF16, bias=bias, act=ops.QVIT_ACT_GELU)),
         ("i8", dict(out_kind=ops.QVIT_OUT_I8, bias=bias, next_q=(0.3, 2.1, None))),
         ("i8+relu", dict(out_kind=ops.QVIT_OUT_I8, bias=bias, act=ops.QVIT_ACT_RELU, next_q=(0.3, 2.1, None))),
         ("i8+gelu", dict(out_kind=ops.QVIT_OUT_I8, bias=bias, act=ops.QVIT_ACT_GELU, next_q=(0.3, 2.1, None))),
         ("i8+gelu nonlinear-q", dict(out_kind=ops.QVIT_OUT_I8, bias=bias, act=ops.QVIT_ACT_GELU, next_q=(0.3, 2.1, 1.0)))]
from quantized_vit_b200 import _lib
for cg in (1, 21):
  _lib.lib().qvit_gemm_set_cta_group(cg)
  print(f"--- cta_group {cg}")
  for name, kw in cases:
    kw = dict(kw, scale_a=0.1, scale_w=0.01)
    if kw["out_kind"] != ops.QVIT_OUT_NONE:
        kw["out"] = ops.gemm_i8(a, w, K, N, **kw)
    med, best = timeit(lambda: ops.gemm_i8(a, w, K, N, **kw), iters=10)
    print(f"{name:22s} {med*1e3:8.1f} us  {2.0*M*K*N/(med*1e-3)/1e12:7.1f} TOPS", flush=True)
_lib.lib().qvit_gemm_set_cta_group(0)
