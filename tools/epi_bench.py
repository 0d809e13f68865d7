import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops
from tools.quick_bench import timeit
M, K, N = 50432, int(os.environ.get("K", 768)), int(os.environ.get("N", 3072))
a = torch.randint(-7, 8, (M, K), dtype=torch.int8, device="cuda")
w = torch.randint(-7, 8, (N, K), dtype=torch.int8, device="cuda")
bias = torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda")
T = lambda v: torch.tensor([v], dtype=torch.float32, device="cuda")
cases = [("none", dict(out_kind=ops.QVIT_OUT_NONE, backend=ops.QVIT_GEMM_TCGEN05)),
         ("i32", dict(out_kind=ops.QVIT_OUT_I32)),
         ("f32", dict(out_kind=ops.QVIT_OUT_F32, bias=bias)),
         ("f32+res", dict(out_kind=ops.QVIT_OUT_F32, bias=bias, residual=res)),
         ("f32+gelu", dict(out_kind=ops.QVIT_OUT_F32, bias=bias, act=ops.QVIT_ACT_GELU)),
         ("f32+relu", dict(out_kind=ops.QVIT_OUT_F32, bias=bias, act=ops.QVIT_ACT_RELU)),
         ("bf16", dict(out_kind=ops.QVIT_OUT_BF16, bias=bias)),
         ("bf16+gelu", dict(out_kind=ops.QVIT_OUT_BF16, bias=bias, act=ops.QVIT_ACT_GELU)),
         ("i8", dict(out_kind=ops.QVIT_OUT_I8, bias=bias, next_q=(T(0.3), T(2.1), None))),
         ("i8+relu", dict(out_kind=ops.QVIT_OUT_I8, bias=bias, act=ops.QVIT_ACT_RELU, next_q=(T(0.3), T(2.1), None))),
         ("i8+gelu", dict(out_kind=ops.QVIT_OUT_I8, bias=bias, act=ops.QVIT_ACT_GELU, next_q=(T(0.3), T(2.1), None))),
         ("i8+gelu A8", dict(out_kind=ops.QVIT_OUT_I8, bias=bias, act=ops.QVIT_ACT_GELU, next_q=(T(2.1 / 127), T(2.1), None))),
         ("i8+gelu nonlinear-q", dict(out_kind=ops.QVIT_OUT_I8, bias=bias, act=ops.QVIT_ACT_GELU, next_q=(T(0.3), T(2.1), T(0.9))))]
from quantized_vit_b200 import _lib
modes = [int(v) for v in sys.argv[1:]] or [1]
sel = os.environ.get("EPI_CASES")
for cg in modes:
  hint = 49 * K
  _lib.lib().qvit_gemm_set_cta_group(cg)
  print(f"--- mode {cg} acc_abs_max {hint}")
  for name, kw in cases:
    if sel and name not in sel.split(","):
        continue
    kw = dict(kw, scale_a=T(0.1), scale_w=T(0.01), acc_abs_max=hint)
    if kw["out_kind"] != ops.QVIT_OUT_NONE:
        kw["out"] = ops.gemm_i8(a, w, K, N, **kw)
    med, best = timeit(lambda: ops.gemm_i8(a, w, K, N, **kw), iters=10, graph=True)
    print(f"{name:22s} {med*1e3:8.1f} us  {2.0*M*K*N/(med*1e-3)/1e12:7.1f} TOPS", flush=True)
_lib.lib().qvit_gemm_set_cta_group(0)
