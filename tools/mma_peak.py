import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops, _lib
from tools.quick_bench import timeit
L = _lib.lib()
for (M, K, N) in [(50432, 768, 3072), (50432, 4096, 4096), (148 * 128 * 4, 16384, 256)]:
    a = torch.randint(-7, 8, (M, K), dtype=torch.int8, device="cuda")
    w = torch.randint(-7, 8, (N, K), dtype=torch.int8, device="cuda")
    for mode in (1, 2, 11, 12):
        L.qvit_gemm_set_cta_group(mode)
        f = lambda: ops.gemm_i8(a, w, K, N, out_kind=ops.QVIT_OUT_NONE, backend=ops.QVIT_GEMM_TCGEN05)
        med, best = timeit(f, iters=10, graph=True)
        print(f"M={M} K={K} N={N} mode={mode:2d}: {med*1e3:8.1f} us  {2.0*M*K*N/(med*1e-3)/1e12:7.1f} TOPS", flush=True)
L.qvit_gemm_set_cta_group(0)
# library int8 GEMM for comparison (cuBLASLt through torch._int_mm)
a = torch.randint(-7, 8, (8192, 8192), dtype=torch.int8, device="cuda")
b = torch.randint(-7, 8, (8192, 8192), dtype=torch.int8, device="cuda")
try:
    med, best = timeit(lambda: torch._int_mm(a, b.t()), iters=10, graph=True)
    print(f"torch._int_mm 8192^3: {med*1e3:.1f} us {2*8192**3/(med*1e-3)/1e12:.1f} TOPS")
except Exception as e:
    print("torch._int_mm failed:", e)
x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
med, best = timeit(lambda: x @ x, iters=10, graph=True)
print(f"bf16 matmul 8192^3: {med*1e3:.1f} us {2*8192**3/(med*1e-3)/1e12:.1f} TFLOPS")
