import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops
from tools.quick_bench import timeit
B, T, H = 256, 197, 12
qkv = torch.randn(B, T, 3 * H * 64, device="cuda")
med, best = timeit(lambda: ops.attention_f32(qkv, H), iters=10, graph=True)
fl = 4.0 * B * H * T * T * 64
print(f"attention_f32 (3 x bf16 split) B={B}: {med*1e3:.1f} us  ({fl/med/1e9:.1f} TFLOP/s fp32-equivalent)")
q, k, v = (qkv.view(B, T, 3, H, 64)[:, :, j].transpose(1, 2) for j in range(3))
med, best = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), iters=10, graph=True)
print(f"library fp32 sdpa: {med*1e3:.1f} us")
