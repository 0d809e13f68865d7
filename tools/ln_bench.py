"""LayerNorm -> quantize kernel timing at the ViT-B step's shape ([50432, 768] fp32 -> int8 codes), CUDA events over 40 calls
alternating between two inputs (each 155 MB: larger than L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops

M, D = int(os.environ.get("M", 50432)), int(os.environ.get("D", 768))
xs = [torch.randn(M, D, device="cuda") for _ in range(2)]
g, b = torch.rand(D, device="cuda") + 0.5, torch.randn(D, device="cuda") * 0.1
d, qm = torch.tensor([2.5 / 7], device="cuda"), torch.tensor([2.5], device="cuda")
for _ in range(5):
    ops.layernorm_quantize(xs[0], g, b, 1e-6, d, qm, None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(40):
    ops.layernorm_quantize(xs[i & 1], g, b, 1e-6, d, qm, None)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 40 * 1e3
print(f"layernorm_quantize [{M} x {D}]: {us:.1f} us -> {(M * D * 5) / us / 1e6:.2f} TB/s algorithmic (4 B read + 1 B written per element)")
