import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import _lib
B, T, H = 64, 197, 12
qkv = torch.randn(B, T, 3 * H * 64, device="cuda")
out = torch.zeros(B, T, H * 64, device="cuda")
dbg = torch.zeros(B, H, 256, 512, device="cuda")
for _ in range(2):
    _lib.check(_lib.lib().qvit_attention_f32_debug(qkv.data_ptr(), B, T, H, 64, 0.125, out.data_ptr(), dbg.data_ptr(), 2, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
for which, off in (("first tile of a pair", 500), ("second tile of a pair", 508)):
    t = dbg[0, 0, 255, off:off + 6].tolist()
    print(which)
    prev = 0
    for n, v in zip(["A: issue S MMA + V^T conversion", "A: wait S MMA", "B: softmax + P planes + barrier", "C: issue PV MMA + Q/K conversion",
                     "C: wait PV MMA", "D: epilogue + barrier"], t):
        print(f"  {n:36s} {v - prev:8.0f} cycles   (cum {v:8.0f})")
        prev = v
