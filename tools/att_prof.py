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
t = dbg[0, 0, 255, 500:507].tolist()
names = ["Q conv done", "K landed+conv done", "S MMA done", "softmax+P store done", "V landed+transpose done", "PV MMA done", "epilogue+sync done"]
prev = 0
for n, v in zip(["Q conv (global loads)", "wait K + K conv", "S MMA", "softmax + P planes", "wait V + V transpose", "PV MMA", "epilogue + sync"], t):
    print(f"{n:28s} {v - prev:8.0f} cycles   (cum {v:8.0f})")
    prev = v
