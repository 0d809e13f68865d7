"""Kernel-time table of one QAT step (ViT-B/16 W4A4, batch 128, one GPU): torch.profiler, top kernels by device time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from quantized_vit_b200 import parallel
from quantized_vit_b200.engine.vit_module import VisionTransformer
from quantized_vit_b200.quantization import model_to_quantize_model

qtype = sys.argv[1] if len(sys.argv) > 1 else "symmetric+linear"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
model = model_to_quantize_model(VisionTransformer(num_classes=1000), num_bits=4, quant_type=qtype,
                                quant_mode="weight_and_activation").cuda().train()
with torch.no_grad():
    for m in model.modules():
        if hasattr(m, "q_m_act"):
            m.q_m_act.fill_(2.5); m.d_quant_act.fill_(2.5 / 7)
red = parallel.GradientAllReducer(model.named_parameters())
x = torch.randn(batch, 3, 224, 224, device="cuda")
y = torch.randint(0, 1000, (batch,), device="cuda")
crit = torch.nn.CrossEntropyLoss()


def step():
    red.zero_grad()
    crit(model(x), y).backward()
    red.reduce(); red.clip_(1.0)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device time {tot / 1e3:.2f} ms over {sum(r[2] for r in rows)} launches")
for k, t, c in rows[:40]:
    print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  x{c:<4d} {k[:130]}")
