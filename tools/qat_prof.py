import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from quantized_vit_b200.engine.vit_module import VisionTransformer
from quantized_vit_b200.quantization import model_to_quantize_model
torch.manual_seed(0)
model = model_to_quantize_model(VisionTransformer(num_classes=1000), num_bits=4, quant_type="symmetric+linear",
                                quant_mode="weight_and_activation").cuda().train()
with torch.no_grad():
    for m in model.modules():
        if hasattr(m, "q_m_act"):
            m.q_m_act.fill_(2.5); m.d_quant_act.fill_(2.5 / 7)
x = torch.randn(128, 3, 224, 224, device="cuda"); y = torch.randint(0, 1000, (128,), device="cuda")
crit = torch.nn.CrossEntropyLoss()
def step():
    model.zero_grad(set_to_none=True)
    crit(model(x), y).backward()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=70))
