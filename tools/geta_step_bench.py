"""Developer benchmark: the fused quantizer-scalar step (qvit_geta_quant_step) against the reference's per-parameter loop
(restated in oracle/ref_geta_step.py) running on the same GPU tensors - ViT-B/16, non-linear W&A quantizers: 300 scalars."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200.engine.vit_module import VisionTransformer
from quantized_vit_b200.quantization import GetaQuantParamStepper, model_to_quantize_model
from oracle import ref_geta_step

torch.manual_seed(0)
model = model_to_quantize_model(VisionTransformer(num_classes=1000), num_bits=4, quant_type="symmetric+nonlinear",
                                quant_mode="weight_and_activation").cuda()
quant = {n: p for n, p in model.named_parameters() if any(t in n for t in ("d_quant", "q_m", "t_quant"))}
for p in quant.values():
    p.grad = torch.randn_like(p) * 0.1
hp = dict(variant="adamw", lr=1e-3, lr_quant=1e-3, first_momentum=0.9, second_momentum=0.999, weight_decay=0.01)
st = GetaQuantParamStepper(model.named_parameters(), **hp)
ref = ref_geta_step.GetaQuantStepRef(**hp)
print(f"{len(quant)} quantizer scalars in {len(st.layer_names)} layers")
for name, fn in (("fused (1 launch)", lambda: st.step("range")), ("reference loop on GPU tensors", lambda: ref.step(quant, "range"))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    print(f"{name:32s} {(time.perf_counter() - t0) / n * 1e3:8.3f} ms per step (wall clock, host + device)")
