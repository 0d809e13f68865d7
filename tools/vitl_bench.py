"""Config 4: ViT-L/16 W4A8 inference, 128 images per GPU (batch 1024 = 128 x 8 GPUs)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200.engine import ViTInferenceEngine
from quantized_vit_b200.engine.synthetic import vit_state_dict
from tools.quick_bench import timeit
sd = vit_state_dict(embed_dim=1024, depth=24, num_heads=16, num_bits=4, act_bits=8, calibrate_to=3.0)
eng = ViTInferenceEngine(sd, depth=24, num_heads=16)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
xs, ys, g = eng.capture(B)
xs.copy_(torch.randn_like(xs))
med, best = timeit(lambda: g.replay(), iters=10)
print(json.dumps({"config": "ViT-L/16 W4A8 inference", "batch_per_gpu": B, "ms_per_step": med, "img_per_s": B / med * 1e3,
                  "gemm_tops": eng.gemm_ops_per_image() * B / (med * 1e-3) / 1e12, "flags": int(eng.flags.item())}))
