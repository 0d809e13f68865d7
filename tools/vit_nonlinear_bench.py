"""Developer benchmark: ViT-B/16 W4A4 through the engine with NON-LINEAR quantizers (t_quant present; what train.py trains),
256 images, CUDA-graph replay - next to the linear configuration of bench.py."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200.engine import ViTInferenceEngine
from tests import fixtures
from oracle import ref_geta
from tools.quick_bench import timeit

for tval in (None, 0.9):
    sd = fixtures.vit_state_dict(seed=0)
    names = [k[:-7] for k in sd if k.endswith(".weight") and sd[k].dim() >= 2]
    for n in names:
        d, qm = ref_geta.init_quant_params(sd[n + ".weight"], 4)
        sd[n + ".d_quant_wt"], sd[n + ".q_m_wt"] = d, qm
        sd[n + ".d_quant_act"], sd[n + ".q_m_act"] = d.clone(), qm.clone()
        if tval is not None:
            sd[n + ".t_quant_wt"] = torch.tensor([tval])
            sd[n + ".t_quant_act"] = torch.tensor([tval])
            r = torch.exp(tval * torch.log(qm.abs() + 1e-6))            # keep 4-bit codes: d = q_m^t / 7
            sd[n + ".d_quant_wt"] = r / 7
            sd[n + ".d_quant_act"] = r / 7
    eng = ViTInferenceEngine(sd, depth=12, num_heads=12)
    xs, ys, g = eng.capture(256)
    xs.copy_(torch.randn(256, 3, 224, 224, device="cuda"))
    med, best = timeit(lambda: g.replay(), iters=10)
    print(f"ViT-B/16 W4A4 {'linear' if tval is None else f'non-linear (t = {tval})'}: {med:.2f} ms per 256 images -> {256 / med * 1e3:.0f} img/s, flags {int(eng.flags.item())}")
