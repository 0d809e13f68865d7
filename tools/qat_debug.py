"""Where does the model-level QAT forward first leave the reference?  For every quantized layer of the depth-2 D = 768 golden
model: deviation of its input / output from the oracle's taps and the number of activation codes that differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ref_geta, ref_models
from tests import fixtures
import tests.test_gpu_qat as T
from quantized_vit_b200.quantization import QuantizeConv2d, QuantizeLinear

tag = sys.argv[1] if len(sys.argv) > 1 else "nl"
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"qat_vit_d768_{tag}.npz"))
img, patch, dim, depth, heads, classes = [int(v) for v in g["cfg"]]
cfg = dict(img=img, patch=patch, dim=dim, depth=depth, heads=heads, classes=classes)
sd = fixtures.vit_state_dict(img, patch, dim, depth, heads, classes, seed=int(g["fill_seed"]))
for k, v in zip(g["q.names"], g["q.values"]):
    sd[str(k)] = torch.tensor([float(v)], dtype=torch.float32)
x = fixtures.vit_input(2, img, seed=1)
taps = {"__layers__": True}
torch.set_num_threads(os.cpu_count())
ref_models.vit_forward(sd, x, depth, heads, patch, taps=taps)
model = T._our_model(sd, cfg, tag)
caps = {}
for name, mod in model.named_modules():
    if isinstance(mod, (QuantizeLinear, QuantizeConv2d)):
        mod.register_forward_hook(lambda m, i, o, name=name: caps.__setitem__(name, (i[0].detach().cpu(), o.detach().cpu())))
model(x.cuda())
for name in caps:
    xi, yo = caps[name]
    xr = taps.get(f"{name}.in") if name != "patch_embed.proj" else x
    yr = taps.get(f"{name}.y")
    if name == "patch_embed.proj":
        yr = None
    t = sd.get(f"{name}.t_quant_act")
    ca = ref_geta.sym_codes(xi, sd[f"{name}.d_quant_act"], sd[f"{name}.q_m_act"], t)
    cr = ref_geta.sym_codes(xr, sd[f"{name}.d_quant_act"], sd[f"{name}.q_m_act"], t)
    flips = (ca.reshape(-1) != cr.reshape(-1))
    per_img = flips.reshape(2, -1).sum(1).tolist() if flips.numel() % 2 == 0 else None
    ein = float((xi.reshape(-1) - xr.reshape(-1)).abs().max() / xr.abs().max())
    eout = float((yo.reshape(-1) - yr.reshape(-1)).abs().max() / yr.abs().max()) if yr is not None else float("nan")
    print(f"{name:24s} input dev {ein:.2e}  input-code flips per image {per_img}  output dev {eout:.2e}")
