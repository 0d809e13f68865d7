import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ref_models
from tests import fixtures
from quantized_vit_b200.engine import ViTInferenceEngine
torch.set_num_threads(os.cpu_count())
for name in ("vit_b16_w4a4_init", "vit_b16_w4a4_calib"):
    g = np.load(f"tests/golden/{name}.npz")
    img, patch, dim, depth, heads, classes = [int(v) for v in g["cfg"]]
    sd = fixtures.vit_state_dict(img, patch, dim, depth, heads, classes, seed=int(g["fill_seed"]))
    for k, v in zip(g["q.names"], g["q.values"]):
        sd[str(k)] = torch.tensor([float(v)])
    x = fixtures.vit_input(int(g["batch"]), img)
    t_ref = {}
    ref = ref_models.vit_forward(sd, x, depth, heads, patch, taps=t_ref)
    print(name, "oracle==golden", bool(np.array_equal(ref.numpy(), g["logits"])), "margin", float(g["margin"].min()), "max|logit|", float(np.abs(g["logits"]).max()))
    for prec, att in (("fp32", "math"), ("fp32", "sdpa"), ("bf16", "sdpa")):
        t_eng = {}
        eng = ViTInferenceEngine(sd, depth=depth, num_heads=heads, patch_size=patch, precision=prec, attention=att)
        out = eng.forward(x.cuda(), taps=t_eng).cpu()
        errs = []
        for k in ("embed", "blocks.0.attn.proj.in", "blocks.0.out", "blocks.5.out", "blocks.11.out"):
            a, b = t_eng[k].cpu().double(), t_ref[k].double()
            errs.append(f"{k.replace('blocks.','b')}={float((a-b).abs().max()/b.abs().max()):.1e}")
        rel = float((out.double()-ref.double()).abs().max()/ref.abs().max())
        print(f"  {prec}/{att}: logits rel {rel:.3e} top1 same {bool((out.argmax(-1)==ref.argmax(-1)).all())} | " + " ".join(errs), flush=True)
