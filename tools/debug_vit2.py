import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from oracle import ref_models
from tests import fixtures
from quantized_vit_b200.engine import ViTInferenceEngine
torch.set_num_threads(os.cpu_count())
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

def fq(x, d, qm):
    a = x.abs()
    out = d * torch.round(a / d)
    out = torch.where(a <= 0, torch.zeros_like(out), out)
    out = torch.where(a >= qm, (d * torch.round(qm.abs() / d)).expand_as(out), out)
    return torch.sign(x) * out

def torch_gpu_forward(sd, x, depth, heads, patch, taps):
    sd = {k: v.cuda() for k, v in sd.items()}
    x = x.cuda()
    def ql(p, y):
        return F.linear(fq(y, sd[p + ".d_quant_act"], sd[p + ".q_m_act"]), fq(sd[p + ".weight"], sd[p + ".d_quant_wt"], sd[p + ".q_m_wt"]), sd[p + ".bias"])
    p = "patch_embed.proj"
    h = F.conv2d(fq(x, sd[p + ".d_quant_act"], sd[p + ".q_m_act"]), fq(sd[p + ".weight"], sd[p + ".d_quant_wt"], sd[p + ".q_m_wt"]), sd[p + ".bias"], stride=patch)
    h = h.flatten(2).transpose(1, 2)
    B = h.shape[0]
    h = torch.cat((sd["cls_token"].expand(B, -1, -1), h), 1) + sd["pos_embed"]
    D = h.shape[-1]
    taps["embed"] = h
    for i in range(depth):
        p = f"blocks.{i}"
        y = F.layer_norm(h, (D,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-6)
        qkv = ql(f"{p}.attn.qkv", y)
        N = qkv.shape[1]
        qkv = qkv.reshape(B, N, 3, heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        a = ((q @ k.transpose(-2, -1)) * (q.shape[-1] ** -0.5)).softmax(-1)
        y = (a @ v).transpose(1, 2).reshape(B, N, -1)
        taps[f"{p}.attn.proj.in"] = y
        h = h + ql(f"{p}.attn.proj", y)
        y = F.layer_norm(h, (D,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-6)
        y = F.gelu(ql(f"{p}.mlp.fc1", y))
        h = h + ql(f"{p}.mlp.fc2", y)
        taps[f"{p}.out"] = h
    h = F.layer_norm(h, (D,), sd["norm.weight"], sd["norm.bias"], 1e-6)
    return ql("head", h[:, 0])

for name in ("vit_b16_w4a4_calib", "vit_b16_w4a4_init"):
    g = np.load(f"tests/golden/{name}.npz")
    img, patch, dim, depth, heads, classes = [int(v) for v in g["cfg"]]
    sd = fixtures.vit_state_dict(img, patch, dim, depth, heads, classes, seed=int(g["fill_seed"]))
    for k, v in zip(g["q.names"], g["q.values"]):
        sd[str(k)] = torch.tensor([float(v)])
    x = fixtures.vit_input(int(g["batch"]), img)
    t_ref, t_gpu, t_eng = {}, {}, {}
    ref = ref_models.vit_forward(sd, x, depth, heads, patch, taps=t_ref)
    with torch.no_grad():
        gpu = torch_gpu_forward(sd, x, depth, heads, patch, t_gpu).cpu()
    eng = ViTInferenceEngine(sd, depth=depth, num_heads=heads, patch_size=patch, precision="fp32", attention="math")
    out = eng.forward(x.cuda(), taps=t_eng).cpu()
    def rel(a, b):
        a, b = a.cpu().double(), b.cpu().double()
        return float((a - b).abs().max() / b.abs().max())
    print(name)
    for k in t_ref:
        print(f"  {k:26s} torchGPU-vs-CPUref {rel(t_gpu[k], t_ref[k]):.2e}   engine-vs-CPUref {rel(t_eng[k], t_ref[k]):.2e}   engine-vs-torchGPU {rel(t_eng[k], t_gpu[k]):.2e}")
    print(f"  logits: torchGPU-vs-CPUref {rel(gpu, ref):.2e} engine-vs-CPUref {rel(out, ref):.2e} engine-vs-torchGPU {rel(out, gpu):.2e}; top1 ref {ref.argmax(-1).tolist()} gpu {gpu.argmax(-1).tolist()} eng {out.argmax(-1).tolist()}", flush=True)
