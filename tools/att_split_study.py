"""CPU study (numpy, no GPU): attention core on real ViT-B qkv values (oracle taps of the golden fixtures) under
  (a) the current exact 3 x bf16 split with 6 product terms, and
  (b) a 2 x fp16 split (hi = fp16(x * S), lo = fp16(x * S - hi), S a power of two from the static bound) with 3 terms,
both with fp32 accumulation (emulated by float32 matmuls), against float64.  Reports max-norm error and the number of
`proj`-input 4-bit codes that differ from the reference's codes."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_geta, ref_models
from tests import fixtures

def split_bf16(x):
    t = torch.from_numpy(x)
    p1 = t.to(torch.bfloat16).float(); r = t - p1
    p2 = r.to(torch.bfloat16).float(); r = r - p2
    p3 = r.to(torch.bfloat16).float()
    return [p.numpy() for p in (p1, p2, p3)]

def split_f16(x):
    t = torch.from_numpy(x)
    hi = t.to(torch.float16).float()
    lo = (t - hi).to(torch.float16).float()
    return hi.numpy(), lo.numpy()

def att(q, k, v, mode, scale):
    # q,k,v: [T, 64] float32
    mm = lambda a, b: (a.astype(np.float32) @ b.astype(np.float32))
    if mode == "f64":
        s = (q.astype(np.float64) @ k.astype(np.float64).T) * scale
        p = np.exp(s - s.max(1, keepdims=True)); return (p @ v.astype(np.float64)) / p.sum(1, keepdims=True)
    if mode == "bf16x3":
        Q, K, V = split_bf16(q), split_bf16(k), split_bf16(v)
        terms = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]
        s = sum(mm(Q[a], K[b].T) for a, b in terms)
    else:
        sq, sk = 2.0 ** mode_scales["q"], 2.0 ** mode_scales["k"]
        Qh, Ql = split_f16(q * np.float32(sq)); Kh, Kl = split_f16(k * np.float32(sk))
        s = (mm(Qh, Kh.T) + (mm(Qh, Kl.T) + mm(Ql, Kh.T))) * np.float32(1.0 / (sq * sk))
    s = s.astype(np.float32)
    mx = s.max(1, keepdims=True)
    p = np.exp2((s - mx) * np.float32(scale * 1.4426950408889634)).astype(np.float32)
    rs = p.sum(1, keepdims=True, dtype=np.float32)
    if mode == "bf16x3":
        P = split_bf16(p)
        o = mm(P[0], V[0]) + mm(P[1], V[0]) + mm(P[2], V[0]) + mm(P[0], V[1]) + mm(P[1], V[1]) + mm(P[0], V[2])
    else:
        sv = 2.0 ** mode_scales["v"]
        Ph, Pl = split_f16(p * np.float32(1024.0)); Vh, Vl = split_f16(v * np.float32(sv))
        o = (mm(Ph, Vh) + (mm(Ph, Vl) + mm(Pl, Vh))) * np.float32(1.0 / (1024.0 * sv))
    return o / rs

mode_scales = {}
def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "vit_b16_w4a4_calib"
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name + ".npz"))
    img, patch, dim, depth, heads, classes = [int(v) for v in g["cfg"]]
    sd = fixtures.vit_state_dict(img, patch, dim, depth, heads, classes, seed=int(g["fill_seed"]))
    for k, v in zip(g["q.names"], g["q.values"]):
        sd[str(k)] = torch.tensor([float(v)], dtype=torch.float32)
    x = fixtures.vit_input(int(g["batch"]), img, seed=1)
    taps = {"__layers__": True}
    torch.set_num_threads(os.cpu_count())
    ref_models.vit_forward(sd, x, depth, heads, patch, taps=taps)
    for blk in (0, depth // 2, depth - 1):
        p = f"blocks.{blk}"
        qkv = taps[f"{p}.attn.qkv.y"].numpy()          # [B, T, 3D]
        o_ref = taps[f"{p}.attn.proj.in"]
        B, T, _ = qkv.shape
        # static bound-based scales: |y| <= d_a d_w sat_a sum_k |w_code| + |b|, scale = 2^(15 - ceil(log2 bound)) per part
        w = sd[f"{p}.attn.qkv.weight"]; b = sd[f"{p}.attn.qkv.bias"]
        wc = ref_geta.sym_codes(w, sd[f"{p}.attn.qkv.d_quant_wt"], sd[f"{p}.attn.qkv.q_m_wt"]).abs().sum(1).float()
        sat_a = ref_geta.saturation_code(sd[f"{p}.attn.qkv.d_quant_act"], sd[f"{p}.attn.qkv.q_m_act"])
        bound = sd[f"{p}.attn.qkv.d_quant_act"].abs() * sd[f"{p}.attn.qkv.d_quant_wt"].abs() * sat_a * wc + b.abs()
        for i, part in enumerate("qkv"):
            bb = float(bound[i * dim:(i + 1) * dim].max())
            mode_scales[part] = 15 - int(np.ceil(np.log2(bb)))
        amax = [float(np.abs(qkv[..., i * dim:(i + 1) * dim]).max()) for i in range(3)]
        want = ref_geta.sym_codes(o_ref.reshape(-1, dim), sd[f"{p}.attn.proj.d_quant_act"], sd[f"{p}.attn.proj.q_m_act"])
        res = {}
        for mode in ("f64", "bf16x3", "f16x2"):
            o = np.zeros((B, T, dim))
            for bi in range(B):
                for h in range(heads):
                    q, k, v = (qkv[bi, :, i * dim + h * 64: i * dim + (h + 1) * 64] for i in range(3))
                    o[bi, :, h * 64:(h + 1) * 64] = att(q, k, v, mode, 64 ** -0.5)
            res[mode] = o
        r64 = res["f64"]
        line = f"{name} {p}: bound-scales 2^{mode_scales} amax {['%.2f' % a for a in amax]}"
        for mode in ("bf16x3", "f16x2"):
            err = np.abs(res[mode] - r64).max() / np.abs(r64).max()
            codes = ref_geta.sym_codes(torch.from_numpy(res[mode].astype(np.float32)).reshape(-1, dim), sd[f"{p}.attn.proj.d_quant_act"], sd[f"{p}.attn.proj.q_m_act"])
            line += f" | {mode}: err {err:.2e} flips {int((codes != want).sum())}"
        e_ref = np.abs(o_ref.numpy() - r64).max() / np.abs(r64).max()
        c64 = ref_geta.sym_codes(torch.from_numpy(r64.astype(np.float32)).reshape(-1, dim), sd[f"{p}.attn.proj.d_quant_act"], sd[f"{p}.attn.proj.q_m_act"])
        line += f" | reference(fp32 CPU) vs f64: err {e_ref:.2e}; f64 flips {int((c64 != want).sum())} of {want.numel()}"
        print(line, flush=True)

if __name__ == "__main__":
    main()
