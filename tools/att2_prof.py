"""Timeline of qvit_attention_f16x2 (CTA 0): per tile, compute-warp stamps [start, S ready, softmax done, epilogue(t-1) done]
and the control thread's stamp after issuing P V of tile t; plus graph-timed kernel duration at B = 256."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops
from tools.quick_bench import timeit
B, T, H = 256, 197, 12
qkv = torch.randn(B, T, 3 * H * 64, device="cuda")
D = H * 64
exps = [ops.f16x2_exponent(float(qkv[..., i * D:(i + 1) * D].abs().max())) for i in range(3)]
col_exp = torch.cat([torch.full((D,), e, dtype=torch.int32) for e in exps]).cuda()
planes = ops.split2_f16(qkv.reshape(B * T, 3 * D), col_exp)
d, qm = torch.tensor([0.05], device="cuda"), torch.tensor([0.35], device="cuda")
prof = torch.zeros(256, dtype=torch.int64, device="cuda")
ops.attention_f16x2(planes, B, T, H, exps, d, qm, None, prof=prof)
torch.cuda.synchronize()
p = prof.cpu().tolist()
t0 = p[0]
print("tile: start  S_ready  softmax_done  epi_done | control: PV issued   (cycles since kernel start of CTA 0)")
for t in range(12):
    print(f"{t:3d}: {p[4*t]-t0:7d} {p[4*t+1]-t0:7d} {p[4*t+2]-t0:7d} {p[4*t+3]-t0:7d} | {p[64+t]-t0:7d}   "
          f"softmax {p[4*t+2]-p[4*t+1]:5d}  epi {p[4*t+3]-p[4*t+2]:5d}  wait_S {p[4*t+1]-p[4*t]:5d} | "
          f"ld {p[128+4*t]-p[4*t+1]:5d} max+bar {p[129+4*t]-p[128+4*t]:5d} exp+split+st {p[130+4*t]-p[129+4*t]:5d} bar2 {p[131+4*t]-p[130+4*t]:5d} | epi(t-1): wait_O {(p[192+4*(t-1)]-p[4*t+2]) if t else 0:5d} ldO {(p[193+4*(t-1)]-p[192+4*(t-1)]) if t else 0:5d} rest {(p[4*t+3]-p[193+4*(t-1)]) if t else 0:5d}")
med, best = timeit(lambda: ops.attention_f16x2(planes, B, T, H, exps, d, qm, None), iters=10, graph=True)
print(f"attention_f16x2 B={B}: {med*1e3:.1f} us (best {best*1e3:.1f})")
