"""Cycle timeline of the attention backward kernel (csrc/attention_train.cu): CTA 0, first 8 tiles, control thread and compute
thread 0.  Also times forward (split + kernel) and backward with CUDA events at the QAT step's shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops

B, T, H = int(os.environ.get("B", 128)), 197, 12
qkv = torch.randn(B, T, 3 * H * 64, device="cuda")
g = torch.randn(B, T, H * 64, device="cuda") * 1e-3
out, lse = ops.attention_train_fwd(qkv, H)
prof = torch.zeros(256, dtype=torch.int64, device="cuda")
ops.attention_train_bwd(qkv, out, lse, g, H, prof=prof)
torch.cuda.synchronize()
p = prof.cpu().view(8, 32)
base = int(p[0, 16])
print("tile | control: wait_ops  issue_sdp01 | per chunk: (wait_pds, issue_out, wait_mma) ... commit | compute: load  per chunk: (wait_sdp, ld+bar, math+st) ... wait_out epi")
for t in range(8):
    c, k = p[t, :16].tolist(), p[t, 16:].tolist()
    ctl = f"{c[1]-c[0]:6d} {c[2]-c[1]:5d} |"
    prev = c[2]
    for ch in range(4):
        a, b_, d = c[3 + ch * 3], c[4 + ch * 3], (c[5 + ch * 3] if ch < 2 else 0)
        ctl += f" ({a-prev:5d},{b_-a:5d},{(d-b_) if d else 0:5d})"
        prev = d if d else b_
    ctl += f" {c[15]-prev:4d}"
    cmp_ = f"{k[1]-k[0]:6d} |"
    prev = k[1]
    for ch in range(4):
        a, b_, d = k[2 + ch * 3], k[3 + ch * 3], k[4 + ch * 3]
        cmp_ += f" ({a-prev:5d},{b_-a:5d},{d-b_:5d})"
        prev = d
    cmp_ += f" {k[14]-prev:5d} {k[15]-k[14]:5d}"
    print(f"{t} @{k[0]-base:7d} total {k[15]-k[0]:6d} | {ctl} | {cmp_}")

def timeit(f, n=10):
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

print(f"forward (split + kernel) {timeit(lambda: ops.attention_train_fwd(qkv, H)):.1f} us, backward (dstat + kernel) "
      f"{timeit(lambda: ops.attention_train_bwd(qkv, out, lse, g, H)):.1f} us  at B={B}")
