"""Developer tool: clock64 timeline of CTA 0 of the tensor-core GEMM (mode + 50 of qvit_gemm_set_cta_group)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops, _lib

M, K, N = 50432, int(os.environ.get("K", 768)), int(os.environ.get("N", 3072))
a = torch.randint(-7, 8, (M, K), dtype=torch.int8, device="cuda")
w = torch.randint(-7, 8, (N, K), dtype=torch.int8, device="cuda")
bias = torch.randn(N, device="cuda")
cases = {"none": dict(out_kind=ops.QVIT_OUT_NONE, backend=ops.QVIT_GEMM_TCGEN05),
         "i32": dict(out_kind=ops.QVIT_OUT_I32),
         "bf16": dict(out_kind=ops.QVIT_OUT_BF16, bias=bias),
         "f32": dict(out_kind=ops.QVIT_OUT_F32, bias=bias),
         "i8+relu": dict(out_kind=ops.QVIT_OUT_I8, bias=bias, act=ops.QVIT_ACT_RELU, next_q=(0.3, 2.1, None)),
         "i8": dict(out_kind=ops.QVIT_OUT_I8, bias=bias, next_q=(0.3, 2.1, None)),
         "i8+gelu": dict(out_kind=ops.QVIT_OUT_I8, bias=bias, act=ops.QVIT_ACT_GELU, next_q=(0.3, 2.1, None))}
L = _lib.lib()
for mode in [int(v) for v in sys.argv[1:]] or [51]:
    for name in os.environ.get("EPI_CASES", "none,i8").split(","):
        kw = dict(cases[name], scale_a=0.1, scale_w=0.01, acc_abs_max=49 * K)
        L.qvit_gemm_set_cta_group(mode)
        for _ in range(2):
            ops.gemm_i8(a, w, K, N, **kw)
        torch.cuda.synchronize()
        reps = int(os.environ.get("REPS", 1))
        for _ in range(reps - 1):
            ops.gemm_i8(a, w, K, N, **kw)
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 996)()
        n = L.qvit_gemm_read_profile(buf, 996)
        print(f"    redo chunks so far (all CTAs, cumulative): {buf[516 + 3 * 159]} of {(M // 128) * (N // 32) * 4} warp-chunks per launch")
        ctas = [(buf[516 + 3 * i], buf[517 + 3 * i], buf[518 + 3 * i]) for i in range(148)]
        s0 = min(c[0] for c in ctas)
        durs = sorted((c[1] - c[0]) / 1e3 for c in ctas)
        print(f"    all CTAs: start spread {(max(c[0] for c in ctas) - s0) / 1e3:.1f} us, end max {(max(c[1] for c in ctas) - s0) / 1e3:.1f} us, "
              f"duration min/median/max {durs[0]:.1f}/{durs[74]:.1f}/{durs[-1]:.1f} us, distinct SMs {len(set(c[2] for c in ctas))}")
        late = [(i, (c[0] - s0) / 1e3, (c[1] - c[0]) / 1e3, c[2]) for i, c in enumerate(ctas) if (c[0] - s0) > 5000]
        print(f"    CTAs starting > 5 us late: {late[:12]}")
        slow = sorted(((c[1] - c[0]) / 1e3, i, c[2]) for i, c in enumerate(ctas))[-6:]
        print(f"    slowest CTAs (us, cta, smid): {slow}")
        c0, g0, c1, g1 = buf[512], buf[513], buf[514], buf[515]
        print(f"    CTA 0 of the last launch: {c1 - c0} clks in {(g1 - g0) / 1e3:.1f} us -> SM clock {(c1 - c0) / max(g1 - g0, 1) * 1e3:.0f} MHz (after {reps} back-to-back launches)")
        t = [buf[i] for i in range(n)]
        t0 = t[0]
        print(f"--- mode {mode} {name}: per tile [mma: acc free, issued | epi: full, ld0, math0, staged0, ld_last, done] (clks from start)")
        for i in range(6, 10):
            r = [v - t0 for v in t[8 * i: 8 * i + 8]]
            print(f"tile {i:2d}: " + " ".join(f"{v:8d}" for v in r))
L.qvit_gemm_set_cta_group(0)
