"""Developer micro-benchmarks (not the contract bench - see bench.py): GEMM sweep in kernel mode and ViT engine step."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200 import ops
from quantized_vit_b200.engine import ViTInferenceEngine
from tests import fixtures


def timeit_graph(fn, iters=20, warm=3):
    """Device time per call with NO host launch latency inside: `iters` calls captured into one CUDA graph, replayed
    three times, best replay / iters.  (Event-bracketing a single eager call of a < 100 us kernel measures the Python +
    ctypes + tensor-map-encode latency of the wrapper, ~60 us, not the kernel.)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                fn()
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / iters)
    ts.sort()
    return ts[1], ts[0]


def timeit(fn, iters=20, warm=3, flush=None, graph=False):
    if graph:
        return timeit_graph(fn, iters=iters, warm=warm)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def gemm_sweep():
    flush = torch.empty(256 << 20, dtype=torch.int8, device="cuda")
    for (K, N) in [(768, 2304), (768, 3072), (3072, 768), (768, 768), (1024, 4096)]:
        for B in (16, 256):
            M = 197 * B
            a = torch.randint(-7, 8, (M, K), dtype=torch.int8, device="cuda")
            w = torch.randint(-7, 8, (N, K), dtype=torch.int8, device="cuda")
            bias = torch.randn(N, device="cuda")
            for kind, nm in ((ops.QVIT_OUT_NONE, "none"), (ops.QVIT_OUT_I8, "i8"), (ops.QVIT_OUT_BF16, "bf16"), (ops.QVIT_OUT_F32, "f32")):
                kw = dict(out_kind=kind, bias=bias, scale_a=0.1, scale_w=0.01)
                if kind == ops.QVIT_OUT_I8:
                    kw.update(next_q=(0.3, 2.1, None), act=ops.QVIT_ACT_GELU)
                if kind == ops.QVIT_OUT_NONE:
                    kw["backend"] = ops.QVIT_GEMM_TCGEN05
                else:
                    kw["out"] = ops.gemm_i8(a, w, K, N, **kw)
                med, best = timeit(lambda: ops.gemm_i8(a, w, K, N, **kw), flush=flush if M * K < (64 << 20) else None)
                tops = 2.0 * M * K * N / (med * 1e-3) / 1e12
                print(f"gemm M={M:6d} K={K:4d} N={N:4d} out={nm:4s} med {med*1e3:8.1f} us best {best*1e3:8.1f} us  {tops:8.1f} TOPS", flush=True)


def vit(batch=256, precision="fp32"):
    sd = fixtures.vit_state_dict(seed=0)
    from oracle import ref_geta
    # W4A4, activation ranges as initialised (q_m_act = max|W|): fixture A of SURVEY.md 8d
    names = [k[:-7] for k in sd if k.endswith(".weight") and sd[k].dim() >= 2]
    for n in names:
        d, qm = ref_geta.init_quant_params(sd[n + ".weight"], 4)
        sd[n + ".d_quant_wt"], sd[n + ".q_m_wt"] = d, qm
        sd[n + ".d_quant_act"], sd[n + ".q_m_act"] = d.clone(), qm.clone()
    eng = ViTInferenceEngine(sd, depth=12, num_heads=12, precision=precision)
    x = torch.randn(batch, 3, 224, 224, device="cuda")
    med, best = timeit(lambda: eng(x), iters=10)
    print(f"vit-b16 w4a4 {precision} eager  B={batch}: {med:.2f} ms  -> {batch/med*1e3:.0f} img/s", flush=True)
    xs, ys, g = eng.capture(batch)
    med, best = timeit(lambda: g.replay(), iters=10)
    print(f"vit-b16 w4a4 {precision} graph  B={batch}: {med:.2f} ms  -> {batch/med*1e3:.0f} img/s  "
          f"({eng.gemm_ops_per_image()*batch/med*1e-9:.0f} TOPS in quantized GEMMs)", flush=True)
    # per-kernel profile of one step
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng(x); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("gemm", "all"):
        gemm_sweep()
    if what in ("vit", "all"):
        vit(256, "fp32")
        vit(256, "bf16")
