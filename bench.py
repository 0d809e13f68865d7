#!/usr/bin/env python
"""Benchmark of the hot path: ViT-B/16 W4A4 inference (BASELINE.json configs[1]) through the fused engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one forward pass of ViT-B/16 (quantized patch-embed + 48 QuantizeLinear + head, W4A4 symmetric-linear,
weight_and_activation) over one batch of 256 synthetic 224x224 images PER GPU (weak scaling: batch-sharded
replicas, no collective on the forward path - SURVEY.md section 8e).  `value` is whole-job images/s with the
batch resident in HBM (CUDA-graph replay, device-timed with CUDA events, max over ranks); `e2e` is the same metric
through the public host-facing call `ViTInferenceEngine.infer_many()` (pinned host batches in, host logits out) with the
host<->device copies inside the timed region.
`roofline` describes the dominant kernel (the tcgen05 int8 GEMM): algorithmic 2*M*K*N ops of the quantized layers
divided by the summed CUDA-event duration of the GEMM launches of a step, against the dense int8 peak MEASURED IN THIS RUN
(cuBLASLt / torch._int_mm at 8192^3, sustained loop; MEASURED_PEAKS.json holds no int8 entry - the 2 x bf16 and nominal
fractions are printed beside it).  `traffic` is read from the committed ncu capture (profiles/roofline_traffic.json).
`cpu_baseline` is the reference's own modules (verbatim copy in the git-ignored oracle/_ref, kind "reference"; the oracle
port when that copy did not travel, kind "port") timed on this host's cores on a bounded sub-batch.
`--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ViT-B/16 W4A4 inference throughput"
UNIT = "img/s"
BATCH = 256
IMG = 224
CFG = dict(embed_dim=768, depth=12, num_heads=12, patch=16, img=IMG, classes=1000)


def workload_config(n_gpus: int, extra=None):
    c = {"workload": "ViT-B/16 W4A4 (GETA symmetric-linear, weight_and_activation, num_bits=4; 50 quantized layers) "
                     "inference, 224x224 synthetic images, random-init weights, activation ranges as initialised "
                     "(q_m_act = max|W|)",
         "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus, "parallelism": f"batch-sharded replicas x{n_gpus}, no collective",
         "l2": "activations of one step (155 MB per [M,768] fp32 tensor) exceed the 126 MB L2; no explicit flush"}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_throughput(sub_batch: int, repeats: int, warmup: int = 1):
    """images/s of the oracle port of the reference's PyTorch-CPU forward (oracle/ref_models.vit_forward) on all host
    cores, on a sub-batch of the same synthetic workload."""
    import torch
    from oracle import ref_geta, ref_models
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    D, depth, hid, classes, patch = CFG["embed_dim"], CFG["depth"], 4 * CFG["embed_dim"], CFG["classes"], CFG["patch"]
    sd = {"cls_token": torch.randn(1, 1, D, generator=g) * 0.02, "pos_embed": torch.randn(1, (IMG // patch) ** 2 + 1, D, generator=g) * 0.02,
          "patch_embed.proj.weight": torch.randn(D, 3, patch, patch, generator=g) * 0.02, "patch_embed.proj.bias": torch.zeros(D),
          "norm.weight": torch.ones(D), "norm.bias": torch.zeros(D),
          "head.weight": torch.randn(classes, D, generator=g) * 0.02, "head.bias": torch.zeros(classes)}
    for i in range(depth):
        p = f"blocks.{i}"
        for n in ("norm1", "norm2"):
            sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"] = torch.ones(D), torch.zeros(D)
        for n, (o, k) in {"attn.qkv": (3 * D, D), "attn.proj": (D, D), "mlp.fc1": (hid, D), "mlp.fc2": (D, hid)}.items():
            sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"] = torch.randn(o, k, generator=g) * 0.02, torch.zeros(o)
    for name in [k[:-7] for k in list(sd) if k.endswith(".weight") and sd[k].dim() >= 2]:
        d, qm = ref_geta.init_quant_params(sd[name + ".weight"], 4)
        sd[name + ".d_quant_wt"], sd[name + ".q_m_wt"] = d, qm
        sd[name + ".d_quant_act"], sd[name + ".q_m_act"] = d.clone(), qm.clone()
    x = torch.randn(sub_batch, 3, IMG, IMG, generator=g)
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            ref_models.vit_forward(sd, x, depth, CFG["num_heads"], patch)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sub_batch / statistics.median(times), statistics.median(times), torch.get_num_threads()


REF_COPY = os.path.join(ROOT, "oracle", "_ref")


def reference_modules_throughput(sub_batch: int, repeats: int, warmup: int = 1):
    """images/s of the UNMODIFIED reference (vit_model.py + only_train_once/quantization, copied verbatim into the
    git-ignored oracle/_ref by oracle/make_ref.py) on all host cores: the factory train.py uses
    (vit_base_patch16_224_in21k(num_classes=1000, has_logits=False), train.py:21,233) converted with
    model_to_quantize_model(num_bits=4, symmetric+linear, weight_and_activation), eval, torch.no_grad()."""
    import torch
    from oracle import _refload as R
    R.use_root(REF_COPY)
    vm, qm = R.vit_model(), R.quant_model()
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = vm.vit_base_patch16_224_in21k(num_classes=CFG["classes"], has_logits=False)
    model = qm.model_to_quantize_model(model, num_bits=4, quant_type="symmetric+linear", quant_mode="weight_and_activation").eval()
    x = torch.randn(sub_batch, 3, IMG, IMG, generator=torch.Generator().manual_seed(1))
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            model(x)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sub_batch / statistics.median(times), statistics.median(times), torch.get_num_threads()


def cpu_arm(sub_batch: int, repeats: int, warmup: int = 1):
    """(img/s, s per forward, threads, kind): the reference's own modules when oracle/_ref travelled with the snapshot,
    else the oracle port (op-for-op restatement, bit-equal on the goldens)."""
    if os.path.isdir(os.path.join(REF_COPY, "QViT_with_GETA", "only_train_once", "quantization")):
        return (*reference_modules_throughput(sub_batch, repeats, warmup), "reference")
    return (*cpu_reference_throughput(sub_batch, repeats, warmup), "port")


def run_reference(args, rank: int):
    if rank != 0:
        return
    sub = 16
    t0 = time.perf_counter()
    ips, sec, threads, kind = cpu_arm(sub, repeats=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fake-quant values through F.linear/F.conv2d on the host CPU)", "data": "synthetic",
            "config": workload_config(args.gpus, {"cpu_sample": f"each step = one forward of a {sub}-image sub-batch"}),
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"{sub}-image sub-batch of the 256-image step, median of {max(1, args.steps)} forwards "
                                       f"({time.perf_counter() - t0:.0f} s of CPU work)"},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- measured int8 peak
def measure_int8_peak(dev, seconds: float = 1.5):
    """Dense int8 tensor-core throughput of THIS GPU right now: cuBLASLt (torch._int_mm) at 8192^3, as SURVEY.md 8(d) asks
    (MEASURED_PEAKS.json only holds bf16).  Returns (burst TOP/s = best single launch of 10, sustained TOP/s = a
    back-to-back loop of `seconds` under the power cap).  The GEMMs of a step are timed inside a long step, so the
    sustained figure is the roofline denominator; the burst one is printed beside it."""
    import torch
    n = 8192
    a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
    ops_per = 2.0 * n ** 3
    try:
        for _ in range(3):
            torch._int_mm(a, b)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch._int_mm(a, b)
        e1.record()
        torch.cuda.synchronize()
        sust = e0.elapsed_time(e1) / reps
        return ops_per / (best * 1e-3) / 1e12, ops_per / (sust * 1e-3) / 1e12
    except Exception as exc:  # noqa: BLE001
        sys.stderr.write(f"bench.py: int8 peak measurement unavailable ({exc})\n")
        return None, None


def profiled_traffic():
    """dram read+write bytes per launch of the dominant kernel from the committed ncu capture (profiles/roofline_traffic.json,
    written by tools/ncu_summary.py from an `ncu --set full` report)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return float(t["dram_bytes_per_launch"]), t.get("note", "")
    except Exception:
        return None, "no committed ncu capture"


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from quantized_vit_b200 import _lib
    from quantized_vit_b200.engine import ViTInferenceEngine
    from quantized_vit_b200.engine.synthetic import vit_state_dict
    _lib.lib()          # fail loudly here if libqvit_b200.so is missing

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sd = vit_state_dict(**CFG, num_bits=4, seed=0, device=dev)
    eng = ViTInferenceEngine(sd, depth=CFG["depth"], num_heads=CFG["num_heads"], patch_size=CFG["patch"], device=dev,
                             precision="fp32")
    g = torch.Generator(device="cpu").manual_seed(1 + rank)
    x_host = torch.randn(BATCH, 3, IMG, IMG, generator=g).pin_memory()
    xs, ys, graph = eng.capture(BATCH, IMG)
    xs.copy_(x_host)
    # launches of OUR kernels per step (C-ABI calls of one eager forward; the graph replays exactly these)
    c0 = _lib.CALLS
    eng.forward(xs)
    calls_per_step = _lib.CALLS - c0
    for _ in range(max(args.warmup, 3)):
        graph.replay()
    if os.environ.get("QVIT_NCU_RANGE"):
        # `ncu --profile-from-start off`: expose exactly ONE eager step (same kernels the graph replays) to the profiler
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        eng.forward(xs)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        graph.replay()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)

    # end to end through the public host-facing API
    # (two distinct pinned batches alternate so that no copy can be elided)
    x_host2 = torch.randn(BATCH, 3, IMG, IMG, generator=g).pin_memory()
    host_batches = [x_host if (i & 1) == 0 else x_host2 for i in range(args.steps)]
    eng.infer_many(host_batches)          # warm-up: staging buffers, graph, pinned result pool (untimed)
    ref_single = eng.infer(x_host).clone()
    barrier()
    t0 = time.perf_counter()
    logits_all = eng.infer_many(host_batches)
    barrier()
    e2e_s = time.perf_counter() - t0
    logits_host = logits_all[0]
    assert torch.equal(logits_host, ref_single), "pipelined and single-call inference disagree"
    flags = int(eng.flags.item())

    # per-kernel timing of the dominant kernel (eager pass, events on the launching stream)
    eng.gemm_events = []
    for _ in range(2):
        eng.gemm_events.clear()
        try:
            torch.cuda._sleep(20_000_000)  # ~10 ms of device-side spin: the host enqueues the whole eager step meanwhile, so
        except Exception:                  # no launch gap can leak into an event pair (each brackets exactly one kernel)
            pass
        eng.forward(xs)
    torch.cuda.synchronize()
    gemm_ms = sum(a.elapsed_time(b) for _, _, a, b in eng.gemm_events)
    gemm_ops = sum(o for _, o, _, _ in eng.gemm_events)
    n_gemm = len(eng.gemm_events)
    eng.gemm_events = None
    gemm_top_per_image = eng.gemm_ops_per_image(IMG) / 1e12

    int8_burst, int8_sust = measure_int8_peak(dev) if rank == 0 else (None, None)

    # secondary workloads (rank 0 of a single-GPU run only; short graph-replay timings, reported beside the headline):
    # fixture B of SURVEY.md 8d (calibrated activation ranges: the fc1 epilogue's exact-redo rate depends on the data) and
    # BASELINE config 4 (ViT-L/16 W4A8, 128 images per GPU)
    extras = None
    if world == 1 and not args.no_extras:
        extras = {}

        def graph_ms(engine, batch, steps=8):
            xs_, ys_, g_ = engine.capture(batch, IMG)
            xs_.copy_(torch.randn(batch, 3, IMG, IMG, generator=torch.Generator().manual_seed(3)).to(dev))
            for _ in range(3):
                g_.replay()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                g_.replay()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / steps
        del eng
        torch.cuda.empty_cache()
        sd_b = vit_state_dict(**CFG, num_bits=4, calibrate_to=2.5, seed=0, device=dev)
        eng_b = ViTInferenceEngine(sd_b, depth=CFG["depth"], num_heads=CFG["num_heads"], patch_size=CFG["patch"], device=dev)
        t_b = graph_ms(eng_b, BATCH)
        extras["vit_b16_w4a4_calibrated_ranges"] = {"ms_per_step": t_b, "img_per_s": BATCH / t_b * 1e3, "batch": BATCH,
                                                    "note": "fixture B: q_m_act = 2.5 for every layer (activations spread over all 15 codes)"}
        del eng_b, sd_b
        torch.cuda.empty_cache()
        cfg_l = dict(embed_dim=1024, depth=24, num_heads=16, patch=16, img=IMG, classes=1000)
        sd_l = vit_state_dict(**cfg_l, num_bits=4, act_bits=8, calibrate_to=3.0, seed=0, device=dev)
        eng_l = ViTInferenceEngine(sd_l, depth=24, num_heads=16, patch_size=16, device=dev)
        t_l = graph_ms(eng_l, 128, steps=5)
        extras["vit_l16_w4a8_batch128"] = {"ms_per_step": t_l, "img_per_s": 128 / t_l * 1e3, "batch": 128,
                                           "gemm_tops": eng_l.gemm_ops_per_image(IMG) * 128 / (t_l * 1e-3) / 1e12,
                                           "note": "BASELINE config 4 per-GPU share (batch 1024 = 128 x 8 GPUs)", "flags": int(eng_l.flags.item())}
        del eng_l, sd_l
        torch.cuda.empty_cache()

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_sust = peaks.get("bf16_tflops_sustained")
    derived = 2.0 * bf16_sust if bf16_sust else 2.0 * 1400.0
    peak_tops = int8_sust if int8_sust else derived
    traffic, traffic_note = profiled_traffic()
    achieved = gemm_ops / (gemm_ms * 1e-3) / 1e12
    value = world * BATCH * args.steps / (ms * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        tc = time.perf_counter()
        ips, sec, threads, kind = cpu_arm(32, repeats=12, warmup=1)
        cpu = {"value": ips, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"32-image sub-batch of the 256-image step, median of 12 forwards ({time.perf_counter() - tc:.0f} s of CPU work)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 codes x int8 codes -> int32 (tcgen05 kind::i8); fp32 epilogues, LayerNorm, residual stream and attention",
            "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": world * BATCH * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": logits_host.numel() * 4,
                    "api": "ViTInferenceEngine.infer_many(pinned host batches) -> host logits; H2D of step i+1 and D2H of "
                           "step i-1 overlap the forward of step i (all copies inside the timed region)"},
            "gpu_launches": calls_per_step * args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tops, "unit": "TOP/s", "frac": achieved / peak_tops,
                         "traffic": traffic, "traffic_note": traffic_note,
                         "frac_alt": {"vs_cublaslt_int8_burst": (achieved / int8_burst) if int8_burst else None,
                                      "vs_2x_bf16_sustained_measured_peaks": achieved / derived, "vs_nominal_4500": achieved / 4500.0},
                         "peak_alt": {"cublaslt_int8_8192_sustained": int8_sust, "cublaslt_int8_8192_burst": int8_burst,
                                      "2x_bf16_sustained_measured_peaks": derived, "nominal": 4500.0},
                         "kernel": "gemm_i8_tc_kernel", "launches_per_step": n_gemm,
                         "kernel_ms_per_step": gemm_ms, "kernel_share_of_step": gemm_ms / (ms / args.steps),
                         "peak_source": ("dense int8 measured in this run: cuBLASLt (torch._int_mm) 8192^3, back-to-back loop of 1.5 s "
                                         "(sustained; MEASURED_PEAKS.json has no int8 entry)" if int8_sust else
                                         "2 x bf16_tflops_sustained of MEASURED_PEAKS.json (derived; int8 measurement unavailable)")},
            "cpu_baseline": cpu, "clocks": clocks, "quantizer_flags": flags,
            "gemm_top_per_image": gemm_top_per_image, "other_configs": extras}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (fixture B, ViT-L/16 W4A8)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one process per GPU)")
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
