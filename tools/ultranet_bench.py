"""Config 1: UltraNet W4A4 batch-1 latency (CUDA-graph replay, median of 200) + batch throughput,
beside the oracle port of the reference CPU path on this host."""
import json, os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantized_vit_b200.engine import UltraNetEngine
from quantized_vit_b200.engine.synthetic import ultranet_state_dict

sd = ultranet_state_dict()
out = {}
for conv in ("tc", "simt"):
    eng = UltraNetEngine(sd, input_bits=8, conv=conv)
    for B in (1, 64, 256):
        x = torch.round(torch.rand(B, 3, 160, 320, device="cuda") * 255) / 255
        xs, ys, g = eng.capture(B)
        xs.copy_(x)
        for _ in range(10):
            g.replay()
        ts = []
        for _ in range(200 if B == 1 else 20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        med = statistics.median(ts)
        out[f"{conv}_batch{B}"] = {"latency_us": med * 1e3, "img_per_s": B / med * 1e3, "tops": B * 2 * UltraNetEngine.macs_per_image() / med / 1e9}
if "--layers" in sys.argv:      # per-layer device time at batch 256 (eager, CUDA events around each launch)
    from quantized_vit_b200 import ops
    eng = UltraNetEngine(sd, input_bits=8, conv="tc")
    x = torch.round(torch.rand(256, 3, 160, 320, device="cuda") * 255) / 255
    taps = []
    eng(x, taps=taps)
    acc_scale = 1.0 / (15 * 7)
    for i in range(1, 9):
        L = eng.layers[i]
        prev = taps[i - 1]
        last = i == 8
        for name, fn in (("tc", lambda: eng._layer(L, prev, acc_scale, f32_out=last)),
                         ("simt", lambda: ops.ultra_conv_bn_act(prev, L["codes_ohwi"], L["pad"], acc_scale, None if last else L["scale"], L["bias"], 15,
                                                                False if last else L["pool"], f32_out=last))):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            macs = 256 * L["C"] * L["O"] * L["kh"] * L["kw"] * prev.shape[1] * prev.shape[2]
            out[f"layer{i}_{name}_b256"] = {"us": statistics.median(ts) * 1e3, "tops": 2 * macs / statistics.median(ts) / 1e9}
from oracle import ref_models
torch.set_num_threads(os.cpu_count())
sdc = {k: v.cpu() for k, v in sd.items()}
xc = torch.round(torch.rand(1, 3, 160, 320) * 255) / 255
with torch.no_grad():
    ref_models.ultranet_features(sdc, xc)
    t = []
    for _ in range(10):
        t0 = time.perf_counter(); ref_models.ultranet_features(sdc, xc); t.append(time.perf_counter() - t0)
out["cpu_port_batch1"] = {"latency_us": statistics.median(t) * 1e6, "img_per_s": 1 / statistics.median(t), "cores": torch.get_num_threads()}
out["gop_per_image"] = 2 * UltraNetEngine.macs_per_image() / 1e9
print(json.dumps(out))
