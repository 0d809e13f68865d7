"""Config 3: ViT-B/16 4-bit QAT forward+backward through the drop-in modules, gradient + step-size
all-reduce over NCCL.  Single process or torchrun; global batch 128 split evenly.  Prints one JSON line (rank 0)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from quantized_vit_b200 import parallel
from quantized_vit_b200.engine.vit_module import VisionTransformer
from quantized_vit_b200.quantization import model_to_quantize_model, check_nan_flags


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    qtype = sys.argv[1] if len(sys.argv) > 1 else "symmetric+linear"
    global_batch, steps = 128, int(sys.argv[2]) if len(sys.argv) > 2 else 5
    if len(sys.argv) > 3:                    # optional: bf16 planes of the gradient operand (3 = exact, 2 = 16 significant bits)
        from quantized_vit_b200.quantization import quant_layers
        quant_layers.GRADIENT_PLANES = int(sys.argv[3])
    if os.environ.get("QVIT_CG"):            # A/B switch: 1 = single-CTA tiles everywhere (qvit_gemm_set_cta_group)
        from quantized_vit_b200 import _lib
        _lib.lib().qvit_gemm_set_cta_group(int(os.environ["QVIT_CG"]))
    torch.manual_seed(0)
    model = VisionTransformer(num_classes=1000)
    model = model_to_quantize_model(model, num_bits=4, quant_type=qtype, quant_mode="weight_and_activation").cuda().train()
    with torch.no_grad():                    # calibrated-ish activation ranges so that gradients are informative
        for m in model.modules():
            if hasattr(m, "q_m_act"):
                m.q_m_act.fill_(2.5); m.d_quant_act.fill_(2.5 / 7)
    use_graph = os.environ.get("QAT_GRAPH", "0") in ("1", "2")
    # graph mode: forward + backward are replayed from ONE CUDA graph and the bucket all-reduces are issued after it (no NCCL
    # call inside the capture: hook-launched collectives under stream capture hung the 2-rank run), so the hooks stay off
    # (QAT_GRAPH=2: the bucket hooks stay on and record external events inside the capture; the all-reduces are launched after
    # graph.replay() on a side stream behind those events, overlapping the rest of the replayed backward)
    red = parallel.GradientAllReducer(model.named_parameters(), overlap=(not use_graph) or os.environ.get("QAT_GRAPH") == "2")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(global_batch, 3, 224, 224, generator=g)
    y = torch.randint(0, 1000, (global_batch,), generator=torch.Generator().manual_seed(2))
    xs, ys = parallel.shard_batch(x, rank, world).cuda(), parallel.shard_batch(y, rank, world).cuda()
    crit = torch.nn.CrossEntropyLoss()

    def step():
        red.zero_grad()                      # persistent bucket views: no flatten / unflatten copies
        loss = crit(model(xs), ys) / 1.0
        loss.backward()                      # bucket all-reduces are launched from the backward hooks
        red.reduce()
        red.clip_(1.0)
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if use_graph:
        # at 16-32 images per GPU the eager step is bound by ~1500 kernel launches, not by the kernels
        def fwd_bwd():
            red.zero_grad()
            loss = crit(model(xs), ys) / 1.0
            loss.backward()
            return loss
        gs = torch.cuda.Stream()
        gs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(gs):
            fwd_bwd()
        torch.cuda.current_stream().wait_stream(gs)
        red.reduce()                                  # drain what the warm-up's hooks launched
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = fwd_bwd()

        def step():                                   # noqa: F811
            graph.replay()
            red.reduce()
            red.clip_(1.0)
            return static_loss
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    flags = check_nan_flags(raise_error=False)
    gn = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in model.parameters() if p.grad is not None))
    dq = model.blocks[0].attn.qkv.d_quant_act.grad.item()
    if rank == 0:
        print(json.dumps({"config": "ViT-B/16 4-bit QAT fwd+bwd+allreduce+clip", "quant_type": qtype, "n_gpus": world,
                          "global_batch": global_batch, "ms_per_step": float(ms), "img_per_s": global_batch / float(ms) * 1e3,
                          "loss": float(loss), "grad_norm_after_clip": float(gn), "grad_d_quant_act_block0_qkv": dq,
                          "nan_flags": flags, "cuda_graph": int(os.environ.get("QAT_GRAPH", "0")),
                          "gradient_planes": __import__("quantized_vit_b200.quantization.quant_layers", fromlist=["x"]).GRADIENT_PLANES}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
