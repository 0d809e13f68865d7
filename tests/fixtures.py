"""Seeded synthetic fixtures shared by CPU and GPU tests.  No reference access."""
import torch

from oracle import ref_models


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def vit_param_shapes(img, patch, dim, depth, heads, classes, mlp_ratio=4):
    n_tok = (img // patch) ** 2 + 1
    shapes = {"cls_token": (1, 1, dim), "pos_embed": (1, n_tok, dim),
              "patch_embed.proj.weight": (dim, 3, patch, patch), "patch_embed.proj.bias": (dim,)}
    hid = int(dim * mlp_ratio)
    for i in range(depth):
        p = f"blocks.{i}"
        shapes.update({f"{p}.norm1.weight": (dim,), f"{p}.norm1.bias": (dim,),
                       f"{p}.attn.qkv.weight": (3 * dim, dim), f"{p}.attn.qkv.bias": (3 * dim,),
                       f"{p}.attn.proj.weight": (dim, dim), f"{p}.attn.proj.bias": (dim,),
                       f"{p}.norm2.weight": (dim,), f"{p}.norm2.bias": (dim,),
                       f"{p}.mlp.fc1.weight": (hid, dim), f"{p}.mlp.fc1.bias": (hid,),
                       f"{p}.mlp.fc2.weight": (dim, hid), f"{p}.mlp.fc2.bias": (dim,)})
    shapes.update({"norm.weight": (dim,), "norm.bias": (dim,), "head.weight": (classes, dim), "head.bias": (classes,)})
    return shapes


def vit_state_dict(img=224, patch=16, dim=768, depth=12, heads=12, classes=1000, seed=0):
    sd = {k: torch.empty(s) for k, s in vit_param_shapes(img, patch, dim, depth, heads, classes).items()}
    return ref_models.fill_state_dict_(sd, seed=seed, weight_std=0.02)


def vit_input(batch, img=224, seed=1):
    return torch.randn(batch, 3, img, img, generator=_gen(seed))


_ULTRA_CONVS = [(0, 3, 16, 3), (4, 16, 32, 3), (8, 32, 64, 3), (12, 64, 64, 3), (16, 64, 64, 3), (19, 64, 64, 3),
                (22, 64, 64, 3), (25, 64, 64, 3), (28, 64, 36, 1)]


def ultranet_state_dict():
    """Same recipe as oracle/make_golden.py::golden_ultranet (seeded fill + BN affine ranges)."""
    sd = {}
    for idx, cin, cout, k in _ULTRA_CONVS:
        sd[f"layers.{idx}.weight"] = torch.empty(cout, cin, k, k)
        if idx == 28:
            sd[f"layers.{idx}.bias"] = torch.empty(cout)
        else:
            bn = idx + 1
            sd[f"layers.{bn}.weight"] = torch.empty(cout)
            sd[f"layers.{bn}.bias"] = torch.empty(cout)
            sd[f"layers.{bn}.running_mean"] = torch.empty(cout)
            sd[f"layers.{bn}.running_var"] = torch.empty(cout)
    ref_models.fill_state_dict_(sd, seed=11, weight_std=0.3)
    for k in list(sd.keys()):
        g = _gen(ref_models._seed_for(k, 12))
        if k.endswith(".weight") and sd[k].dim() == 1:
            sd[k].copy_(torch.rand(sd[k].shape, generator=g) * 0.2 + 0.2)
        if k.endswith(".bias") and sd[k].dim() == 1 and not k.startswith("layers.28"):
            sd[k].copy_(torch.rand(sd[k].shape, generator=g) * 0.4 + 0.3)
    return sd


def ultranet_input(batch=1, seed=1):
    x = torch.rand(batch, 3, 160, 320, generator=_gen(seed))
    return torch.round(x * 255) / 255
