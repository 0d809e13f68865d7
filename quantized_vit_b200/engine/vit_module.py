"""A compact nn.Module VisionTransformer with the reference's parameter names (cls_token, pos_embed,
patch_embed.proj, blocks.N.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}, norm, head) so that state_dicts are
interchangeable with ``QViT_with_GETA/vit_model.py`` and ``model_to_quantize_model`` swaps its Linear/Conv2d layers.
It is the CALLER of the hot path used for the QAT configuration (autograd through the drop-in modules); inference
throughput goes through ``ViTInferenceEngine`` instead.  Structure follows vit_model.py:46-328 (no dist token, no
pre-logits, dropout / drop-path 0 as in the reference defaults)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class _LayerNormFunction(torch.autograd.Function):
    """nn.LayerNorm forward / backward on the fused kernels of csrc/layernorm.cu (one pass each; the library backward runs
    three kernels and was 5 % of the QAT step)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        from .. import ops
        y, mean, rstd = ops.layernorm_fwd(x, weight, bias, eps)
        ctx.save_for_backward(x, weight, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, gy):
        from .. import ops
        x, weight, mean, rstd = ctx.saved_tensors
        gx, dg, db = ops.layernorm_bwd(x, gy, weight, mean, rstd)
        return gx, dg, db, None


class _ResidualLayerNormFunction(torch.autograd.Function):
    """(x, LayerNorm(x)): the block input handed through for the residual connection together with its normalised form, so that
    the backward sees BOTH gradients reaching x and sums them inside the LayerNorm backward kernel (no separate add kernel)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        from .. import ops
        y, mean, rstd = ops.layernorm_fwd(x, weight, bias, eps)
        ctx.save_for_backward(x, weight, mean, rstd)
        return x.view_as(x), y

    @staticmethod
    def backward(ctx, g_pass, gy):
        from .. import ops
        x, weight, mean, rstd = ctx.saved_tensors
        if gy is None:
            return g_pass, None, None, None
        gx, dg, db = ops.layernorm_bwd(x, gy, weight, mean, rstd, add=g_pass)
        return gx, dg, db, None


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm with the same parameters / state_dict; fp32 CUDA inputs whose width the kernels cover take the fused path."""

    def _fused_ok(self, x):
        from .. import ops
        return (x.is_cuda and x.dtype == torch.float32 and self.elementwise_affine and self.bias is not None
                and len(self.normalized_shape) == 1 and ops.layernorm_supported(self.normalized_shape[0]))

    def forward(self, x):
        if self._fused_ok(x):
            return _LayerNormFunction.apply(x, self.weight, self.bias, self.eps)
        return super().forward(x)

    def with_passthrough(self, x):
        """(x', LayerNorm(x)) with x' == x: use x' for the residual connection (see _ResidualLayerNormFunction)."""
        if self._fused_ok(x) and torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            return _ResidualLayerNormFunction.apply(x, self.weight, self.bias, self.eps)
        return x, self.forward(x)


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_c, embed_dim):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.proj = nn.Conv2d(in_c, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)          # vit_model.py:100


class _AttentionCoreFunction(torch.autograd.Function):
    """softmax(q k^T / sqrt(d)) v straight on the qkv layer's output (no permuted copies), forward and backward on the tcgen05
    kernels of csrc/attention_train.cu."""

    @staticmethod
    def forward(ctx, qkv, num_heads):
        from .. import ops
        out, lse = ops.attention_train_fwd(qkv, num_heads)
        ctx.save_for_backward(qkv, out, lse)
        ctx.num_heads = num_heads
        return out

    @staticmethod
    def backward(ctx, g):
        from .. import ops
        qkv, out, lse = ctx.saved_tensors
        return ops.attention_train_bwd(qkv, out, lse, g, ctx.num_heads), None


def _linear_plus(layer, x, residual):
    """residual + layer(x); inside the layer's GEMM epilogue when it is a QuantizeLinear."""
    if residual is None:
        return layer(x)
    if getattr(layer, "fuses_pre_act", False):
        return layer(x, residual=residual)
    return residual + layer(x)


class ViTAttention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x, residual=None):
        """residual: the block input; `residual + proj(...)` (vit_model.py:206) happens in proj's GEMM epilogue when it can."""
        B, N, C = x.shape
        from .. import ops
        qkv = self.qkv(x)
        if qkv.is_cuda and qkv.dtype == torch.float32 and ops.attention_train_supported(N, C // self.num_heads):
            o = _AttentionCoreFunction.apply(qkv, self.num_heads)
        else:
            qkv = qkv.reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
            o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])  # softmax(q k^T / sqrt(d)) v, vit_model.py:141-149
            o = o.transpose(1, 2).reshape(B, N, C)
        return _linear_plus(self.proj, o, residual)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)

    def forward(self, x, residual=None):
        h = self.fc1(x)
        if getattr(self.fc2, "fuses_pre_act", False) and getattr(self.act, "approximate", None) == "none":
            # QuantizeLinear: GELU fused into fc2's quantizer kernels (forward and backward), residual into its GEMM epilogue
            return self.fc2(h, pre_act="gelu", residual=residual)
        return _linear_plus(self.fc2, self.act(h), residual)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1, self.attn = LayerNorm(dim, eps=1e-6), ViTAttention(dim, num_heads)
        self.norm2, self.mlp = LayerNorm(dim, eps=1e-6), Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        # x + attn(norm1(x)); x + mlp(norm2(x))  (vit_model.py:206-207): the additions run in the proj / fc2 GEMM epilogues and
        # the two gradients that meet at each block input are summed inside the LayerNorm backward kernel
        xp, y = self.norm1.with_passthrough(x)
        x = self.attn(y, residual=xp)
        xp, y = self.norm2.with_passthrough(x)
        return self.mlp(y, residual=xp)


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_c=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0):
        super().__init__()
        self.num_classes, self.embed_dim = num_classes, embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_c, embed_dim)
        n = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Linear(embed_dim, num_classes)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        for m in self.modules():                                 # vit_model.py:330-346
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.01)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + self.pos_embed
        x = self.norm(self.blocks(x))
        return self.head(x[:, 0])
