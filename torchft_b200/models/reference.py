"""Plain-PyTorch fp32 reference of the Llama forward/backward (test oracle only).

Numerics tests compare every hand-written kernel path in ``models/llama.py``
against this eager implementation, which shares nothing with it except the
parameter tensors.
"""

from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F


def _rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def _rope(x: torch.Tensor, theta: float) -> torch.Tensor:
    B, S, H, D = x.shape
    inv = 1.0 / (theta ** (torch.arange(0, D, 2, device=x.device, dtype=torch.float32) / D))
    ang = torch.outer(torch.arange(S, device=x.device, dtype=torch.float32), inv)
    f = torch.polar(torch.ones_like(ang), ang)
    xc = torch.view_as_complex(x.reshape(B, S, H, D // 2, 2))
    return torch.view_as_real(xc * f[None, :, None, :]).reshape(B, S, H, D)


def reference_loss(model, tokens: torch.Tensor, targets: torch.Tensor) -> Tuple[float, Dict[str, torch.Tensor]]:
    """Returns (loss, {param_name: fp32 grad}) computed with eager fp32 PyTorch."""

    cfg = model.cfg
    P = {n: p.detach().float().requires_grad_() for n, p in model.named_parameters()}
    B, S = tokens.shape
    x = F.embedding(tokens, P["tok_embeddings"])
    Hq, Hkv, D = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
    for i in range(cfg.n_layers):
        pre = f"layers.{i}."
        n = _rmsnorm(x, P[pre + "attention_norm"], cfg.norm_eps)
        qkv = (n @ P[pre + "wqkv"].t()).view(B, S, Hq + 2 * Hkv, D)
        q, k, v = qkv[:, :, :Hq], qkv[:, :, Hq : Hq + Hkv], qkv[:, :, Hq + Hkv :]
        q, k = _rope(q, cfg.rope_theta), _rope(k, cfg.rope_theta)
        rep = Hq // Hkv
        k = k.repeat_interleave(rep, dim=2)
        v = v.repeat_interleave(rep, dim=2)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal=True)
        o = o.transpose(1, 2).reshape(B, S, cfg.dim)
        h = x + o @ P[pre + "wo"].t()
        n2 = _rmsnorm(h, P[pre + "ffn_norm"], cfg.norm_eps)
        gu = n2 @ P[pre + "w13"].t()
        a = F.silu(gu[..., : cfg.ffn_dim]) * gu[..., cfg.ffn_dim :]
        x = h + a @ P[pre + "w2"].t()
    hfin = _rmsnorm(x, P["norm"], cfg.norm_eps)
    logits = hfin @ P["output"].t()
    loss = F.cross_entropy(logits.view(B * S, -1), targets.reshape(-1))
    loss.backward()
    return float(loss.item()), {n: p.grad for n, p in P.items()}
