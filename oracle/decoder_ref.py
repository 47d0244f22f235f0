"""ORACLE (test infrastructure, not product code) — fp32 torch-CPU restatement of SegFormerStrongDecoder and the
SigLIP2_MTL heads (Siglip2sidafrozen.py:698-742,773-803), evaluation mode.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned by tests/golden/decoder_golden.npz, which oracle/make_golden.py produced by instantiating the reference's OWN
classes (extracted from the source text) with seeded weights.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


def init_decoder_state(in_dim: int, K: int, embed_dim: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded weights in the reference's key layout (decoder.* and cls_head.*)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std: torch.randn(*s, generator=g) * std
    E = embed_dim
    sd = {"cls_head.weight": rn(3, in_dim, std=in_dim ** -0.5), "cls_head.bias": rn(3, std=0.1)}
    for i in range(K):
        sd[f"decoder.projs.{i}.proj.weight"] = rn(E, in_dim, std=in_dim ** -0.5)
        sd[f"decoder.projs.{i}.proj.bias"] = rn(E, std=0.1)
        sd[f"decoder.smooth.{i}.0.weight"] = rn(E, 1, 3, 3, std=1.0 / 3.0)
        sd[f"decoder.smooth.{i}.0.bias"] = rn(E, std=0.1)
        sd[f"decoder.smooth.{i}.1.weight"] = rn(E, E, 1, 1, std=E ** -0.5)
        sd[f"decoder.smooth.{i}.1.bias"] = rn(E, std=0.1)
    sd["decoder.fuse_attn.0.weight"] = rn(E * K // 4, E * K, 1, 1, std=(E * K) ** -0.5)
    sd["decoder.fuse_attn.0.bias"] = rn(E * K // 4, std=0.1)
    sd["decoder.fuse_attn.2.weight"] = rn(E * K, E * K // 4, 1, 1, std=(E * K // 4) ** -0.5)
    sd["decoder.fuse_attn.2.bias"] = rn(E * K, std=0.1)
    sd["decoder.fuse.0.weight"] = rn(E, E * K, 1, 1, std=(E * K) ** -0.5)
    sd["decoder.fuse.0.bias"] = rn(E, std=0.1)
    sd["decoder.head.weight"] = rn(1, E, 1, 1, std=E ** -0.5)
    sd["decoder.head.bias"] = rn(1, std=0.1)
    return sd


def decoder_forward(sd: Dict[str, torch.Tensor], hidden_list: Sequence[torch.Tensor], grid: int, target_size: int) -> torch.Tensor:
    """hidden_list: K tensors [B, N, C] fp32 -> [B, 1, S, S]   (Siglip2sidafrozen.py:726-742)."""
    feats: List[torch.Tensor] = []
    for i, h in enumerate(hidden_list):
        x = F.linear(h, sd[f"decoder.projs.{i}.proj.weight"], sd[f"decoder.projs.{i}.proj.bias"]).transpose(1, 2)
        B, E, N = x.shape
        x = x.reshape(B, E, grid, grid)
        x = F.conv2d(x, sd[f"decoder.smooth.{i}.0.weight"], sd[f"decoder.smooth.{i}.0.bias"], padding=1, groups=E)
        x = F.gelu(F.conv2d(x, sd[f"decoder.smooth.{i}.1.weight"], sd[f"decoder.smooth.{i}.1.bias"]))
        feats.append(x)
    x = torch.cat(feats, dim=1)
    a = F.gelu(F.conv2d(x, sd["decoder.fuse_attn.0.weight"], sd["decoder.fuse_attn.0.bias"]))
    x = torch.sigmoid(F.conv2d(a, sd["decoder.fuse_attn.2.weight"], sd["decoder.fuse_attn.2.bias"])) * x
    x = F.conv2d(x, sd["decoder.fuse.0.weight"], sd["decoder.fuse.0.bias"])
    x = F.interpolate(x, size=(target_size, target_size), mode="bilinear", align_corners=False)
    return F.conv2d(x, sd["decoder.head.weight"], sd["decoder.head.bias"])


def cls_head(sd: Dict[str, torch.Tensor], pooled: torch.Tensor) -> torch.Tensor:
    return F.linear(pooled, sd["cls_head.weight"], sd["cls_head.bias"])
