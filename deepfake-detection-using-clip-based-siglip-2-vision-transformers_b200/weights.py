"""Seeded random weights in the reference's layouts (no pretrained checkpoints are reachable offline;
BASELINE.json asks for random-init weights of the named architecture).  Product-side helper for bench.py and
examples — tests use the oracle's own initialiser instead."""
from __future__ import annotations

import math
from typing import Dict

import torch

from .engine import VisionArch


def random_vision_state_dict(a: VisionArch, seed: int = 0, device="cpu", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """HF SiglipVisionModel key layout (without the `vision_model.` prefix); O(1) activations everywhere."""
    g = torch.Generator(device=device).manual_seed(seed)
    D, I, P, N = a.hidden_size, a.intermediate_size, a.patch_size, a.tokens

    def rn(*shape, std=1.0):
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * std).to(dtype)

    sd: Dict[str, torch.Tensor] = {}

    def lin(prefix, out_f, in_f, gain=1.0):
        sd[prefix + ".weight"] = rn(out_f, in_f, std=gain / math.sqrt(in_f))
        sd[prefix + ".bias"] = rn(out_f, std=0.1)

    def ln(prefix):
        sd[prefix + ".weight"] = 1.0 + rn(D, std=0.1)
        sd[prefix + ".bias"] = rn(D, std=0.1)

    sd["embeddings.patch_embedding.weight"] = rn(D, 3, P, P, std=1.0 / math.sqrt(3 * P * P))
    sd["embeddings.patch_embedding.bias"] = rn(D, std=0.1)
    sd["embeddings.position_embedding.weight"] = rn(N, D, std=0.5)
    for i in range(a.num_hidden_layers):
        p = f"encoder.layers.{i}"
        ln(p + ".layer_norm1")
        ln(p + ".layer_norm2")
        lin(p + ".self_attn.q_proj", D, D, 1.5)
        lin(p + ".self_attn.k_proj", D, D, 1.5)
        lin(p + ".self_attn.v_proj", D, D)
        lin(p + ".self_attn.out_proj", D, D, 0.5)
        lin(p + ".mlp.fc1", I, D)
        lin(p + ".mlp.fc2", D, I, 0.5)
    ln("post_layernorm")
    sd["head.probe"] = rn(1, 1, D)
    sd["head.attention.in_proj_weight"] = rn(3 * D, D, std=1.5 / math.sqrt(D))
    sd["head.attention.in_proj_bias"] = rn(3 * D, std=0.1)
    lin("head.attention.out_proj", D, D)
    ln("head.layernorm")
    lin("head.mlp.fc1", I, D)
    lin("head.mlp.fc2", D, I, 0.5)
    return sd


def random_classifier_head(kind: str, D: int, seed: int = 1) -> Dict[str, torch.Tensor]:
    """`BinaryClassifier` head state dict: kind 'A' (classifier.{0,2,5}) or 'B' (se.{0,2} + classifier.{0,2,5,7})."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std  # noqa: E731
    sd = {}
    if kind == "B":
        sd["se.0.weight"], sd["se.0.bias"] = rn(D // 16, D, std=2.0), rn(D // 16, std=0.1)
        sd["se.2.weight"], sd["se.2.bias"] = rn(D, D // 16, std=1.0 / math.sqrt(D // 16)), rn(D, std=0.1)
    sd["classifier.0.weight"], sd["classifier.0.bias"] = 1.0 + rn(D, std=0.1), rn(D, std=0.1)
    sd["classifier.2.weight"], sd["classifier.2.bias"] = rn(D // 2, D, std=1.0 / math.sqrt(D)), rn(D // 2, std=0.1)
    if kind == "A":
        sd["classifier.5.weight"], sd["classifier.5.bias"] = rn(1, D // 2, std=1.0 / math.sqrt(D // 2)), rn(1, std=0.1)
    else:
        sd["classifier.5.weight"], sd["classifier.5.bias"] = rn(D // 4, D // 2, std=1.0 / math.sqrt(D // 2)), rn(D // 4, std=0.1)
        sd["classifier.7.weight"], sd["classifier.7.bias"] = rn(1, D // 4, std=1.0 / math.sqrt(D // 4)), rn(1, std=0.1)
    return sd


def random_fast_classifier_head(model_size: str, D: int, seed: int = 4) -> Dict[str, torch.Tensor]:
    """`FastBinaryClassifier` head (H-D, cifake_binary_classifier.py:643-687) with the reference's key names: layer_norm,
    attention (qkv + proj for tiny / small, in_proj + out_proj for large, none for medium), size-dependent classifier."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std  # noqa: E731
    sd = {"layer_norm.weight": 1.0 + rn(D, std=0.1), "layer_norm.bias": rn(D, std=0.1)}
    att = {"tiny": ("qkv", "proj"), "small": ("qkv", "proj"), "large": ("in_proj_", "out_proj")}.get(model_size)
    if att:
        sep = "" if att[0].endswith("_") else "."
        sd[f"attention.{att[0]}{sep}weight"], sd[f"attention.{att[0]}{sep}bias"] = rn(3 * D, D, std=D ** -0.5), rn(3 * D, std=0.1)
        sd[f"attention.{att[1]}.weight"], sd[f"attention.{att[1]}.bias"] = rn(D, D, std=D ** -0.5), rn(D, std=0.1)
    dims, idx = {"tiny": ([D, 1], [1]), "small": ([D, D // 4, 1], [0, 3])}.get(model_size, ([D, D // 2, D // 4, 1], [0, 3, 6]))
    for i, n in enumerate(idx):
        sd[f"classifier.{n}.weight"], sd[f"classifier.{n}.bias"] = rn(dims[i + 1], dims[i], std=dims[i] ** -0.5), rn(dims[i + 1], std=0.1)
    return sd


def random_freq_mlp_g2(seed: int = 2) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std  # noqa: E731
    sd = {"normer.mean": rn(24, std=0.5), "normer.std": 0.5 + torch.rand(24, generator=g),
          "contrast.alpha": 1.0 + rn(24, std=0.2), "contrast.beta": rn(24, std=0.2), "band.gates": rn(4),
          "head.weight": rn(1, 24, std=0.4), "head.bias": rn(1, std=0.1), "temp.T": torch.tensor(1.3)}
    for b in range(2):
        sd[f"blocks.{b}.norm.weight"], sd[f"blocks.{b}.norm.bias"] = 1.0 + rn(24, std=0.1), rn(24, std=0.1)
        sd[f"blocks.{b}.fc1.weight"], sd[f"blocks.{b}.fc1.bias"] = rn(64, 24, std=0.3), rn(64, std=0.1)
        sd[f"blocks.{b}.fc2.weight"], sd[f"blocks.{b}.fc2.bias"] = rn(24, 64, std=0.2), rn(24, std=0.1)
    return sd


def random_fusion_g2(seed: int = 3) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std  # noqa: E731
    return {"mlp.0.weight": rn(32, 3, std=0.6), "mlp.0.bias": rn(32, std=0.3), "mlp.2.weight": rn(2, 32, std=0.4),
            "mlp.2.bias": rn(2, std=0.1), "temp.T": torch.tensor(0.9)}
