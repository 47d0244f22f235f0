"""SigLIP2_MTL inference (Siglip2sidafrozen.py:747-803): SigLIP-2 vision encoder with per-layer
hidden-state taps + a 3-class head on the pooled embedding + the SegFormer-style tamper-mask decoder (:698-742).

    cls_logit [B,3], seg_logits [B,1,S,S] = model(pixel_values)

Every Linear / 1x1 convolution is a tcgen05 GEMM on token-major [B·N, C] bf16 matrices (erf-GELU, sigmoid and the
`fuse_attn(x) * x` gate run in the GEMM epilogue, the four branch outputs are written straight into their column slice
of the concatenated [B·N, 4E] matrix), the depthwise 3x3 convolution and the head + bilinear resize are small streaming
kernels (csrc/decoder.cu).  Evaluation mode only (dropout off).  Other square input sizes (the reference calls the
encoder with `interpolate_pos_encoding=True`, Siglip2sidafrozen.py:787, for its progressive-resize schedule :975-987) run
on a per-grid engine whose position table is the bicubic resample of the trained one (HF:modeling_siglip.py:137-173).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch

from . import ops
from .engine import ARCHS, SiglipEngine, VisionArch, canonicalize_state_dict


class SegFormerStrongDecoder:
    """SegFormerStrongDecoder (Siglip2sidafrozen.py:698-742) on token-major bf16 activations; evaluation mode."""

    def __init__(self, num_inputs: int, embed_dim: int = 256, device=None):
        self.K, self.embed_dim = num_inputs, embed_dim
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._w: Dict[str, torch.Tensor] = {}

    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        """Keys {projs.i.proj, smooth.i.0, smooth.i.1, fuse_attn.0, fuse_attn.2, fuse.0, head}.{weight,bias}."""
        E, K, dev = self.embed_dim, self.K, self.device
        bf = lambda t: t.detach().to(dev, torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
        g = lambda k: sd[prefix + k]
        w = {}
        for i in range(K):
            w[f"proj{i}.w"], w[f"proj{i}.b"] = bf(g(f"projs.{i}.proj.weight")), f32(g(f"projs.{i}.proj.bias"))
            w[f"dw{i}.w"], w[f"dw{i}.b"] = f32(g(f"smooth.{i}.0.weight").reshape(E, 9)), f32(g(f"smooth.{i}.0.bias"))
            w[f"pw{i}.w"], w[f"pw{i}.b"] = bf(g(f"smooth.{i}.1.weight").reshape(E, E)), f32(g(f"smooth.{i}.1.bias"))
        w["att0.w"], w["att0.b"] = bf(g("fuse_attn.0.weight").reshape(E * K // 4, E * K)), f32(g("fuse_attn.0.bias"))
        w["att2.w"], w["att2.b"] = bf(g("fuse_attn.2.weight").reshape(E * K, E * K // 4)), f32(g("fuse_attn.2.bias"))
        w["fuse.w"], w["fuse.b"] = bf(g("fuse.0.weight").reshape(E, E * K)), f32(g("fuse.0.bias"))
        w["head.w"], w["head.b"] = f32(g("head.weight").reshape(E)), float(g("head.bias").reshape(()))
        self._w = w
        return self

    @torch.no_grad()
    def forward(self, hidden_list: Sequence[torch.Tensor], B: int, grid: int, target_size: int) -> torch.Tensor:
        """K tensors [B·grid², C] (bf16) -> seg logits [B,1,S,S] fp32."""
        w, E, K = self._w, self.embed_dim, self.K
        assert len(hidden_list) == K
        M = B * grid * grid
        cat = torch.empty((M, E * K), dtype=torch.bfloat16, device=self.device)
        for i, h in enumerate(hidden_list):
            p = ops.gemm_bf16(h, w[f"proj{i}.w"], bias=w[f"proj{i}.b"])                       # LinearProj
            d = ops.dwconv3x3_bf16(p, w[f"dw{i}.w"], w[f"dw{i}.b"], B, grid, grid)            # depthwise 3x3
            ops.gemm_bf16(d, w[f"pw{i}.w"], bias=w[f"pw{i}.b"], act=2, out=cat[:, i * E:(i + 1) * E])  # 1x1 + GELU
        a = ops.gemm_bf16(cat, w["att0.w"], bias=w["att0.b"], act=2)
        gated = ops.gemm_bf16(a, w["att2.w"], bias=w["att2.b"], act=3, residual=cat, residual_op=1)   # sigmoid(.) * x
        f = ops.gemm_bf16(gated, w["fuse.w"], bias=w["fuse.b"])
        return ops.seg_head_upsample(f, w["head.w"], w["head.b"], B, grid, grid, target_size)

    __call__ = forward


class SigLIP2_MTL:
    def __init__(self, arch: VisionArch | str = "siglip2-base-patch16-224", device: int = 0, max_batch: int = 8,
                 seg_layers: Sequence[int] = (2, 6, 10, -1), embed_dim: int = 256):
        self.arch = ARCHS[arch] if isinstance(arch, str) else arch
        self.device = torch.device("cuda", device)
        self.max_batch = max_batch
        self.engine = SiglipEngine(self.arch, device, max_batch)
        self._engines = {self.arch.grid: self.engine}
        self._enc_sd: Dict[str, torch.Tensor] = {}
        self.seg_layers, self.embed_dim = tuple(seg_layers), embed_dim
        self.decoder = SegFormerStrongDecoder(len(self.seg_layers), embed_dim, self.device)
        self._w: Dict[str, torch.Tensor] = {}
        g = int(math.isqrt(self.arch.tokens))
        if g * g != self.arch.tokens:
            raise ValueError(f"Cannot reshape {self.arch.tokens} tokens into square grid.")
        self.grid = g

    def eval(self):
        return self

    def to(self, device):
        return self

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        """Keys as written by the reference's trainer: encoder.vision_model.*, cls_head.{weight,bias} (or cls_head.1.* when
        it was built with dropout), decoder.*"""
        enc = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
        self._enc_sd = canonicalize_state_dict(enc)
        self.engine.load_state_dict(self._enc_sd)
        for g in [g for g in self._engines if g != self.arch.grid]:   # engines of other grids hold the old weights
            self._engines.pop(g).close()
        ck = "cls_head.1" if "cls_head.1.weight" in sd else "cls_head"
        self._w = {"cls.w": sd[f"{ck}.weight"].detach().to(self.device, torch.float32).contiguous(),
                   "cls.b": sd[f"{ck}.bias"].detach().to(self.device, torch.float32).contiguous()}
        self.decoder.load_state_dict(sd, prefix="decoder.")
        return self

    def _engine_for(self, S: int) -> SiglipEngine:
        """Engine for S x S inputs: the native one, or one per other patch grid with an interpolated position table."""
        from .dropin import interpolated_position_table

        a = self.arch
        g = S // a.patch_size
        if g not in self._engines:
            arch = VisionArch(g * a.patch_size, a.patch_size, a.hidden_size, a.intermediate_size, a.num_hidden_layers,
                              a.num_attention_heads, a.layer_norm_eps)
            sd = dict(self._enc_sd)
            sd["embeddings.position_embedding.weight"] = interpolated_position_table(
                self._enc_sd["embeddings.position_embedding.weight"], g)
            mb = max(1, min(self.max_batch, self.max_batch * a.tokens // (g * g)))
            self._engines[g] = SiglipEngine(arch, self.device.index, mb).load_state_dict(sd)
        return self._engines[g]

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor):
        x = pixel_values.to(self.device)
        B = x.shape[0]
        u8 = x.dtype == torch.uint8
        Hh, Ww = (int(x.shape[1]), int(x.shape[2])) if u8 else (int(x.shape[-2]), int(x.shape[-1]))
        if Hh != Ww:
            raise ValueError("Cannot reshape tokens into square grid. Try using square input images.")  # Siglip2sidafrozen.py:799-801
        S = Hh
        eng = self._engine_for(S)
        N, grid = eng.arch.tokens, eng.arch.grid
        cls, seg = [], []
        for b0 in range(0, B, eng.max_batch):
            xb = x[b0:b0 + eng.max_batch]
            nb = xb.shape[0]
            pooled, _, hidden = eng.forward_hidden(xb)
            cls.append(ops.linear_small(pooled, self._w["cls.w"], self._w["cls.b"]))
            last = hidden.shape[0] - 1
            idxs = [(i + 1 if i >= 0 else last) for i in self.seg_layers]
            feats = [hidden[i].reshape(nb * N, self.arch.hidden_size) for i in idxs]
            seg.append(self.decoder(feats, nb, grid, S))
        return torch.cat(cls), torch.cat(seg)

    __call__ = forward
