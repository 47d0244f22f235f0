"""DetectionPipeline — the batched form of the reference's `detect_core` (deepfake-detector-v2/app.py:1329-1412)
and of the extraction loops of train_fusion_head_only.py:329-347:

    images (u8 NHWC) ──SigLIP engine──> pooled ──classifier head──> z_sig ┐
    gray256 (f32)    ──freq feature kernels──> 24-d features ─────────────┴─ score epilogue ──> z, CORAL, p_blend

One call = (7·L + 13) backbone launches + head + 2 freq kernels + 1 score-epilogue kernel, all enqueue-only on the
current stream; `detect()` adds the pinned H2D copies and one D2H read of the packed scores, which is what the
reference loops do per batch (inference_ai_human_images.py:274,296).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .engine import ARCHS, SiglipEngine, VisionArch
from .scoring import EPS_TRAINER, FreqFeatureExtractor, ScoringStack

PACKED_FIELDS = ("z_sig", "z_freq", "z", "z_scaled", "p_raw", "p_coral", "entropy", "p_blend", "risk_idx",
                 "risk_p0", "risk_p1", "risk_p2", "risk_p3", "risk_p4")


def head_params_from_state(sd: Dict[str, torch.Tensor], dim: int, device) -> "ops.HeadParams":
    """`BinaryClassifier` heads: H-A keys classifier.{0,2,5} (inference_ai_human_images.py:131-138,
    no eps on the norm) or H-B keys se.{0,2} + classifier.{0,2,5,7} (train_fusion_head_only.py:84-99, +1e-6)."""
    t = {"ln_g": sd["classifier.0.weight"], "ln_b": sd["classifier.0.bias"], "w1": sd["classifier.2.weight"],
         "b1": sd["classifier.2.bias"], "w2": sd["classifier.5.weight"], "b2": sd["classifier.5.bias"]}
    if "se.0.weight" in sd:
        t.update({"se_w1": sd["se.0.weight"], "se_b1": sd["se.0.bias"], "se_w2": sd["se.2.weight"],
                  "se_b2": sd["se.2.bias"], "w3": sd["classifier.7.weight"], "b3": sd["classifier.7.bias"]})
        return ops.HeadParams(2, dim, 1e-6, t, device)
    return ops.HeadParams(1, dim, 0.0, t, device)


class DetectionPipeline:
    def __init__(self, arch: VisionArch | str, backbone_state: Dict[str, torch.Tensor],
                 head_state: Dict[str, torch.Tensor], scoring: ScoringStack, device: int = 0, max_batch: int = 64,
                 freq_eps: float = EPS_TRAINER, freq_zscore: Optional[bool] = None, fuse_ln: bool = False):
        self.arch = ARCHS[arch] if isinstance(arch, str) else arch
        self.device = torch.device("cuda", device)
        self.engine = SiglipEngine(self.arch, device, max_batch, fuse_ln=fuse_ln).load_state_dict(backbone_state)
        self.head = head_params_from_state(head_state, self.arch.hidden_size, self.device)
        self.scoring = scoring
        # G1 heads were trained on z-scored vectors (app.py:840-846), G2 on raw ones + learned normaliser
        zs = (scoring.gen == 1) if freq_zscore is None else freq_zscore
        self.freq = FreqFeatureExtractor(self.device, eps=freq_eps, zscore=zs)
        self._pin: Dict[str, torch.Tensor] = {}

    # ---- device-resident path ---------------------------------------------------------------------
    def detect_device(self, images: torch.Tensor, gray256: torch.Tensor, resize_mode: int = 0) -> Dict[str, torch.Tensor]:
        pooled, _ = self.engine(images, resize_mode=resize_mode)
        _, z_sig, _ = ops.head_fwd(self.head, pooled)
        feats = self.freq.from_gray(gray256)
        out = self.scoring(z_sig, feats=feats)
        out["pooled"] = pooled
        return out

    @staticmethod
    def pack(out: Dict[str, torch.Tensor]) -> torch.Tensor:
        """[B, 14] fp32 score records (PACKED_FIELDS) — the unit that is all-gathered across ranks."""
        cols = [out[k].float() for k in PACKED_FIELDS[:8]] + [out["risk_idx"].float()]
        return torch.cat([torch.stack(cols, 1), out["risk_probs"]], 1)

    # ---- host-buffer path (what a caller of the reference loops sees) -----------------------------------
    def detect(self, images_host: torch.Tensor, gray256_host: torch.Tensor, resize_mode: int = 0) -> np.ndarray:
        """Host (ideally pinned) u8 NHWC images + f32 gray256 -> numpy [B,14] score records."""
        img = images_host.to(self.device, non_blocking=True)
        gray = gray256_host.to(self.device, non_blocking=True)
        packed = self.pack(self.detect_device(img, gray, resize_mode))
        key = f"out{packed.shape[0]}"
        if key not in self._pin:
            self._pin[key] = torch.empty(packed.shape, dtype=torch.float32).pin_memory()
        self._pin[key].copy_(packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._pin[key].numpy()
