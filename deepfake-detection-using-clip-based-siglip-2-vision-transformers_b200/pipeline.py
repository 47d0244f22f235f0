"""DetectionPipeline — the batched form of the reference's `detect_core` (deepfake-detector-v2/app.py:1329-1412)
and of the extraction loops of train_fusion_head_only.py:329-347:

    images (u8 NHWC) ──SigLIP engine──> pooled ──classifier head──> z_sig ┐
       └─ gray256 kernels (luma, CLAHE, bicubic 256²) ─ freq feature kernels ─> 24-d features ─┴─ score epilogue ──> z, CORAL
    (gray256 may also be handed in, e.g. when the caller's images are not at the model resolution)

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
    """`BinaryClassifier` / `FastBinaryClassifier` heads, recognised by their key names:
      H-A  classifier.{0,2,5}                 inference_ai_human_images.py:131-138 (no eps on the L2 norm)
      H-B  se.{0,2} + classifier.{0,2,5,7}    train_fusion_head_only.py:84-99 (norm + 1e-6)
      H-D  layer_norm + [attention.*] + classifier.{0,3[,6]} or {1}   cifake_binary_classifier.py:643-684,728-749
           (attention over ONE token is exactly proj(v(x)): softmax of a single score is 1)"""
    G, N = ops.ACT_GELU, ops.ACT_NONE
    if "layer_norm.weight" in sd:  # H-D
        layers = []
        if "attention.qkv.weight" in sd:        # LightweightAttention: rows [2D,3D) of the fused qkv are v
            layers += [(sd["attention.qkv.weight"][2 * dim:], sd["attention.qkv.bias"][2 * dim:], N),
                       (sd["attention.proj.weight"], sd["attention.proj.bias"], N)]
        elif "attention.in_proj_weight" in sd:  # nn.MultiheadAttention
            layers += [(sd["attention.in_proj_weight"][2 * dim:], sd["attention.in_proj_bias"][2 * dim:], N),
                       (sd["attention.out_proj.weight"], sd["attention.out_proj.bias"], N)]
        idx = sorted(int(k.split(".")[1]) for k in sd if k.startswith("classifier.") and k.endswith(".weight"))
        for n, i in enumerate(idx):
            layers.append((sd[f"classifier.{i}.weight"], sd[f"classifier.{i}.bias"], G if n + 1 < len(idx) else N))
        return ops.HeadParams(1, dim, 0.0, device, ln=(sd["layer_norm.weight"], sd["layer_norm.bias"]), layers=layers)
    ln = (sd["classifier.0.weight"], sd["classifier.0.bias"])
    if "se.0.weight" in sd:
        se = (sd["se.0.weight"], sd["se.0.bias"], sd["se.2.weight"], sd["se.2.bias"])
        layers = [(sd["classifier.2.weight"], sd["classifier.2.bias"], G), (sd["classifier.5.weight"], sd["classifier.5.bias"], G),
                  (sd["classifier.7.weight"], sd["classifier.7.bias"], N)]
        return ops.HeadParams(2, dim, 1e-6, device, ln=ln, se=se, layers=layers)
    layers = [(sd["classifier.2.weight"], sd["classifier.2.bias"], G), (sd["classifier.5.weight"], sd["classifier.5.bias"], N)]
    return ops.HeadParams(1, dim, 0.0, device, ln=ln, layers=layers)


class DetectionPipeline:
    def __init__(self, arch: VisionArch | str, backbone_state: Dict[str, torch.Tensor],
                 head_state: Dict[str, torch.Tensor], scoring: ScoringStack, device: int = 0, max_batch: int = 64,
                 freq_eps: float = EPS_TRAINER, freq_zscore: Optional[bool] = None, fuse_ln: bool = True):
        self.arch = ARCHS[arch] if isinstance(arch, str) else arch
        self.device = torch.device("cuda", device)
        self.engine = SiglipEngine(self.arch, device, max_batch, fuse_ln=fuse_ln).load_state_dict(backbone_state)
        self.head = head_params_from_state(head_state, self.arch.hidden_size, self.device)
        self.scoring = scoring
        # G1 heads were trained on z-scored vectors (app.py:840-846), G2 on raw ones + learned normaliser
        zs = (scoring.gen == 1) if freq_zscore is None else freq_zscore
        self.freq = FreqFeatureExtractor(self.device, eps=freq_eps, zscore=zs)
        self._pin: Dict[str, torch.Tensor] = {}
        self._gray_scratch: Optional[torch.Tensor] = None

    # ---- device-resident path ---------------------------------------------------------------------
    def detect_device(self, images: torch.Tensor, gray256: Optional[torch.Tensor] = None, resize_mode: int = 0,
                      clahe: bool = True, pil_resize: Optional[str] = None) -> Dict[str, torch.Tensor]:
        """images: u8 NHWC on the device.  gray256 None = derive it from the same (original-size) pixels on the device
        (train_fusion_head_only.py:142-148 with clahe=True; app.py:736-749 with DETECT_USE_CLAHE for clahe).
        pil_resize "bilinear" / "bicubic": images of another size first go through the PIL-exact `Resize((S,S))` of the
        reference's transforms (inference_ai_human_images.py:200-204) instead of the in-model resize (resize_mode)."""
        S = self.arch.image_size
        model_in = images
        if pil_resize is not None and tuple(images.shape[1:3]) != (S, S):
            model_in = ops.resize_u8(images, S, S, pil_resize)
        pooled, _ = self.engine(model_in, resize_mode=resize_mode)
        _, z_sig, _ = ops.head_fwd(self.head, pooled)
        if gray256 is None:
            need = ops._lib.load().dfd_gray256_scratch_bytes(*images.shape[:3])
            if self._gray_scratch is None or self._gray_scratch.numel() < need:
                self._gray_scratch = torch.empty((need,), dtype=torch.uint8, device=self.device)
            gray256 = ops.gray256_from_rgb(images, clahe, scratch=self._gray_scratch)
        feats = self.freq.from_gray(gray256)
        out = self.scoring(z_sig, feats=feats)
        out["pooled"] = pooled
        return out

    @staticmethod
    def pack(out: Dict[str, torch.Tensor]) -> torch.Tensor:
        """[B, 14] fp32 score records (PACKED_FIELDS) — the unit that is all-gathered across ranks."""
        cols = [out[k].float() for k in PACKED_FIELDS[:8]] + [out["risk_idx"].float()]
        return torch.cat([torch.stack(cols, 1), out["risk_probs"]], 1)

    # ---- detect_core: multicrop, one batched call (deepfake-detector-v2/app.py:1329-1412, 1418-1430) ---------------
    MULTICROP_WEIGHTS = (0.4, 0.4, 0.05, 0.05, 0.05, 0.05)

    def make_multicrops(self, pil):
        """The reference's 6 views: full image, bicubic S x S resize, four quadrants; weights .4/.4/.05x4."""
        from PIL import Image

        S = self.arch.image_size
        w, h = pil.size
        w2, h2 = max(1, w // 2), max(1, h // 2)
        return [pil, pil.resize((S, S), Image.BICUBIC), pil.crop((0, 0, w2, h2)), pil.crop((w2, 0, w, h2)),
                pil.crop((0, h2, w2, h)), pil.crop((w2, h2, w, h))]

    @torch.no_grad()
    def detect_core(self, pils, multicrop: bool = True, clahe: bool = False, preprocess=None, freq_temp: float = 1.25):
        """Batched `detect_core`: every crop of every image goes through ONE backbone / feature batch, crop logits
        are combined with the reference's weights (on logits, app.py:1346-1347), then one score epilogue per image.
        Returns a list of dicts with the reference's keys."""
        import numpy as np

        from .dropin import make_preprocess
        from .scoring import pil_to_gray256

        pre = preprocess or make_preprocess(self.arch.image_size, "bilinear")
        views, wts = [], []
        for pil in pils:
            pil = pil.convert("RGB")
            cs = self.make_multicrops(pil) if multicrop else [pil]
            views += cs
            wts.append(list(self.MULTICROP_WEIGHTS) if multicrop else [1.0])
        nv = len(wts[0])
        x = torch.stack([pre(v) for v in views]).to(self.device, non_blocking=True)
        gray = torch.from_numpy(np.stack([pil_to_gray256(v, clahe) for v in views])).to(self.device, non_blocking=True)
        pooled, _ = self.engine(x)
        z_sig = ops.head_fwd(self.head, pooled)[1]
        z_freq = self.scoring(torch.zeros_like(z_sig), feats=self.freq.from_gray(gray))["z_freq"]
        w = torch.tensor(wts, dtype=torch.float32, device=self.device)
        zs = (z_sig.view(-1, nv) * w).sum(1).contiguous()
        zf = (z_freq.view(-1, nv) * w).sum(1).contiguous()
        out = self.scoring(zs, z_freq=zf)
        host = {k: v.cpu().numpy() for k, v in out.items()}
        res = []
        for i in range(len(pils)):
            zsi, zfi = float(host["z_sig"][i]), float(host["z_freq"][i])
            res.append({"z_sig": zsi, "z_freq": zfi, "z_scaled": float(host["z_scaled"][i]),
                        "p_fake_raw": float(host["p_raw"][i]), "p_fake_coral": float(host["p_coral"][i]),
                        "p_blend": float(host["p_blend"][i]), "visual_prob": 1.0 / (1.0 + np.exp(-zsi)),
                        "freq_prob": 1.0 / (1.0 + np.exp(-zfi / freq_temp)), "p_or": None, "p_moe": None,
                        "risk_idx": int(host["risk_idx"][i]), "risk_probs": torch.from_numpy(host["risk_probs"][i].copy()),
                        "entropy": float(host["entropy"][i])})
        return res

    # ---- host-buffer path (what a caller of the reference loops sees) -----------------------------------
    def detect(self, images_host: torch.Tensor, gray256_host: Optional[torch.Tensor] = None, resize_mode: int = 0,
               clahe: bool = True, pil_resize: Optional[str] = None) -> np.ndarray:
        """Host (ideally pinned) u8 NHWC images [+ f32 gray256; None = computed on the device from the same
        pixels] -> numpy [B,14] score records."""
        img = images_host.to(self.device, non_blocking=True)
        gray = None if gray256_host is None else gray256_host.to(self.device, non_blocking=True)
        packed = self.pack(self.detect_device(img, gray, resize_mode, clahe, pil_resize))
        key = f"out{packed.shape[0]}"
        if key not in self._pin:
            self._pin[key] = torch.empty(packed.shape, dtype=torch.float32).pin_memory()
        self._pin[key].copy_(packed, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        # a fresh array per call: the pinned buffer is reused by the next call of the same batch size, and callers collect
        # per-batch results (`all_probs.extend(...)` in the reference loops) — 28 KB per 512 images
        return self._pin[key].numpy().copy()
