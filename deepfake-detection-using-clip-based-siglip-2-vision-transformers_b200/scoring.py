"""Host side of the scoring stack: constants of the 256² frequency grid, the gray256 producer, loaders for the
reference's artefact files, and reference-shaped classes (FreqMLP, FusionHead, AdaptiveFusionHead,
CoralCalibrator) whose arithmetic runs in libdfd's CUDA kernels.

Reference: train_fusion_head_only.py:142-317; deepfake-detector-v2/app.py:213-255,601-709,736-846,1265-1412.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import ops

EPS_TRAINER = 1e-8   # train_fusion_head_only.py:51, deepfake-detector-v2/app.py
EPS_APPV3 = 1e-6     # appv3.py:570
FREQ_TEMP = 1.25     # deepfake-detector-v2/app.py:280
_N = 256


def build_freq_tables() -> tuple:
    """Per-pixel lookup tables over the fft-SHIFTED 256x256 grid ([sy, sx], CPU), built with the very torch ops the
    reference uses so every comparison / bin edge rounds identically (train_fusion_head_only.py:156-169,193-197):
      band   u8  0: r<=r1, 1: r1<r<=r2, 2: r>r2
      rbin   i8  log-radius bin 0..38 (bucketize(r+1, logspace(0, log10(rmax+1), 40)) - 1), -1 = not counted
      sector i8  0..7 for a0 <= atan2(dy,dx) < a0+pi/4, -1 = in no sector (angle == pi)
    """
    yy, xx = torch.meshgrid(torch.arange(_N), torch.arange(_N), indexing="ij")
    cy = cx = _N // 2
    r = torch.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)
    rmax = float(r.max())
    r1, r2 = 0.15 * rmax, 0.45 * rmax
    band = torch.zeros((_N, _N), dtype=torch.uint8)
    band[(r > r1) & (r <= r2)] = 1
    band[r > r2] = 2
    rb = torch.logspace(math.log10(1.0), math.log10(rmax + 1.0), 40)
    ridx = (torch.bucketize(r.flatten() + 1.0, rb) - 1).reshape(_N, _N)
    rbin = torch.where((ridx >= 0) & (ridx < len(rb) - 1), ridx, torch.full_like(ridx, -1)).to(torch.int8)
    ang = torch.atan2(yy - cy, xx - cx)
    sector = torch.full((_N, _N), -1, dtype=torch.int8)
    for k, a0 in enumerate(np.linspace(-math.pi, math.pi, 8, endpoint=False)):
        sector[(ang >= a0) & (ang < a0 + math.pi / 4)] = k
    return band.contiguous(), rbin.contiguous(), sector.contiguous()


def pack_freq_luts(band: torch.Tensor, rbin: torch.Tensor, sector: torch.Tensor) -> torch.Tensor:
    """The table dfd_freq_features reads (include/dfd.h): int32 [256*256 + 48] — one word per grid position, TRANSPOSED
    (index sx*256 + sy) = band | (rbin & 0xff) << 8 | (sector & 0xff) << 16, then the populations of the 40 log-radius
    bins and the 8 sectors (the denominators of the reference's masked `.mean()`s, which depend on geometry only)."""
    word = band.to(torch.int32) | ((rbin.to(torch.int32) & 0xFF) << 8) | ((sector.to(torch.int32) & 0xFF) << 16)
    logcnt = torch.bincount(rbin[rbin >= 0].to(torch.int64), minlength=40)[:40]
    seccnt = torch.bincount(sector[sector >= 0].to(torch.int64), minlength=8)[:8]
    return torch.cat([word.t().contiguous().flatten(), logcnt.to(torch.int32), seccnt.to(torch.int32)]).contiguous()


def build_freq_luts(device) -> torch.Tensor:
    return pack_freq_luts(*build_freq_tables()).to(device)


def pil_to_gray256(pil, clahe: bool) -> np.ndarray:
    """Stage-1 boundary: gray256 is produced on the host with the reference's own integer path —
    exif transpose, PIL 'L' luma, optional cv2 CLAHE(2.0, 8x8), PIL bicubic 256² — then /255 in fp32
    (train_fusion_head_only.py:142-148; app.py:736-749)."""
    from PIL import Image, ImageOps

    g = ImageOps.exif_transpose(pil).convert("L")
    if clahe:
        import cv2

        g = Image.fromarray(cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(np.array(g, dtype=np.uint8)))
    g = g.resize((_N, _N), Image.BICUBIC)
    return np.asarray(g, dtype=np.float32) / 255.0


class FreqFeatureExtractor:
    """extract_freq_vector on the GPU for batches of gray256 images (train_fusion_head_only.py:224-226 raw,
    app.py:840-846 z-scored)."""

    def __init__(self, device, eps: float = EPS_TRAINER, zscore: bool = False, clahe: bool = True):
        self.device = torch.device(device)
        self.eps, self.zscore, self.clahe = eps, zscore, clahe
        self.luts = build_freq_luts(self.device)
        self._scratch = None

    def from_gray(self, gray256: torch.Tensor) -> torch.Tensor:
        need = ops._lib.load().dfd_freq_scratch_bytes(gray256.shape[0])
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty((need,), dtype=torch.uint8, device=self.device)
        return ops.freq_features(gray256, self.luts, self.eps, self.zscore, self._scratch)

    def __call__(self, pils: Sequence) -> torch.Tensor:
        g = np.stack([pil_to_gray256(p, self.clahe) for p in pils])
        return self.from_gray(torch.from_numpy(g).to(self.device, non_blocking=True))


# ---- artefact loaders ------------------------------------------------------------------------------------
def _logit(p: float) -> float:
    p = min(max(p, 1e-6), 1 - 1e-6)
    return math.log(p / (1 - p))


def load_coral(cutpoints_path: Optional[str], temp_path: Optional[str]) -> tuple:
    """Returns (cut logits[4], temperature).  Accepts the G1 dict-of-quantile-probabilities file
    (app.py:1271-1279), the G2 list-of-logits file (coral.py:381-386), and {'temperature'} / {'temp'} / bare
    float temperature files (app.py:237-243).  Missing cutpoints -> the app's fallback .32/.47/.61/.75."""
    cuts = None
    if cutpoints_path and os.path.exists(cutpoints_path):
        with open(cutpoints_path) as f:
            cuts = json.load(f)
    if isinstance(cuts, dict) and cuts:
        c = [_logit(cuts[k]) for k in ("q25", "q50", "q75", "max")]
    elif isinstance(cuts, list) and len(cuts) == 4:
        c = [float(v) for v in cuts]
    else:
        c = [_logit(v) for v in (0.32, 0.47, 0.61, 0.75)]
    temp = 1.0
    if temp_path and os.path.exists(temp_path):
        with open(temp_path) as f:
            t = json.load(f)
        temp = float(t.get("temperature", t.get("temp", 1.0))) if isinstance(t, dict) else float(t)
    return c, temp


def fit_coral_cutpoints_shipped(probs: np.ndarray) -> Dict[str, float]:
    """The rule behind the shipped coral_cutpoints.json: quantiles (.25,.5,.75) and max of the per-sample fused
    probabilities in coral_bins.npy."""
    q = np.quantile(probs, [0.25, 0.5, 0.75]).astype(np.float32)  # the shipped values are float32-rounded
    return {"q25": float(q[0]), "q50": float(q[1]), "q75": float(q[2]), "max": float(np.float32(np.max(probs)))}


def fit_coral_cutpoints(logits, labels=None, num_classes: int = 5) -> list:
    """coral.py:300-322: sorted fused logits at ranks floor(q*n), q in (.15,.35,.55,.75)."""
    s = np.sort(np.asarray(logits.detach().cpu() if hasattr(logits, "detach") else logits))
    return [float(s[int(q * len(s))]) for q in (0.15, 0.35, 0.55, 0.75)]


def write_coral_artifacts(out_prefix: str, fused_logits, temperature: float = 1.0) -> dict:
    """coral.py:375-397 ("fit_coral_v5"): `<prefix>_cutpoints.json` = list of 4 logit-space cutpoints,
    `<prefix>_temp.json` = {"temperature": 1.0}, `<prefix>_bins.npy` = np.histogram(logits, 50)[0].  The files load
    back through load_coral()."""
    lg = np.asarray(fused_logits.detach().cpu() if hasattr(fused_logits, "detach") else fused_logits, dtype=np.float32)
    cuts = fit_coral_cutpoints(lg)
    with open(out_prefix + "_cutpoints.json", "w") as f:
        json.dump(cuts, f, indent=2)
    with open(out_prefix + "_temp.json", "w") as f:
        json.dump({"temperature": temperature}, f, indent=2)
    np.save(out_prefix + "_bins.npy", np.histogram(lg, bins=50)[0])
    return {"cutpoints": cuts, "temperature": temperature}


def detect_generation(freq_state: Optional[dict], fusion_state: Optional[dict]) -> int:
    """G1: keys net.* / fc.*; G2: normer.* / mlp.* (SURVEY.md App. B)."""
    for sd, g1, g2 in ((freq_state, "net.1.weight", "normer.mean"), (fusion_state, "fc.weight", "mlp.0.weight")):
        if sd is not None:
            if g1 in sd:
                return 1
            if g2 in sd:
                return 2
    raise KeyError("cannot tell the head generation from the state dict keys")


class ScoringStack:
    """FreqMLP + fusion + temperature + CORAL as ONE warp-level kernel (dfd_score_epilogue).

    `from_dir(path)` reads the reference's artefact names: freq_mlp.safetensors, fusion_head.safetensors,
    coral_cutpoints.json, coral_temp.json (siglip/ in the reference repo)."""

    def __init__(self, device, freq_state: dict, fusion_state: dict, cut_logits, coral_temp: float,
                 freq_temp: float = FREQ_TEMP):
        self.gen = detect_generation(freq_state, fusion_state)
        if detect_generation(freq_state, None) != detect_generation(None, fusion_state):
            raise ValueError("FreqMLP and fusion head are of different generations")
        self.device = torch.device(device)
        self.cut_logits, self.coral_temp = [float(c) for c in cut_logits], float(coral_temp)
        self.params = ops.ScoreParams(self.gen, self.device, freq_state=freq_state, fusion_state=fusion_state,
                                      coral_cuts_logit=self.cut_logits, coral_temp=coral_temp, freq_temp=freq_temp)

    @classmethod
    def from_dir(cls, path: str, device, prefix: str = "coral") -> "ScoringStack":
        from safetensors.torch import load_file

        cuts, temp = load_coral(os.path.join(path, f"{prefix}_cutpoints.json"), os.path.join(path, f"{prefix}_temp.json"))
        return cls(device, load_file(os.path.join(path, "freq_mlp.safetensors")),
                   load_file(os.path.join(path, "fusion_head.safetensors")), cuts, temp)

    def __call__(self, z_sig: torch.Tensor, feats: Optional[torch.Tensor] = None,
                 z_freq: Optional[torch.Tensor] = None) -> dict:
        return ops.score_epilogue(self.params, z_sig.contiguous().float(), feats, z_freq)


# ---- reference-shaped classes -----------------------------------------------------------------------------
class _StateHolder:
    """Minimal nn.Module-like surface the reference scripts touch: load_state_dict / state_dict / eval / to."""

    _keys: tuple = ()

    def __init__(self):
        self._state: Dict[str, torch.Tensor] = {}
        self._device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        self._params = None

    def load_state_dict(self, sd, strict: bool = True):
        missing = [k for k in self._keys if k not in sd]
        extra = [k for k in sd if k not in self._keys]
        if strict and (missing or extra):
            raise RuntimeError(f"state dict mismatch: missing {missing}, unexpected {extra}")
        self._state = {k: sd[k].detach().clone().float() for k in self._keys if k in sd}
        self._params = None
        return self

    def state_dict(self):
        return {k: v.clone() for k, v in self._state.items()}

    def eval(self):
        return self

    def train(self, mode: bool = True):
        return self

    def to(self, device):
        d = torch.device(device)
        if d.type != "cuda":
            raise RuntimeError("dfd heads run on CUDA only (there is no CPU fallback)")
        self._device, self._params = d, None
        return self


class FreqMLP(_StateHolder):
    """G1 (app.py:615-628; eval-time 0.001·randn jitter NOT applied — deterministic) or G2
    (train_fusion_head_only.py:282-301).  forward([B,24]) -> logits [B]."""

    _G1 = ("net.0.weight", "net.0.bias", "net.1.weight", "net.1.bias", "net.3.weight", "net.3.bias")
    _G2 = ("normer.mean", "normer.std", "contrast.alpha", "contrast.beta", "band.gates") + tuple(
        f"blocks.{b}.{n}" for b in range(2) for n in ("norm.weight", "norm.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")
    ) + ("head.weight", "head.bias", "temp.T")

    def __init__(self, dim: int = 24, hidden: int = 64, num_bands: int = 4, in_dim: Optional[int] = None,
                 hid: Optional[int] = None):
        super().__init__()
        assert (in_dim or dim) == 24 and (hid or hidden) == 64 and num_bands == 4
        self._keys = self._G2

    def load_state_dict(self, sd, strict: bool = True):
        self._keys = self._G1 if "net.1.weight" in sd else self._G2
        return super().load_state_dict(sd, strict)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        gen = 1 if self._keys is self._G1 else 2
        if self._params is None:
            self._params = ops.ScoreParams(gen, x.device, freq_state=self._state,
                                           fusion_state=_IDENTITY_FUSION[gen])
        z_dummy = torch.zeros((x.shape[0],), dtype=torch.float32, device=x.device)
        return ops.score_epilogue(self._params, z_dummy, x.contiguous().float())["z_freq"]

    __call__ = forward


_IDENTITY_FUSION = {
    1: {"fc.weight": torch.zeros(1, 2), "fc.bias": torch.zeros(1)},
    2: {"mlp.0.weight": torch.zeros(32, 3), "mlp.0.bias": torch.zeros(32), "mlp.2.weight": torch.zeros(2, 32),
        "mlp.2.bias": torch.zeros(2), "temp.T": torch.tensor(1.0)},
}


class FusionHead(_StateHolder):
    """G1: Linear(2,1) on [p_sig, p_freq] (app.py:691-696).  forward([B,2]) -> [B,1]."""

    _keys = ("fc.weight", "fc.bias")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        w = self._state["fc.weight"].to(x.device).reshape(-1)
        # two multiply-adds per sample: done by the score epilogue in the pipeline; here (probability inputs given
        # directly) the same arithmetic is expressed through the kernel by inverting the sigmoids
        eps = 1e-7
        p = x.float().clamp(eps, 1 - eps)
        z_sig = torch.log(p[:, 0] / (1 - p[:, 0])).contiguous()
        z_freq = (torch.log(p[:, 1] / (1 - p[:, 1])) * FREQ_TEMP).contiguous()
        prm = ops.ScoreParams(1, x.device, fusion_state=self._state)
        return ops.score_epilogue(prm, z_sig, None, z_freq)["z"].unsqueeze(1)

    __call__ = forward


class CoralCalibrator:
    """app.py:1269-1297 — scalar API plus a batched overload; arithmetic in dfd_score_epilogue."""

    def __init__(self, cuts: Optional[dict] = None, device=None):
        self.c = torch.tensor([_logit(cuts[k]) for k in ("q25", "q50", "q75", "max")] if cuts
                              else [_logit(v) for v in (0.32, 0.47, 0.61, 0.75)], dtype=torch.float32)
        self._device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        # identity fusion: z = 1·z_freq-path disabled; we feed z directly through gen-2 weights that pass z_sig
        self._params = None

    def _run(self, z: torch.Tensor) -> dict:
        if self._params is None:
            self._params = _passthrough_params(self._device, self.c.tolist(), 1.0)
        z = torch.as_tensor(z, dtype=torch.float32, device=self._device).reshape(-1).contiguous()
        return ops.score_epilogue(self._params, z, None, z.clone())

    @torch.no_grad()
    def probs(self, z_scaled) -> torch.Tensor:
        p = self._run(z_scaled)["risk_probs"]
        return p[0] if p.shape[0] == 1 and not (torch.is_tensor(z_scaled) and z_scaled.dim() > 0) else p

    @torch.no_grad()
    def predict(self, z_scaled):
        out = self._run(z_scaled)
        if out["risk_idx"].numel() == 1 and not (torch.is_tensor(z_scaled) and z_scaled.dim() > 0):
            return int(out["risk_idx"].item()), out["risk_probs"][0].cpu()
        return out["risk_idx"], out["risk_probs"]


def _passthrough_params(device, cut_logits, coral_temp: float) -> "ops.ScoreParams":
    """gen-2 fusion weights with zero MLP => softmax weights (.5,.5); with z_freq == z_sig == z the fused logit
    is exactly z (0.5z + 0.5z), so the epilogue applies only temperature + CORAL."""
    fus = {"mlp.0.weight": torch.zeros(32, 3), "mlp.0.bias": torch.zeros(32), "mlp.2.weight": torch.zeros(2, 32),
           "mlp.2.bias": torch.zeros(2), "temp.T": torch.tensor(1.0 - 1e-6)}
    return ops.ScoreParams(2, device, fusion_state=fus, coral_cuts_logit=cut_logits, coral_temp=coral_temp)
