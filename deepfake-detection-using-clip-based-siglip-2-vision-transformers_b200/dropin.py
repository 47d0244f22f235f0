"""Drop-in Python surface of the detection hot path: the names, arguments and return shapes the
reference's entry points use, backed by the dfd engine instead of open_clip / timm / HF modules.

    open_clip.create_model_and_transforms(name, pretrained, device) -> (model, None, preprocess)
        inference_ai_human_images.py:124-128, train_fusion_head_only.py:81-83, coral.py:87-89
    BinaryClassifier(model_size, device)  .backbone .preprocess .resolution .classifier  forward(x) -> [B]
        inference_ai_human_images.py:111-152 (head H-A) ; train_fusion_head_only.py:78-109 (SE + head H-B)
    run_inference / few_shot_prototype
        inference_ai_human_images.py:250-318, 477-541
    transformers.SiglipVisionModel(...)(pixel_values=...) -> .pooler_output / .last_hidden_state
        Siglip2sidafrozen.py:753,787-788

`install_import_shims()` registers `open_clip` and `pywt` stand-ins in sys.modules so the reference scripts can be
imported unmodified on a box without those packages.
"""
from __future__ import annotations

import sys
import types
from typing import Dict, Iterable, Optional

import numpy as np
import torch

from . import ops
from .engine import ARCHS, SiglipEngine, VisionArch, arch_from_state_dict, canonicalize_state_dict
from .pipeline import head_params_from_state
from .weights import random_classifier_head, random_vision_state_dict


def _as_device(device) -> torch.device:
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError("the dfd backbone runs on CUDA only (sm_100a); there is no CPU fallback")
    return torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device())


def make_preprocess(resolution: int, interpolation: str = "bicubic"):
    """PIL -> float32 [3,S,S] in [-1,1]: Resize((S,S)) + ToTensor + Normalize(.5,.5) — what open_clip's SigLIP
    transform and inference_ai_human_images.py:200-204 produce.  (Host side, as in the reference.)"""
    from torchvision import transforms
    from torchvision.transforms import InterpolationMode

    mode = {"bicubic": InterpolationMode.BICUBIC, "bilinear": InterpolationMode.BILINEAR}[interpolation]
    return transforms.Compose([
        transforms.Lambda(lambda im: im.convert("RGB")),
        transforms.Resize((resolution, resolution), interpolation=mode),
        transforms.ToTensor(),
        transforms.Normalize([0.5] * 3, [0.5] * 3),
    ])


class VisionTower:
    """The object `open_clip.create_model_and_transforms` returns, reduced to what the reference touches:
    `encode_image`, `embed_dim`, `eval`, `to`, `load_state_dict`, `parameters`."""

    def __init__(self, arch: VisionArch, device, max_batch: int = 64, state_dict: Optional[dict] = None, seed: int = 0):
        self.arch = arch
        self.device = _as_device(device)
        self.embed_dim = arch.hidden_size
        self.engine = SiglipEngine(arch, self.device.index, max_batch)
        sd = state_dict if state_dict is not None else random_vision_state_dict(arch, seed, device=self.device)
        self.engine.load_state_dict(sd)
        self.resize_mode = ops.RESIZE_NONE

    def encode_image(self, x: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """x: float32 [B,3,H,W] normalised to [-1,1] (or uint8 [B,H,W,3]) -> float32 [B,D] pooled embeddings."""
        x = x.to(self.device, non_blocking=True)
        if x.dtype in (torch.float16, torch.bfloat16, torch.float64):
            x = x.float()
        mode = self.resize_mode
        hw = x.shape[1:3] if x.dtype == torch.uint8 else x.shape[2:4]
        gp, P = self.arch.grid * self.arch.patch_size, self.arch.patch_size
        if not all(gp <= s < gp + P for s in hw) and mode == ops.RESIZE_NONE:
            mode = ops.RESIZE_BILINEAR
        pooled, _ = self.engine(x, resize_mode=mode)
        f = pooled.float()
        return f / f.norm(dim=-1, keepdim=True) if normalize else f

    def load_state_dict(self, sd: dict, strict: bool = False):
        canon = canonicalize_state_dict(sd)
        if canon:
            self.engine.load_state_dict(sd)
        elif strict:
            raise RuntimeError("no vision-tower tensors in the state dict")
        return self

    def eval(self):
        return self

    def train(self, mode: bool = True):
        return self

    def to(self, *a, **k):
        return self

    def parameters(self) -> Iterable[torch.Tensor]:
        return iter(())

    def named_parameters(self):
        return iter(())

    __call__ = encode_image


def create_model_and_transforms(model_name: str, pretrained: Optional[str] = None, device="cuda", max_batch: int = 64,
                                state_dict: Optional[dict] = None, **_):
    """open_clip-shaped factory.  No pretrained weights are reachable offline: unless `state_dict` is given the
    tower gets seeded random weights of the named architecture (`pretrained` is accepted and ignored)."""
    if model_name not in ARCHS:
        raise KeyError(f"unknown model '{model_name}'; known: {sorted(ARCHS)}")
    arch = ARCHS[model_name]
    model = VisionTower(arch, device, max_batch=max_batch, state_dict=state_dict)
    return model, None, make_preprocess(arch.image_size)


class _Head:
    """Holder of classifier-head tensors with the reference's key names (classifier.N.*, se.N.*)."""

    def __init__(self, state: Dict[str, torch.Tensor], prefix: str):
        self.prefix = prefix
        self._state = state

    def state_dict(self):
        return {k[len(self.prefix):]: v for k, v in self._state.items() if k.startswith(self.prefix)}


class BinaryClassifier:
    """Both reference variants behind one class:
      head='A'  inference_ai_human_images.py:111-152 — f/||f||, LN -> Linear(D,D/2) -> GELU -> Linear(D/2,1)
      head='B'  train_fusion_head_only.py:78-109     — nearest resize if needed, f/(||f||+1e-6), SE gate, 3-layer MLP
    `model_size` follows the reference's tables; `arch` overrides it with any name in ARCHS."""

    SIZES = {"small": "ViT-B-16-SigLIP-384", "medium": "ViT-L-16-SigLIP-384", "large": "ViT-L-16-SigLIP-384",
             "so400m": "google/siglip2-so400m-patch14-384", "base224": "google/siglip2-base-patch16-224"}

    def __init__(self, model_size: str = "large", device="cuda", head: str = "A", arch: Optional[str] = None,
                 max_batch: int = 64, backbone_state: Optional[dict] = None, seed: int = 0):
        name = arch or self.SIZES[model_size]
        self.device = _as_device(device)
        self.backbone, _, self.preprocess = create_model_and_transforms(name, None, self.device, max_batch,
                                                                         backbone_state)
        self.arch = self.backbone.arch
        self.resolution = self.arch.image_size
        self.head_kind = head
        self._head_state = {k: v.clone() for k, v in random_classifier_head(head, self.arch.hidden_size, seed + 1).items()}
        self._params = None
        if head == "B":
            self.backbone.resize_mode = ops.RESIZE_NEAREST  # F.interpolate default (train_fusion_head_only.py:103-104)

    # reference attribute surface
    @property
    def classifier(self):
        return _Head(self._head_state, "classifier.")

    @property
    def se(self):
        return _Head(self._head_state, "se.")

    def eval(self):
        return self

    def train(self, mode: bool = True):
        return self

    def to(self, *a, **k):
        return self

    def parameters(self):
        return iter(())

    def state_dict(self):
        return dict(self._head_state)

    def load_state_dict(self, sd: dict, strict: bool = True):
        """Accepts the reference checkpoints: `backbone.*` (+ ignored `backbone.text.*`) go to the
        engine, `classifier.*` / `se.*` to the head.  Non-strict loading drops shape mismatches like
        train_fusion_head_only.py:111-123."""
        sd = {k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k: v for k, v in sd.items()}
        missing, loaded = [], 0
        for k in self._head_state:
            if k in sd and tuple(sd[k].shape) == tuple(self._head_state[k].shape):
                self._head_state[k] = sd[k].detach().float().cpu().clone()
                loaded += 1
            else:
                missing.append(k)
        bb = {k: v for k, v in sd.items() if k.startswith("backbone.") and not k.startswith("backbone.text.")}
        if bb:
            self.backbone.load_state_dict(bb)
        elif strict:
            missing.append("backbone.*")
        if strict and missing:
            raise RuntimeError(f"missing keys: {missing}")
        self._params = None
        return types.SimpleNamespace(missing_keys=missing, unexpected_keys=[])

    def _head(self):
        if self._params is None:
            self._params = head_params_from_state(self._head_state, self.arch.hidden_size, self.device)
        return self._params

    def _pooled(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(self.device, non_blocking=True)
        if x.dtype not in (torch.uint8, torch.float32):
            x = x.float()
        hw = x.shape[1:3] if x.dtype == torch.uint8 else x.shape[2:4]
        mode = ops.RESIZE_NONE
        gp, P = self.arch.grid * self.arch.patch_size, self.arch.patch_size
        if not all(gp <= s < gp + P for s in hw):
            mode = self.backbone.resize_mode or ops.RESIZE_BILINEAR
        return self.backbone.engine(x, resize_mode=mode)[0]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.head_fwd(self._head(), self._pooled(x))[1]

    __call__ = forward

    def prototype_probs(self, x: torch.Tensor, prototypes: Dict[str, torch.Tensor]) -> torch.Tensor:
        """P(fake) = softmax([-||f-p_real||, -||f-p_fake||])[1]  (inference_ai_human_images.py:288-295), fused into the
        head kernel."""
        pr = torch.stack([prototypes["real"], prototypes["fake"]]).to(self.device, torch.float32).contiguous()
        return ops.head_fwd(self._head(), self._pooled(x), prototypes=pr)[2]

    def features(self, x: torch.Tensor) -> torch.Tensor:
        """L2-normalised embeddings f32 [B,D] (`encode_image` + `/ norm`)."""
        return ops.head_fwd(self._head(), self._pooled(x), want_features=True)[0]


class FastBinaryClassifier(BinaryClassifier):
    """cifake_binary_classifier.py:597-749 (BASELINE config 4): bilinear align_corners=False resize inside the model
    (32x32 CiFake images -> S), f/||f|| -> LayerNorm -> one-token attention -> size-dependent classifier (head H-D).
    State-dict keys follow the reference: layer_norm.*, attention.*, classifier.*."""

    CONFIGS = {"tiny": "ViT-B-16-SigLIP-256", "small": "ViT-B-16-SigLIP-384", "medium": "ViT-L-16-SigLIP-384",
               "large": "ViT-SO400M-16-SigLIP2-512"}

    def __init__(self, model_size: str = "small", device="cuda", arch: Optional[str] = None, max_batch: int = 64,
                 backbone_state: Optional[dict] = None, head_state: Optional[dict] = None, **_):
        name = arch or self.CONFIGS[model_size]
        self.model_size = model_size
        self.device = _as_device(device)
        self.backbone, _, self.preprocess = create_model_and_transforms(name, None, self.device, max_batch, backbone_state)
        self.arch = self.backbone.arch
        self.resolution = self.arch.image_size
        self.feature_dim = self.arch.hidden_size
        self.head_kind = "D"
        self.backbone.resize_mode = ops.RESIZE_BILINEAR
        self._head_state = {k: v.detach().float().cpu().clone() for k, v in (head_state or {}).items()}
        self._params = None

    def load_state_dict(self, sd: dict, strict: bool = True):
        sd = {k[len("_orig_mod."):] if k.startswith("_orig_mod.") else k: v for k, v in sd.items()}
        head = {k: v.detach().float().cpu().clone() for k, v in sd.items()
                if k.startswith(("layer_norm.", "attention.", "classifier."))}
        if head:
            self._head_state = head
        bb = {k: v for k, v in sd.items() if k.startswith("backbone.") and not k.startswith("backbone.text.")}
        if bb:
            self.backbone.load_state_dict(bb)
        if strict and (not head or not bb):
            raise RuntimeError("missing keys: " + ("head " if not head else "") + ("backbone.*" if not bb else ""))
        self._params = None
        return types.SimpleNamespace(missing_keys=[], unexpected_keys=[])


@torch.no_grad()
def run_inference(model: BinaryClassifier, dataloader, device=None, use_amp: bool = True, desc: str = "Inference",
                  invert_logits: bool = False, prototypes: Optional[dict] = None):
    """inference_ai_human_images.py:250-318: loader of (images, labels, filenames) -> (labels, P(fake), filenames).
    One D2H read per batch, like the reference loop."""
    all_labels, all_probs, all_files = [], [], []
    for images, labels, filenames in dataloader:
        if prototypes is not None:
            probs = model.prototype_probs(images, prototypes)
        else:
            z = model(images)
            probs = torch.sigmoid(-z if invert_logits else z)
        all_probs.extend(probs.cpu().numpy())
        all_labels.extend(np.asarray(labels))
        all_files.extend(filenames)
    return np.array(all_labels), np.array(all_probs), all_files


@torch.no_grad()
def few_shot_prototype(model: BinaryClassifier, support_loader, device=None, use_amp: bool = True) -> dict:
    """inference_ai_human_images.py:477-541: L2-normalised class means of L2-normalised features.  Class sums are
    accumulated on the device; with torch.distributed initialised they are all-reduced so every rank gets the
    same prototypes."""
    from . import distributed

    D = model.arch.hidden_size
    sums = torch.zeros(2, D, device=model.device)
    counts = torch.zeros(2, device=model.device)
    for images, labels, _ in support_loader:
        f = model.features(images)
        lab = (torch.as_tensor(labels).to(model.device) != 0).long()
        sums.index_add_(0, lab, f)
        counts.index_add_(0, lab, torch.ones_like(lab, dtype=torch.float32))
    bucket = torch.cat([sums.reshape(-1), counts])
    distributed.all_reduce_sum_(bucket)
    sums, counts = bucket[: 2 * D].reshape(2, D), bucket[2 * D:]
    means = sums / counts.clamp_min(1.0)[:, None]
    protos = means / means.norm(dim=-1, keepdim=True)
    return {"real": protos[0], "fake": protos[1]}


class SiglipVisionModel:
    """HF-shaped wrapper: `SiglipVisionModel.from_state_dict(sd)(pixel_values=x)` -> .pooler_output [B,D] f32,
    .last_hidden_state [B,N,D], and with output_hidden_states=True the tuple of L+1 per-layer states
    (Siglip2sidafrozen.py:753,787-793)."""

    def __init__(self, arch: VisionArch, state_dict: dict, device="cuda", max_batch: int = 32):
        self.arch = arch
        self.device = _as_device(device)
        self.config = types.SimpleNamespace(hidden_size=arch.hidden_size, image_size=arch.image_size,
                                            patch_size=arch.patch_size, num_hidden_layers=arch.num_hidden_layers,
                                            num_attention_heads=arch.num_attention_heads,
                                            intermediate_size=arch.intermediate_size)
        self.engine = SiglipEngine(arch, self.device.index, max_batch).load_state_dict(state_dict)

    @classmethod
    def from_state_dict(cls, state_dict: dict, device="cuda", max_batch: int = 32, num_heads: Optional[int] = None):
        return cls(arch_from_state_dict(state_dict, num_heads), state_dict, device, max_batch)

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def __call__(self, pixel_values: torch.Tensor, output_hidden_states: bool = False,
                 interpolate_pos_encoding: bool = False, **_):
        gp, P = self.arch.grid * self.arch.patch_size, self.arch.patch_size
        if not all(gp <= s < gp + P for s in pixel_values.shape[-2:]):
            raise ValueError(f"pixel_values sides must be in [{gp}, {gp + P}) (position-embedding interpolation "
                             "for other grids is not built)")
        x = pixel_values.to(self.device).float()
        if output_hidden_states:  # tuple of L+1 tensors, as SigLIP2_MTL consumes them (Siglip2sidafrozen.py:790-793)
            pooled, last, hid = self.engine.forward_hidden(x)
            return types.SimpleNamespace(pooler_output=pooled.float(), last_hidden_state=last.float(),
                                         hidden_states=tuple(h.float() for h in hid))
        pooled, last = self.engine(x, want_last_hidden=True)
        return types.SimpleNamespace(pooler_output=pooled.float(), last_hidden_state=last.float(), hidden_states=None)


def install_import_shims() -> None:
    """Make `import open_clip` / `import pywt` resolve to dfd-backed stand-ins (the reference scripts import both at
    module level; neither package is installed here)."""
    oc = types.ModuleType("open_clip")
    oc.create_model_and_transforms = create_model_and_transforms
    oc.__doc__ = "dfd stand-in for open_clip (vision tower only)"
    sys.modules.setdefault("open_clip", oc)

    pw = types.ModuleType("pywt")

    def dwt2(x, wavelet):
        if wavelet not in ("db1", "haar"):
            raise ValueError("dfd pywt stand-in only implements the db1 / Haar wavelet")
        a = np.asarray(x)
        p, q, r, s = a[0::2, 0::2], a[0::2, 1::2], a[1::2, 0::2], a[1::2, 1::2]
        return (p + q + r + s) * 0.5, ((p + q - r - s) * 0.5, (p - q + r - s) * 0.5, (p - q - r + s) * 0.5)

    pw.dwt2 = dwt2
    sys.modules.setdefault("pywt", pw)
