"""Drop-in for train_fusion_head_only.py (entry point #2): train ONLY the AdaptiveFusionHead on (z_freq, z_sig)
logits of a frozen SigLIP classifier + FreqMLP, and write `fusion_head.safetensors` in the reference's layout
(mlp.0.weight[32,3], mlp.0.bias[32], mlp.2.weight[2,32], mlp.2.bias[2], temp.T[]).

Differences from the reference, on purpose:
  * the two extraction loops (train_fusion_head_only.py:329-347: one image, one sync at a time) run batched on the
    GPU engine, and so does their preprocessing: the host decodes the files, then per-channel CLAHE, the PIL Resize and the
    gray256 stage are the bit-exact device kernels (dfd_clahe_u8 / dfd_resize_u8 / dfd_gray256; `on_device=False` keeps the
    reference's own cv2 / PIL / torchvision chain);
  * the head's forward+backward is one CUDA kernel (dfd_fusion_fwd_bwd) that returns the batch's loss/gradient
    partial sums; with torch.distributed initialised every rank takes a contiguous shard of each mini-batch and
    one all-reduce of the flat 196-float bucket (195 grads + loss) precedes the identical clip + AdamW step;
  * the reference sets torch.set_grad_enabled(False) globally and then calls loss.backward(), which raises
     — the behaviour reproduced here is "reference loop under enable_grad".
"""
from __future__ import annotations

import argparse
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

from . import distributed, ops
from .scoring import FreqFeatureExtractor, FreqMLP, pil_to_gray256

IMG_EXTS = (".jpg", ".jpeg", ".png", ".bmp", ".webp", ".tif", ".tiff", ".gif", ".jfif", ".heic", ".heif")
IMG_SIZE = 384


class AdaptiveFusionHead:
    """x=[z_f, z_s, |z_f-z_s|] -> Linear(3,32) -> GELU -> Linear(32,2) -> softmax -> w0 z_f + w1 z_s -> /(T+1e-6)
    (train_fusion_head_only.py:303-317).  Parameters live in ONE flat fp32 device tensor in state-dict order."""

    NAMES = ops.FUSION_PARAM_ORDER
    SHAPES = ((32, 3), (32,), (2, 32), (2,), ())

    def __init__(self, hidden_dim: int = 32, device=None):
        assert hidden_dim == 32, "the fused kernel is specialised for the reference's hidden_dim=32"
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        # same default initialisation (and RNG consumption) as the reference's nn.Sequential + TemperatureScaler
        l0, l1 = torch.nn.Linear(3, hidden_dim), torch.nn.Linear(hidden_dim, 2)
        sd = {"mlp.0.weight": l0.weight, "mlp.0.bias": l0.bias, "mlp.2.weight": l1.weight, "mlp.2.bias": l1.bias,
              "temp.T": torch.tensor(1.0)}
        self.flat = torch.nn.Parameter(self._flatten(sd).to(self.device), requires_grad=False)

    @classmethod
    def _flatten(cls, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
        return torch.cat([sd[k].detach().float().reshape(-1).cpu() for k in cls.NAMES]).contiguous()

    def parameters(self):
        return [self.flat]

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out, o = {}, 0
        for k, shp in zip(self.NAMES, self.SHAPES):
            n = int(np.prod(shp)) if shp else 1
            out[k] = self.flat.data[o:o + n].reshape(shp).clone()
            o += n
        return out

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        miss = [k for k in self.NAMES if k not in sd]
        if miss:
            raise RuntimeError(f"missing keys: {miss}")
        self.flat.data.copy_(self._flatten(sd).to(self.device))
        return self

    def train(self, mode: bool = True):
        return self

    def eval(self):
        return self

    def to(self, device):
        return self

    @torch.no_grad()
    def forward(self, z_freq: torch.Tensor, z_sig: torch.Tensor) -> torch.Tensor:
        zf = z_freq.to(self.device, torch.float32).contiguous()
        zs = z_sig.to(self.device, torch.float32).contiguous()
        y = torch.zeros_like(zf)
        return ops.fusion_fwd_bwd(self.flat.data, zf, zs, y, want_logits=True)[2]

    __call__ = forward

    @torch.no_grad()
    def loss_and_grad(self, z_freq, z_sig, y, global_batch: Optional[int] = None):
        """(mean BCE-with-logits over the GLOBAL batch, d loss / d flat params), all-reduced over ranks."""
        n = global_batch if global_batch is not None else z_freq.numel()
        if z_freq.numel() > 0:
            loss, grads, _ = ops.fusion_fwd_bwd(self.flat.data, z_freq.contiguous(), z_sig.contiguous(), y.contiguous(),
                                                1.0 / n)
        else:  # a rank may own an empty shard of a ragged last mini-batch
            loss, grads = torch.zeros(1, device=self.device), torch.zeros(195, device=self.device)
        bucket = torch.cat([grads, loss])
        distributed.all_reduce_sum_(bucket)
        return bucket[195], bucket[:195]


def list_images(folder: str) -> List[str]:
    out = []
    for root, _, files in os.walk(folder):
        out += [os.path.join(root, n) for n in files if n.lower().endswith(IMG_EXTS)]
    return sorted(out)


def apply_clahe(pil):
    import cv2
    from PIL import Image

    arr = np.array(pil, dtype=np.uint8)
    clahe = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8))
    for c in range(3):
        arr[:, :, c] = clahe.apply(arr[:, :, c])
    return Image.fromarray(arr)


def make_preprocess(img_size: int = IMG_SIZE):
    """train_fusion_head_only.py:60-74: per-channel CLAHE -> Resize -> ToTensor -> Normalize(.5,.5) (host side)."""
    from torchvision import transforms

    return transforms.Compose([transforms.Lambda(apply_clahe), transforms.Resize((img_size, img_size)),
                               transforms.ToTensor(), transforms.Normalize([0.5] * 3, [0.5] * 3)])


def _decode_rgb_u8(path: str, exif_transpose: bool = False) -> np.ndarray:
    """Decoded pixels of a file, [H,W,3] u8.  exif_transpose: the gray256 stage honours the EXIF orientation
    (`ImageOps.exif_transpose`, train_fusion_head_only.py:143), the SigLIP preprocess does not (:60-74)."""
    from PIL import Image, ImageOps

    with Image.open(path) as pil:
        rgb = pil.convert("RGB")
        if exif_transpose:
            rgb = ImageOps.exif_transpose(rgb)
        return np.array(rgb, dtype=np.uint8)   # a writable copy: torch.from_numpy refuses read-only buffers


def preprocess_on_device(rgb_u8: np.ndarray, img_size: int, device) -> torch.Tensor:
    """train_fusion_head_only.py:60-74 on the GPU, for one decoded image [H,W,3] u8: ONE upload, per-channel CLAHE
    (dfd_clahe_u8, bit-exact with cv2) -> PIL-exact bilinear Resize((S,S)) (dfd_resize_u8) -> u8 [S,S,3] on the device.
    ToTensor + Normalize(.5,.5) happen inside the patch kernel when the batch enters the engine, so the backbone sees exactly
    the bf16 pixels the host chain `make_preprocess` would have produced."""
    img = torch.from_numpy(np.ascontiguousarray(rgb_u8)).to(device, non_blocking=True)[None]
    return ops.resize_u8(ops.clahe_u8(img), img_size, img_size, "bilinear")[0]


@torch.no_grad()
def extract_siglip_logits(siglip, paths: Sequence[str], batch_size: int = 32, preprocess=None,
                          on_device: Optional[bool] = None) -> torch.Tensor:
    """z_sig of every image (train_fusion_head_only.py:339-347, batched).  on_device (default: whenever no custom `preprocess`
    is passed): the reference's preprocess runs on the GPU (`preprocess_on_device`); the host only decodes the files.
    on_device=False keeps the reference's own host chain (cv2 + PIL + torchvision)."""
    from PIL import Image

    if on_device is None:
        on_device = preprocess is None
    out = []
    if on_device:
        S, dev = siglip.resolution, siglip.device
        with torch.cuda.device(dev):
            for i in range(0, len(paths), batch_size):
                xs = torch.stack([preprocess_on_device(_decode_rgb_u8(p), S, dev) for p in paths[i:i + batch_size]])
                out.append(siglip(xs).float().cpu())
        return torch.cat(out) if out else torch.zeros(0)
    pre = preprocess or make_preprocess(siglip.resolution)
    for i in range(0, len(paths), batch_size):
        xs = []
        for p in paths[i:i + batch_size]:
            with Image.open(p) as pil:
                xs.append(pre(pil.convert("RGB")))
        out.append(siglip(torch.stack(xs)).float().cpu())
    return torch.cat(out) if out else torch.zeros(0)


@torch.no_grad()
def extract_freq_logits(freq_model: FreqMLP, paths: Sequence[str], device, batch_size: int = 64,
                        on_device: bool = True) -> torch.Tensor:
    """z_freq of every image (train_fusion_head_only.py:329-337, batched).  on_device: gray256 (Pillow luma, cv2 CLAHE, Pillow
    bicubic resize — dfd_gray256, bit-exact) is produced on the GPU from one upload per decoded image; on_device=False runs the
    libraries themselves on the host (`scoring.pil_to_gray256`)."""
    from PIL import Image

    fx = FreqFeatureExtractor(device, zscore=False, clahe=True)
    out = []
    for i in range(0, len(paths), batch_size):
        if on_device:
            with torch.cuda.device(fx.device):
                gray = torch.cat([ops.gray256_from_rgb(torch.from_numpy(_decode_rgb_u8(p, True)).to(fx.device, non_blocking=True)[None],
                                                       True) for p in paths[i:i + batch_size]])
        else:
            gs = []
            for p in paths[i:i + batch_size]:
                with Image.open(p) as pil:
                    gs.append(pil_to_gray256(pil.convert("RGB"), clahe=True))
            gray = torch.from_numpy(np.stack(gs)).to(fx.device)
        feats = fx.from_gray(gray)
        out.append(freq_model(feats).float().cpu())
    return torch.cat(out) if out else torch.zeros(0)


def _safe_auc(y: np.ndarray, score: np.ndarray) -> float:
    try:
        from sklearn.metrics import roc_auc_score

        return float(roc_auc_score(y, score))
    except ValueError:
        return float("nan")


def fit_fusion_head(z_freq: torch.Tensor, z_sig: torch.Tensor, labels: torch.Tensor, batch_size: int = 32,
                    epochs: int = 5, lr: float = 5e-4, device=None, head: Optional[AdaptiveFusionHead] = None,
                    verbose: bool = True):
    """The training loop of train_fusion_head_only.py:402-455 on cached logits.  Returns (head, best_state, best_auc)."""
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    head = head or AdaptiveFusionHead(device=device)
    dev = head.device
    loader = DataLoader(TensorDataset(torch.stack([z_freq, z_sig], 1).float(), labels.float()), batch_size=batch_size,
                        shuffle=True, drop_last=False)  # identical permutation on every rank (same torch seed)
    optim = torch.optim.AdamW(head.parameters(), lr=lr)
    best_auc, best_state = -float("inf"), None
    for ep in range(1, epochs + 1):
        losses = []
        for xb, yb in loader:
            lo, hi = distributed.shard_bounds(xb.shape[0], world, rank)
            xs, ys = xb[lo:hi].to(dev), yb[lo:hi].to(dev)
            loss, grads = head.loss_and_grad(xs[:, 0], xs[:, 1], ys, global_batch=xb.shape[0])
            gnorm = grads.norm()                      # clip_grad_norm_(max_norm=5.0) after the all-reduce
            grads = grads * torch.clamp(5.0 / (gnorm + 1e-6), max=1.0)
            head.flat.grad = grads
            optim.step()
            losses.append(loss)
        probs = torch.sigmoid(head(z_freq, z_sig)).cpu().numpy()
        auc = _safe_auc(labels.numpy(), probs)
        acc = float(((probs >= 0.5) == labels.numpy()).mean())
        if verbose and rank == 0:
            print(f"[fusion] epoch {ep:03d}/{epochs} loss={float(torch.stack(losses).mean()):.4f} acc={acc:.3f} auc={auc:.3f}")
        if auc > best_auc:
            best_auc, best_state = auc, {k: v.cpu() for k, v in head.state_dict().items()}
    return head, best_state, best_auc


def train_fusion_head(real_dir: str, fake_dir: str, best_model_path: str, freq_mlp_path: str, fusion_out: str,
                      batch_size: int = 32, epochs: int = 5, device="cuda", arch: str = "ViT-L-16-SigLIP-384",
                      extract_batch: int = 32):
    from safetensors.torch import load_file, save_file

    from .dropin import BinaryClassifier

    real_paths, fake_paths = list_images(real_dir), list_images(fake_dir)
    if not real_paths or not fake_paths:
        raise SystemExit("No images found under real/fake dirs.")
    all_paths = real_paths + fake_paths
    labels = torch.cat([torch.zeros(len(real_paths)), torch.ones(len(fake_paths))])
    print(f"[data] {len(real_paths)} real, {len(fake_paths)} fake, total {len(all_paths)}")

    siglip = BinaryClassifier(device=device, head="B", arch=arch, max_batch=extract_batch)
    siglip.load_state_dict(load_file(best_model_path), strict=False)  # shape-filtered, text tower dropped
    freq_model = FreqMLP()
    freq_model.load_state_dict(load_file(freq_mlp_path), strict=True)

    print("[stage] Extracting frequency logits...")
    z_freq = extract_freq_logits(freq_model, all_paths, siglip.device)
    print("[stage] Extracting SigLIP logits...")
    z_sig = extract_siglip_logits(siglip, all_paths, extract_batch)
    print("[stage] Training AdaptiveFusionHead...")
    _, best_state, best_auc = fit_fusion_head(z_freq, z_sig, labels, batch_size, epochs, device=siglip.device)
    if best_state is None:
        print("[fusion] WARNING: no best_state, not saving.")
        return None
    if not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0:
        save_file({k: v.contiguous() for k, v in best_state.items()}, fusion_out)
        print(f"[fusion] Saved trained fusion head to: {fusion_out}")
        print(f"[fusion] Best AUC on training set: {best_auc:.3f}")
    return best_state


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="Train ONLY fusion_head.safetensors")
    ap.add_argument("--real-dir", type=str, required=True, help="Folder with REAL images.")
    ap.add_argument("--fake-dir", type=str, required=True, help="Folder with FAKE images.")
    ap.add_argument("--best-model", type=str, required=True, help="Path to best_model.safetensors.")
    ap.add_argument("--freq-mlp", type=str, required=True, help="Path to freq_mlp.safetensors (v5 FreqMLP).")
    ap.add_argument("--fusion-out", type=str, required=True, help="Output path for fusion_head.safetensors.")
    ap.add_argument("--batch-size", type=int, default=32, help="Batch size.")
    ap.add_argument("--epochs", type=int, default=5, help="Fusion training epochs.")
    ap.add_argument("--arch", type=str, default="ViT-L-16-SigLIP-384", help="Backbone architecture name (dfd.engine.ARCHS).")
    return ap.parse_args(argv)


def main(argv=None):
    a = parse_args(argv)
    distributed.init_from_env()
    train_fusion_head(a.real_dir, a.fake_dir, a.best_model, a.freq_mlp, a.fusion_out, a.batch_size, a.epochs, arch=a.arch)


if __name__ == "__main__":
    main()
