"""Drop-in for "FreqMLP trainer.py" (SURVEY.md §8f.3): train the generation-2 FreqMLP on the 24-d frequency features and
write `freq_mlp.safetensors` in the reference's layout (normer.{mean,std}, contrast.{alpha,beta}, band.gates,
blocks.{0,1}.{norm,fc1,fc2}.*, head.*, temp.T — 6 494 parameters + 2 buffers).

Differences from the reference, on purpose:
  * feature extraction ("FreqMLP trainer.py":212-217: one PIL image at a time, FFT + SRM on the CPU) runs batched on the
    GPU kernels (`FreqFeatureExtractor`); gray256 (CLAHE + bicubic) is bit-exact either way;
  * the model's forward + backward is one CUDA kernel (dfd_freqmlp_fwd_bwd) that returns the mini-batch's loss / gradient
    partial sums; with torch.distributed initialised every rank takes a contiguous shard of each mini-batch and one
    all-reduce of the flat 6 495-float bucket precedes the identical clip(5.0) + AdamW step;
  * dropout masks come from a counter hash instead of torch's Philox stream (same rate, different bits).
"""
from __future__ import annotations

import argparse
import os
import random
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset

from . import distributed, ops
from .scoring import FreqFeatureExtractor, pil_to_gray256
from .train_fusion import _safe_auc, list_images


class TrainableFreqMLP:
    """The reference's FreqMLP ("FreqMLP trainer.py":275-301) with its parameters in ONE flat fp32 device tensor in
    state-dict order (ops.FREQMLP_PARAM_ORDER) plus the two FeatureNormalizer buffers."""

    NAMES = ops.FREQMLP_PARAM_ORDER
    SHAPES = ops.FREQMLP_PARAM_SHAPES

    def __init__(self, device=None, dropout: float = 0.05):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dropout = dropout
        # the reference's default initialisation, in its module construction order (same RNG consumption)
        sd: Dict[str, torch.Tensor] = {"contrast.alpha": torch.ones(24), "contrast.beta": torch.zeros(24),
                                       "band.gates": torch.zeros(4)}
        for b in range(2):
            fc1, fc2 = torch.nn.Linear(24, 64), torch.nn.Linear(64, 24)
            sd.update({f"blocks.{b}.norm.weight": torch.ones(24), f"blocks.{b}.norm.bias": torch.zeros(24),
                       f"blocks.{b}.fc1.weight": fc1.weight, f"blocks.{b}.fc1.bias": fc1.bias,
                       f"blocks.{b}.fc2.weight": fc2.weight, f"blocks.{b}.fc2.bias": fc2.bias})
        head = torch.nn.Linear(24, 1)
        sd.update({"head.weight": head.weight, "head.bias": head.bias, "temp.T": torch.tensor(1.0)})
        self.flat = torch.nn.Parameter(self._flatten(sd).to(self.device), requires_grad=False)
        self.mean = torch.zeros(24, device=self.device)
        self.std = torch.ones(24, device=self.device)
        self._step = 0

    @classmethod
    def _flatten(cls, sd) -> torch.Tensor:
        return torch.cat([sd[k].detach().float().reshape(-1).cpu() for k in cls.NAMES]).contiguous()

    def parameters(self):
        return [self.flat]

    def fit_normalization(self, feats: torch.Tensor):
        """FeatureNormalizer.fit ("FreqMLP trainer.py":225-227): mean, unbiased std + 1e-6 (forward adds 1e-6 again)."""
        self.mean = feats.float().mean(dim=0).to(self.device).contiguous()
        self.std = (feats.float().std(dim=0) + 1e-6).to(self.device).contiguous()

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out, o = {"normer.mean": self.mean.clone(), "normer.std": self.std.clone()}, 0
        for k, shp in zip(self.NAMES, self.SHAPES):
            n = int(np.prod(shp)) if shp else 1
            out[k] = self.flat.data[o:o + n].reshape(shp).clone()
            o += n
        return out

    def load_state_dict(self, sd, strict: bool = True):
        miss = [k for k in self.NAMES + ("normer.mean", "normer.std") if k not in sd]
        if miss:
            raise RuntimeError(f"missing keys: {miss}")
        self.flat.data.copy_(self._flatten(sd).to(self.device))
        self.mean = sd["normer.mean"].detach().float().to(self.device).contiguous()
        self.std = sd["normer.std"].detach().float().to(self.device).contiguous()
        return self

    @torch.no_grad()
    def forward(self, feats: torch.Tensor) -> torch.Tensor:
        """Eval-mode logits [B]."""
        x = feats.to(self.device, torch.float32).contiguous()
        return ops.freqmlp_fwd_bwd(self.flat.data, self.mean, self.std, x, want_grads=False)[2]

    __call__ = forward

    @torch.no_grad()
    def loss_and_grad(self, feats, y, global_batch: Optional[int] = None, train: bool = True):
        """(mean BCE-with-logits over the GLOBAL batch, d loss / d flat params), all-reduced over ranks."""
        n = global_batch if global_batch is not None else feats.shape[0]
        self._step += 1     # on every rank, also on one with an empty shard: the dropout streams stay aligned by step
        if feats.shape[0] > 0:
            # the kernel hashes (seed, LOCAL sample index, block, feature): fold the rank into the seed so that sample i of
            # every rank's shard does not share one dropout mask (the reference drops every sample independently)
            seed = (self._step * 2654435761 + distributed.rank() * 0x85EBCA6B) & 0xFFFFFFFF
            loss, grads, _ = ops.freqmlp_fwd_bwd(self.flat.data, self.mean, self.std, feats.contiguous(), y.contiguous(),
                                                 1.0 / n, self.dropout if train else 0.0, seed=seed)
        else:  # a rank may own an empty shard of a ragged last mini-batch
            loss = torch.zeros(1, device=self.device)
            grads = torch.zeros(ops.FREQMLP_NUM_PARAMS, device=self.device)
        bucket = torch.cat([grads, loss])
        distributed.all_reduce_sum_(bucket)
        return bucket[-1], bucket[:-1]


@torch.no_grad()
def extract_freq_matrix(paths: Sequence[str], device, batch_size: int = 64) -> torch.Tensor:
    """"FreqMLP trainer.py":212-217 batched: gray256 with CLAHE (host, the libraries themselves — images differ in
    size) then the GPU feature kernels.  Returns [len(paths), 24] on the CPU."""
    from PIL import Image

    fx = FreqFeatureExtractor(device, zscore=False, clahe=True)
    out = []
    for i in range(0, len(paths), batch_size):
        gs = []
        for p in paths[i:i + batch_size]:
            with Image.open(p) as pil:
                gs.append(pil_to_gray256(pil.convert("RGB"), clahe=True))
        out.append(fx.from_gray(torch.from_numpy(np.stack(gs)).to(fx.device)).float().cpu())
    return torch.cat(out) if out else torch.zeros(0, 24)


def fit_freq_mlp(features: torch.Tensor, labels: torch.Tensor, epochs: int = 100, batch_size: int = 8, lr: float = 1e-3,
                 device=None, model: Optional[TrainableFreqMLP] = None, dropout: float = 0.05, verbose: bool = True):
    """The training loop of "FreqMLP trainer.py":352-396 on a feature matrix.  Returns (model, best_state, best_auc)."""
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    model = model or TrainableFreqMLP(device=device, dropout=dropout)
    dev = model.device
    model.fit_normalization(features)
    loader = DataLoader(TensorDataset(features.float(), labels.float()), batch_size=batch_size, shuffle=True,
                        drop_last=False)  # identical permutation on every rank (same torch seed)
    optim = torch.optim.AdamW(model.parameters(), lr=lr)
    feats_dev = features.float().to(dev)
    best_auc, best_state = -float("inf"), None
    for ep in range(1, epochs + 1):
        losses = []
        for xb, yb in loader:
            lo, hi = distributed.shard_bounds(xb.shape[0], world, rank)
            loss, grads = model.loss_and_grad(xb[lo:hi].to(dev), yb[lo:hi].to(dev), global_batch=xb.shape[0])
            gnorm = grads.norm()                      # clip_grad_norm_(max_norm=5.0) after the all-reduce
            model.flat.grad = grads * torch.clamp(5.0 / (gnorm + 1e-6), max=1.0)
            optim.step()
            losses.append(loss)
        probs = torch.sigmoid(model(feats_dev)).cpu().numpy()
        auc = _safe_auc(labels.numpy(), probs)
        acc = float(((probs >= 0.5) == labels.numpy()).mean())
        if verbose and rank == 0:
            print(f"[freq] epoch {ep:03d}/{epochs} loss={float(torch.stack(losses).mean()):.4f} acc={acc:.3f} auc={auc:.3f}")
        if auc > best_auc:
            best_auc, best_state = auc, {k: v.cpu() for k, v in model.state_dict().items()}
    return model, best_state, best_auc


def prepare_paths(real_dir: str, fake_dir: str, limit: int, seed: int) -> Tuple[List[str], List[str]]:
    real_paths, fake_paths = list_images(real_dir), list_images(fake_dir)
    if not real_paths or not fake_paths:
        raise SystemExit("No images found under the provided directories.")
    rng = random.Random(seed)
    lim = lambda ps: ps if limit <= 0 or len(ps) <= limit else rng.sample(ps, limit)
    real_paths, fake_paths = lim(real_paths), lim(fake_paths)
    print(f"[data] using {len(real_paths)} real and {len(fake_paths)} fake samples")
    return real_paths, fake_paths


def train_freq_mlp(real_paths: Sequence[str], fake_paths: Sequence[str], epochs: int, batch_size: int, lr: float,
                   save_path: str, device="cuda"):
    from safetensors.torch import save_file

    dev = torch.device(device if device != "cuda" else f"cuda:{torch.cuda.current_device()}")
    print(f"[freq] extracting features for {len(real_paths)} real / {len(fake_paths)} fake images")
    real_feats, fake_feats = extract_freq_matrix(real_paths, dev), extract_freq_matrix(fake_paths, dev)
    features = torch.cat([real_feats, fake_feats], dim=0)
    labels = torch.cat([torch.zeros(len(real_feats)), torch.ones(len(fake_feats))])
    _, best_state, best_auc = fit_freq_mlp(features, labels, epochs, batch_size, lr, device=dev)
    if best_state is not None and save_path and (not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0):
        save_file({k: v.contiguous() for k, v in best_state.items()}, save_path)
        print(f"[freq] saved best model to {save_path} (AUC={best_auc:.3f})")
    return best_state


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="Train upgraded FreqMLP (24-d FFT+SRM) for Deepfake Detection v5")
    ap.add_argument("--real-dir", type=str, default=os.environ.get("REAL_DIR", ""), help="Folder that contains REAL images.")
    ap.add_argument("--fake-dir", type=str, default=os.environ.get("FAKE_DIR", ""), help="Folder that contains FAKE images.")
    ap.add_argument("--limit", type=int, default=int(os.environ.get("SAMPLE", 0)), help="Samples per class (0 = use all).")
    ap.add_argument("--epochs", type=int, default=100, help="Number of training epochs.")
    ap.add_argument("--batch-size", type=int, default=8, help="Batch size.")
    ap.add_argument("--lr", type=float, default=1e-3, help="Learning rate.")
    ap.add_argument("--freq-out", type=str, default="freq_mlp.safetensors", help="Output path for the freq MLP weights.")
    ap.add_argument("--seed", type=int, default=1337, help="Random seed.")
    return ap.parse_args(argv)


def main(argv=None):
    a = parse_args(argv)
    if not a.real_dir or not a.fake_dir:
        raise SystemExit("Please provide both --real-dir and --fake-dir (or set REAL_DIR / FAKE_DIR).")
    random.seed(a.seed)
    np.random.seed(a.seed)
    torch.manual_seed(a.seed)
    distributed.init_from_env()
    real_paths, fake_paths = prepare_paths(a.real_dir, a.fake_dir, a.limit, a.seed)
    train_freq_mlp(real_paths, fake_paths, a.epochs, a.batch_size, a.lr, a.freq_out)


if __name__ == "__main__":
    main()
