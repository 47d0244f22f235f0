"""CiFake evaluation driver (BASELINE config 4): `evaluate` of cifake_binary_classifier.py:893-953 on top of
`dropin.FastBinaryClassifier` — 32x32 images resampled to the model resolution INSIDE the model (bilinear,
align_corners=False, :716-717; here: inside the patch kernel), head H-D, BCE-with-logits loss, sigmoid, and the
reference's metric tuple.

The reference moves every batch to the device as float NCHW (`channels_last`), runs the model under autocast, and does
one `.item()` and two `.cpu()` reads per batch.  Here a batch may be float NCHW (what the reference's loaders yield) or
uint8 NHWC (4x fewer bytes over PCIe; ToTensor + Normalize happen in the patch kernel); logits stay on the device and
are read back ONCE at the end.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops


def test_time_augmentation(model, images: torch.Tensor, device=None, gpu_transform=None, n_tta: int = 5,
                           generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """cifake_binary_classifier.py:755-786: base logits + n_tta randomly flipped / 95 %-cropped views, averaged.
    The mirrored view costs no copy (DFD_FLIP_H read in the patch kernel); a crop is a strided view that the in-model
    bilinear resample (the same one that takes 32 -> S) scales back up, instead of a separate F.interpolate to (h, w)
    followed by the model's own resize — so cropped views are close to, not bit-equal with, the reference's."""
    preds = [model(images)]
    flags0 = model.tta_flags
    for _ in range(n_tta):
        flip = bool(torch.rand(1, generator=generator) > 0.5)
        view = images
        if bool(torch.rand(1, generator=generator) > 0.5):
            h, w = (images.shape[1:3] if images.dtype == torch.uint8 else images.shape[-2:])
            cs = int(0.95 * min(h, w))
            top = int(torch.randint(0, h - cs + 1, (1,), generator=generator))
            left = int(torch.randint(0, w - cs + 1, (1,), generator=generator))
            view = (images[:, top:top + cs, left:left + cs] if images.dtype == torch.uint8
                    else images[:, :, top:top + cs, left:left + cs]).contiguous()
        model.tta_flags = ops.FLIP_H if flip else 0
        try:
            preds.append(model(view))
        finally:
            model.tta_flags = flags0
    return torch.stack(preds).mean(dim=0)


@torch.no_grad()
def evaluate(model, dataloader, criterion=None, device=None, gpu_transform=None, use_tta: bool = False, ema=None):
    """cifake_binary_classifier.py:893-953.  Returns the reference's tuple
    (avg_loss, accuracy, balanced_acc, precision, recall, f1, auc, mcc, confusion_matrix, labels, probs).
    `criterion` defaults to BCE-with-logits (the reference passes nn.BCEWithLogitsLoss); it is applied per batch and
    averaged over batches like the reference's running `total_loss / len(dataloader)`.  `ema` follows the reference's
    protocol (`apply_shadow()` before, `restore()` after) when given."""
    from sklearn.metrics import (accuracy_score, balanced_accuracy_score, confusion_matrix, matthews_corrcoef,
                                 precision_recall_fscore_support, roc_auc_score)

    if ema is not None:
        ema.apply_shadow()
    model.eval()
    dev = model.device
    crit = criterion if criterion is not None else torch.nn.BCEWithLogitsLoss()
    logits_all, labels_all, losses = [], [], []
    for images, labels in dataloader:
        images = images.to(dev, non_blocking=True)
        labels = torch.as_tensor(labels).to(dev, non_blocking=True)
        if gpu_transform is not None:
            images = gpu_transform(images)
        z = test_time_augmentation(model, images, dev, gpu_transform, n_tta=5) if use_tta else model(images)
        losses.append(crit(z, labels.float()))
        logits_all.append(z)
        labels_all.append(labels)
    if ema is not None:
        ema.restore()
    z = torch.cat(logits_all)
    # ONE device -> host read for the whole evaluation (the reference syncs three times per batch)
    probs = torch.sigmoid(z).float().cpu().numpy()
    y = torch.cat(labels_all).cpu().numpy()
    avg_loss = float(torch.stack(losses).mean().cpu())
    preds = (probs > 0.5).astype(int)
    precision, recall, f1, _ = precision_recall_fscore_support(y, preds, average="binary", zero_division=0)
    auc = roc_auc_score(y, probs) if len(np.unique(y)) > 1 else float("nan")
    return (avg_loss, accuracy_score(y, preds), balanced_accuracy_score(y, preds), precision, recall, f1, auc,
            matthews_corrcoef(y, preds), confusion_matrix(y, preds), y, probs)


@torch.no_grad()
def throughput_sweep(model, batches=(256, 512, 1024, 2048, 4096, 8192), side: int = 32, iters: int = 3, device=None):
    """BASELINE config 4's sweep: images/s of the 32 -> S path per batch size, uint8 NHWC inputs resident on the device,
    CUDA-event timed.  Returns {batch: images_per_second}."""
    dev = model.device
    out = {}
    for B in batches:
        x = torch.randint(0, 256, (B, side, side, 3), dtype=torch.uint8, device=dev)
        model(x)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            model(x)
        e1.record()
        torch.cuda.synchronize(dev)
        out[B] = B * iters / (e0.elapsed_time(e1) / 1e3)
    return out
