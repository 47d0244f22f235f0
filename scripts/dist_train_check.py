"""Head-only training under real NCCL data parallelism (BASELINE config 5), to be launched with torchrun on >= 2 B200s:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_train_check.py

Every rank trains the AdaptiveFusionHead and the FreqMLP on the same seeded data, taking its shard of each mini-batch and
all-reducing the gradient bucket; rank 0 then re-trains single-process (world-size-1 code path, same seeds) in the same
process and compares the final parameters: data parallelism must not change the result beyond fp32 summation order."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfd import distributed, train_freq, train_fusion  # noqa: E402


def data():
    rng = np.random.default_rng(0)
    n = 203
    y = torch.from_numpy((rng.random(n) > 0.5).astype(np.float32))
    zs = torch.from_numpy((rng.normal(0, 2, n) + 2.0 * (y.numpy() - 0.5)).astype(np.float32))
    zf = torch.from_numpy((rng.normal(0, 2, n) + 1.0 * (y.numpy() - 0.5)).astype(np.float32))
    feats = torch.from_numpy((rng.normal(0.2, 0.7, (n, 24)) + 0.6 * (y.numpy()[:, None] - 0.5)).astype(np.float32))
    return y, zs, zf, feats


def run(dev):
    y, zs, zf, feats = data()
    torch.manual_seed(11)
    head, _, auc1 = train_fusion.fit_fusion_head(zf, zs, y, batch_size=32, epochs=3, device=dev, verbose=False)
    torch.manual_seed(5)
    fm, _, auc2 = train_freq.fit_freq_mlp(feats, y, epochs=3, batch_size=8, lr=1e-3, device=dev, dropout=0.0, verbose=False)
    return head.flat.data.clone(), fm.flat.data.clone(), auc1, auc2


def main():
    rank, world, local = distributed.init_from_env("nccl")
    assert world >= 2, "launch with torchrun on at least 2 GPUs"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    h_dp, f_dp, a1, a2 = run(dev)
    # every rank must hold identical parameters after the identical all-reduced steps
    for t in (h_dp, f_dp):
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, t), "ranks diverged"
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        h_1, f_1, _, _ = run(dev)          # torch.distributed no longer initialised: single-process path
        eh, ef = float((h_dp - h_1).abs().max()), float((f_dp - f_1).abs().max())
        print(f"world {world}: fusion head max|dp - single| = {eh:.2e} (auc {a1:.3f}), FreqMLP = {ef:.2e} (auc {a2:.3f})")
        assert eh < 1e-4 and ef < 1e-3, (eh, ef)
        print("dist_train_check ok")


if __name__ == "__main__":
    main()
