"""Data parallelism over the GPUs of one box: one process per GPU, images are independent units.

  * inference: the global batch is split contiguously over ranks (no data-path collective), each rank runs the
    whole detection pipeline on its shard, then ONE all-gather of the packed [B_local,14] score records
    (NCCL over NVLink; tens of KB, latency bound) — SURVEY.md §8e.
  * head-only training: every rank computes the AdaptiveFusionHead loss/gradient partial sums of its shard with
    dfd_fusion_fwd_bwd (already scaled by 1/B_global), one all-reduce(sum) of the flat [196] bucket (195 grads +
    loss), then the identical clip + AdamW step on every rank.

The collective calls go through torch.distributed so the same host logic runs on `gloo` for the CPU tests.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the default process group if world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of n units: rank r gets [lo, hi); the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_records(local: torch.Tensor, counts: List[int] | None = None) -> torch.Tensor:
    """All-gather of [B_local, F] score records into [sum B, F] (rank order == image order for shard_bounds).
    Ragged shards are padded to the largest one for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if counts is None:
        c = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        cl = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(cl, c)
        counts = [int(t.item()) for t in cl]
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous())
    if all(cn == mx for cn in counts):
        return out
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], 0)


def all_reduce_sum_(bucket: torch.Tensor) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
    return bucket


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class WorkQueue:
    """A queue of independent work units (sub-batches of images) that every rank drains: an atomic counter in the process
    group's store (rank 0's TCPStore; one `add` per unit, ~0.1 ms against tens of ms of GPU work per unit).  Images are
    independent, so a GPU that settles at a lower power-capped clock simply takes fewer units: aggregate throughput is the
    SUM of the ranks' rates instead of world x the slowest one, and no per-step rendezvous is needed.  Without a process
    group: a local counter.  Keys must be unique per queue (every rank constructs the queue with the same key)."""

    def __init__(self, key: str):
        self.key, self._local, self._store = key, 0, None
        if dist.is_initialized() and dist.get_world_size() > 1:
            self._store = dist.distributed_c10d._get_default_store()

    def next(self) -> int:
        """Index of the next unclaimed unit (0, 1, 2, ... across all ranks; each index is handed out exactly once)."""
        if self._store is None:
            self._local += 1
            return self._local - 1
        return int(self._store.add(self.key, 1)) - 1
