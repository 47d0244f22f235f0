"""CPU tests of the multi-rank host logic on the gloo backend (world_size 2): contiguous sharding, ragged
all-gather of score records, gradient-bucket all-reduce == full-batch gradient."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank)})
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from dfd import distributed
    from oracle import scoring_ref as S

    r, w, _ = distributed.init_from_env("gloo")
    assert (r, w) == (rank, world)
    res = {}
    # 1. ragged all-gather keeps image order
    n = 11
    lo, hi = distributed.shard_bounds(n, world, rank)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 14) + 0.5
    full = distributed.all_gather_records(local)
    res["gather_ok"] = bool(torch.equal(full[:, 0], torch.arange(n, dtype=torch.float32) + 0.5)) and full.shape == (n, 14)
    # 2. equal shards take the single-collective path
    full2 = distributed.all_gather_records(torch.full((4, 3), float(rank)))
    res["gather_eq_ok"] = full2.shape == (8, 3) and float(full2[:4].mean()) == 0.0 and float(full2[4:].mean()) == 1.0
    # 3. DP gradient identity: sum over ranks of shard partial sums (scaled by 1/B_global) == full-batch gradient
    rng = np.random.default_rng(0)
    B = 37
    zf, zs = rng.normal(0, 2, B).astype(np.float32), rng.normal(0, 2, B).astype(np.float32)
    y = (rng.random(B) > 0.5).astype(np.float32)
    sd = S.init_fusion_g2(3)
    lo, hi = distributed.shard_bounds(B, world, rank)
    l_loc, g_loc, _ = S.fusion_loss_and_grads(sd, zf[lo:hi], zs[lo:hi], y[lo:hi])
    nloc = hi - lo
    bucket = torch.cat([torch.from_numpy(g_loc) * nloc / B, torch.tensor([l_loc * nloc / B], dtype=torch.float64)])
    distributed.all_reduce_sum_(bucket)
    l_full, g_full, _ = S.fusion_loss_and_grads(sd, zf, zs, y)
    res["grad_err"] = float(np.abs(bucket[:195].numpy() - g_full).max())
    res["loss_err"] = abs(float(bucket[195]) - l_full)
    # 4. the same identity for the FreqMLP trainer's 6 495-float bucket
    f = rng.normal(0.3, 0.8, (B, 24)).astype(np.float32)
    fsd = S.init_freq_mlp_g2(7)
    lf_loc, gf_loc, _ = S.freq_mlp_g2_loss_and_grads(fsd, f[lo:hi], y[lo:hi])
    fb = torch.cat([torch.from_numpy(gf_loc) * nloc / B, torch.tensor([lf_loc * nloc / B], dtype=torch.float64)])
    distributed.all_reduce_sum_(fb)
    lf_full, gf_full, _ = S.freq_mlp_g2_loss_and_grads(fsd, f, y)
    res["fgrad_err"] = float(np.abs(fb[:-1].numpy() - gf_full).max())
    res["floss_err"] = abs(float(fb[-1]) - lf_full)
    res["max"] = distributed.max_over_ranks(float(rank + 1), "cpu")
    # 5. work queue: the ranks drain 23 units between them, each unit exactly once; the records of a unit are written by its
    #    owner into a zero-filled slab and ONE sum over ranks reassembles them (bench.py's dynamic all-gather)
    wq = distributed.WorkQueue("test_queue")
    slab = torch.zeros(23, 3)
    taken = []
    while True:
        u = wq.next()
        if u >= 23:
            break
        taken.append(u)
        slab[u] = torch.tensor([u + 0.25, rank, 1.0])
    distributed.all_reduce_sum_(slab)
    res["queue_ok"] = bool(torch.equal(slab[:, 0], torch.arange(23.0) + 0.25)) and bool((slab[:, 2] == 1.0).all())
    res["queue_taken"] = taken
    distributed.barrier()
    q.put((rank, res))
    dist.destroy_process_group()


def test_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert out[r]["gather_ok"] and out[r]["gather_eq_ok"]
        assert out[r]["grad_err"] < 1e-12 and out[r]["loss_err"] < 1e-12
        assert out[r]["fgrad_err"] < 1e-12 and out[r]["floss_err"] < 1e-12
        assert out[r]["max"] == 2.0
        assert out[r]["queue_ok"]
    assert sorted(out[0]["queue_taken"] + out[1]["queue_taken"]) == list(range(23))


def test_work_queue_without_process_group_is_a_local_counter():
    from dfd import distributed

    q = distributed.WorkQueue("solo")
    assert [q.next() for _ in range(4)] == [0, 1, 2, 3]


def test_shard_bounds_cover_everything():
    from dfd import distributed

    for n in (0, 1, 7, 512, 513):
        for w in (1, 2, 3, 8):
            spans = [distributed.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bench_unit_plan_covers_every_image_once_with_a_short_tail():
    """bench.py:unit_plan — the work queue's units: contiguous, each aligned to its own size (so that it is a slice of a rank's
    resident batch), whole batches on one GPU, and on several GPUs a tail of half- and quarter-size units (2 x world each)."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                          "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for world, steps, B in ((1, 5, 512), (2, 5, 512), (4, 3, 512), (8, 20, 512), (8, 5, 256), (3, 2, 512)):
        sub = B if world == 1 else B // 2
        plan = bench.unit_plan(steps * world * B, world, B, sub)
        assert plan[0][0] == 0 and plan[-1][0] + plan[-1][1] == steps * world * B
        assert all(a[0] + a[1] == b[0] for a, b in zip(plan, plan[1:]))
        assert all(off % n == 0 and off % B + n <= B for off, n in plan)
        sizes = [n for _, n in plan]
        if world == 1:
            assert sizes == [B] * steps
        else:
            assert sizes.count(sub // 2) == 2 * world and sizes.count(sub // 4) >= 2 * world   # (a remainder goes out as quarter units)
            assert sizes == sorted(sizes, reverse=True)      # big units first, the short ones at the end
