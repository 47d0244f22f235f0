"""What the residual path of the GEMM epilogue costs at the in-step size (B = 512 so400m images, M = 373 248 rows), measured
like the step runs: 20 back-to-back launches per variant (power-capped clock), CUDA events.  Development aid."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfd import ops  # noqa: E402

DEV = "cuda:0"
M = 373248


def run(name, n, k, **kw):
    a = torch.randn(M, k, device=DEV).to(torch.bfloat16)
    w = (torch.randn(n, k, device=DEV) / math.sqrt(k)).to(torch.bfloat16)
    bias = torch.randn(n, device=DEV)
    out = torch.randn(M, n, device=DEV).to(torch.bfloat16)
    stats = torch.empty(((n + 63) // 64, M, 2), device=DEV)
    args = dict(bias=bias, out=out)
    if kw.get("res"):
        args["residual"] = out
    if kw.get("stats"):
        args["stats_out"] = stats
    for _ in range(5):
        ops.gemm_bf16(a, w, **args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm_bf16(a, w, **args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name:28s} N={n} K={k}: {ms:7.3f} ms {2.0 * M * n * k / ms / 1e9:7.1f} TF/s  variant {ops.gemm_last_variant()}", flush=True)


for rep in range(2):
    for n, k, tag in ((1152, 4304, "fc2"), (1152, 1152, "out"), (3456, 1152, "qkv-shape")):
        run(f"{tag} bias only", n, k)
        run(f"{tag} bias + residual", n, k, res=True)
        run(f"{tag} bias + residual + stats", n, k, res=True, stats=True)
