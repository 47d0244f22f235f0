"""Tiny driver for ncu: a few launches of selected kernels at so400m shapes (B=64)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfd import ops  # noqa: E402

DEV = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "outres"
M, D, I = int(os.environ.get("DFD_PROF_M", 46656)), 1152, 4304
torch.manual_seed(0)
if which in ("outres", "qkv", "fc1", "fc2"):
    n, k = {"outres": (D, D), "qkv": (3 * D, D), "fc1": (I, D), "fc2": (D, I)}[which]
    a = torch.randn(M, k, device=DEV).to(torch.bfloat16)
    w = (torch.randn(n, k, device=DEV) / math.sqrt(k)).to(torch.bfloat16)
    bias = torch.randn(n, device=DEV)
    res = torch.randn(M, n, device=DEV).to(torch.bfloat16) if which in ("outres", "fc2") else None
    out = torch.empty(M, n, dtype=torch.bfloat16, device=DEV)
    for _ in range(4):
        ops.gemm_bf16(a, w, bias=bias, act=1 if which == "fc1" else 0, residual=res, out=out)
elif which == "attn":
    B, N, H, hd = 64, 729, 16, 72
    qkv = torch.randn(B * N, 3 * H * hd, device=DEV).to(torch.bfloat16)
    impl = int(sys.argv[2]) if len(sys.argv) > 2 else None
    for _ in range(4):
        ops.attention_bf16(qkv, B, N, H, hd, impl=impl)
elif which == "engine":
    from dfd import engine, weights

    name = "siglip2-so400m-patch14-384"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    eng = engine.SiglipEngine(engine.ARCHS[name], 0, max_batch=B).load_state_dict(weights.random_vision_state_dict(engine.ARCHS[name], 0, DEV))
    img = torch.randint(0, 256, (B, 384, 384, 3), dtype=torch.uint8, device=DEV)
    for _ in range(2):
        eng(img)
torch.cuda.synchronize()
print("done", which)
