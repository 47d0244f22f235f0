"""Where do the N = 1152 residual GEMMs lose against qkv / fc1?  Times the fc2 (K = 4304) and out-projection (K = 1152) shapes
with and without the residual epilogue and with N = 1024 / 1152 / 1280 (no tail / half-width tail / five full tiles), CUDA
events, L2 flushed.  A development aid like kbench.py."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from dfd import ops  # noqa: E402
from kbench import DEV, timeit  # noqa: E402


def main():
    images = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    m = images * 729
    for k in (4304, 1152):
        a = torch.randn(m, k, device=DEV).to(torch.bfloat16)
        for n in (1024, 1152, 1280):
            w = (torch.randn(n, k, device=DEV) / math.sqrt(k)).to(torch.bfloat16)
            bias = torch.randn(n, device=DEV)
            out = torch.empty(m, n, dtype=torch.bfloat16, device=DEV)
            res = torch.randn(m, n, device=DEV).to(torch.bfloat16)
            stats = torch.empty((n + 63) // 64, m, 2, device=DEV)
            for label, kw in (("bias", {}), ("bias+res", {"residual": res}), ("bias+res+stats", {"residual": res, "stats_out": stats})):
                med, best = timeit(lambda: ops.gemm_bf16(a, w, bias=bias, out=out, **kw), iters=7)
                tf = 2.0 * m * n * k / med / 1e9
                print(f"M={m} K={k} N={n} {label:15s} {med:7.3f} ms {tf:7.1f} TF/s  best {best:.3f}  variant {ops.gemm_last_variant()}",
                      flush=True)
            med, _ = timeit(lambda: torch.matmul(a, w.t(), out=out), iters=7)
            print(f"M={m} K={k} N={n} cuBLAS          {med:7.3f} ms {2.0 * m * n * k / med / 1e9:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
