import sys, os, torch
sys.path.insert(0, "/root/repo")
from dfd import ops
sys.path.insert(0, "/root/repo/scripts")
import kbench
B, N, H, hd = 64, 729, 16, 72
qkv = torch.randn(B * N, 3 * H * hd, device="cuda:0").to(torch.bfloat16)
fl = 4.0 * N * N * H * hd * B
for sc in (None, -0.1, -10.0):
    med, best = kbench.timeit(lambda: ops.attention_bf16(qkv, B, N, H, hd, scale=sc, impl=3))
    print(f"impl 3 scale={sc}: {med:.3f} ms {fl/med/1e9:.1f} TF/s")
