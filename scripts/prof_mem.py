"""One launch of every memory-bound kernel of the detect step at so400m-384 / B=512 shapes, for an `ncu --set full`
capture (scripts/collect_profiles.sh).  Timing lives in scripts/kbench.py mem — never read times from a profiled run."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dfd import engine, ops, scoring, weights  # noqa: E402
from dfd.pipeline import head_params_from_state  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
arch = engine.ARCHS["siglip2-so400m-patch14-384"]
S, P, D, N, H = arch.image_size, arch.patch_size, arch.hidden_size, arch.tokens, arch.num_attention_heads
img = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device=DEV)
ops.patchify(img, S, P)
kv = torch.randn(B * N, 2 * D, device=DEV).to(torch.bfloat16)
ops.map_attention_bf16(kv, torch.randn(D, device=DEV), B, N, H, D // H)
gray = ops.gray256_from_rgb(img, True)
ops.freq_features(gray, scoring.build_freq_luts(torch.device(DEV)))
head = head_params_from_state(weights.random_classifier_head("B", D, 1), D, torch.device(DEV))
ops.head_fwd(head, torch.randn(B, D, device=DEV).to(torch.bfloat16))
torch.cuda.synchronize()
print("ok")
