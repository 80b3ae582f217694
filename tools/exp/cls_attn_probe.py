"""cls_attention (one query row per image over the packed qkv buffer, the CLS-only last concept block) at the bench shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tools.gpu_perf_probe import timeit  # noqa: E402
from vitcap_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, N, heads = 512, 577, 12
qkv = torch.randn(B, N, 2304, device=dev).to(torch.bfloat16)
q = torch.randn(B, 768, device=dev).to(torch.bfloat16)
out = torch.empty(B, 768, device=dev, dtype=torch.bfloat16)
for r in range(3):
    ms = timeit(lambda: ops.cls_attention(q, qkv, out, B, N, heads, 0.125), iters=50, warm=5)
    print("cls_attention B=%d N=%d: %.1f us  (%.0f GB/s of K+V)" % (B, N, ms * 1e3, B * N * 1536 * 2 / ms / 1e6), flush=True)
