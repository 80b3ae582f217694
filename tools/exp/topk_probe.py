"""tag_topk (sigmoid -> top-50 of 30522 concept logits per image, modeling_bert.py:1429-1432) at the bench shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tools.gpu_perf_probe import timeit  # noqa: E402
from vitcap_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, V, K = 512, 30522, 50
torch.manual_seed(0)
lg = torch.randn(B, 30528, device=dev) * 2.0 - 4.0
idx = torch.zeros(B, K, device=dev, dtype=torch.int32)
pr = torch.zeros(B, K, device=dev)
ln = torch.zeros(B, device=dev, dtype=torch.int32)
for r in range(3):
    ms = timeit(lambda: ops.tag_topk(lg, V, K, 0.2, idx, pr, ln), iters=50, warm=5)
    print("tag_topk B=%d V=%d K=%d: %.1f us" % (B, V, K, ms * 1e3), flush=True)
ref = lg[:, :V].topk(K, dim=1)          # (sigmoid is monotonic; its fp32 rounding could tie neighbours)
print("indices equal torch.topk:", bool(torch.equal(idx.long(), ref.indices)))
