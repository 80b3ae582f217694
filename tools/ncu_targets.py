"""Launches single hot kernels at the bench shapes (B=512) for `ncu --set full -k regex:...` captures.
usage: python tools/ncu_targets.py [attn|gemm_qkv|gemm_qkv_fold|gemm_fc1_fold|gemm_proj|gemm_fc1|gemm_fc2|x3_fc1|x3_fc2|x3_vocab|dattn|ln] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitcap_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "attn"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B = int(os.environ.get("VC_B", "512"))
M = B * 577


def gemm(N, K, act, resid, outf32):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    out = torch.randn(M, N, device=dev, dtype=torch.float32 if outf32 else torch.bfloat16)
    for _ in range(reps):
        ops.linear(a, w, b, out, act=act, resid=out if resid else None)


if what == "attn":
    qkv = torch.randn(B, 577, 2304, device=dev).to(torch.bfloat16)
    out = torch.empty(B, 577, 768, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        ops.attention(qkv, out, B, 577, 12, 0.125)
elif what == "gemm_qkv":
    gemm(2304, 768, 0, False, False)
elif what in ("gemm_qkv_fold", "gemm_fc2_emit"):
    hid = torch.randn(M, 3072, device=dev).to(torch.bfloat16)
    w2 = (torch.randn(768, 3072, device=dev) * 0.02).to(torch.bfloat16)
    b2 = torch.randn(768, device=dev)
    x = torch.randn(M, 768, device=dev)
    xb = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    stats = torch.empty(M, 3, 2, device=dev)
    wq = (torch.randn(2304, 768, device=dev) * 0.02).to(torch.bfloat16)
    bq, cq = torch.randn(2304, device=dev), torch.randn(2304, device=dev)
    qkv = torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        ops.linear_ln_emit(hid, w2, b2, x, x, xb, stats)
        ops.linear_ln_fold(xb, wq, bq, cq, stats, 3, 1e-6, qkv)
elif what in ("gemm_fc1_fold", "gemm_proj_emit"):
    # the bench's dominant pair inside a ViT block: proj + residual emitting (bf16 row, statistics) -> fc1 + GELU folding norm2
    att = torch.randn(M, 768, device=dev).to(torch.bfloat16)
    wp = (torch.randn(768, 768, device=dev) * 0.02).to(torch.bfloat16)
    bp = torch.randn(768, device=dev)
    x = torch.randn(M, 768, device=dev)
    xb = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    stats = torch.empty(M, 3, 2, device=dev)
    w1 = (torch.randn(3072, 768, device=dev) * 0.02).to(torch.bfloat16)
    b1, c1 = torch.randn(3072, device=dev), torch.randn(3072, device=dev)
    hid = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        ops.linear_ln_emit(att, wp, bp, x, x, xb, stats)
        ops.linear_ln_fold(xb, w1, b1, c1, stats, 3, 1e-6, hid, act=ops.ACT_GELU)
elif what == "gemm_proj":
    gemm(768, 768, 0, True, True)
elif what == "gemm_fc1":
    gemm(3072, 768, 1, False, False)
elif what == "gemm_fc2":
    gemm(768, 3072, 0, True, True)
elif what in ("x3_fc1", "x3_fc2", "x3_vocab"):
    # decode-step GEMMs on split-bf16 operands (K' = 3K): rows = 2B (token + MASK row per sequence) or B (MASK rows)
    Mx, N, K, act, resid = {"x3_fc1": (2 * B, 3072, 768, 1, False), "x3_fc2": (2 * B, 768, 3072, 0, True),
                            "x3_vocab": (B, 30522, 768, 0, False)}[what]
    a = torch.randn(Mx, 3 * K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, 3 * K, device=dev) * 0.02).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    ldo = (N + 63) // 64 * 64
    out = torch.randn(Mx, ldo, device=dev)
    for _ in range(reps):
        ops.linear(a, w, b, out[:, :N], act=act, resid=out[:, :N] if resid else None, ldo=ldo)
elif what == "dattn":
    ctx = torch.randn(B, 578, 2304, device=dev).to(torch.bfloat16)
    sq = torch.randn(20, 2 * B, 2304, device=dev).to(torch.bfloat16)
    out = torch.empty(2 * B, 768, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        ops.decode_attention(ctx, sq, None, out, B, 578, 12, 1, 10, 0.125)
elif what == "ln":
    x = torch.randn(M, 768, device=dev)
    g, b = torch.randn(768, device=dev), torch.randn(768, device=dev)
    o = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    for _ in range(reps):
        ops.layernorm(x, g, b, 1e-6, out_t=o)
torch.cuda.synchronize()
