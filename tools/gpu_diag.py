"""GPU-side diagnostics for the tcgen05 kernels: prints error structure (which rows/columns/k-slices are wrong)
so a descriptor or swizzle mistake can be localised from one run. Usage: python tools/gpu_diag.py [gemm|attn|all]"""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitcap_b200 import ops  # noqa: E402


def summarize(name, out, ref, tol):
    err = (out.float() - ref.float()).abs()
    bad = err > tol
    print("[%s] shape %s max_err %.4g mean_err %.4g bad %.2f%% ref_absmax %.3g out_absmax %.3g" % (
        name, tuple(out.shape), float(err.max()), float(err.mean()), 100.0 * float(bad.float().mean()),
        float(ref.abs().max()), float(out.float().abs().max())), flush=True)
    if bad.any():
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("   bad rows: n=%d first %s last %s" % (len(rows), rows[:8].tolist(), rows[-4:].tolist()))
        print("   bad cols: n=%d first %s last %s" % (len(cols), cols[:8].tolist(), cols[-4:].tolist()))
        # coarse 32x32 block map
        M, N = bad.shape
        bm = min(16, (M + 31) // 32)
        bn = min(16, (N + 31) // 32)
        for i in range(bm):
            line = ""
            for j in range(bn):
                blk = bad[i * 32:(i + 1) * 32, j * 32:(j + 1) * 32]
                f = float(blk.float().mean()) if blk.numel() else 0
                line += "." if f == 0 else ("#" if f > 0.9 else "+")
            print("   " + line)
        r0, c0 = int(rows[0]), int(cols[0])
        print("   sample out[%d,%d:%d]=%s" % (r0, c0, c0 + 6, out[r0, c0:c0 + 6].float().tolist()))
        print("   sample ref[%d,%d:%d]=%s" % (r0, c0, c0 + 6, ref[r0, c0:c0 + 6].float().tolist()))
    return not bool(bad.any())


def diag_gemm():
    dev = torch.device("cuda:0")
    ok = True
    for (M, N, K) in [(128, 64, 64), (128, 256, 64), (128, 256, 128), (256, 512, 768), (300, 200, 64), (1154, 2304, 768)]:
        for tile_n in (64, 128, 256):
            g = torch.Generator().manual_seed(M + N + K)
            a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
            w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(dev)
            out = torch.zeros(M, (N + 7) // 8 * 8, device=dev)
            try:
                ops.linear(a, w, None, out[:, :N], ldo=out.stride(0), impl="tc", tile_n=tile_n)
                torch.cuda.synchronize()
            except Exception:
                traceback.print_exc()
                return False
            ref = a.float() @ w.float().t()
            ok &= summarize("gemm_tc M%d N%d K%d tn%d" % (M, N, K, tile_n), out[:, :N], ref, 1e-3)
    # identity-like probes: A = one-hot rows -> out rows = rows of W^T (locates k-slice / swizzle errors)
    M, N, K = 128, 64, 64
    a = torch.zeros(M, K, dtype=torch.bfloat16, device=dev)
    for i in range(M):
        a[i, i % K] = 1.0
    w = (torch.arange(N * K, device=dev).float().view(N, K) / 64.0).to(torch.bfloat16)
    out = torch.zeros(M, N, device=dev)
    ops.linear(a, w, None, out, impl="tc", tile_n=64)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    ok &= summarize("gemm_tc onehot", out, ref, 1e-3)
    return ok


def diag_attn():
    dev = torch.device("cuda:0")
    ok = True
    for (B, N, heads) in [(1, 128, 1), (1, 64, 1), (1, 256, 1), (1, 577, 2), (2, 578, 12)]:
        g = torch.Generator().manual_seed(N)
        qkv = torch.randn(B, N, 3 * heads * 64, generator=g).to(torch.bfloat16).to(dev)
        out = torch.zeros(B, N, heads * 64, device=dev, dtype=torch.bfloat16)
        try:
            ops.attention(qkv, out, B, N, heads, 0.125)
            torch.cuda.synchronize()
        except Exception:
            traceback.print_exc()
            return False
        q, k, v = qkv.float().view(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
        a = torch.softmax((q @ k.transpose(-1, -2)) * 0.125, dim=-1)
        ref = (a @ v).transpose(1, 2).reshape(B, N, heads * 64)
        ok &= summarize("attn_tc B%d N%d h%d" % (B, N, heads), out.view(B * N, -1), ref.view(B * N, -1), 3e-2)
    return ok


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
    res = {}
    if what in ("gemm", "all"):
        res["gemm"] = diag_gemm()
    if what in ("attn", "all"):
        try:
            res["attn"] = diag_attn()
        except Exception:
            traceback.print_exc()
            res["attn"] = False
    print("DIAG", res)
