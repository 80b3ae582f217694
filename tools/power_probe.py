#!/usr/bin/env python
"""Steady-state throughput, SM clock and board power of single kernels (each looped for ~2 s, nvidia-smi sampled every
100 ms): energy per flop of the hand-written GEMM / attention kernels next to cuBLAS / cuDNN on the same shapes. Explains
which kernels run into the board's power cap and how efficiently they use it.  usage: python tools/power_probe.py [B]"""
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitcap_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
M = B * 577


class Smi:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "100"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._rd, daemon=True).start()

    def _rd(self):
        for line in self.p.stdout:
            try:
                a, b = line.split(",")
                self.rows.append((time.time(), float(a), float(b)))
            except ValueError:
                pass

    def window(self, t0, t1):
        r = [(c, p) for t, c, p in self.rows if t0 <= t <= t1]
        if not r:
            return float("nan"), float("nan")
        return statistics.median(x[0] for x in r), statistics.median(x[1] for x in r)


def run(name, fn, flops, smi, seconds=2.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    n = max(5, int(seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
    t0 = time.time()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1) / n
    clk, pw = smi.window(t0 + 0.5, t1)
    tf = flops / ms / 1e9
    print("%-34s %8.3f ms %8.1f TFLOP/s  SM %6.0f MHz  %6.0f W  %6.3f pJ/flop" % (name, ms, tf, clk, pw, pw / (tf * 1e12) * 1e12 if tf else 0),
          flush=True)
    time.sleep(1.0)


def main():
    smi = Smi()
    time.sleep(0.5)
    print(torch.cuda.get_device_name(0), "B =", B)
    for (N, K, act, resid, outf32, name) in [(2304, 768, 0, False, False, "qkv"), (3072, 768, 1, False, False, "fc1+gelu"),
                                             (3072, 768, 0, False, False, "fc1 (no gelu)"), (768, 3072, 0, True, True, "fc2+res"),
                                             (768, 768, 0, True, True, "proj+res")]:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
        b = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.float32 if outf32 else torch.bfloat16)
        if resid:
            out.normal_()
        fl = 2.0 * M * N * K
        run("ours   " + name, lambda: ops.linear(a, w, b, out, act=act, resid=out if resid else None), fl, smi)
        if name == "proj+res":
            for tn in (256, 512):
                run("ours   proj+res tile_n=%d" % tn,
                    lambda: ops.linear(a, w, b, out, act=act, resid=out, impl="tc", tile_n=tn), fl, smi)
        ob = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        run("cublas " + name.split()[0] + " (plain bf16)", lambda: torch.matmul(a, w.t(), out=ob), fl, smi)
        del a, w, out, ob
    qkv = torch.randn(B, 577, 2304, device=dev).to(torch.bfloat16)
    out = torch.empty(B, 577, 768, device=dev, dtype=torch.bfloat16)
    fl = 4.0 * B * 12 * 577 * 577 * 64
    run("ours   attention", lambda: ops.attention(qkv, out, B, 577, 12, 0.125), fl, smi)
    q, k, v = qkv.view(B, 577, 3, 12, 64).permute(2, 0, 3, 1, 4)
    run("cudnn  SDPA", lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), fl, smi)
    x = torch.randn(M, 768, device=dev)
    g, bb = torch.randn(768, device=dev), torch.randn(768, device=dev)
    o = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    run("ours   layernorm (GB/s as TFLOP)", lambda: ops.layernorm(x, g, bb, 1e-6, out_t=o), M * 768 * 6 * 1e3, smi)
    smi.p.terminate()


if __name__ == "__main__":
    main()
