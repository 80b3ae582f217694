#!/usr/bin/env python
"""Latency of one captioning call at small batch sizes: eager launches + decode-loop graph (the throughput path) against
graph_forward=True (the whole forward as one CUDA graph). Wall clock per call with a device synchronisation after every call
(what a serving caller sees), images already on the device, greedy 20 tokens, ViT-B/16-384, bf16.

    python tools/latency_probe.py [B ...]
    VITCAP_PROBE_EOS_BIAS=2.0 python tools/latency_probe.py 1 8     # EOS-planted weights: captions end after a few tokens and
                                                                    # the captured loop leaves through its conditional nodes
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import synth  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402

DEV = torch.device("cuda", 0)


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64]
    cfg = vcfg.variant("16_384")
    eos_bias = float(os.environ.get("VITCAP_PROBE_EOS_BIAS", "0"))
    sd = synth.make_state_dict(cfg, seed=0, eos_bias=eos_bias)
    print("eos_bias %.2f, early exit %s" % (eos_bias, os.environ.get("VITCAP_EARLY_EXIT", "1")), flush=True)
    models = {}
    for name, gf in (("eager+decode graph", False), ("whole-forward graph", True)):
        m = FastImageCaptioning(cfg, mode="bf16", max_batch=max(sizes), graph_forward=gf)
        m.load_state_dict(sd)
        models[name] = m.to(DEV)
    print("%-6s %-22s %10s %10s %10s" % ("batch", "path", "median ms", "min ms", "images/s"), flush=True)
    for B in sizes:
        data = {k: v.to(DEV) for k, v in synth.make_text_inputs(cfg, B).items()}
        imgs = [synth.make_images(cfg, B, seed=s).to(DEV) for s in range(4)]
        res = {}
        for name, m in models.items():
            for i in range(4):
                out = m(dict(data, image=imgs[i % 4]))
            torch.cuda.synchronize()
            ts = []
            for i in range(30):
                t0 = time.perf_counter()
                out = m(dict(data, image=imgs[i % 4]))
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            ts.sort()
            res[name] = out
            hit = out[0][:, 0] == 102
            mean_len = float(torch.where(hit.any(1), hit.float().argmax(1) + 1, torch.full_like(out[0][:, 0, 0], 20)).float().mean())
            print("%-6d %-22s %10.3f %10.3f %10.1f   mean caption tokens %.1f" % (B, name, ts[len(ts) // 2], ts[0],
                                                                              B / ts[len(ts) // 2] * 1e3, mean_len), flush=True)
        a, b = res["eager+decode graph"], res["whole-forward graph"]
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


if __name__ == "__main__":
    main()
