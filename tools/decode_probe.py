"""Decode-loop probe (B images, greedy, 20 tokens): times the captured loop for a list of split-K settings, or -- `eager N` --
runs N eager decode steps after a warm-up so that `ncu -k regex:...` sees the loop's kernels one by one.
    python tools/decode_probe.py 512 sweep 3,6,6 2,6,6 6,6,12
    python tools/decode_probe.py 512 prec bf16x3 fp16 bf16        # captured loop and full forward per decode_precision
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_dec|finish_ln|token_step|decode_att|embed_ln" \
        -c 200 --csv --log-file gpurun_out/dec_launches.csv python tools/decode_probe.py 512 eager"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gpu_perf_probe import timeit  # noqa: E402
from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import synth  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402

dev = torch.device("cuda:0")


def precisions(B, names):
    cfg = vcfg.variant("16_384")
    sd = synth.make_state_dict(cfg, seed=0)
    data = {k: v.to(dev) for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = synth.make_images(cfg, B, seed=1).to(dev)
    print(torch.cuda.get_device_name(0), "B =", B, flush=True)
    models, ref = {}, None
    for n in names:
        m = FastImageCaptioning(cfg, mode="bf16", max_batch=B, decode_precision=n)
        m.load_state_dict(sd)
        models[n] = m.to(dev)
        ids, _ = models[n](data)
        models[n](data)
        ref = ids if ref is None else ref
        print("%s: %.4f of all tokens equal those of %s" % (n, float((ids == ref).float().mean()), names[0]), flush=True)
    for rnd in range(3):                                   # alternate: the boards drift under their power cap
        for n in names:
            m = models[n]
            t_dec = timeit(lambda: m.engine.greedy_or_sample(B, 1, 20, 101, 0, [102], 103), iters=8, warm=2)
            t_all = timeit(lambda: m(data), iters=4, warm=1)
            print("decode_precision=%-9s round %d: decode(graph) %.2f ms  full forward %.2f ms (%.1f images/s)  kernels=%s" % (
                n, rnd, t_dec, t_all, B / t_all * 1e3, m.engine.stats.get("graph_kernels")), flush=True)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    what = sys.argv[2] if len(sys.argv) > 2 else "sweep"
    if what == "prec":
        return precisions(B, sys.argv[3:] or ["bf16x3", "fp16", "bf16"])
    cfg = vcfg.variant("16_384")
    sd = synth.make_state_dict(cfg, seed=0)
    m = FastImageCaptioning(cfg, mode="bf16", max_batch=B, use_cuda_graph=(what != "eager"))
    m.load_state_dict(sd)
    m = m.to(dev)
    eng = m.engine
    data = {k: v.to(dev) for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = synth.make_images(cfg, B, seed=1).to(dev)
    ref_ids, ref_lp = m(data)
    torch.cuda.synchronize()
    if what == "eager":
        eng.greedy_or_sample(B, 1, 20, 101, 0, [102], 103)
        torch.cuda.synchronize()
        return
    configs = [tuple(int(v) for v in c.split(",")) for c in sys.argv[3:]] or [eng.dec_splits]
    print(torch.cuda.get_device_name(0), "B =", B, flush=True)
    for rnd in range(2):
        for c in configs:
            eng.dec_splits = c
            eng._dec_ws.clear()
            ids, lp = m(data)
            m(data)
            torch.cuda.synchronize()
            t_dec = timeit(lambda: eng.greedy_or_sample(B, 1, 20, 101, 0, [102], 103), iters=8, warm=2)
            print("splits o,f,t = %s round %d: decode(graph) %.2f ms  ids_equal=%s max|dlp| %.2g  kernels=%s" % (
                c, rnd, t_dec, bool(torch.equal(ids, ref_ids)), float((lp - ref_lp).abs().max()), eng.stats.get("graph_kernels")), flush=True)


if __name__ == "__main__":
    main()
