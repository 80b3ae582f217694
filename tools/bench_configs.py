#!/usr/bin/env python
"""Device-resident timings of the BASELINE.json configurations that are NOT the bench line (bench.py times configs[2]):

    configs[1]  encoder + concept head top-50 only, batch 256           (images/s and encoder TFLOP/s)
    configs[2]  greedy captioning, batch 512, with a 12-layer decoder   (the BASELINE.json wording; 4 layers is what ships)
    configs[3]  beam search, 4 beams, batch 256
    configs[4]  sampling, 5 sequences per image, batch 512, 16_224 variant (+ the greedy baseline pass SCST also runs)

One JSON line per configuration on stdout. Synthetic images resident in HBM, random-init weights, bf16 operands; W warm-up
calls, K timed calls between CUDA events (inputs and activations are far larger than L2).

    python tools/bench_configs.py [--only 1,2,3,4] [--steps 5] [--warmup 3]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import synth  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402

DEV = torch.device("cuda", 0)


def timed(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def build(variant, B, dec_layers=4, **extra_kw):
    cfg = vcfg.variant(variant, dec_layers=dec_layers)
    sd = synth.make_state_dict(cfg, seed=0)
    extra = synth.default_test_extra_input(cfg, **extra_kw)
    m = FastImageCaptioning(cfg, test_extra_input=extra, mode="bf16", max_batch=B)
    m.load_state_dict(sd)
    m = m.to(DEV)
    data = {k: v.to(DEV) for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = synth.make_images(cfg, B, seed=1234).to(DEV)
    return cfg, m, data


def block_gflop(cfg):
    n, h, f = cfg.n_tokens, cfg.hidden, cfg.inter
    return (2 * n * h * 3 * h + 4 * n * n * h + 2 * n * h * h + 4 * n * h * f) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="1,2,3,4")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    only = {int(x) for x in args.only.split(",")}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained") or 1400.0

    def emit(**kw):
        kw.update(steps=args.steps, warmup=args.warmup, dtype="bf16", data="synthetic, device resident", n_gpus=1)
        print(json.dumps(kw), flush=True)

    if 1 in only:
        B = 256
        cfg, m, data = build("16_384", B)
        ms = timed(lambda: m.forward_tags(data["image"], caption_branch=False), args.warmup, args.steps)
        # what runs: patch embed + 8 shared blocks + 4 concept blocks (the last one for the CLS row only: K/V projection of
        # all rows, everything else for one row) + pooler + concept head; the caption branch is skipped (tags only)
        n, h = cfg.n_tokens, cfg.hidden
        g = 2 * cfg.n_patches * cfg.patch_dim * h / 1e9 + 11 * block_gflop(cfg) + 2 * n * h * 2 * h / 1e9 + 0.049
        emit(config="BASELINE.json configs[1]: ViT-B/16-384 encoder + concept head top-50 only (tags only: 8 shared + 4 concept "
                    "blocks, caption branch skipped), batch 256", metric="images/s", value=B / ms * 1e3, ms_per_step=ms,
             algorithmic_gflop_per_image=g, tflops=g * B / ms, frac_of_sustained_bf16_peak=g * B / ms / peak)
        ms2 = timed(lambda: m.forward_tags(data["image"]), args.warmup, args.steps)
        g2 = 2 * cfg.n_patches * cfg.patch_dim * h / 1e9 + 15 * block_gflop(cfg) + 2 * n * h * 2 * h / 1e9 + 0.049
        emit(config="BASELINE.json configs[1], both branches (16 blocks: what the captioner needs; SURVEY.md 8d's 147.78 GFLOP/img "
                    "less the CLS-only shortcut of the last concept block), batch 256", metric="images/s", value=B / ms2 * 1e3,
             ms_per_step=ms2, algorithmic_gflop_per_image=g2, tflops=g2 * B / ms2, frac_of_sustained_bf16_peak=g2 * B / ms2 / peak)
        del m
    if 2 in only:
        B = 512
        cfg, m, data = build("16_384", B, dec_layers=12)
        ms = timed(lambda: m(data), args.warmup, args.steps)
        emit(config="BASELINE.json configs[2] with a 12-layer decoder (config.decoder_layer = 12; the shipped model has 4, "
                    "modeling_bert.py:1342-1346): greedy, batch 512, max_len 20", metric="images/s", value=B / ms * 1e3,
             ms_per_step=ms)
        del m
    if 3 in only:
        B = 256
        cfg, m, data = build("16_384", B, num_beams=4)
        ms = timed(lambda: m(data), args.warmup, args.steps)
        emit(config="BASELINE.json configs[3]: beam search, 4 beams, num_keep_best 1, length_penalty 1, batch 256, max_len 20",
             metric="images/s", value=B / ms * 1e3, ms_per_step=ms)
        del m
    if 4 in only:
        B = 512
        cfg, m, data = build("16_224", B, do_sample=True, num_return_sequences=5)
        ms = timed(lambda: m(data), args.warmup, args.steps)
        emit(config="BASELINE.json configs[4]: sampling, 5 sequences per image, top_k 0, top_p 1, temperature 1, batch 512, "
                    "16_224 variant", metric="images/s", value=B / ms * 1e3, ms_per_step=ms, sequences_per_s=5 * B / ms * 1e3)
        del m
        cfg, m, data = build("16_224", B)
        ms = timed(lambda: m(data), args.warmup, args.steps)
        emit(config="BASELINE.json configs[4], the greedy baseline pass of SCST: batch 512, 16_224 variant", metric="images/s",
             value=B / ms * 1e3, ms_per_step=ms)


if __name__ == "__main__":
    main()
