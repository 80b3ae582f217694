"""Experiment: does running the 512-image step as micro-batches on two streams (lanes) hide the HBM-bound kernels of one
lane under the tensor-bound kernels of the other?  Prints ms per 512 images for several splits."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from vitcap_b200 import config as vcfg, synth
from vitcap_b200.model import FastImageCaptioning

dev = torch.device("cuda", 0)
cfg = vcfg.variant("16_384", dec_layers=4)
sd = synth.make_state_dict(cfg, seed=0)
extra = synth.default_test_extra_input(cfg)
B = 512

def make(b):
    m = FastImageCaptioning(cfg, test_extra_input=extra, mode="bf16", max_batch=b)
    m.load_state_dict(sd)
    return m.to(dev)

img = synth.make_images(cfg, B, seed=1234).to(dev)
text = {k: v.to(dev) for k, v in synth.make_text_inputs(cfg, B).items()}

def run(chunks, lanes, steps=4, warm=2):
    """chunks: list of chunk sizes (sum = B); chunk i runs on lane i % lanes."""
    models = [make(max(chunks[i::lanes])) for i in range(lanes)]
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    offs = [sum(chunks[:i]) for i in range(len(chunks))]
    def step():
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(main)
        outs = []
        for i, (o, c) in enumerate(zip(offs, chunks)):
            l = i % lanes
            streams[l].wait_event(ev)
            with torch.cuda.stream(streams[l]):
                d = {k: v[o:o + c] for k, v in text.items()}
                d["image"] = img[o:o + c]
                outs.append(models[l](d))
        for s in streams:
            main.wait_stream(s)
        return outs
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        outs = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ids = torch.cat([o[0] for o in outs], 0)
    del models
    torch.cuda.empty_cache()
    return ms, ids

base_ms, base_ids = run([512], 1)
print("1 lane  [512]            %.2f ms  %.0f img/s" % (base_ms, B / base_ms * 1e3), flush=True)
for chunks, lanes in [([256, 256], 1), ([256, 256], 2), ([128] * 4, 2), ([192, 128, 192], 2), ([256, 128, 128], 2), ([171, 171, 170], 3),
                      ([128] * 4, 4), ([64] * 8, 2)]:
    ms, ids = run(chunks, lanes)
    print("%d lanes %-18s %.2f ms  %.0f img/s  ids equal to 1-lane: %.4f" % (lanes, chunks, ms, B / ms * 1e3,
          (ids == base_ids).float().mean().item()), flush=True)
