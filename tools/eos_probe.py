"""Early exit of the decode loop (modeling_utils.py:865-867) on EOS-planted weights: for a list of eos_bias values prints the
mean caption length, the captured decode loop's time and the full step's time (B images, greedy, 20 tokens). Finished captions
are skipped by the decode-step attention (the loop's HBM-bound 60 %), so the loop time follows the mean length.
    python tools/eos_probe.py 512 0 0.5 1.0 1.5 2.0"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gpu_perf_probe import timeit  # noqa: E402
from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import synth  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402

dev = torch.device("cuda:0")


def caption_lengths(ids, eos=102):
    """Tokens produced per caption, BOS and the closing EOS included (ids [R, 1, L])."""
    row = ids[:, 0]
    hit = row == eos
    first = torch.where(hit.any(1), hit.float().argmax(1) + 1, torch.full_like(row[:, 0], row.shape[1]))
    return first.float()


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    biases = [float(v) for v in sys.argv[2:]] or [0.0, 1.0, 2.0]
    cfg = vcfg.variant("16_384")
    m = FastImageCaptioning(cfg, mode="bf16", max_batch=B)
    m.load_state_dict(synth.make_state_dict(cfg, seed=0))
    m = m.to(dev)
    data = {k: v.to(dev) for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = synth.make_images(cfg, B, seed=1).to(dev)
    print(torch.cuda.get_device_name(0), "B =", B, flush=True)
    for eb in biases:
        m.load_state_dict(synth.make_state_dict(cfg, seed=0, eos_bias=eb))
        ids, _ = m(data)
        torch.cuda.synchronize()
        ln = caption_lengths(ids)
        eng = m.engine
        for early in (True, False, True, False):
            eng.early_exit = early
            eng._dec_ws.clear()                          # drop the captured loops: the next call captures with / without IF nodes
            m(data)
            m(data)
            t_dec = timeit(lambda: eng.greedy_or_sample(B, 1, 20, 101, 0, [102], 103), iters=8, warm=2)
            t_all = timeit(lambda: m(data), iters=8, warm=2)
            print("eos_bias %.3f early_exit=%d: mean length %.2f (min %d max %d, %d of %d reach 20)  decode loop %.2f ms  full step %.2f ms "
                  "(%.1f images/s)" % (eb, early, float(ln.mean()), int(ln.min()), int(ln.max()), int((ln >= 20).sum()), B, t_dec, t_all,
                                       B / t_all * 1e3), flush=True)
        eng.early_exit = True


if __name__ == "__main__":
    main()
