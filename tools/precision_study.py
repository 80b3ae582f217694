"""CPU study (no GPU, no kernels): how many greedy tokens of the fast mode's ARITHMETIC agree with the fp32 reference algorithm
for different operand formats of the decode-step MLP and the vocabulary head. The fast mode is represented by its executable
spec, oracle/port.py QuantPortModel (pinned to the kernels by tests/test_fullsize_gpu.py); the variants only replace its
`lin_x3` (the GEMMs of BertIntermediate / BertOutput of a decode step and of BertLMPredictionHead):

    bf16x3   split-bf16 operands, three products (decode_precision='bf16x3', the default until session 4 of round 2)
    bf16     plain bf16 operands, one product (decode_precision='bf16')
    fp16     operands rounded to IEEE half (11-bit significand), one product -- same tensor-core rate and bytes as plain bf16
             (decode_precision='fp16', the default)
    fp16x2   activations split into two halves (hi + lo), weights one half: two products
    fp16all  NOT a shipped mode: EVERY bf16 rounding point of the fast mode (encoder, prefill, K/V cache, attention
             probabilities, decode steps) replaced by an IEEE-half rounding -- what half-precision storage of all operands would
             buy (candidate for a later round; the kernels store bf16 today)

Counting as tests/test_fullsize_gpu.py::test_bf16_mode_token_agreement_fullsize_vs_oracle (same weights, same images):
tokens produced under an identical prefix, every row up to and including its first divergence.

    python tools/precision_study.py [images=96] [chunk=16] [variants=bf16x3,bf16,fp16,fp16x2]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port  # noqa: E402
from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import synth  # noqa: E402


def q_f16(x):
    return x.half().float()


class Fp16Decode(port.QuantPortModel):
    def lin_x3(self, a, wkey, bkey):
        key = ("f16", wkey)
        if key not in self._wq:
            self._wq[key] = q_f16(self.sd[wkey])
        out = F.linear(q_f16(a), self._wq[key])
        return out + self.sd[bkey] if bkey else out


class Fp16x2Decode(port.QuantPortModel):
    def lin_x3(self, a, wkey, bkey):
        key = ("f16", wkey)
        if key not in self._wq:
            self._wq[key] = q_f16(self.sd[wkey])
        hi = q_f16(a)
        lo = q_f16(a - hi)
        out = F.linear(hi, self._wq[key]) + F.linear(lo, self._wq[key])
        return out + self.sd[bkey] if bkey else out


class AllHalf(port.QuantPortModel):
    """Every operand rounding of QuantPortModel as an IEEE-half rounding (port.q_bf16 is swapped while this model runs)."""

    def run(self, fn):
        saved = port.q_bf16
        port.q_bf16 = port.q_f16
        try:
            return fn()
        finally:
            port.q_bf16 = saved


def build(name, cfg, sd):
    if name == "fp16all":
        return AllHalf(cfg, sd, decode_f16=True)
    if name == "bf16x3":
        return port.QuantPortModel(cfg, sd, decode_x3=True)
    if name == "bf16":
        return port.QuantPortModel(cfg, sd, decode_x3=False)
    if name == "fp16":
        return Fp16Decode(cfg, sd, decode_x3=True)
    if name == "fp16x2":
        return Fp16x2Decode(cfg, sd, decode_x3=True)
    raise ValueError(name)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    variants = (sys.argv[3] if len(sys.argv) > 3 else "bf16x3,bf16,fp16,fp16x2").split(",")
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = vcfg.variant("16_384")
    sd = synth.make_state_dict(cfg, seed=0, vocab_gain=1.0, eos_bias=1.0)
    extra = synth.default_test_extra_input(cfg)
    images = synth.make_images(cfg, 192, seed=321)[:n]          # the agreement test's images
    ref = port.PortModel(cfg, sd)
    models = {v: build(v, cfg, sd) for v in variants}
    stats = {v: dict(same=0, agree=0, div=0, worst=0.0) for v in variants}
    t0 = time.time()
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        data = synth.make_text_inputs(cfg, hi - lo)
        data["image"] = images[lo:hi]
        trace, r_info = [], {}
        with torch.no_grad():
            r_ids, _ = port.caption(ref, data, extra, algorithm="cached", trace=trace, info=r_info)
        r = r_ids[:, 0].numpy()
        gaps = torch.stack([tr.float().topk(2).values for tr in trace])
        gaps = (gaps[..., 0] - gaps[..., 1]).numpy()
        for v in variants:
            info = {}
            with torch.no_grad():
                call = lambda: port.caption(models[v], data, extra, algorithm="cached", info=info)      # noqa: E731
                ids, _ = models[v].run(call) if hasattr(models[v], "run") else call()
            a = ids[:, 0].numpy()
            s_ = stats[v]
            s_.setdefault("cap_err", []).append(float((info["cap"] - r_info["cap"]).norm() / r_info["cap"].norm()))
            s = stats[v]
            for row in range(hi - lo):
                neq = np.nonzero(a[row] != r[row])[0]
                if len(neq) == 0:
                    n_tok = int((r[row] != 0).sum()) - 1
                    s["same"] += n_tok
                    s["agree"] += n_tok
                    continue
                t = int(neq[0])
                s["same"] += t
                s["agree"] += t - 1
                s["div"] += 1
                s["worst"] = max(s["worst"], float(gaps[t - 1, row]))
        line = "  ".join("%s %d/%d (%.4f, %d rows, worst gap %.3g, caption features vs fp32 %.2e)"
                         % (v, s["agree"], s["same"], s["agree"] / max(1, s["same"]), s["div"], s["worst"],
                            sum(s["cap_err"]) / len(s["cap_err"])) for v, s in stats.items())
        print("images %d..%d  %.0f s  %s" % (lo, hi, time.time() - t0, line), flush=True)


if __name__ == "__main__":
    main()
