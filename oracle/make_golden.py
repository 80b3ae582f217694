"""TEST INFRASTRUCTURE ONLY. Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_loader.py shims) on synthetic weights/images.

Run in the build container (the reference tree is not on the GPU box):

    python -m oracle.make_golden            # all cases (~6 min on 8 vCPU)
    python -m oracle.make_golden g1 g3      # selected cases

The reference ships no golden vectors of its own, so these files ARE the pin:
they hold what the reference itself computes for (state_dict, images, decode flags) that
``vitcap_b200.synth`` reproduces bit-for-bit anywhere from the recorded seeds.
"""
import json
import os
import sys
import time

import numpy as np
import torch

from oracle import ref_loader
from vitcap_b200 import config as vcfg
from vitcap_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (variant, cfg overrides, weight kwargs, batch, image seed, decode overrides, torch seed)
    "g1_greedy_16_384": ("16_384", {}, dict(seed=0), 2, 1234, {}, None),
    "g2_beam4_16_384": ("16_384", {}, dict(seed=0), 2, 1234, dict(num_beams=4), None),
    "g3_greedy_eos_16_224": ("16_224", {}, dict(seed=0, eos_bias=1.9), 6, 1234, {}, None),
    "g4_beam3_keep3_16_224": ("16_224", {}, dict(seed=0, eos_bias=1.9), 3, 1234,
                              dict(num_beams=3, num_keep_best=3, length_penalty=0.6), None),
    "g5_sample5_16_224": ("16_224", {}, dict(seed=0, eos_bias=1.8), 2, 1234,
                          dict(do_sample=True, num_return_sequences=5), 77),
    "g6_greedy_32_384": ("32_384", {}, dict(seed=3), 2, 99, {}, None),
    "g7_greedy_dec12_16_224": ("16_224", dict(dec_layers=12), dict(seed=1), 2, 5, {}, None),
    "g8_sample_filtered_16_224": ("16_224", {}, dict(seed=0, eos_bias=1.8), 2, 1234,
                                  dict(do_sample=True, num_return_sequences=2, temperature=0.7, top_k=50, top_p=0.9), 78),
    "g9_greedy_refinit_16_224": ("16_224", {}, dict(seed=4, style="reference"), 2, 11, {}, None),
    # visible od/tag label region: text_b label strings per sample -> 0 / 4 / 50 visible slots
    "g10_greedy_labels_16_224": ("16_224", {}, dict(seed=0, eos_bias=1.7), 4, 1234, {}, None,
                                 ["", "dog cat person", " ".join(["tree"] * 60), " ".join(["sky", "grass"] * 11)]),
    "g11_beam3_labels_16_224": ("16_224", {}, dict(seed=0, eos_bias=1.8), 3, 1234,
                                dict(num_beams=3, num_keep_best=2, length_penalty=0.8), None,
                                ["a man riding a wave on top of a surfboard", " ".join(["sky", "grass", "field"] * 7),
                                 " ".join(["tree"] * 60)]),
    # tag_bias moves topk_len[0] below 50 so that the label-embedding recipe (modeling_bert.py:1435) flips mid-caption
    # (g12, g13) or is the raw one from the first step (g14)
    "g12_greedy_labels_flip_16_224": ("16_224", {}, dict(seed=0, eos_bias=1.7, tag_bias=-3.176), 3, 1234, {}, None,
                                      [" ".join(["sky", "grass"] * 11), " ".join(["tree"] * 60), "dog cat person"]),
    "g13_beam3_labels_flip_16_224": ("16_224", {}, dict(seed=0, eos_bias=1.75, tag_bias=-3.16), 2, 1234,
                                     dict(num_beams=3, num_keep_best=1), None,
                                     [" ".join(["tree"] * 60), " ".join(["sky", "grass", "field"] * 7)]),
    "g14_greedy_labels_raw_16_224": ("16_224", {}, dict(seed=0, eos_bias=1.7, tag_bias=-3.3), 3, 1234, {}, None,
                                     [" ".join(["tree"] * 60), "", " ".join(["sky", "grass"] * 11)]),
}


def run_case(name):
    variant, over, wkw, B, iseed, dec, tseed = CASES[name][:7]
    text_b = CASES[name][7] if len(CASES[name]) > 7 else None
    cfg = vcfg.variant(variant, **over)
    sd = synth.make_state_dict(cfg, **wkw)
    ref, tok = ref_loader.build_reference(variant, decoder_layer=over.get("dec_layers"))
    ref.load_state_dict(sd, strict=True)
    extra = synth.default_test_extra_input(cfg, **dec)
    ref.test_extra_input = extra
    img = synth.make_images(cfg, B, seed=iseed)
    ti = ref_loader.reference_text_inputs(tok, B, text_b)
    n_label = (ti["attention_mask"][:, 0, cfg.max_seq_a:].sum(1)).tolist()
    mine = synth.make_text_inputs(cfg, B, n_label=n_label)
    for k in ti:
        if text_b is not None and k == "input_ids":      # label word pieces: synth uses a filler id (never reaches the output)
            assert torch.equal(ti[k][:, :cfg.max_seq_a], mine[k][:, :cfg.max_seq_a])
            assert torch.equal(ti[k] == 0, mine[k] == 0)
            continue
        assert torch.equal(ti[k], mine[k]) and ti[k].dtype == mine[k].dtype, k

    cap = {"step_top": [], "n_calls": 0}

    def enc_hook(mod, inp, out):
        if "cap" not in cap:
            cap["cap"], cap["tagf"] = out[0].detach().clone(), out[1].detach().clone()
            cap["img_feats"] = inp[0].detach().clone()

    def tag_hook(mod, inp, out):
        if "tag_logit" not in cap:
            cap["tag_logit"] = out.detach().clone()

    def cls_hook(mod, inp, out):
        cap["n_calls"] += 1
        cur = cap["n_calls"]           # logits row index == cur_len == call number
        row = out[:, cur].detach()
        v, i = row.topk(4, dim=-1)
        cap["step_top"].append((v.clone(), i.clone()))

    h1 = ref.module.bert.encoder.register_forward_hook(enc_hook)
    h2 = ref.module.bert.tag_logit.register_forward_hook(tag_hook)
    h3 = ref.module.cls.register_forward_hook(cls_hook)
    data = dict(ti)
    data["image"] = img
    data["key"] = ["k%d" % i for i in range(B)]
    if tseed is not None:
        torch.manual_seed(tseed)
    t0 = time.time()
    with torch.no_grad():
        ids, lp = ref(data)
    dt = time.time() - t0
    for h in (h1, h2, h3):
        h.remove()
    E = extra["num_beams"] * extra["num_return_sequences"]
    tl = cap["tag_logit"][::E]
    prob, idx = torch.sigmoid(tl).topk(cfg.topk, dim=1)
    out = {
        "ids": ids.numpy(), "logprobs": lp.numpy(),
        "tag_logit_head": tl[:, :512].numpy(), "tag_logit_sum": tl.double().sum(1).numpy(),
        "tag_topk_idx": idx.numpy(), "tag_topk_prob": prob.numpy(),
        "tag_topk_len": (prob >= cfg.tag_thresh).sum(1).numpy(),
        "img_feats_s": cap["img_feats"][::E, ::29, ::37].numpy(),
        "cap_feats_s": cap["cap"][::E, ::29, ::37].numpy(),
        "tag_feats_s": cap["tagf"][::E, ::29, ::37].numpy(),
        "step_top_val": torch.stack([v for v, _ in cap["step_top"]]).numpy(),
        "step_top_idx": torch.stack([i for _, i in cap["step_top"]]).numpy(),
        "meta": np.array(json.dumps({
            "case": name, "variant": variant, "cfg_overrides": over, "weights": wkw, "batch": B,
            "image_seed": iseed, "decode": dec, "torch_seed": tseed, "torch": torch.__version__,
            "text_b": text_b, "n_label": n_label,
            "n_model_calls": cap["n_calls"], "reference_seconds": round(dt, 1),
            "reference_threads": torch.get_num_threads(),
        })),
    }
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("%s: %.1fs, %d model calls, ids[0]=%s lp=%s" % (name, dt, cap["n_calls"], ids[0, 0].tolist(), lp.flatten().tolist()[:4]),
          flush=True)


if __name__ == "__main__":
    sel = sys.argv[1:]
    for name in CASES:
        if sel and not any(name.startswith(s) for s in sel):
            continue
        run_case(name)
