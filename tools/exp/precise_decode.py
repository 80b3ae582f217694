"""Experiment (measurement script, not a test): how much of the bf16 fast mode's token disagreement with the fp32
oracle comes from the DECODE-step GEMM operands? The encoder, the prefill and the decode attention stay bf16; the decode-step
Linear layers and the vocabulary head run on the fp32 CUDA-core kernel with fp32 weights.   python tools/exp/precise_decode.py [B]
Result (256 images, profiles/r01_precise_decode_exp_s9.log): fc1 + fc2 + head + vocab in fp32 -> 99.5 %; that became the
split-bf16 decode path (engine._decode_layers, decode_precision="bf16x3")."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.test_fullsize_gpu import DEV, _data, _oracle_on_gpu  # noqa: E402
from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import ops, synth  # noqa: E402
from vitcap_b200.engine import PackedWeights  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402


def agreement(ids, ref_ids, B):
    a, r = ids[:, 0].cpu().numpy(), ref_ids[:, 0].cpu().numpy()
    tot = agree = div = 0
    for row in range(B):
        neq = np.nonzero(a[row] != r[row])[0]
        if len(neq) == 0:
            n = int((r[row] != 0).sum()) - 1
            tot += n
            agree += n
        else:
            t = int(neq[0])
            tot += t
            agree += t - 1
            div += 1
    return agree, tot, div


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    cfg = vcfg.variant("16_384")
    sd = synth.make_state_dict(cfg, seed=0, eos_bias=1.0)
    extra = synth.default_test_extra_input(cfg)
    data = _data(cfg, B, seed=321)
    ref_ids, _, _ = _oracle_on_gpu(cfg, sd, data, extra)
    m = FastImageCaptioning(cfg, test_extra_input=extra, mode="bf16", max_batch=128, use_cuda_graph=False, decode_precision="bf16")
    m.load_state_dict(sd)
    m = m.to(DEV)
    ids, _ = m(data)
    print("bf16 decode GEMMs : %d/%d agree, %d/%d rows diverge" % (*agreement(ids, ref_ids, B), B), flush=True)

    eng = m.engine
    w32 = PackedWeights(cfg, sd, "fp32", torch.device(DEV))
    H = cfg.hidden
    f32 = torch.float32

    PRECISE = set()
    LAYER = [-1]                       # decoder layer being evaluated ("fc1@3" selects one layer's GEMM)

    def lin(name, a_f, a_t, p32, p16, wk, bk, out, **kw):
        """fp32 CUDA-core GEMM when `name` is in PRECISE, else the bf16 tensor-core GEMM on the bf16 copy of the operand."""
        if name in PRECISE or "%s@%d" % (name, LAYER[0]) in PRECISE:
            ops.linear(a_f, p32[wk], p32[bk], out, **kw)
        else:
            ops.linear(a_t if a_t is not None else a_f.to(torch.bfloat16), p16[wk], p16[bk], out, **kw)

    def decode_layers(ws, Bc, E, cur_len, anc, mask_id, labels=False, head=True):
        w = eng.w
        R = ws["R"]
        e_f, e_t = ws["e_f"], ws["e_t"]
        ops.embed_ln(ws["ids"], cur_len, mask_id, w.word, w.pos, w.type0, w.emb_ln_w, w.emb_ln_b, cfg.bert_ln_eps, e_f, e_t, R)
        scale = 1.0 / math.sqrt(cfg.head_dim)
        qkv_f = torch.empty(2 * R, 3 * H, device=DEV, dtype=f32)
        hid_f = torch.empty(2 * R, cfg.inter, device=DEV, dtype=f32)
        for l, (p, q) in enumerate(zip(w32.dec, w.dec)):
            sq = ws["step_qkv"][l]
            LAYER[0] = l
            lin("qkv", e_f, e_t, p, q, "qkv_w", "qkv_b", qkv_f)
            sq[cur_len - 1].copy_(qkv_f)                                   # the caption-row K/V cache stays bf16
            ops.decode_attention(eng._enc_ws["ctx_qkv"][l], sq, anc, ws["att"], Bc, cfg.n_ctx, cfg.heads, E, cur_len, scale)
            lin("o", ws["att"].float(), ws["att"], p, q, "o_w", "o_b", ws["tmp"], resid=e_f)
            ops.layernorm(ws["tmp"], p["ln1_w"], p["ln1_b"], cfg.bert_ln_eps, out_t=ws["a_t"], out_f=ws["a_f"], rows=2 * R)
            lin("fc1", ws["a_f"], ws["a_t"], p, q, "i_w", "i_b", hid_f, act=ops.ACT_GELU)
            lin("fc2", hid_f, None, p, q, "f_w", "f_b", ws["tmp"], resid=ws["a_f"])
            ops.layernorm(ws["tmp"], p["ln2_w"], p["ln2_b"], cfg.bert_ln_eps, out_t=e_t, out_f=e_f, rows=2 * R)
        LAYER[0] = -1
        if head:
            hp, hq = w32.cls_head, w.cls_head
            rows_f = e_f[1::2].contiguous()
            th = torch.empty(R, H, device=DEV, dtype=f32)
            lin("head_t", rows_f, None, hp, hq, "t_w", "t_b", th, act=ops.ACT_GELU)
            th2 = torch.empty(R, H, device=DEV, dtype=f32)
            ops.layernorm(th, hp["ln_w"], hp["ln_b"], cfg.bert_ln_eps, out_t=ws["head_t"], out_f=th2, rows=R)
            lin("vocab", th2, ws["head_t"], hp, hq, "dec_w", "bias", ws["logits"][:, :cfg.vocab], ldo=ws["logits"].stride(0))

    eng._decode_layers = decode_layers
    sels = (["qkv", "o", "fc1", "fc2", "head_t", "vocab"], [], ["head_t", "vocab"], ["vocab"], ["fc1", "fc2"], ["qkv", "o"],
            ["fc1", "fc2", "head_t", "vocab"], ["qkv", "o", "fc1", "fc2"])
    if len(sys.argv) > 2 and sys.argv[2] == "layers":
        hv = ["head_t", "vocab"]
        sels = (hv + ["fc1", "fc2"], hv + ["fc1@3", "fc2@3"], hv + ["fc1@2", "fc2@2", "fc1@3", "fc2@3"], hv + ["fc2"], hv + ["fc1"],
                ["vocab", "fc1", "fc2"], hv + ["fc1@0", "fc2@0"], hv + ["fc2@2", "fc2@3"])
    for sel in sels:
        PRECISE.clear()
        PRECISE.update(sel)
        ids2, _ = m(data)
        print("fp32 GEMMs %-44s: %d/%d agree, %d/%d rows diverge" % (",".join(sel) or "(none)", *agreement(ids2, ref_ids, B), B),
              flush=True)


if __name__ == "__main__":
    main()
