"""BASELINE.json configurations at their FULL sizes on the GPU, checked through size-independent properties (the CPU oracle
cannot run 512 images): images are independent, so every image's result in the big batch must equal -- bit for bit, the
kernels are deterministic and their per-row arithmetic does not depend on the batch size -- its result in a small batch
(which test_e2e_gpu.py pins to the reference goldens / the oracle), plus the structural invariants of each search mode."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import synth  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402

DEV = "cuda:0"
_CACHE = {}


def _model(variant, eos_bias, extra_kw, max_batch, mode="bf16", seed=0):
    key = (variant, eos_bias, seed)
    if key not in _CACHE:
        _CACHE.clear()                                   # one 217 M-parameter state_dict alive at a time
        cfg = vcfg.variant(variant)
        _CACHE[key] = (cfg, synth.make_state_dict(cfg, seed=seed, eos_bias=eos_bias))
    cfg, sd = _CACHE[key]
    extra = synth.default_test_extra_input(cfg, **extra_kw)
    m = FastImageCaptioning(cfg, test_extra_input=extra, mode=mode, max_batch=max_batch)
    m.load_state_dict(sd)
    return cfg, m.to(DEV)


def _data(cfg, B, seed, lo=0, hi=None):
    hi = B if hi is None else hi
    d = synth.make_text_inputs(cfg, hi - lo)
    img = synth.make_images(cfg, B, seed=seed)[lo:hi].contiguous()
    d["image"] = img
    return {k: v.to(DEV) for k, v in d.items()}


def _check_caption_structure(ids, lp, bos=101, eos=102, pad=0, pad_inside_ok=False):
    ids = ids.cpu().numpy()
    lp = lp.cpu().numpy()
    assert (ids[..., 0] == bos).all()
    assert np.isfinite(lp).all() and (lp <= 1e-6).all()
    flat = ids.reshape(-1, ids.shape[-1])
    for row in flat:
        w = np.nonzero(row == eos)[0]
        assert len(w) >= 1, row                          # every caption ends with [SEP] (forced at the last slot at the latest)
        assert (row[w[0] + 1:] == pad).all(), row        # and is PAD-filled after it
        if not pad_inside_ok:                            # (a sampler may legitimately draw vocabulary id 0 = [PAD])
            assert (row[1:w[0]] != pad).all(), row


def test_config2_encoder_tags_b256():
    """BASELINE configs[1]: ViT-B/16-384 encoder + concept head top-50, batch 256, bf16."""
    cfg, m = _model("16_384", 0.0, {}, 256)
    B = 256
    data = _data(cfg, B, seed=1234)
    lg, idx, pr, n = m.forward_tags(data["image"])
    assert lg.shape == (B, cfg.vocab) and idx.shape == (B, 50) and pr.shape == (B, 50) and n.shape == (B,)
    pr_c, idx_c = pr.cpu(), idx.cpu()
    assert bool((pr_c[:, :-1] >= pr_c[:, 1:]).all())                       # sorted like torch.topk(sorted=True)
    assert all(len(set(r.tolist())) == 50 for r in idx_c)                  # distinct vocabulary ids
    assert torch.equal(n.cpu(), (pr_c >= 0.2).sum(1))                      # topk_len, modeling_bert.py:1432
    # the selection is the true top-50 of the logits this very run produced (selected on the logits: sigmoid is monotonic
    # but rounds distinct logits to equal fp32 probabilities, where torch.topk's tie order is unspecified)
    ref_l, ref_i = lg.topk(50, dim=1)
    assert torch.equal(idx_c, ref_i.cpu())
    np.testing.assert_allclose(pr_c.numpy(), torch.sigmoid(ref_l).cpu().numpy(), atol=1e-6)
    # batch independence: images 40..47 alone give bitwise the same logits
    lg8, idx8, _, _ = m.forward_tags(data["image"][40:48].contiguous())
    assert torch.equal(lg8, lg[40:48]) and torch.equal(idx8, idx[40:48])
    cap, tag = m.encode_features(data["image"][:64].contiguous())
    assert torch.isfinite(cap).all() and torch.isfinite(tag).all()


def test_config3_greedy_b512_batch_independent():
    """BASELINE configs[2]: full greedy captioning, batch 512, bf16; EOS planted so early stop / PAD fill / forced EOS all occur."""
    cfg, m = _model("16_384", 1.9, {}, 512)
    B = 512
    data = _data(cfg, B, seed=99)
    ids, lp = m(data)
    assert ids.shape == (B, 1, 20) and ids.dtype == torch.int64 and lp.shape == (B, 1)
    _check_caption_structure(ids, lp)
    ids2, lp2 = m(data)                                                     # CUDA-graph replay: identical
    assert torch.equal(ids, ids2) and torch.equal(lp, lp2)
    for lo in (0, 300, 504):
        sub = _data(cfg, B, seed=99, lo=lo, hi=lo + 8)
        i8, l8 = m(sub)
        assert torch.equal(i8, ids[lo:lo + 8]), lo
        np.testing.assert_allclose(l8.cpu().numpy(), lp[lo:lo + 8].cpu().numpy(), atol=1e-6)
    lens = (ids[:, 0] != 0).sum(1)
    assert int(lens.min()) < 20                                             # some captions stop early


def test_config4_beam4_b256():
    """BASELINE configs[3]: beam search, 4 beams, batch 256 (context K/V shared by the beams, ancestor-table reorder)."""
    cfg, m = _model("16_384", 1.9, dict(num_beams=4, num_keep_best=2, length_penalty=0.8), 256)
    B = 256
    data = _data(cfg, B, seed=7)
    ids, lp = m(data)
    assert ids.shape == (B, 2, 20) and lp.shape == (B, 2)
    lpc = lp.cpu()
    filled = lpc > -1e4
    assert bool(filled[:, 0].all())
    assert bool((lpc[:, 0] >= lpc[:, 1]).all())                             # best hypothesis first (modeling_utils.py:1086)
    _check_caption_structure(ids[:, :1], lp[:, :1])
    sub = _data(cfg, B, seed=7, lo=100, hi=104)
    i4, l4 = m(sub)
    assert torch.equal(i4, ids[100:104])
    np.testing.assert_allclose(l4.cpu().numpy(), lp[100:104].cpu().numpy(), atol=1e-6)
    # the two kept hypotheses of an image are different sequences
    idc = ids.cpu()
    assert bool((idc[:, 0] != idc[:, 1]).any(dim=1)[filled[:, 1]].all())


def test_config5_sampling_k5_b512_16_224():
    """BASELINE configs[4]: SCST-style sampling, 5 samples per image, batch 512, the 16_224 variant."""
    cfg, m = _model("16_224", 1.5, dict(do_sample=True, num_return_sequences=5), 512)
    B, K = 512, 5
    data = _data(cfg, B, seed=5)
    m.sample_seed = 1234
    m._sample_calls = 0
    ids, lp = m(data)
    assert ids.shape == (B * K, 1, 20) and lp.shape == (B * K, 1)
    _check_caption_structure(ids, lp, pad_inside_ok=True)
    per_img = ids.view(B, K, 20)
    distinct = [len({tuple(s.tolist()) for s in per_img[b]}) for b in range(0, B, 37)]
    assert np.mean(distinct) > 1.5                                          # the samples of an image differ
    m._sample_calls = 0
    ids2, lp2 = m(data)                                                     # same seed and call index: same samples
    assert torch.equal(ids, ids2)
    ids3, _ = m(data)                                                       # next call index: fresh noise
    assert not torch.equal(ids, ids3)


def _oracle_on_gpu(cfg, sd, data, extra):
    """The CPU oracle's cached algorithm (fp32, TF32 off) executed by torch on the GPU, so that it can run the full-size model
    on a few dozen images: returns (ids (B,1,L), per-step logits list)."""
    from oracle import port
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    pm = port.PortModel(cfg, {k: v.to(DEV) for k, v in sd.items()})
    torch.set_default_device(DEV)                      # the port builds its masks / position ids with bare factory calls
    try:
        trace = []
        with torch.no_grad():
            ids, lp = port.caption(pm, data, extra, algorithm="cached", trace=trace)
    finally:
        torch.set_default_device("cpu")
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return ids, lp, trace


@pytest.mark.parametrize("precision,vocab_gain,min_agree,max_gap", [("bf16x3", 1.0, 0.99, 2.5e-2), ("bf16", 1.0, 0.975, 2.5e-2),
                                                                    ("fp16", 1.0, 0.99, 2.5e-2),
                                                                    ("bf16x3", 4.0, 0.98, 1e-1), ("bf16", 4.0, 0.96, 1e-1)])
def test_bf16_mode_token_agreement_fullsize_vs_oracle(precision, vocab_gain, min_agree, max_gap):
    """North-star criterion for the fast mode: >= 99 % greedy-token agreement with the fp32 reference algorithm. Full-size
    ViT-B/16-384 model, 192 images, oracle = oracle/port.py (cached, fp32) run by torch on the same GPU. Agreement is counted
    over the tokens produced under an identical prefix (every token of a row up to and including its first divergence: the
    teacher-forced condition); every divergence must sit on a reference near-tie (top-1/top-2 logit gap < max_gap).
    vocab_gain 1 = the bench's own weights (synth seed 0; median top-2 gap 0.08).
    precision 'bf16x3' (the default: decode-step MLP and vocabulary head on split-bf16 operands): measured
    99.60 % over 192 images (3481/3495; 14 rows diverge), 99.45 % at vocab_gain 4 -- the >= 99 % of the north star is ASSERTED.
    precision 'bf16' (plain bf16 operands everywhere): 98.4-98.6 % over 256 images (59-65 rows diverge, all at reference gaps
    <= 1e-2): bf16 operands leave the logits with ~0.9 % relative error (0.005 absolute at their 0.55 standard deviation) and 1.5 %
    of this model's argmax decisions have a top-2 gap below that; the folded LayerNorms do not move the figure. vocab_gain 4 scales
    logits AND their error by 4 (97.3-98.2 %): the flip rate is set by relative precision, not by peakiness. max_gap = 4.5 % of the
    logits' standard deviation (0.55 x gain); the asserted floors leave room for sampling noise; measured values are printed."""
    cfg = vcfg.variant("16_384")
    sd = synth.make_state_dict(cfg, seed=0, vocab_gain=vocab_gain, eos_bias=1.0)
    extra = synth.default_test_extra_input(cfg)
    B = int(os.environ.get("VITCAP_AGREE_B", "192"))     # (environment override for A/B measurements of numerics changes)
    data = _data(cfg, B, seed=321)
    m = FastImageCaptioning(cfg, test_extra_input=extra, mode="bf16", max_batch=min(B, 128), decode_precision=precision)
    m.load_state_dict(sd)
    m = m.to(DEV)
    ids, lp = m(data)
    ref_ids, ref_lp, trace = _oracle_on_gpu(cfg, sd, data, extra)
    a = ids[:, 0].cpu().numpy()
    r = ref_ids[:, 0].cpu().numpy()
    same_prefix_tokens, agree, diverged, worst_gap = 0, 0, 0, 0.0
    near_ties = 0                                           # same-prefix decisions whose reference top-2 gap is below max_gap
    gaps = torch.stack([tr.float().topk(2).values for tr in trace])          # (steps, B, 2)
    gaps = (gaps[..., 0] - gaps[..., 1]).cpu().numpy()
    for row in range(B):
        neq = np.nonzero(a[row] != r[row])[0]
        n_same = (int((r[row] != 0).sum()) - 1) if len(neq) == 0 else int(neq[0])
        near_ties += int((gaps[:n_same, row] < max_gap).sum())
        if len(neq) == 0:
            n_tok = int((r[row] != 0).sum()) - 1           # generated tokens (BOS excluded, PAD fill excluded)
            same_prefix_tokens += n_tok
            agree += n_tok
            continue
        t = int(neq[0])                                     # position t was produced at decode step t-1
        top2 = trace[t - 1][row].float().topk(2).values
        gap = float(top2[0] - top2[1])
        worst_gap = max(worst_gap, gap)
        assert gap < max_gap, "row %d diverges at position %d where the reference gap is %.3g" % (row, t, gap)
        same_prefix_tokens += t                             # positions 1..t were produced under the reference's prefix
        agree += t - 1
        diverged += 1
    frac = agree / same_prefix_tokens
    print("%s decode vs fp32 oracle, vocab_gain %.0f: %d/%d same-prefix tokens agree (%.4f), %d/%d rows diverge, largest excused "
          "gap %.3g; gap-aware: %d of the decisions are reference near-ties (gap < %.3g), the other %d agree 100 %%"
          % (precision, vocab_gain, agree, same_prefix_tokens, frac, diverged, B, worst_gap, near_ties, max_gap, same_prefix_tokens - near_ties))
    assert frac >= min_agree
    if diverged == 0:
        np.testing.assert_allclose(lp.cpu().numpy(), ref_lp.cpu().numpy(), atol=3e-2)


def _on_gpu(fn):
    """Runs fn() with torch's default device on the GPU and TF32 off (the oracle builds masks / position ids with bare factory
    calls and must accumulate in true fp32)."""
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_default_device(DEV)
    try:
        with torch.no_grad():
            return fn()
    finally:
        torch.set_default_device("cpu")
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def _tapped_forward(cfg, sd, extra, data, B, decode_precision=None):
    m = FastImageCaptioning(cfg, test_extra_input=extra, mode="bf16", max_batch=B, decode_precision=decode_precision)
    m.load_state_dict(sd)
    m = m.to(DEV)
    taps = {}
    m.engine.tap = lambda name, i, t: taps.__setitem__((name, i), t.detach().clone())
    try:
        ids, lp = m(data)                                     # production path: CLS-only last concept block, eager decode loop
    finally:
        m.engine.tap = None
    return m, taps, ids, lp


def test_bf16_mode_every_kernel_vs_quantisation_matched_oracle():
    """North-star criterion 2 for the fast mode ("encoder features and logits within 1e-3 relative error in the bf16 mode"),
    stated where it can hold: PER KERNEL. oracle/port.py QuantPortModel is the fp32 algorithm with its operands rounded to bf16
    exactly where the kernels round; every stage of it is fed the CUDA path's OWN input of that stage (teacher forcing), so what
    is compared is one kernel's arithmetic -- summation order, MUFU ex2 / tanh, one-pass statistics -- not the operand
    quantisation and not the drift of everything upstream. Full-size ViT-B/16-384 model, the benchmarked configuration
    (folded LayerNorms, CLS-only last concept block, split-bf16 decode GEMMs); asserted <= 1e-3 norm-wise for every kernel
    of the encoder, the concept head, the prefill layers and the first decode step (measured 2e-6 .. 1.4e-4: the floor is the
    fraction of bf16 roundings that flip because two fp32 summations differ by ~3e-6).
    End to end the same comparison CANNOT stay below 1e-3 for any pair of implementations: a bf16 rounding turns a relative
    input perturbation e into ~0.04 sqrt(e) of output error, so even 1e-7 reaches the quantisation-noise level after a few
    stages -- see test_bf16_mode_end_to_end_vs_quantisation_matched_oracle, which measures exactly that drift for the oracle
    against itself. Reference: vision_transformer.py:233-250, modeling_bert.py:303-437, 540-563, 1415-1432."""
    from oracle import port
    cfg = vcfg.variant("16_384")
    sd = synth.make_state_dict(cfg, seed=0, eos_bias=1.0)
    extra = synth.default_test_extra_input(cfg)
    B = 8
    data = _data(cfg, B, seed=2024)
    m, taps, ids, lp = _tapped_forward(cfg, sd, extra, data, B, decode_precision="bf16x3")
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    N, C, H, heads, d, F_ = cfg.n_tokens, cfg.n_ctx, cfg.hidden, cfg.heads, cfg.head_dim, cfg.inter
    split_at = cfg.enc_blocks - cfg.split_blocks
    errs = []

    def check(name, got, want):
        e = _rel(got.float().reshape(want.shape), want)
        errs.append((name, e))

    def run():
        qm = port.QuantPortModel(cfg, sd_dev)
        check("patch embed", taps[("patch", 0)], qm.patch_embed(data["image"]))

        def block(kind, i, prefix, x_in, fold1):
            t = lambda n: taps[(kind + "." + n, i)]                                   # noqa: E731
            eps = cfg.vit_ln_eps
            n1 = (prefix + "norm1.weight", prefix + "norm1.bias", eps, prefix + "attn.qkv.weight", prefix + "attn.qkv.bias")
            o_qkv = qm.lin_fold(x_in, *n1) if fold1 else qm.lin_ln(x_in, *n1)[0]
            k_qkv = t("qkv").float().view(B, N, 3 * H)
            check("%s %d qkv (%s)" % (kind, i, "folded norm1" if fold1 else "LayerNorm kernel"), k_qkv, port.q_bf16(o_qkv))
            kq = k_qkv.reshape(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
            k_att = t("att").float().view(B, N, H)
            check("%s %d attention" % (kind, i), k_att, qm.attend(kq[0], kq[1], kq[2], d ** -0.5))
            k_mid = t("mid").view(B, N, H)
            check("%s %d proj + residual" % (kind, i), k_mid, x_in + qm.lin(k_att, prefix + "attn.proj.weight", prefix + "attn.proj.bias"))
            n2 = (prefix + "norm2.weight", prefix + "norm2.bias", eps, prefix + "mlp.fc1.weight", prefix + "mlp.fc1.bias")
            k_hid = t("hid").float().view(B, N, F_)
            check("%s %d fc1 + GELU (folded norm2)" % (kind, i), k_hid, port.q_bf16(port.gelu_fast(qm.lin_fold(k_mid, *n2))))
            k_out = taps[(kind, i)]
            check("%s %d fc2 + residual" % (kind, i), k_out, k_mid + qm.lin(k_hid, prefix + "mlp.fc2.weight", prefix + "mlp.fc2.bias"))
            return k_out

        x = taps[("patch", 0)]
        trunk_out = None
        for i in range(cfg.enc_blocks):
            if i == split_at:
                trunk_out = x
            x = block("block", i, "module.bert.encoder.blocks.%d." % i, x, fold1=i > 0)
        xt = trunk_out
        for j in range(cfg.split_blocks - 1):
            xt = block("tag_block", j, "module.bert.encoder.tag_blocks.%d." % j, xt, fold1=True)
        # the CLS-only last concept block (engine._vit_block_cls_only), kernel by kernel
        prefix = "module.bert.encoder.tag_blocks.%d." % (cfg.split_blocks - 1)
        eps = cfg.vit_ln_eps
        o_qkv, o_h = qm.lin_ln(xt, prefix + "norm1.weight", prefix + "norm1.bias", eps, prefix + "attn.qkv.weight", prefix + "attn.qkv.bias")
        o_qkv = port.q_bf16(o_qkv)
        k_kv = taps[("cls.kv", 0)].float().view(B, N, 2 * H)
        check("cls-only block k|v", k_kv, o_qkv[..., H:])
        k_q = taps[("cls.q", 0)].float()
        check("cls-only block q", k_q, o_qkv[:, 0, :H])
        kk = k_kv.view(B, N, 2, heads, d).permute(2, 0, 3, 1, 4)
        k_att = taps[("cls.att", 0)].float()
        check("cls-only block attention (fp32 probabilities)", k_att,
              qm.attend(k_q.view(B, 1, heads, d).permute(0, 2, 1, 3), kk[0], kk[1], d ** -0.5, round_p=False)[:, 0])
        k_mid = taps[("cls.mid", 0)]
        check("cls-only block proj + residual", k_mid, xt[:, 0] + qm.lin(k_att, prefix + "attn.proj.weight", prefix + "attn.proj.bias"))
        k_hid = taps[("cls.hid", 0)].float()
        check("cls-only block fc1 + GELU", k_hid, port.q_bf16(port.gelu_fast(
            qm.lin_ln(k_mid, prefix + "norm2.weight", prefix + "norm2.bias", eps, prefix + "mlp.fc1.weight", prefix + "mlp.fc1.bias")[0])))
        k_cls = taps[("tag_block", cfg.split_blocks - 1)][:, 0]
        check("cls-only block fc2 + residual", k_cls, k_mid + qm.lin(k_hid, prefix + "mlp.fc2.weight", prefix + "mlp.fc2.bias"))
        # concept head
        k_pool = taps[("tag.pooled", 0)].float()
        check("pooler (tanh)", k_pool, port.q_bf16(torch.tanh(qm.lin(port.q_bf16(k_cls), "module.bert.pooler.dense.weight",
                                                                      "module.bert.pooler.dense.bias"))))
        check("concept logits", taps[("tag.logits", 0)], qm.head("module.bert.tag_logit.predictions.", k_pool))
        # decoder prefill over the context rows, kernel by kernel (engine._prefill_folded: no stand-alone LayerNorm; raw1 / raw2
        # are the pre-LayerNorm rows, the normalised streams exist only inside the GEMMs that fold / re-apply them)
        ctx = taps[("prefill.in", 0)].view(B, C, H)
        check("context assembly", ctx, torch.cat([k_cls.unsqueeze(1), x], dim=1))
        L = cfg.dec_layers
        Kc, Vc = [], []
        lnorm = torch.nn.functional.layer_norm
        hs = lambda t_: t_.reshape(B, C, heads, d).permute(0, 2, 1, 3)                  # noqa: E731
        resid, raw2, ln2_prev = ctx, None, None              # layer 0: the residual is the context itself
        for l in range(L):
            p = "module.bert.decoder.layer.%d." % l
            if l == 0:
                q_, k_, v_ = qm.qkv_rows(l, ctx)
            else:
                q_, k_, v_ = (hs(port.q_bf16(qm.lin_fold(raw2, ln2_prev[0], ln2_prev[1], cfg.bert_ln_eps,
                                                         p + "attention.self.%s.weight" % n, p + "attention.self.%s.bias" % n)))
                              for n in ("query", "key", "value"))
            k_qkv = taps[("prefill.qkv", l)].float().view(B, C, 3 * H)
            kq, kk_, kv = hs(k_qkv[..., :H]), hs(k_qkv[..., H:2 * H]), hs(k_qkv[..., 2 * H:])
            Kc.append(kk_)
            Vc.append(kv)
            how = "plain" if l == 0 else "folded output LayerNorm of layer %d" % (l - 1)
            check("prefill %d k|v (%s)" % (l, how), torch.cat([kk_, kv], 1), torch.cat([k_, v_], 1))
            if l == L - 1:
                break
            check("prefill %d q (%s)" % (l, how), kq, q_)
            k_att = taps[("prefill.att", l)].float().view(B, C, H)
            check("prefill %d attention" % l, k_att, qm.attend(kq, kk_, kv, 1.0 / (d ** 0.5)))
            ln1 = (p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias")
            ln2 = (p + "output.LayerNorm.weight", p + "output.LayerNorm.bias")
            k_raw1 = taps[("prefill.raw1", l)].view(B, C, H)
            check("prefill %d o-proj + residual%s" % (l, "" if l == 0 else " (LayerNorm re-applied to the raw tile)"), k_raw1,
                  qm.lin(k_att, p + "attention.output.dense.weight", p + "attention.output.dense.bias") + resid)
            k_hid = taps[("prefill.hid", l)].float().view(B, C, F_)
            check("prefill %d intermediate + GELU (folded attention-output LayerNorm)" % l, k_hid, port.q_bf16(port.gelu_fast(
                qm.lin_fold(k_raw1, ln1[0], ln1[1], cfg.bert_ln_eps, p + "intermediate.dense.weight", p + "intermediate.dense.bias"))))
            a_norm = lnorm(k_raw1, (H,), qm.p(ln1[0]), qm.p(ln1[1]), cfg.bert_ln_eps)
            raw2 = taps[("prefill.raw2", l)].view(B, C, H)
            check("prefill %d output + residual (LayerNorm re-applied to the raw tile)" % l, raw2,
                  qm.lin(k_hid, p + "output.dense.weight", p + "output.dense.bias") + a_norm)
            resid, ln2_prev = lnorm(raw2, (H,), qm.p(ln2[0]), qm.p(ln2[1]), cfg.bert_ln_eps), ln2
        # first decode step from the kernel's OWN context K/V cache: rows [BOS, MASK] through the four layers and the head
        inp = torch.tensor([[int(extra["bos_token_id"]), int(extra["mask_token_id"])]]).expand(B, 2)
        e = qm.embeddings(inp, torch.tensor([[0, 1]]).expand(B, 2))
        for l in range(L):
            q_, k_, v_ = qm.qkv_rows(l, e)
            add = torch.zeros(1, 1, 2, C + 2)
            add[..., 0, -1] = port.NEG_MASK
            e = qm.bert_layer_from_kv(l, e, q_, torch.cat([Kc[l], k_], 2), torch.cat([Vc[l], v_], 2), add, step=True)
        check("decode step 1: embeddings .. vocabulary logits (4 layers + head, from the kernel's K/V cache)",
              taps[("logits", 1)], qm.head("module.cls.predictions.", e[:, 1]))

    _on_gpu(run)
    for name, e in errs:
        print("  %-86s %.3g" % (name, e))
    worst = max(errs, key=lambda t: t[1])
    print("bf16 mode, kernel by kernel vs quantisation-matched oracle: %d stages, worst %.3g (%s)" % (len(errs), worst[1], worst[0]))
    for name, e in errs:
        assert e <= 1e-3, (name, e)


def test_fp16_decode_step_vs_quantisation_matched_oracle():
    """decode_precision='fp16' at full size, the kernel-level statement: the first decode step (embeddings, four layers on the
    kernel's OWN context K/V cache, head, vocabulary logits through the one-plane + bias form of vc_dec_linear) against
    QuantPortModel(decode_f16=True) -- the same arithmetic with the MLP / head operands rounded to IEEE half where the kernels
    round (finish_ln operand copies, GELU epilogue, packed weights): <= 1e-3 norm-wise, and the greedy token of every row equals
    the arg max of those logits (the arg-max epilogue and the logits plane are the same GEMM)."""
    from oracle import port
    cfg = vcfg.variant("16_384")
    sd = synth.make_state_dict(cfg, seed=0, eos_bias=1.0)
    extra = synth.default_test_extra_input(cfg)
    B = 8
    data = _data(cfg, B, seed=2024)
    m, taps, ids, lp = _tapped_forward(cfg, sd, extra, data, B, decode_precision="fp16")
    assert m.engine.decode_f16
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    C, H, heads, d, L = cfg.n_ctx, cfg.hidden, cfg.heads, cfg.head_dim, cfg.dec_layers
    res = {}

    def run():
        qm = port.QuantPortModel(cfg, sd_dev, decode_f16=True)
        hs = lambda t_: t_.reshape(B, C, heads, d).permute(0, 2, 1, 3)                  # noqa: E731
        inp = torch.tensor([[int(extra["bos_token_id"]), int(extra["mask_token_id"])]]).expand(B, 2)
        e = qm.embeddings(inp, torch.tensor([[0, 1]]).expand(B, 2))
        for l in range(L):
            k_qkv = taps[("prefill.qkv", l)].float().view(B, C, 3 * H)
            q_, k_, v_ = qm.qkv_rows(l, e)
            add = torch.zeros(1, 1, 2, C + 2)
            add[..., 0, -1] = port.NEG_MASK
            e = qm.bert_layer_from_kv(l, e, q_, torch.cat([hs(k_qkv[..., H:2 * H]), k_], 2),
                                      torch.cat([hs(k_qkv[..., 2 * H:]), v_], 2), add, step=True)
        res["want"] = qm.head("module.cls.predictions.", e[:, 1])

    _on_gpu(run)
    got = taps[("logits", 1)].float()
    err = _rel(got.reshape(res["want"].shape), res["want"])
    print("fp16 decode step 1 vs its quantisation-matched oracle: %.3g" % err)
    assert err <= 1e-3
    assert torch.equal(got.argmax(-1).reshape(-1).cpu(), ids[:, 0, 1].cpu())


@pytest.mark.parametrize("precision", ["fp16", "bf16x3"])
def test_bf16_mode_end_to_end_vs_quantisation_matched_oracle(precision):
    """(precision: operands of the decode-step MLP / vocabulary head; the quantised oracle rounds them the same way.)
    The same comparison END TO END (32 images, full-size model, benchmarked configuration): caption / concept features, concept
    logits and the vocabulary logits of all 19 decode steps against (a) QuantPortModel and (b) the fp32 oracle.
      * vocabulary logits (rows whose prefix equals the oracle's): <= 1e-3 against the quantised oracle, the north star's
        figure (measured 6e-4: the decode steps keep their MLP / head operands as split bf16 pairs);
      * encoder features / concept logits: two bf16-storage implementations of one spec cannot stay within 1e-3 over 16 blocks
        (every bf16 rounding amplifies a perturbation e to ~0.04 sqrt(e)). The yardstick is the oracle against ITSELF with fp64
        instead of fp32 accumulation (acc64): the CUDA path must be no further from the oracle than 1.5 x that self-distance;
      * against the fp32 oracle (operand quantisation included): round-1 measured values + 30 % (5.0e-3 / 5.0e-3 / 8.6e-3)."""
    from oracle import port
    cfg = vcfg.variant("16_384")
    sd = synth.make_state_dict(cfg, seed=0, eos_bias=1.0)
    extra = synth.default_test_extra_input(cfg)
    B = int(os.environ.get("VITCAP_QPARITY_B", "32"))
    data = _data(cfg, B, seed=2024)
    m, taps, ids, lp = _tapped_forward(cfg, sd, extra, data, B, decode_precision=precision)
    f16 = precision == "fp16"
    taps = {k: v for k, v in taps.items() if k[0] in ("block", "tag_block", "logits")}
    tag_logits, tag_idx, tag_prob, tag_n = m.forward_tags(data["image"])
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    n_enc, n_tag = cfg.enc_blocks, cfg.split_blocks

    def run_oracle(model):
        trace, info = [], {}
        out = port.caption(model, data, extra, algorithm="cached", trace=trace, info=info)
        return out, trace, info

    (q_ids, q_lp), q_trace, q_info = _on_gpu(lambda: run_oracle(port.QuantPortModel(cfg, sd_dev, decode_f16=f16)))
    (d_ids, d_lp), d_trace, d_info = _on_gpu(lambda: run_oracle(port.QuantPortModel(cfg, sd_dev, acc64=True, decode_f16=f16)))
    (f_ids, f_lp), f_trace, f_info = _on_gpu(lambda: run_oracle(port.PortModel(cfg, sd_dev)))
    k_cap, k_cls = taps[("block", n_enc - 1)], taps[("tag_block", n_tag - 1)][:, 0]
    rows = {"caption features": (k_cap, "cap", None), "concept CLS feature": (k_cls, "tag_feats", 0), "concept logits": (tag_logits, "tag", None)}
    e = {}
    for name, (got, key, row) in rows.items():
        def pick(info):
            v = info[key][0] if key == "tag" else info[key]
            return v[:, row] if row is not None else v
        e[name] = (_rel(got, pick(q_info)), _rel(pick(d_info), pick(q_info)), _rel(got, pick(f_info)))
    a, qa, fa = ids[:, 0], q_ids[:, 0].to(ids.device), f_ids[:, 0].to(ids.device)
    worst_q, worst_f, rows_q = 0.0, 0.0, []
    for step in range(1, cfg.max_seq_a):
        got = taps[("logits", step)]
        same_q = (a[:, :step] == qa[:, :step]).all(dim=1)
        same_f = (a[:, :step] == fa[:, :step]).all(dim=1)
        rows_q.append(int(same_q.sum()))
        if step - 1 < len(q_trace) and bool(same_q.any()):
            worst_q = max(worst_q, _rel(got[same_q], q_trace[step - 1][same_q]))
        if step - 1 < len(f_trace) and bool(same_f.any()):
            worst_f = max(worst_f, _rel(got[same_f], f_trace[step - 1][same_f]))
    for name, (vq, vself, vf) in e.items():
        print("bf16 mode end to end, %-20s vs quantised oracle %.3g (oracle vs its fp64-accumulating self %.3g)   vs fp32 oracle %.3g"
              % (name, vq, vself, vf))
    print("decode_precision %s:" % precision)
    print("bf16 mode end to end, vocabulary logits (worst of 19 steps) vs quantised oracle %.3g   vs fp32 oracle %.3g" % (worst_q, worst_f))
    print("greedy tokens equal to the quantised oracle's: %.4f, to the fp32 oracle's: %.4f (rows with identical prefix per step: "
          "min %d of %d)" % (float((a == qa).float().mean()), float((a == fa).float().mean()), min(rows_q), B))
    assert worst_q <= 1e-3
    assert min(rows_q) >= B // 2                              # the per-step comparison really covered the batch
    for name, (vq, vself, vf) in e.items():
        assert vq <= 1.5 * vself + 2e-4, (name, vq, vself)
    assert e["caption features"][2] <= 6.5e-3 and e["concept CLS feature"][2] <= 8.5e-3
    assert e["concept logits"][2] <= 1.15e-2 and worst_f <= 3e-3
    # concept top-50 of the production path against the quantised oracle: same set up to near-ties of the oracle's logits
    q_logit = q_info["tag"][0]
    kth = q_logit.topk(cfg.topk, dim=1).values[:, -1]
    tol = 2.0 * e["concept logits"][0] * float(q_logit.pow(2).mean().sqrt()) * 4      # 4 sigma of the measured logit distance
    for b in range(B):
        for v in set(q_info["tag"][2][b].tolist()) ^ set(tag_idx[b].tolist()):
            assert abs(float(q_logit[b, v] - kth[b])) < tol, (b, v, tol)


def test_config2_tags_b256_vs_oracle_slice():
    """BASELINE configs[1] at its full size against the ORACLE (not against itself): top-50 concept indices of the B = 256 fast
    path on ViT-B/16-384, checked on a 32-image slice against oracle/port.py run by torch on the GPU -- the fp32 oracle
    and the quantisation-matched oracle, gap-aware (an index may differ only where the oracle's logit is within 8 x the measured
    rms logit distance of its 50th; at most 2 % of the entries). Images are independent, so the slice inherits the B = 256
    arithmetic bit for bit (test_config2_encoder_tags_b256)."""
    from oracle import port
    cfg, m = _model("16_384", 0.0, {}, 256)
    B, lo, n = 256, 96, 32
    data = _data(cfg, B, seed=1234)
    lg, idx, pr, cnt = m.forward_tags(data["image"])
    sd_dev = {k: v.to(DEV) for k, v in _CACHE[("16_384", 0.0, 0)][1].items()}
    img = data["image"][lo:lo + n].contiguous()
    for model_cls, tol_rel in ((port.QuantPortModel, 1.0e-2), (port.PortModel, 1.15e-2)):
        r_cap, r_tag, r_logit, r_prob, r_idx, r_n = _on_gpu(lambda: port.encode_tags(model_cls(cfg, sd_dev), img))
        err = _rel(lg[lo:lo + n], r_logit)
        # an index may differ only at a near-tie: within 8 x the measured rms logit distance of the oracle's 50th logit
        tol_abs = 8.0 * err * float(r_logit.pow(2).mean().sqrt())
        kth = r_logit.topk(cfg.topk, dim=1).values[:, -1]
        swapped = 0
        for b in range(n):
            miss = set(r_idx[b].tolist()) ^ set(idx[lo + b].tolist())
            swapped += len(miss) // 2
            for v in miss:
                assert abs(float(r_logit[b, v] - kth[b])) < tol_abs, (model_cls.__name__, b, v, tol_abs)
        print("configs[1] B=256 slice vs %s: tag logits rel %.3g, %d of %d top-50 entries swapped at near-ties (< %.3g)"
              % (model_cls.__name__, err, swapped, n * cfg.topk, tol_abs))
        assert err <= tol_rel
        assert swapped <= n * cfg.topk // 50                      # at most 2 % of the entries


def test_config4_beam4_b256_vs_oracle_slice():
    """BASELINE configs[3] at its full size against the oracle: beam-4 ids of the B = 256 fast path on a 16-image slice against
    the quantisation-matched oracle's beam search (oracle/port.py beam_search over QuantPortModel). Beam search has no
    per-row teacher forcing, so the comparison is on whole hypotheses: the best hypothesis must be identical, or its score
    must be within 2e-3 of the oracle's best (a near-tie between two hypotheses)."""
    from oracle import port
    kw = dict(num_beams=4, num_keep_best=1, length_penalty=1.0)
    cfg, m = _model("16_384", 1.9, kw, 256)
    B, lo, n = 256, 64, 16
    data = _data(cfg, B, seed=7)
    ids, lp = m(data)
    sd_dev = {k: v.to(DEV) for k, v in _CACHE[("16_384", 1.9, 0)][1].items()}
    sub = {k: v[lo:lo + n].contiguous() for k, v in data.items()}
    extra = synth.default_test_extra_input(cfg, **kw)
    q_ids, q_lp = _on_gpu(lambda: port.caption(port.QuantPortModel(cfg, sd_dev), sub, extra, algorithm="cached"))
    same = (ids[lo:lo + n, 0].cpu() == q_ids[:, 0].cpu()).all(dim=1)
    d = (lp[lo:lo + n, 0].cpu() - q_lp[:, 0].cpu()).abs()
    print("configs[3] beam-4 B=256 slice vs quantised oracle: %d of %d best hypotheses identical, max |score diff| %.3g"
          % (int(same.sum()), n, float(d.max())))
    assert bool((same | (d < 2e-3)).all())
    assert int(same.sum()) >= n - 2
    assert float(d[same].max()) < 1e-3
