"""SCST training-side use: vitcap_b200.scst.sequence_logprobs -- ONE teacher-forced, differentiable
pass over the module's parameters -- against the reference's own way of obtaining ``sample_logprobs`` with gradients
(legacy pipeline tagger_caption_uni_pipeline_expanding.py:447-468): autograd through the 19 full-model calls of ``generate``,
restated by oracle/port.py's ``faithful`` algorithm (pinned to the live reference by tests/test_oracle_golden.py). Same sampled
tokens, same rewards: the sequence log-probs AND the parameter gradients of the SCST loss must agree. Runs on the CPU (the
gradient pass is plain PyTorch by design, see the module docstring)."""
import numpy as np
import pytest
import torch

from oracle import port
from vitcap_b200 import config as vcfg
from vitcap_b200 import scst, synth
from vitcap_b200.model import FastImageCaptioning


def _sample_with_the_oracle(cfg, sd, data, extra, seed):
    """Sampled captions (cached algorithm, torch.multinomial under a fixed seed) + the tokens chosen at every step."""
    chosen = []
    gen = torch.Generator().manual_seed(seed)

    def rec(logits, cur_len):
        nxt = torch.multinomial(torch.softmax(logits, dim=-1), num_samples=1, generator=gen).squeeze(1)
        chosen.append(nxt.clone())
        return nxt
    with torch.no_grad():
        ids, lp = port.caption(port.PortModel(cfg, sd), data, extra, algorithm="cached", sampler=rec)
    R, L = ids.shape[0], ids.shape[2]
    raw = torch.full((R, L), int(extra["pad_token_id"]), dtype=torch.long)
    raw[:, 0] = int(extra["bos_token_id"])
    unf = torch.ones(R, dtype=torch.bool)
    for t, nxt in enumerate(chosen, start=1):
        raw[:, t] = torch.where(unf, nxt, raw[:, t])
        unf = unf & (nxt != int(extra["eos_token_ids"][0]))
    return ids, lp, raw, chosen


@pytest.mark.parametrize("temperature,top_k", [(1.0, 0), (0.8, 40)])
def test_sequence_logprobs_and_scst_gradients_match_the_19_call_formulation(temperature, top_k):
    torch.manual_seed(0)
    cfg = vcfg.tiny(enc_blocks=2, split_blocks=1, dec_layers=2, vocab=600, inter=768)
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=2.5)
    B, K = 2, 2
    data = synth.make_text_inputs(cfg, B)
    data["image"] = synth.make_images(cfg, B, seed=3)
    extra = synth.default_test_extra_input(cfg, do_sample=True, num_return_sequences=K, temperature=temperature, top_k=top_k)
    ids, lp_ref, raw, chosen = _sample_with_the_oracle(cfg, sd, data, extra, seed=5)
    assert int((raw[:, 1:] == 102).any(1).sum()) >= 1 and int((ids[:, 0, -1] == 102).sum()) >= 1     # early ends AND full-length rows
    reward = torch.tensor([0.7, -0.4, 1.3, -1.1])

    # (a) the reference's formulation: autograd through every full-model call of generate (faithful algorithm), tokens forced
    pm = port.PortModel(cfg, sd)
    pm.sd = {k: v.clone().requires_grad_(True) for k, v in pm.sd.items()}
    it = iter(chosen)
    ids_f, lp_f = port.caption(pm, data, extra, algorithm="faithful", sampler=lambda logits, cur_len: next(it))
    assert torch.equal(ids_f, ids)
    loss_f = -(lp_f[:, 0] * reward).mean()                 # ScstRewardCriterion.forward, utils_caption_evaluate.py:196-198
    loss_f.backward()

    # (b) one teacher-forced pass over the module's own parameters
    m = FastImageCaptioning(cfg, test_extra_input=extra, mode="fp32")
    m.load_state_dict(sd)
    lp_m = scst.sequence_logprobs(m, data["image"], raw, K, temperature=temperature, top_k=top_k)
    assert lp_m.requires_grad
    np.testing.assert_allclose(lp_m.detach().numpy(), lp_f[:, 0].detach().numpy(), atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(lp_m.detach().numpy(), lp_ref[:, 0].numpy(), atol=2e-5, rtol=1e-5)
    loss_m = -(lp_m * reward).mean()
    loss_m.backward()

    P = scst.reference_params(m)
    word = "module.bert.embeddings.word_embeddings.weight"
    checked = 0
    g_scale = max(float(v.grad.norm()) for v in pm.sd.values() if v.grad is not None)
    for key, g_f in ((k, v.grad) for k, v in pm.sd.items()):
        if key == "module.cls.predictions.decoder.weight" or key not in P:
            continue
        if key == word:                                    # tied in the module (modeling_bert.py:728-730): the gradients add up
            g_f = g_f + pm.sd["module.cls.predictions.decoder.weight"].grad
        g_m = P[key].grad
        if g_f is None or float(g_f.abs().max()) == 0.0:
            # parameters the SCST loss does not reach: concept head (only its top-k INDICES are used), dead pooler, extra
            # embeddings, the timm classifier, and -- with no visible label region -- nothing else
            assert g_m is None or float(g_m.abs().max()) == 0.0, key
            continue
        # (the key biases get a mathematically zero gradient -- a constant added to every key of a softmax row -- which is
        # rounding noise in both formulations: hence the absolute term)
        err = float((g_m - g_f).norm())
        assert err < 2e-3 * float(g_f.norm()) + 1e-6 * g_scale, (key, err, float(g_f.norm()))
        checked += 1
    assert checked > 60
    for must in ("image_encoder.module.patch_embed.proj.weight", "module.bert.encoder.blocks.0.attn.qkv.weight",
                 "module.bert.encoder.tag_blocks.0.mlp.fc2.weight", "module.bert.decoder.layer.1.output.dense.weight", word,
                 "module.cls.predictions.bias"):
        assert float(P[must].grad.abs().max()) > 0.0, must


def test_filter_restates_top_k_top_p_filtering():
    lg = torch.randn(5, 3, 200) * 3.0
    for top_k, top_p in ((10, 1.0), (0, 0.7), (25, 0.4)):
        ref = port.top_k_top_p_filtering(lg.reshape(15, 200).clone(), top_k=top_k, top_p=top_p).reshape(5, 3, 200)
        got = scst._filter(lg.clone(), top_k, top_p)
        assert torch.equal(torch.isinf(ref), torch.isinf(got))
        assert torch.equal(ref[~torch.isinf(ref)], got[~torch.isinf(got)])
