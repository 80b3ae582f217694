"""SCST training-side use on the GPU (vitcap_b200.scst.ScstSampler): the greedy baseline and the K sampled captions come from the
CUDA kernels, the log-probs with gradients from one teacher-forced pass over the module's parameters. The two must describe
the same captions: the differentiable log-probs equal the kernels' own (exact mode: 2e-4; fast mode: bf16 operand error), and
the SCST loss back-propagates to the parameters the reference's loss reaches."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import scst, synth  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402

DEV = "cuda:0"


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-4), ("bf16", 3e-2)])
def test_scst_sampler_logprobs_with_gradients(mode, tol):
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=1.0)
    B, K = 4, 3
    data = {k: v.to(DEV) for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = synth.make_images(cfg, B, seed=3).to(DEV)
    m = FastImageCaptioning(cfg, mode=mode, sample_seed=123)
    m.load_state_dict(sd)
    m = m.to(DEV)
    greedy_ref, _ = m(data)
    out = scst.ScstSampler(m, num_return_sequences=K, temperature=0.9)(data)
    assert torch.equal(out["greedy_ids"], greedy_ref[:, 0])
    assert out["sample_ids"].shape == (B * K, 20) and out["sample_ids"].dtype == torch.int64
    lp = out["sample_logprobs"]
    assert lp.requires_grad and not out["sample_ids"].requires_grad           # legacy pipeline :463-464
    np.testing.assert_allclose(lp.detach().cpu().numpy(), out["sample_logprobs_kernels"].cpu().numpy(), atol=tol, rtol=tol)
    # the raw ids differ from the returned ones only by the forced EOS of captions that ran to max_length
    raw = m.last_raw_ids
    diff = raw != out["sample_ids"]
    assert not bool(diff[:, :-1].any())
    assert bool((out["sample_ids"][:, -1][diff[:, -1]] == 102).all())
    assert len({tuple(r.tolist()) for r in out["sample_ids"][:K]}) > 1         # the samples of one image differ
    # ScstRewardCriterion.forward (utils_caption_evaluate.py:196-198) with stand-in rewards
    reward = torch.linspace(-1.0, 1.0, B * K, device=DEV)
    loss = -(lp * reward).mean()
    loss.backward()
    P = scst.reference_params(m)
    for key in ("image_encoder.module.patch_embed.proj.weight", "module.bert.encoder.blocks.0.attn.qkv.weight",
                "module.bert.decoder.layer.0.intermediate.dense.weight", "module.bert.embeddings.word_embeddings.weight",
                "module.cls.predictions.transform.dense.weight"):
        g = P[key].grad
        assert g is not None and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0.0, key
    # an optimizer step invalidates the packed kernel weights (parameter version counters): the next call re-packs
    with torch.no_grad():
        P["module.cls.predictions.bias"].add_(0.01)
    ids2, _ = m(data)
    assert ids2.shape == greedy_ref.shape
