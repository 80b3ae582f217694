"""Pins the CPU oracle (oracle/port.py) to what the unmodified reference computed
(tests/golden/*.npz, produced by oracle/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import port
from tests.helpers import golden_setup, load_golden

CACHED_CASES = ["g1_greedy_16_384", "g3_greedy_eos_16_224", "g4_beam3_keep3_16_224", "g6_greedy_32_384",
                "g7_greedy_dec12_16_224", "g9_greedy_refinit_16_224",
                # visible od/tag label region: 'ln' recipe until the last step, flip at step 8,
                # 'raw' recipe throughout
                "g10_greedy_labels_16_224", "g12_greedy_labels_flip_16_224", "g14_greedy_labels_raw_16_224"]


def _run(name, algorithm):
    z, meta = load_golden(name)
    cfg, sd, data, extra = golden_setup(meta)
    pm = port.PortModel(cfg, sd)
    info, trace = {}, []
    if meta["torch_seed"] is not None:
        torch.manual_seed(meta["torch_seed"])
    with torch.no_grad():
        ids, lp = port.caption(pm, data, extra, algorithm=algorithm, info=info, trace=trace)
    return z, meta, cfg, ids, lp, info, trace


@pytest.mark.parametrize("name", CACHED_CASES)
def test_cached_port_matches_reference(name):
    z, meta, cfg, ids, lp, info, trace = _run(name, "cached")
    assert np.array_equal(ids.numpy(), z["ids"])
    np.testing.assert_allclose(lp.numpy(), z["logprobs"], rtol=0, atol=5e-5)
    logit, prob, idx, n = info["tag"]
    assert np.array_equal(idx.numpy(), z["tag_topk_idx"])
    np.testing.assert_allclose(prob.numpy(), z["tag_topk_prob"], atol=1e-5)
    assert np.array_equal(n.numpy(), z["tag_topk_len"])
    np.testing.assert_allclose(logit[:, :512].numpy(), z["tag_logit_head"], atol=2e-5)
    np.testing.assert_allclose(info["img_feats"][:, ::29, ::37].numpy(), z["img_feats_s"], atol=1e-5)
    np.testing.assert_allclose(info["cap"][:, ::29, ::37].numpy(), z["cap_feats_s"], atol=5e-4, rtol=1e-4)
    np.testing.assert_allclose(info["tag_feats"][:, ::29, ::37].numpy(), z["tag_feats_s"], atol=5e-4, rtol=1e-4)
    if meta["decode"].get("num_beams", 1) == 1:
        # per-step top-4 logits of the step the reference actually ran
        for s, t in enumerate(trace):
            v, i = t.topk(4, dim=-1)
            np.testing.assert_allclose(v.numpy(), z["step_top_val"][s], atol=5e-5)


@pytest.mark.parametrize("name", ["g2_beam4_16_384", "g11_beam3_labels_16_224", "g13_beam3_labels_flip_16_224"])
def test_cached_port_beam(name):
    z, meta, cfg, ids, lp, info, trace = _run(name, "cached")
    assert np.array_equal(ids.numpy(), z["ids"])
    np.testing.assert_allclose(lp.numpy(), z["logprobs"], atol=5e-5)


def test_label_recipe_flip_points():
    """The step at which the reference switches the label embedding depends on the FIRST sample's tag count
    (modeling_bert.py:1435): recorded counts 50 / 39 / 15 put it at the last step, at cur_len 8, and before step 1."""
    for name, flip in (("g10_greedy_labels_16_224", 19), ("g12_greedy_labels_flip_16_224", 8), ("g14_greedy_labels_raw_16_224", 1)):
        z, meta = load_golden(name)
        t0 = int(z["tag_topk_len"][0])
        first_raw = min(L for L in range(1, 20) if t0 + 20 <= L + 1 + 50)
        assert first_raw == max(1, t0 - 31) == flip


@pytest.mark.parametrize("name", ["g5_sample5_16_224", "g8_sample_filtered_16_224"])
def test_sampling_semantics_with_torch_multinomial(name):
    """do_sample path: same torch.manual_seed + torch.multinomial => same draws as the reference
    (only meaningful on the torch build that made the golden)."""
    z, meta = load_golden(name)
    if meta["torch"] != torch.__version__:
        pytest.skip("golden was drawn with torch %s" % meta["torch"])
    z, meta, cfg, ids, lp, info, trace = _run(name, "cached")
    assert ids.shape == z["ids"].shape == (meta["batch"] * meta["decode"]["num_return_sequences"], 1, 20)
    assert np.array_equal(ids.numpy(), z["ids"])
    np.testing.assert_allclose(lp.numpy(), z["logprobs"], atol=5e-5)


@pytest.mark.parametrize("name", ["g6_greedy_32_384", "g4_beam3_keep3_16_224", "g12_greedy_labels_flip_16_224"])
def test_faithful_port_matches_reference(name):
    """The uncached restatement (what bench.py times as the CPU baseline) reproduces the reference,
    including the number of full-model calls (19 for a 20-token caption: no KV cache is live,
    modeling_bert.py:1072-1073)."""
    z, meta, cfg, ids, lp, info, trace = _run(name, "faithful")
    assert np.array_equal(ids.numpy(), z["ids"])
    np.testing.assert_allclose(lp.numpy(), z["logprobs"], atol=5e-5)
    assert info["n_calls"] == meta["n_model_calls"]


_STAGED_CHECK = r"""
import sys, numpy as np, torch
from oracle import ref_loader
from tests.helpers import golden_setup, load_golden
assert ref_loader.REF_KIND == "env" and ref_loader.REF_ROOT.endswith("_ref"), (ref_loader.REF_ROOT, ref_loader.REF_KIND)
z, meta = load_golden(sys.argv[1])
cfg, sd, data, extra = golden_setup(meta)
ref, tok = ref_loader.build_reference(meta["variant"])
import src.layers.bert.modeling_bert as mb
assert "reference_pyc.zip" in mb.__file__ and mb.__file__.endswith(".pyc"), mb.__file__     # the archive, not /root/reference
ref.load_state_dict(sd, strict=True)
ref.test_extra_input = extra
data = dict(data)
data["key"] = ["k%d" % i for i in range(data["image"].shape[0])]
with torch.no_grad(), ref_loader.on_cpu():
    ids, lp = ref(data)
assert np.array_equal(ids.numpy(), z["ids"]) and np.array_equal(lp.numpy(), z["logprobs"])
print("STAGED-REFERENCE-OK")
"""


def test_staged_reference_bytecode_reproduces_the_golden():
    """oracle/_ref (oracle/build_ref.py: the unmodified reference compiled to bytecode, what bench.py's CPU legs time on the GPU
    box) is the reference: imported without /root/reference it returns the committed golden bit for bit."""
    import os
    import subprocess
    import sys
    from oracle import build_ref
    staged = build_ref.staged_root()
    if staged is None:
        pytest.skip("oracle/_ref not staged (python -m oracle.build_ref needs the reference tree)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VITCAP_REFERENCE_ROOT=staged, PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", _STAGED_CHECK, "g9_greedy_refinit_16_224"], cwd=root, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "STAGED-REFERENCE-OK" in r.stdout, r.stderr[-2000:]


def test_half_storage_oracle_is_eight_times_closer_to_the_fp32_algorithm():
    """port.half_store(): every operand rounding of QuantPortModel as an IEEE-half rounding (the spec of a VITCAP_STORE=fp16
    process). On the tiny model the caption features sit ~8 x closer to the fp32 algorithm than with bf16 roundings (11-bit
    against 8-bit significands), and leaving the block restores the bf16 roundings."""
    from vitcap_b200 import config as vcfg
    from vitcap_b200 import synth
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=1.0)
    data = synth.make_text_inputs(cfg, 2)
    data["image"] = synth.make_images(cfg, 2, seed=3)
    extra = synth.default_test_extra_input(cfg)
    out = {}
    with torch.no_grad():
        for name in ("fp32", "half", "bf16"):
            info = {}
            if name == "fp32":
                port.caption(port.PortModel(cfg, sd), data, extra, algorithm="cached", info=info)
            elif name == "half":
                with port.half_store():
                    port.caption(port.QuantPortModel(cfg, sd, decode_x3=False), data, extra, algorithm="cached", info=info)
            else:
                port.caption(port.QuantPortModel(cfg, sd, decode_x3=False), data, extra, algorithm="cached", info=info)
            out[name] = info["cap"]
    rel = lambda a, b: float((a - b).norm() / b.norm())                     # noqa: E731
    e_half, e_bf16 = rel(out["half"], out["fp32"]), rel(out["bf16"], out["fp32"])
    assert e_half < 1e-3 and 4.0 * e_half < e_bf16 < 1e-2, (e_half, e_bf16)
    assert float(port.q_bf16(torch.tensor([1.003]))[0]) == 1.0              # bf16 roundings are back
