"""Visible od/tag label region (dataset.py:395-408, modeling_bert.py:1435-1489): kernel-level
parity of the label-aware entry points against torch fp32 expressions, and end-to-end parity against golden vectors of the
unmodified reference (both label-embedding recipes, the mid-caption recipe flip, greedy and beam search)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import port  # noqa: E402
from tests.helpers import compare_ids_gap_aware, golden_setup, load_golden  # noqa: E402
from tests.test_decode_kernels_gpu import _da_inputs, _decode_attention_ref  # noqa: E402
from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import ops, synth  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402

DEV = "cuda:0"


def _rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _ref_attention_labels(qkv, heads, scale, n_base, n_extra):
    """softmax(QK^T*scale + mask)V with the label-region mask: rows < n_base see keys < n_base; rows >= n_base see
    keys < n_base + n_extra[b] (the reference adds -10000 to hidden keys, which underflows to exactly 0)."""
    B, N, H3 = qkv.shape
    q, k, v = qkv.float().view(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    ar = torch.arange(N)
    for b in range(B):
        vis = (ar.view(1, N) < n_base) | ((ar.view(N, 1) >= n_base) & (ar.view(1, N) < n_base + int(n_extra[b])))
        s[b].masked_fill_(~vis.unsqueeze(0), -1e30)
    return (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, N, heads * 64)


# (B, n_base, K label rows, heads): ViT-B/16-384 (578 + 50), 16_224 (198 + 50), 32_384 (146 + 50: a chunk that is full for the
# label rows and ragged for the image rows of the same warp), tiny (18 + 50: labels already in the first key chunk)
LAB_CASES = [(3, 578, 50, 12), (4, 198, 50, 12), (3, 146, 50, 12), (4, 18, 50, 2), (2, 130, 62, 1)]


@pytest.mark.parametrize("impl", ["auto", "simt"])
@pytest.mark.parametrize("B,n_base,K,heads", LAB_CASES)
def test_attention_labels_bf16(B, n_base, K, heads, impl):
    N = n_base + K
    qkv = _rnd(B, N, 3 * heads * 64, seed=3).to(torch.bfloat16)
    n_extra = torch.tensor(([0, K, 4, 23] * 2)[:B], dtype=torch.int32).clamp(max=K)
    out = torch.zeros(B, N, heads * 64, device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv.to(DEV), out, B, N, heads, 0.125, impl=impl, n_base=n_base, n_extra=n_extra.to(DEV))
    ref = _ref_attention_labels(qkv, heads, 0.125, n_base, n_extra)
    torch.testing.assert_close(out.float().cpu(), ref, rtol=2e-2, atol=2e-2)
    # the unmasked kernel on the same buffer must differ for the image rows of an image with hidden labels (mask is live)
    out2 = torch.zeros_like(out)
    ops.attention(qkv.to(DEV), out2, B, N, heads, 0.125, impl=impl)
    assert float((out2[0, :n_base].float() - out[0, :n_base].float()).abs().max()) > 1e-3


@pytest.mark.parametrize("B,n_base,K,heads", LAB_CASES)
def test_attention_labels_exact_fp32(B, n_base, K, heads):
    N = n_base + K
    qkv = _rnd(B, N, 3 * heads * 64, seed=5)
    n_extra = torch.tensor(([K, 0, 7, 31] * 2)[:B], dtype=torch.int32).clamp(max=K)
    out = torch.zeros(B, N, heads * 64, device=DEV)
    ops.attention(qkv.to(DEV), out, B, N, heads, 0.125, n_base=n_base, n_extra=n_extra.to(DEV))
    torch.testing.assert_close(out.cpu(), _ref_attention_labels(qkv, heads, 0.125, n_base, n_extra), rtol=1e-4, atol=1e-5)


def test_attention_labels_many_ctas_bf16():
    """Several waves of co-resident CTAs with per-image label counts (barrier phases across tiles, ragged last chunk)."""
    B, n_base, K, heads = 40, 578, 50, 12
    N = n_base + K
    qkv = _rnd(B, N, 3 * heads * 64, seed=11).to(torch.bfloat16)
    n_extra = torch.tensor([(7 * i) % (K + 1) for i in range(B)], dtype=torch.int32)
    out = torch.zeros(B, N, heads * 64, device=DEV, dtype=torch.bfloat16)
    for _ in range(2):
        ops.attention(qkv.to(DEV), out, B, N, heads, 0.125, n_base=n_base, n_extra=n_extra.to(DEV))
    torch.cuda.synchronize()
    torch.testing.assert_close(out.float().cpu(), _ref_attention_labels(qkv, heads, 0.125, n_base, n_extra), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("B,n_base,K,heads,E,cur_len", [(3, 578, 50, 12, 1, 7), (2, 198, 50, 12, 5, 19), (2, 18, 50, 2, 3, 2),
                                                        (2, 146, 50, 12, 8, 11)])
def test_decode_attention_labels(B, n_base, K, heads, E, cur_len, dtype):
    """Cp = n_base + K context rows allocated per image, ctx_vis[b] of them visible: equals the plain kernel contract on a
    context truncated to ctx_vis[b] rows."""
    Cp = n_base + K
    ctx, stepq, anc = _da_inputs(B, Cp, heads, E, cur_len, dtype, seed=21)
    vis = torch.tensor(([n_base, Cp, n_base + 4][:B]), dtype=torch.int32)
    H = heads * 64
    R = B * E
    ref = torch.zeros(2 * R, H)
    for b in range(B):              # reference: one image at a time with its own truncated context
        c = int(vis[b])
        sub_anc = (anc[:, b * E:(b + 1) * E] - b * E)
        r = _decode_attention_ref(ctx[b:b + 1, :c].contiguous(), stepq[:, 2 * b * E:2 * (b + 1) * E].contiguous(), sub_anc, 1, c,
                                  heads, E, cur_len, 0.125)
        ref[2 * b * E:2 * (b + 1) * E] = r
    for impl in (("auto", "simt") if dtype == torch.bfloat16 else ("auto",)):
        out = torch.full((2 * R, H), float("nan"), dtype=dtype, device=DEV)
        ops.decode_attention(ctx.to(DEV), stepq.to(DEV), anc.to(DEV), out, B, Cp, heads, E, cur_len, 0.125, impl=impl,
                             ctx_vis=vis.to(DEV))
        got = out.float().cpu()
        assert torch.isfinite(got).all()
        if dtype == torch.bfloat16:
            assert float((got - ref).abs().max()) < 2.5e-2 and float((got - ref).norm() / ref.norm()) < 6e-3
        else:
            np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("recipe_ln", [False, True])
def test_label_rows_and_pitched_ctx(dtype, recipe_ln):
    B, N, H, K, V = 3, 17, 768, 50, 999
    C, Cp = N + 1, N + 1 + K
    d = lambda t: t.to(DEV)
    cap, tag = d(_rnd(B, N, H, seed=1)), d(_rnd(B, N, H, seed=2))
    word, pos, typ = d(_rnd(V, H, seed=3)), d(_rnd(512, H, seed=4)), d(_rnd(H, seed=5))
    g, be = d(1 + 0.1 * _rnd(H, seed=6)), d(0.1 * _rnd(H, seed=7))
    idx = torch.randint(0, V, (B, K), generator=torch.Generator().manual_seed(8), dtype=torch.int32)
    cf = torch.full((B, Cp, H), 7.0, device=DEV)
    ct = torch.full((B, Cp, H), 7.0, device=DEV, dtype=dtype)
    ops.assemble_ctx(cap, tag, cf, ct, B, N, H, rows_per_image=Cp)
    ops.label_rows(d(idx), 102, recipe_ln, 20, word, pos, typ, g, be, 1e-12, cf, ct, B, Cp, C)
    ref_ctx = torch.cat([tag[:, 0:1], cap], 1)
    assert torch.equal(cf[:, :C], ref_ctx)
    ids = idx.long().clone()
    ids[:, -1] = 102
    lab = word[d(ids)]
    if recipe_ln:
        lab = F.layer_norm(lab + pos[20:20 + K] + typ, (H,), g, be, 1e-12)
        torch.testing.assert_close(cf[:, C:], lab, rtol=1e-5, atol=1e-5)
    else:
        assert torch.equal(cf[:, C:], lab)
    if dtype == torch.bfloat16:          # exact mode keeps one fp32 buffer (the operand copy aliases the residual copy)
        assert torch.equal(ct.float(), cf.to(dtype).float())


def _build(cfg, sd, extra, mode, **kw):
    m = FastImageCaptioning(cfg, test_extra_input=extra, mode=mode, **kw)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


def _to_dev(data):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in data.items()}


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("name", ["g10_greedy_labels_16_224", "g12_greedy_labels_flip_16_224", "g14_greedy_labels_raw_16_224"])
def test_exact_mode_greedy_labels_match_reference_golden(name, graph):
    """Token ids bit-exact and log-probs to 2e-4 against the unmodified reference fed the data layer's label masks:
    'ln' recipe until the last step (g10), recipe flip at cur_len 8 with prefill + caption replay (g12), 'raw' recipe
    throughout (g14); per-sample visible label counts incl. 0 and 50."""
    z, meta = load_golden(name)
    cfg, sd, data, extra = golden_setup(meta)
    m = _build(cfg, sd, extra, "fp32", max_batch=8, use_cuda_graph=graph)
    for rep in range(2 if graph else 1):
        ids, lp = m(_to_dev(data))
        excused = compare_ids_gap_aware(ids.cpu().numpy()[:, 0], z["ids"][:, 0], z["step_top_val"], 2e-4, name)
        if excused == 0:
            np.testing.assert_allclose(lp.cpu().numpy(), z["logprobs"], atol=2e-4)
    # the label region is live: the same images under the eval mask (no visible labels) give different log-probs
    plain = synth.make_text_inputs(cfg, meta["batch"])
    plain["image"] = data["image"]
    ids0, lp0 = m(_to_dev(plain))
    assert float((lp0.cpu() - torch.from_numpy(z["logprobs"])).abs().max()) > 5e-4


@pytest.mark.parametrize("name", ["g11_beam3_labels_16_224", "g13_beam3_labels_flip_16_224"])
def test_exact_mode_beam_labels_match_reference_golden(name):
    z, meta = load_golden(name)
    cfg, sd, data, extra = golden_setup(meta)
    m = _build(cfg, sd, extra, "fp32", max_batch=8)
    for rep in range(2):                       # second call replays the captured graph (flip + replay inside it)
        ids, lp = m(_to_dev(data))
        assert np.array_equal(ids.cpu().numpy(), z["ids"])
        np.testing.assert_allclose(lp.cpu().numpy(), z["logprobs"], atol=2e-4)


def test_chunked_forward_keeps_the_batch_level_recipe():
    """The recipe follows the FIRST sample of the caller's batch (modeling_bert.py:1435) even when forward() feeds the batch to
    the engine in chunks of max_batch."""
    z, meta = load_golden("g12_greedy_labels_flip_16_224")
    cfg, sd, data, extra = golden_setup(meta)
    m = _build(cfg, sd, extra, "fp32", max_batch=2)
    ids, lp = m(_to_dev(data))
    compare_ids_gap_aware(ids.cpu().numpy()[:, 0], z["ids"][:, 0], z["step_top_val"], 2e-4, "chunked")


@pytest.mark.parametrize("name", ["g10_greedy_labels_16_224", "g12_greedy_labels_flip_16_224"])
def test_bf16_mode_greedy_labels_gap_aware(name):
    z, meta = load_golden(name)
    cfg, sd, data, extra = golden_setup(meta)
    m = _build(cfg, sd, extra, "bf16", max_batch=8)
    ids, lp = m(_to_dev(data))
    compare_ids_gap_aware(ids.cpu().numpy()[:, 0], z["ids"][:, 0], z["step_top_val"], 6e-2, name)


def test_tiny_sampling_with_labels_matches_oracle():
    """Sampling (E = 3 sequences per image share the image's label rows) on the tiny model, oracle drawing the same noise."""
    from oracle import philox
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=0.5, tag_bias=-1.5)
    B, K, seed = 3, 3, 99
    data = synth.make_text_inputs(cfg, B, n_label=[50, 0, 9])
    data["image"] = synth.make_images(cfg, B, seed=3)
    extra = synth.default_test_extra_input(cfg, do_sample=True, num_return_sequences=K)
    pm = port.PortModel(cfg, sd)
    with torch.no_grad():
        rids, rlp = port.caption(pm, data, extra, algorithm="cached", sampler=philox.make_sampler(seed))
    m = _build(cfg, sd, extra, "fp32", sample_seed=seed, use_cuda_graph=False)
    ids, lp = m(_to_dev(data))
    agree = float((ids.cpu() == rids).float().mean())
    assert agree >= 0.98, agree
