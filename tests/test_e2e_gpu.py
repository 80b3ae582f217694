"""End-to-end parity of the CUDA caption path (through the drop-in module and the C ABI) against
(a) golden vectors produced by the unmodified reference, (b) the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import philox, port  # noqa: E402
from tests.helpers import compare_ids_gap_aware, golden_setup, load_golden  # noqa: E402
from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import synth  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402

DEV = "cuda:0"
FP32_MIN_GAP = 2e-4     # an fp32 argmax may legitimately flip only where the reference's own top-2 gap is below this
BF16_MIN_GAP = 6e-2     # bf16 operands perturb logits by ~1e-2 (measured)


def build(cfg, sd, extra, mode, **kw):
    m = FastImageCaptioning(cfg, test_extra_input=extra, mode=mode, **kw)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


def to_dev(data):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in data.items()}


def rel_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("name", ["g1_greedy_16_384", "g3_greedy_eos_16_224", "g6_greedy_32_384", "g7_greedy_dec12_16_224",
                                  "g9_greedy_refinit_16_224"])
def test_exact_mode_greedy_matches_reference_golden(name):
    z, meta = load_golden(name)
    cfg, sd, data, extra = golden_setup(meta)
    m = build(cfg, sd, extra, "fp32", max_batch=8)
    data["key"] = list(range(meta["batch"]))
    ids, lp = m(to_dev(data))
    assert ids.dtype == torch.int64 and tuple(ids.shape) == z["ids"].shape
    excused = compare_ids_gap_aware(ids.cpu().numpy()[:, 0], z["ids"][:, 0], z["step_top_val"], FP32_MIN_GAP, name)
    if excused == 0:
        np.testing.assert_allclose(lp.cpu().numpy(), z["logprobs"], atol=2e-4)
    # concept head: indices bit-exact, probabilities close
    lg, idx, pr, n = m.forward_tags(data["image"].to(DEV))
    assert np.array_equal(idx.cpu().numpy(), z["tag_topk_idx"])
    np.testing.assert_allclose(pr.cpu().numpy(), z["tag_topk_prob"], atol=2e-5)
    assert np.array_equal(n.cpu().numpy(), z["tag_topk_len"])
    np.testing.assert_allclose(lg[:, :512].cpu().numpy(), z["tag_logit_head"], atol=1e-4)
    cap, tag = m.encode_features(data["image"].to(DEV))
    np.testing.assert_allclose(cap[:, ::29, ::37].cpu().numpy(), z["cap_feats_s"], atol=2e-3, rtol=1e-3)
    np.testing.assert_allclose(tag[:, ::29, ::37].cpu().numpy(), z["tag_feats_s"], atol=2e-3, rtol=1e-3)


@pytest.mark.parametrize("name", ["g2_beam4_16_384", "g4_beam3_keep3_16_224"])
def test_exact_mode_beam_matches_reference_golden(name):
    z, meta = load_golden(name)
    cfg, sd, data, extra = golden_setup(meta)
    m = build(cfg, sd, extra, "fp32", max_batch=8)
    ids, lp = m(to_dev(data))
    assert tuple(ids.shape) == z["ids"].shape and tuple(lp.shape) == z["logprobs"].shape
    assert np.array_equal(ids.cpu().numpy(), z["ids"])
    np.testing.assert_allclose(lp.cpu().numpy(), z["logprobs"], atol=2e-4)


def _oracle(cfg, sd, data, extra, sampler=None):
    pm = port.PortModel(cfg, sd)
    info, trace = {}, []
    with torch.no_grad():
        ids, lp = port.caption(pm, data, extra, algorithm="cached", info=info, trace=trace, sampler=sampler)
    return ids, lp, info, trace


@pytest.mark.parametrize("mode,gap", [("fp32", FP32_MIN_GAP), ("bf16", BF16_MIN_GAP)])
@pytest.mark.parametrize("graph", [False, True])
def test_tiny_greedy_vs_oracle(mode, gap, graph):
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=1.0)
    B = 5
    data = synth.make_text_inputs(cfg, B)
    data["image"] = synth.make_images(cfg, B, seed=3)
    extra = synth.default_test_extra_input(cfg)
    rids, rlp, info, trace = _oracle(cfg, sd, data, extra)
    m = build(cfg, sd, extra, mode, max_batch=3, use_cuda_graph=graph)     # exercises batch chunking (3 + 2)
    for rep in range(2):                                                   # second call replays the captured graph
        ids, lp = m(to_dev(data))
        top = torch.stack([t.topk(2).values for t in trace]).numpy()
        compare_ids_gap_aware(ids.cpu().numpy()[:, 0], rids.numpy()[:, 0], top, gap, "%s rep%d" % (mode, rep))


@pytest.mark.parametrize("search", ["greedy", "sample", "beam"])
@pytest.mark.parametrize("mode,every", [("fp32", 1), ("bf16", 1), ("bf16", 3)])
def test_captured_decode_loop_exits_early_on_the_device(search, mode, every, monkeypatch):
    """`if cur_unfinished.max() == 0: break` / `if all(done): break` (modeling_utils.py:865-867, 1071-1073) inside the captured
    loop: every step after the first is the body of a conditional graph node. With an EOS planted so strongly that every caption
    ends within a few tokens, (1) ids and log-probs equal the loop without the conditional nodes (VITCAP_EARLY_EXIT=0) bit for
    bit, (2) the steps after the last live one did NOT run: the q|k|v rows they would have written keep a sentinel."""
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=9.0)
    B = 5
    data = synth.make_text_inputs(cfg, B)
    data["image"] = synth.make_images(cfg, B, seed=3)
    kw = {"greedy": {}, "sample": dict(do_sample=True, num_return_sequences=3, temperature=0.8),
          "beam": dict(num_beams=3, num_keep_best=2, length_penalty=0.8)}[search]
    extra = synth.default_test_extra_input(cfg, **kw)
    E = {"greedy": 1, "sample": 3, "beam": 3}[search]
    out = {}
    # every: decode steps per condition (engine._exit_every: 1 for small batches, 2 / 4 for large ones; forced here)
    monkeypatch.setenv("VITCAP_EARLY_EXIT_EVERY", str(every))
    for early in ("0", "1"):
        monkeypatch.setenv("VITCAP_EARLY_EXIT", early)
        m = build(cfg, sd, extra, mode, sample_seed=11, use_cuda_graph=True)
        assert m.engine.early_exit == (early == "1")
        m(to_dev(data))                                   # eager warm-up + capture
        ws = m.engine._decoder_ws(B, E, 20)
        ws["step_qkv"].fill_(7.0)                         # sentinel in every step's q|k|v rows
        ids, lp = m(to_dev(data))                         # replay
        torch.cuda.synchronize()
        out[early] = (ids.clone(), lp.clone(), ws["step_qkv"].float().clone())
    assert torch.equal(out["0"][0], out["1"][0]) and torch.equal(out["0"][1], out["1"][1])
    ids = out["1"][0]
    if search == "beam":
        n_live = 19                                       # beams end when the pool is full and cannot improve: count from the rows
    else:
        eos = int(extra["eos_token_ids"][0])
        first = (ids[:, 0] == eos).float().argmax(1)      # position of the first EOS of every caption
        assert bool((ids[:, 0] == eos).any(1).all()), "the planted EOS must end every caption"
        n_live = int(first.max())                         # steps 1..n_live ran (the step that emits the last EOS included)
        assert n_live <= 6
    touched0 = (out["0"][2] != 7.0).flatten(2).any(2)     # [L, max_len]: did step s+1 write its rows?
    touched1 = (out["1"][2] != 7.0).flatten(2).any(2)
    assert bool(touched0[:, :19].all())                   # without the conditional nodes all 19 steps run
    if search != "beam":
        # step 1 always runs; a group of `every` steps starting at step f runs iff a caption was unfinished after step f - 1
        ran = [s == 1 or (2 + (s - 2) // every * every) <= n_live for s in range(1, 20)]
        assert ran[:n_live] == [True] * n_live and sum(ran) <= n_live + every - 1
        for s_ in range(1, 20):
            assert bool(touched1[:, s_ - 1].all()) == ran[s_ - 1] and bool(touched1[:, s_ - 1].any()) == ran[s_ - 1], (s_, touched1)
    else:
        assert bool(touched1[:, 0].all()) and not bool(touched1[:, 18].any()), touched1
    # PAD after the end, as the reference pads after its break
    if search != "beam":
        pad = int(extra["pad_token_id"])
        for r in range(ids.shape[0]):
            assert bool((ids[r, 0, int(first[r]) + 1:] == pad).all())


def test_tiny_sampling_matches_oracle_with_same_noise():
    """do_sample: Gumbel-max with Philox noise == multinomial(softmax); with the oracle drawing the same counter-based
    noise the sampled ids agree token for token (exact mode)."""
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=0.5)
    B, K, seed = 3, 4, 1234
    data = synth.make_text_inputs(cfg, B)
    data["image"] = synth.make_images(cfg, B, seed=3)
    extra = synth.default_test_extra_input(cfg, do_sample=True, num_return_sequences=K, temperature=0.9)
    rids, rlp, info, trace = _oracle(cfg, sd, data, extra, sampler=philox.make_sampler(seed))
    m = build(cfg, sd, extra, "fp32", sample_seed=seed, use_cuda_graph=False)
    ids, lp = m(to_dev(data))
    assert tuple(ids.shape) == (B * K, 1, 20)
    agree = float((ids.cpu() == rids).float().mean())
    assert agree >= 0.98, agree
    same = (ids.cpu() == rids).all(dim=-1).squeeze(1)
    np.testing.assert_allclose(lp.cpu().numpy()[same.numpy()], rlp.numpy()[same.numpy()], atol=2e-4)
    assert len(set(map(tuple, ids.cpu().numpy()[:K, 0].tolist()))) > 1      # samples of one image differ


def test_tiny_beam_vs_oracle():
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=1.0)
    B = 4
    data = synth.make_text_inputs(cfg, B)
    data["image"] = synth.make_images(cfg, B, seed=3)
    extra = synth.default_test_extra_input(cfg, num_beams=4, num_keep_best=2, length_penalty=0.7)
    rids, rlp, info, trace = _oracle(cfg, sd, data, extra)
    m = build(cfg, sd, extra, "fp32", use_cuda_graph=True)
    for rep in range(2):
        ids, lp = m(to_dev(data))
        assert np.array_equal(ids.cpu().numpy(), rids.numpy())
        np.testing.assert_allclose(lp.cpu().numpy(), rlp.numpy(), atol=2e-4)


def _matched_oracle_f16(cfg, sd, data, extra, sampler=None):
    """oracle/port.py QuantPortModel with the decode-step MLP / vocabulary head on IEEE-half operands: the executable spec of
    decode_precision='fp16'."""
    qm = port.QuantPortModel(cfg, sd, decode_f16=True)
    trace = []
    with torch.no_grad():
        ids, lp = port.caption(qm, data, extra, algorithm="cached", trace=trace, sampler=sampler)
    return ids, lp, trace


@pytest.mark.parametrize("graph", [False, True])
def test_tiny_greedy_fp16_decode_vs_matched_oracle(graph):
    """decode_precision='fp16' (decode-step MLP and vocabulary head as ONE product on IEEE-half operands) against its
    quantisation-matched oracle: identical greedy ids except at near-ties of that oracle, close log-probs; chunked batch,
    eager and captured loop (the arg-max epilogue path: no logits)."""
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=1.0)
    B = 5
    data = synth.make_text_inputs(cfg, B)
    data["image"] = synth.make_images(cfg, B, seed=3)
    extra = synth.default_test_extra_input(cfg)
    rids, rlp, trace = _matched_oracle_f16(cfg, sd, data, extra)
    m = build(cfg, sd, extra, "bf16", max_batch=3, use_cuda_graph=graph, decode_precision="fp16")
    assert m.engine.decode_f16 and not m.engine.decode_x3
    top = torch.stack([t.topk(2).values for t in trace]).numpy()
    for rep in range(2):
        ids, lp = m(to_dev(data))
        excused = compare_ids_gap_aware(ids.cpu().numpy()[:, 0], rids.numpy()[:, 0], top, 2e-2, "fp16 rep%d" % rep)
        if excused == 0:
            np.testing.assert_allclose(lp.cpu().numpy(), rlp.numpy(), atol=5e-3)


def test_tiny_fp16_decode_beam_and_sampling_read_the_half_logits():
    """Beam search and sampling need the logits: with decode_precision='fp16' they come from the decode-step kernel itself
    (VC_DEC_PARTIAL, one plane + bias). Beam search against the matched oracle (most rows identical, the others excused by a
    close score), sampling against the matched oracle drawing the same Philox noise."""
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=1.0)
    B = 4
    data = synth.make_text_inputs(cfg, B)
    data["image"] = synth.make_images(cfg, B, seed=3)
    extra = synth.default_test_extra_input(cfg, num_beams=4, num_keep_best=2, length_penalty=0.7)
    rids, rlp, _ = _matched_oracle_f16(cfg, sd, data, extra)
    m = build(cfg, sd, extra, "bf16", use_cuda_graph=True, decode_precision="fp16")
    for rep in range(2):
        ids, lp = m(to_dev(data))
        same = (ids.cpu() == rids).all(dim=-1)                       # (B, keep)
        assert float(same.float().mean()) >= 0.75, (ids.cpu(), rids)
        np.testing.assert_allclose(lp.cpu().numpy()[same.numpy()], rlp.numpy()[same.numpy()], atol=5e-3)
        # a hypothesis that differs must score within a near-tie of the oracle's
        np.testing.assert_allclose(lp.cpu().numpy(), rlp.numpy(), atol=5e-2)
    K, seed = 4, 1234
    sd2 = synth.make_state_dict(cfg, seed=7, eos_bias=0.5)
    extra2 = synth.default_test_extra_input(cfg, do_sample=True, num_return_sequences=K, temperature=0.9)
    rids, rlp, _ = _matched_oracle_f16(cfg, sd2, data, extra2, sampler=philox.make_sampler(seed))
    m = build(cfg, sd2, extra2, "bf16", sample_seed=seed, use_cuda_graph=False, decode_precision="fp16")
    ids, lp = m(to_dev(data))
    agree = float((ids.cpu() == rids).float().mean())
    assert agree >= 0.9, agree
    same = (ids.cpu() == rids).all(dim=-1).squeeze(1)
    np.testing.assert_allclose(lp.cpu().numpy()[same.numpy()], rlp.numpy()[same.numpy()], atol=5e-3)


def test_bf16_mode_features_and_tags_16_224():
    """Fast mode accuracy on the 16_224 variant: encoder features / tag logits relative error and top-50 overlap."""
    z, meta = load_golden("g3_greedy_eos_16_224")
    cfg, sd, data, extra = golden_setup(meta)
    pm = port.PortModel(cfg, sd)
    with torch.no_grad():
        cap_r, tag_r, logit_r, prob_r, idx_r, n_r = port.encode_tags(pm, data["image"])
    m = build(cfg, sd, extra, "bf16", max_batch=8)
    cap, tag = m.encode_features(data["image"].to(DEV))
    e_cap, e_tag = rel_err(cap, cap_r), rel_err(tag, tag_r)
    lg, idx, pr, n = m.forward_tags(data["image"].to(DEV))
    e_lg = rel_err(lg, logit_r)
    overlap = np.mean([len(set(a) & set(b)) / 50.0 for a, b in zip(idx.cpu().tolist(), idx_r.tolist())])
    print("bf16 rel err: cap %.3g tag %.3g tag_logits %.3g top50 overlap %.3f" % (e_cap, e_tag, e_lg, overlap))
    # against the fp32 arithmetic (operand quantisation included): the measured values (5.0e-3 / 5.0e-3 / 8.3e-3, overlap 0.99)
    # + 30 %; the 1e-3 bound of the north star is held kernel by kernel against the quantisation-matched oracle
    # (tests/test_fullsize_gpu.py::test_bf16_mode_every_kernel_vs_quantisation_matched_oracle)
    assert e_cap < 6.5e-3 and e_tag < 6.5e-3 and e_lg < 1.1e-2
    assert overlap >= 0.95


def test_overlapped_host_loop_matches_direct_forward():
    """vitcap_b200.stream.OverlappedCaptioner (double-buffered H2D, pinned D2H) returns exactly what model(data) returns,
    batch by batch and in order."""
    from vitcap_b200.stream import OverlappedCaptioner
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=1.0)
    extra = synth.default_test_extra_input(cfg)
    m = build(cfg, sd, extra, "fp32", max_batch=4)
    batches = []
    for i in range(5):
        d = synth.make_text_inputs(cfg, 4)
        d["image"] = synth.make_images(cfg, 4, seed=20 + i)
        batches.append({k: v.pin_memory() for k, v in d.items()})
    direct = [m(to_dev(b)) for b in batches]
    oc = OverlappedCaptioner(m, DEV, depth=2, with_tags=True)
    got = list(oc.run(iter(batches)))
    assert len(got) == 5
    for (ids, lp, tidx, tprob), (rids, rlp) in zip(got, direct):
        assert torch.equal(ids, rids.cpu())
        assert torch.equal(lp, rlp.cpu())
        assert tidx.shape == (4, cfg.topk) and tprob.shape == (4, cfg.topk)
    assert len({tuple(g[0].flatten().tolist()) for g in got}) > 1          # different batches, different captions


def _host_transform_tail(u8_bgr_hwc):
    """BGR2RGB -> ToTensor -> Normalize(0.5, 0.5) exactly as torchvision computes them (uni_pipeline.py:1233-1256)."""
    rgb = u8_bgr_hwc[..., [2, 1, 0]]
    t = rgb.permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
    mean = torch.tensor([0.5, 0.5, 0.5]).view(1, 3, 1, 1)
    std = torch.tensor([0.5, 0.5, 0.5]).view(1, 3, 1, 1)
    return t.sub_(mean).div_(std)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_patchify_u8_bit_identical_to_host_transform(dtype):
    from vitcap_b200 import ops
    B, S, p = 3, 64, 16
    u8 = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3))
    ref_img = _host_transform_tail(u8).to(DEV)
    a = torch.empty(B * (S // p) ** 2, 3 * p * p, device=DEV, dtype=dtype)
    b = torch.empty_like(a)
    ops.patchify(ref_img, a, p)
    ops.patchify_u8(u8.to(DEV), b, p, bgr=True)
    assert torch.equal(a, b)
    ops.patchify_u8(u8[..., [2, 1, 0]].contiguous().to(DEV), b, p, bgr=False)
    assert torch.equal(a, b)


def test_uint8_images_caption_like_host_transformed_floats():
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=1.0)
    extra = synth.default_test_extra_input(cfg)
    m = build(cfg, sd, extra, "fp32", max_batch=4)
    B = 6
    u8 = torch.randint(0, 256, (B, cfg.img_size, cfg.img_size, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(5))
    data = synth.make_text_inputs(cfg, B)
    d_f = dict(data, image=_host_transform_tail(u8))
    d_u = dict(data, image=u8)
    ids_f, lp_f = m(to_dev(d_f))
    ids_u, lp_u = m(to_dev(d_u))
    assert torch.equal(ids_f, ids_u) and torch.equal(lp_f, lp_u)


def test_alternating_search_modes_keep_their_workspaces_and_graphs():
    """SCST runs a greedy pass and a sampling pass per batch (legacy pipeline lines 447-462): the engine keeps both decode
    workspaces (and their captured graphs) alive instead of re-allocating and re-capturing on every switch."""
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=1.0)
    B = 3
    data = synth.make_text_inputs(cfg, B)
    data["image"] = synth.make_images(cfg, B, seed=9)
    m = build(cfg, sd, synth.default_test_extra_input(cfg), "fp32", max_batch=4)
    greedy = synth.default_test_extra_input(cfg)
    sample = synth.default_test_extra_input(cfg, do_sample=True, num_return_sequences=5)
    m.sample_seed = 11
    outs = {"g": [], "s": []}
    for rnd in range(3):
        m.test_extra_input = greedy
        outs["g"].append(m(to_dev(data)))
        m.test_extra_input = sample
        if rnd == 2:
            m._sample_calls = 0                  # round 2 repeats the seed of round 0; round 1 draws with the next seed
        outs["s"].append(m(to_dev(data)))
    eng = m.engine
    assert len(eng._dec_ws) == 2
    # one captured loop per mode: the sampling seed lives in device memory, not in the graph
    assert all(len(ws["graphs"]) == 1 for ws in eng._dec_ws.values())
    # round 0 = eager + capture for each mode, rounds 1-2 = replays only
    assert eng.stats.get("graph_replays", 0) == 4
    for ids, lp in outs["g"][1:]:
        assert torch.equal(ids, outs["g"][0][0]) and torch.equal(lp, outs["g"][0][1])
    assert torch.equal(outs["s"][2][0], outs["s"][0][0]) and torch.equal(outs["s"][2][1], outs["s"][0][1])
    assert not torch.equal(outs["s"][1][0], outs["s"][0][0])
    assert tuple(outs["s"][0][0].shape) == (B * 5, 1, cfg.max_seq_a)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_programmatic_dependent_launch_does_not_change_results(mode):
    """Every kernel of the path is launched with programmatic stream serialization and waits (griddepcontrol.wait) for its
    predecessor before touching global memory: greedy, beam and sampling results -- eager first call and CUDA-graph replays --
    must be bit-identical with the attribute off (ordinary stream order)."""
    from vitcap_b200 import ops
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=3, eos_bias=1.0)
    B = 5
    data = synth.make_text_inputs(cfg, B)
    data["image"] = synth.make_images(cfg, B, seed=4)
    runs = {}
    default = ops.get_pdl()
    assert default == 1
    try:
        for pdl in (2, 1, 0):
            ops.set_pdl(pdl)
            outs = []
            for kw in ({}, dict(num_beams=3, num_keep_best=2), dict(do_sample=True, num_return_sequences=4)):
                m = build(cfg, sd, synth.default_test_extra_input(cfg, **kw), mode, max_batch=8)
                m.sample_seed = 5
                for rep in range(3):                   # eager + capture, then two replays
                    m._sample_calls = 0
                    ids, lp = m(to_dev(data))
                    outs.append((ids.cpu(), lp.cpu()))
                lg, idx, pr, n = m.forward_tags(data["image"].to(DEV))
                outs.append((idx.cpu(), lg.cpu()))
            runs[pdl] = outs
    finally:
        ops.set_pdl(default)
    assert len(runs[0]) == len(runs[1]) == len(runs[2]) == 12
    for k in (1, 2):
        for (a0, a1), (b0, b1) in zip(runs[k], runs[0]):
            assert torch.equal(a0, b0) and torch.equal(a1, b1)


@pytest.mark.parametrize("kw", [{}, dict(num_beams=3, num_keep_best=2)])
def test_whole_forward_graph_matches_eager_path(kw):
    """graph_forward=True (small-batch serving): the whole forward of a batch shape is captured once and replayed with new
    images copied into its input buffer; ids, log-probs and the concept top-k must equal the eager path's bit for bit, for
    every new image batch, and a different batch size gets its own graph."""
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=3, eos_bias=1.0)
    extra = synth.default_test_extra_input(cfg, **kw)
    ref = build(cfg, sd, extra, "bf16", max_batch=8)
    fast = build(cfg, sd, extra, "bf16", max_batch=8, graph_forward=True)
    for rnd, (B, seed) in enumerate([(2, 1), (2, 2), (2, 3), (5, 4), (2, 5), (5, 6)]):
        data = synth.make_text_inputs(cfg, B)
        data["image"] = synth.make_images(cfg, B, seed=seed)
        d = to_dev(data)
        i0, l0 = ref(d)
        t0 = ref.last_tags
        i1, l1 = fast(d)
        t1 = fast.last_tags
        assert torch.equal(i0, i1) and torch.equal(l0, l1), rnd
        assert torch.equal(t0[0], t1[0]) and torch.equal(t0[1], t1[1]), rnd
    eng = fast.engine
    assert len(eng.forward_graphs) == 2                       # one per batch shape
    assert eng.stats.get("forward_graph_replays", 0) == 4     # first call of each shape is eager (+ capture)
    # sampling and label-region calls fall back to the eager path
    fast.test_extra_input = synth.default_test_extra_input(cfg, do_sample=True)
    fast(d)
    assert eng.stats.get("forward_graph_replays", 0) == 4


def test_alternating_keep_best_on_one_workspace():
    """num_keep_best 1 -> 3 -> 1 on the same (B, beams, max_len) workspace: every captured graph keeps the beam state it was
    captured with (a rebuilt state would leave the old graph replaying into freed memory)."""
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=7, eos_bias=1.0)
    B = 4
    data = synth.make_text_inputs(cfg, B)
    data["image"] = synth.make_images(cfg, B, seed=3)
    d = to_dev(data)
    outs = {}
    m = build(cfg, sd, synth.default_test_extra_input(cfg, num_beams=3, num_keep_best=1), "fp32", max_batch=8)
    for rnd in range(3):
        for keep in (1, 3):
            m.test_extra_input = synth.default_test_extra_input(cfg, num_beams=3, num_keep_best=keep)
            ids, lp = m(d)
            assert ids.shape[1] == keep
            if keep in outs:
                assert torch.equal(ids, outs[keep][0]) and torch.equal(lp, outs[keep][1]), (rnd, keep)
            outs[keep] = (ids, lp)
    assert len(m.engine._dec_ws) == 1
