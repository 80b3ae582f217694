"""CPU-only checks of the host side: state_dict layout, C-ABI surface, argument validation, sharding."""
import ctypes
import os
import re

import pytest
import torch

from vitcap_b200 import config as vcfg
from vitcap_b200 import ops, synth
from vitcap_b200.model import FastImageCaptioning

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_layout_matches_reference_names():
    """Same 288 keys / shapes as ImageCaptioning(ViTCAP) of the reference, same order."""
    cfg = vcfg.variant("16_384")
    m = FastImageCaptioning(cfg)
    sd = m.state_dict()
    spec = synth.state_dict_spec(cfg)
    assert len(sd) == len(spec) == 288
    assert list(sd.keys()) == [k for k, _, _ in spec]
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    # tied vocabulary matrix (modeling_bert.py:728-730)
    assert sd["module.cls.predictions.decoder.weight"].data_ptr() == sd["module.bert.embeddings.word_embeddings.weight"].data_ptr()


def test_state_dict_roundtrip_tiny():
    cfg = vcfg.tiny()
    sd = synth.make_state_dict(cfg, seed=3)
    m = FastImageCaptioning(cfg)
    r = m.load_state_dict(sd, strict=True)
    assert not r.missing_keys and not r.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    # checkpoint written by the reference Checkpointer: {'model': state_dict} with a leading 'module.' (DDP) is the
    # caller's concern; names below the wrapper are identical
    assert all(k.startswith(("module.", "image_encoder.module.")) for k in sd)


def test_header_symbols_are_exported_and_bound():
    """Every function declared in include/vitcap_b200.h is exported by the shared library and has a ctypes prototype."""
    hdr = open(os.path.join(ROOT, "include", "vitcap_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(vc_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    lib = ops.load_library()
    for n in names:
        assert hasattr(lib, n), n
    simple = {"vc_last_error", "vc_abi_version", "vc_launch_count", "vc_reset_launch_count", "vc_set_pdl", "vc_get_pdl",
              "vc_check_device"}
    assert names - simple == set(ops.SIGNATURES), (names - simple) ^ set(ops.SIGNATURES)
    assert lib.vc_abi_version() == 9
    assert isinstance(lib.vc_last_error(), bytes)


def test_arch_guard_without_a_gpu():
    """On a box without a usable sm_100 device every compute entry point answers VC_ERR_UNSUPPORTED (-2) with a message instead
    of failing inside a launch (no GPU here: the guard itself is what is exercised; tests/test_kernels_gpu.py checks the
    pass-through on a B200 and the refusal under VITCAP_FAKE_CC)."""
    lib = ops.load_library()
    if not torch.cuda.is_available():
        assert lib.vc_check_device() == -2
        assert b"CUDA device" in lib.vc_last_error() or b"sm_100a" in lib.vc_last_error()
        with pytest.raises(RuntimeError, match="-2"):
            ops.check_device()
        # a compute entry point (null pointers are never dereferenced: the guard answers first)
        assert lib.vc_layernorm(1, None, 0, None, None, 1e-6, None, 0, None, 0, 1, 768, None) == -2


def test_library_has_no_torch_or_libcuda_link_dependency():
    import subprocess
    out = subprocess.run(["ldd", ops.LIB_PATH], capture_output=True, text=True).stdout
    out = re.sub(r"\(0x[0-9a-f]+\)", "", out)            # load addresses are random hex: they may spell "c10"
    assert "torch" not in out and "libcuda.so" not in out and "c10" not in out


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        ops.load_library(str(tmp_path / "nope.so"))


def test_cpu_module_refuses_to_run():
    cfg = vcfg.tiny()
    m = FastImageCaptioning(cfg)
    data = synth.make_text_inputs(cfg, 1)
    data["image"] = synth.make_images(cfg, 1)
    with pytest.raises(RuntimeError, match="CUDA device only"):
        m(data)


def test_unsupported_flags_raise():
    cfg = vcfg.tiny()
    m = FastImageCaptioning(cfg)
    feats = torch.zeros(1, cfg.n_tokens, cfg.hidden)
    base = dict(synth.default_test_extra_input(cfg), input_ids=synth.make_text_inputs(cfg, 1)["input_ids"])
    for bad in (dict(use_cbs=True), dict(repetition_penalty=1.2), dict(num_beams=2, do_sample=True),
                dict(num_keep_best=2), dict(head_mask=torch.ones(1))):
        kw = dict(base)
        kw.update(bad)
        with pytest.raises((NotImplementedError, AssertionError)):
            m.module(feats, **kw)
    with pytest.raises(NotImplementedError):
        m.module(feats, is_decode=False)
    # a mask outside the seq2seq family of the data layer is refused, not silently mis-computed
    ti = synth.make_text_inputs(cfg, 1)
    ti["attention_mask"][:, :20, 20:] = 1          # C-L visible without the L-L block
    with pytest.raises(NotImplementedError, match="outside the seq2seq family"):
        m._label_counts(ti["attention_mask"], ti["input_ids"], 20)
    ti = synth.make_text_inputs(cfg, 1)
    ti["attention_mask"][:, 5, 9] = 1              # a caption row looking ahead
    with pytest.raises(NotImplementedError, match="outside the seq2seq family"):
        m._label_counts(ti["attention_mask"], ti["input_ids"], 20)


def test_label_counts_of_the_data_layer_masks():
    """The eval mask (text_b == '') has no visible label slot; text_b with n-1 word pieces + [SEP] shows n of them
    (dataset.py:395-408). Also through the full (70+N)^2 mask of construct_attn_mask (pipeline file lines 57-85)."""
    cfg = vcfg.tiny()
    m = FastImageCaptioning(cfg)
    ti = synth.make_text_inputs(cfg, 2)
    assert m._label_counts(ti["attention_mask"], ti["input_ids"], 20) is None
    ti = synth.make_text_inputs(cfg, 4, n_label=[0, 4, 50, 23])
    assert m._label_counts(ti["attention_mask"], ti["input_ids"], 20).tolist() == [0, 4, 50, 23]
    from oracle import port
    full = port.construct_full_mask(ti["attention_mask"], cfg.n_tokens)
    assert m._label_counts(full, ti["input_ids"], 20).tolist() == [0, 4, 50, 23]
    assert port.label_counts(ti["attention_mask"], 20).tolist() == [0, 4, 50, 23]


def test_split_weight_bf16x3_layout_and_precision():
    """Host-side weight layout of the three-product decode GEMMs: [w_hi | w_hi | w_lo]; hi + lo restores the fp32 weight to
    ~2^-16 relative (the kernel-side operand split is checked on the GPU, tests/test_kernels_gpu.py)."""
    import torch
    from vitcap_b200 import ops
    w = torch.randn(37, 64, generator=torch.Generator().manual_seed(3)) * 0.02
    w3 = ops.split_weight_bf16x3(w)
    assert w3.shape == (37, 192) and w3.dtype == torch.bfloat16 and w3.is_contiguous()
    hi, hi2, lo = w3[:, :64], w3[:, 64:128], w3[:, 128:]
    assert torch.equal(hi, w.to(torch.bfloat16)) and torch.equal(hi, hi2)
    rel = ((hi.float() + lo.float()) - w).abs().max() / w.abs().max()
    assert float(rel) < 2.0 ** -15
    assert float((hi.float() - w).abs().max() / w.abs().max()) > 2.0 ** -11          # what plain bf16 leaves


def test_three_product_split_bf16_arithmetic_spec():
    """Executable statement of what the decode-step split-bf16 GEMMs compute, in plain torch on the CPU:
    with a = hi + lo and w = w_hi + w_lo (each part a bf16 number), hi w_hi + lo w_hi + hi w_lo accumulated in fp32 misses only
    the lo w_lo term (2^-18 relative) -- two orders closer to the fp32 product than the plain bf16 GEMM. The GPU kernels are
    pinned to the same figures in tests/test_kernels_gpu.py."""
    import torch
    from vitcap_b200 import ops
    g = torch.Generator().manual_seed(11)
    a, w = torch.randn(64, 768, generator=g), torch.randn(96, 768, generator=g) * 0.05
    hi = a.to(torch.bfloat16)
    lo = (a - hi.float()).to(torch.bfloat16)
    a3 = torch.cat([hi, lo, hi], 1).float()
    w3 = ops.split_weight_bf16x3(w).float()
    ref = a.double() @ w.double().t()
    scale = float(ref.abs().max())
    err3 = float((a3 @ w3.t() - ref).abs().max()) / scale
    err1 = float((hi.float() @ w.to(torch.bfloat16).float().t() - ref).abs().max()) / scale
    assert err3 < 1e-5 and err1 > 5e-4 and err3 < err1 / 100


def test_decode_precision_option(monkeypatch):
    """decode_precision: 'fp16' by default in the fast mode, overridable by argument or VITCAP_DECODE_PRECISION; the exact mode
    ignores it; anything else is refused at construction."""
    from vitcap_b200 import config as vcfg
    from vitcap_b200.model import FastImageCaptioning
    cfg = vcfg.tiny()
    monkeypatch.delenv("VITCAP_DECODE_PRECISION", raising=False)
    assert FastImageCaptioning(cfg).decode_precision == "fp16"
    assert FastImageCaptioning(cfg, decode_precision="bf16").decode_precision == "bf16"
    assert FastImageCaptioning(cfg, decode_precision="bf16x3").decode_precision == "bf16x3"
    assert FastImageCaptioning(cfg, mode="fp32", decode_precision="bf16x3").decode_precision == "fp32"
    monkeypatch.setenv("VITCAP_DECODE_PRECISION", "bf16")
    assert FastImageCaptioning(cfg).decode_precision == "bf16"
    assert FastImageCaptioning(cfg, decode_precision="bf16x3").decode_precision == "bf16x3"
    with pytest.raises(AssertionError):
        FastImageCaptioning(cfg, decode_precision="fp8")


def test_generate_limits_are_checked_before_any_kernel():
    """num_beams / num_keep_best / max_length beyond what the search kernels hold (search.cu VC_MAX_BEAMS = 8, VC_MAX_LEN = 64,
    keep <= 64) are refused by generate() itself -- on a CPU module, i.e. before the encoder, the prefill or a graph capture."""
    cfg = vcfg.tiny(max_seq=128)
    m = FastImageCaptioning(cfg)
    feats = torch.zeros(1, cfg.n_tokens, cfg.hidden)
    base = dict(synth.default_test_extra_input(cfg), input_ids=synth.make_text_inputs(cfg, 1)["input_ids"])
    for bad, msg in ((dict(num_beams=9), "num_beams <= 8"), (dict(num_beams=4, num_keep_best=65), "num_keep_best <= 64"),
                     (dict(max_length=65), "max_length must be <= 64"), (dict(max_length=1), "max_length must be in")):
        kw = dict(base)
        kw.update(bad)
        with pytest.raises(ValueError, match=msg):
            m.module(feats, **kw)


def test_visible_labels_with_a_non_cls_tag_embedding_are_refused():
    """config.tagemb != 'cls' takes the reference through encode_tag_to_embedding(cls_emb=None) / extra_embeddings
    (modeling_bert.py:1466, 1485), which is not built: a visible label region then raises instead of silently using the 'cls'
    recipes. Without visible labels the setting is irrelevant (the label slots are dead) and accepted."""
    cfg = vcfg.tiny(tagemb="bert")
    m = FastImageCaptioning(cfg)
    ti = synth.make_text_inputs(cfg, 2)
    assert m._label_counts(ti["attention_mask"], ti["input_ids"], 20) is None
    ti = synth.make_text_inputs(cfg, 2, n_label=[3, 0])
    with pytest.raises(NotImplementedError, match="tagemb"):
        m._label_counts(ti["attention_mask"], ti["input_ids"], 20)


def test_packed_weights_are_invalidated_by_moves_and_in_place_updates():
    """The kernel-side weight copies must follow the parameters: .to() / .float() (new storage) and in-place updates
    (optimizer step, manual edit) both force a re-pack at the next use."""
    cfg = vcfg.tiny()
    m = FastImageCaptioning(cfg)
    sentinel = object()
    m._engine, m._packed_version = sentinel, m._param_version()
    m.double()
    assert m._engine is None
    m._engine, m._packed_version = sentinel, m._param_version()
    with torch.no_grad():
        m.module.bert.pooler.dense.weight.mul_(0.5)
    assert m._packed_version != m._param_version()
    with pytest.raises(RuntimeError, match="CUDA device only"):      # the engine property re-packs, which needs the GPU
        m.engine


def test_half_range_bounds_of_the_fp16_decode_operands():
    """decode_precision='fp16' stores LayerNorm / GELU outputs and four weight matrices as IEEE halves. engine.half_range_bounds
    derives worst-case magnitudes from the weights alone; below 65504 they prove that no conversion can saturate. The synthetic
    'stress' and reference-style weights are far inside; the bounds really are bounds (checked against the oracle's activations);
    an absurd LayerNorm gain is reported."""
    import torch
    from oracle import port
    from vitcap_b200 import config as vcfg, synth
    from vitcap_b200.engine import HALF_MAX, half_range_bounds
    cfg = vcfg.tiny()
    for style in ("stress", "reference"):
        sd = synth.make_state_dict(cfg, seed=3, style=style)
        b = half_range_bounds(cfg, sd)
        assert 0 < b["weights"] < 10 and b["layernorm"] < HALF_MAX / 10 and b["gelu"] < HALF_MAX / 2, (style, b)
    # the bounds hold for what the model actually computes: LayerNorm outputs and GELU outputs of a decode step
    sd = synth.make_state_dict(cfg, seed=3)
    b = half_range_bounds(cfg, sd)
    seen = {"ln": 0.0, "gelu": 0.0}

    class Spy(port.QuantPortModel):
        def lin_x3(self, a, wkey, bkey):
            kind = "gelu" if wkey.endswith("output.dense.weight") else "ln"
            seen[kind] = max(seen[kind], float(a.abs().max()))
            return super().lin_x3(a, wkey, bkey)

    data = synth.make_text_inputs(cfg, 2)
    data["image"] = synth.make_images(cfg, 2, seed=5)
    with torch.no_grad():
        port.caption(Spy(cfg, sd, decode_f16=True), data, synth.default_test_extra_input(cfg), algorithm="cached")
    assert 0 < seen["ln"] <= b["layernorm"] and 0 < seen["gelu"] <= b["gelu"], (seen, b)
    sd["module.bert.decoder.layer.0.output.LayerNorm.weight"] = sd["module.bert.decoder.layer.0.output.LayerNorm.weight"] * 1e5
    assert half_range_bounds(cfg, sd)["layernorm"] > HALF_MAX


def test_early_exit_step_groups():
    """engine._step_groups: step 1 unconditional, then `every` steps per condition, covering 1 .. max_len - 1 exactly once."""
    from vitcap_b200.engine import CaptionEngine
    eng = CaptionEngine.__new__(CaptionEngine)
    for n, every in ((1, 1), (16, 1), (17, 2), (128, 2), (129, 4), (512, 4)):
        assert eng._exit_every(n) == every
        for max_len in (2, 3, 7, 20):
            groups = eng._step_groups(max_len, n, True)
            steps = [s for f, l, c in groups for s in range(f, l)]
            assert steps == list(range(1, max_len)) and groups[0] == (1, 2, False)
            assert all(c and l - f <= every for f, l, c in groups[1:])
    assert eng._step_groups(20, 512, False) == [(1, 20, False)]


def test_half_storage_library_exports_the_same_abi():
    """libvitcap_b200_f16.so (the same sources built with -DVC_STORE_F16, selected by VITCAP_STORE=fp16) exports every entry
    point of the header under the same ABI version; the process-wide switch picks it and torch.float16."""
    import ctypes
    import subprocess
    import sys
    from vitcap_b200 import build as b
    assert os.path.exists(b.LIB_F16), "python -m vitcap_b200.build builds both libraries"
    lib = ctypes.CDLL(b.LIB_F16)
    lib.vc_abi_version.restype = ctypes.c_int
    assert lib.vc_abi_version() == 9
    for n in ops.SIGNATURES:
        assert hasattr(lib, n), n
    assert ops.STORE == torch.bfloat16 and not ops.HALF_STORE and ops.LIB_PATH.endswith("libvitcap_b200.so")     # this process
    code = "from vitcap_b200 import ops; import torch; assert ops.HALF_STORE and ops.STORE == torch.float16; " \
           "assert ops.LIB_PATH.endswith('_f16.so'); ops.load_library(); print('ok')"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, VITCAP_STORE="fp16", PYTHONPATH=ROOT),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-1500:]
