"""Synthetic inputs and random-init weights in the reference's ``state_dict`` layout.

There is no network, so neither the released checkpoint nor COCO is available: parity and
throughput are measured on random-init weights and random 384x384 images (BASELINE.json).
The key names/shapes are exactly the 288 tensors of the reference eval module
(``ImageCaptioning(ViTCAP, image_encoder=InputAsDict(timm ViT))``),
so the same dict loads into the reference (``load_state_dict``) and into
``vitcap_b200.FastImageCaptioning``.

numpy's PCG64 stream is used (not torch's RNG) so the tensors are identical on any box.
"""
import numpy as np
import torch

from .config import VitCapConfig

CLS_ID, SEP_ID, PAD_ID, MASK_ID = 101, 102, 0, 103


def _block_keys(prefix, cfg):
    h, f = cfg.hidden, cfg.inter
    return [
        (prefix + "norm1.weight", (h,), "ln_w"), (prefix + "norm1.bias", (h,), "ln_b"),
        (prefix + "attn.qkv.weight", (3 * h, h), "w"), (prefix + "attn.qkv.bias", (3 * h,), "b"),
        (prefix + "attn.proj.weight", (h, h), "w"), (prefix + "attn.proj.bias", (h,), "b"),
        (prefix + "norm2.weight", (h,), "ln_w"), (prefix + "norm2.bias", (h,), "ln_b"),
        (prefix + "mlp.fc1.weight", (f, h), "w"), (prefix + "mlp.fc1.bias", (f,), "b"),
        (prefix + "mlp.fc2.weight", (h, f), "w"), (prefix + "mlp.fc2.bias", (h,), "b"),
    ]


def _emb_keys(prefix, cfg, kind):
    h = cfg.hidden
    return [
        (prefix + "word_embeddings.weight", (cfg.vocab, h), kind),
        (prefix + "position_embeddings.weight", (cfg.max_pos, h), kind),
        (prefix + "token_type_embeddings.weight", (cfg.type_vocab, h), kind),
        (prefix + "LayerNorm.weight", (h,), "ln_w"), (prefix + "LayerNorm.bias", (h,), "ln_b"),
    ]


def _head_keys(prefix, cfg, kind, with_decoder=True):
    h = cfg.hidden
    ks = [
        (prefix + "predictions.bias", (cfg.vocab,), "vb"),
        (prefix + "predictions.transform.dense.weight", (h, h), kind),
        (prefix + "predictions.transform.dense.bias", (h,), "b"),
        (prefix + "predictions.transform.LayerNorm.weight", (h,), "ln_w"),
        (prefix + "predictions.transform.LayerNorm.bias", (h,), "ln_b"),
    ]
    if with_decoder:
        ks.append((prefix + "predictions.decoder.weight", (cfg.vocab, h), kind))
    return ks


def state_dict_spec(cfg: VitCapConfig):
    """[(key, shape, kind)] in the reference's order. ``kind``: w = N(0,.02) matrix,
    u = torch-default Linear/Embedding init, b = bias, ln_w/ln_b = LayerNorm affine, vb = vocab bias."""
    h, f = cfg.hidden, cfg.inter
    spec = []
    spec += _emb_keys("module.bert.embeddings.", cfg, "w")
    spec += _emb_keys("module.bert.extra_embeddings.", cfg, "e")
    for i in range(cfg.enc_blocks):
        spec += _block_keys("module.bert.encoder.blocks.%d." % i, cfg)
    for i in range(cfg.split_blocks):
        spec += _block_keys("module.bert.encoder.tag_blocks.%d." % i, cfg)
    spec += [("module.bert.caption_pooler.dense.weight", (h, h), "w"),
             ("module.bert.caption_pooler.dense.bias", (h,), "b"),
             ("module.bert.pooler.dense.weight", (h, h), "u"),
             ("module.bert.pooler.dense.bias", (h,), "ub")]
    spec += _head_keys("module.bert.tag_logit.", cfg, "u")
    for i in range(cfg.dec_layers):
        p = "module.bert.decoder.layer.%d." % i
        for n in ("query", "key", "value"):
            spec += [(p + "attention.self.%s.weight" % n, (h, h), "w"), (p + "attention.self.%s.bias" % n, (h,), "b")]
        spec += [(p + "attention.output.dense.weight", (h, h), "w"), (p + "attention.output.dense.bias", (h,), "b"),
                 (p + "attention.output.LayerNorm.weight", (h,), "ln_w"), (p + "attention.output.LayerNorm.bias", (h,), "ln_b"),
                 (p + "intermediate.dense.weight", (f, h), "w"), (p + "intermediate.dense.bias", (f,), "b"),
                 (p + "output.dense.weight", (h, f), "w"), (p + "output.dense.bias", (h,), "b"),
                 (p + "output.LayerNorm.weight", (h,), "ln_w"), (p + "output.LayerNorm.bias", (h,), "ln_b")]
    spec += _head_keys("module.cls.", cfg, "w", with_decoder=True)   # decoder.weight is tied, filled below
    spec += [("image_encoder.module.cls_token", (1, 1, h), "w"),
             ("image_encoder.module.pos_embed", (1, cfg.n_tokens, h), "w"),
             ("image_encoder.module.patch_embed.proj.weight", (h, 3, cfg.patch, cfg.patch), "conv"),
             ("image_encoder.module.patch_embed.proj.bias", (h,), "ub"),
             ("image_encoder.module.head.weight", (1000, h), "w"),
             ("image_encoder.module.head.bias", (1000,), "b")]
    return spec


def make_state_dict(cfg: VitCapConfig, seed=0, style="stress", vocab_gain=1.0, eos_bias=0.0, tag_bias=0.0):
    """Random weights in the reference layout.

    style="reference": the reference's own init families (zero biases, unit LayerNorm;
        modeling_bert.py:578-589, vision_transformer.py:387-398, torch defaults for
        pooler/tag_logit/extra_embeddings).
    style="stress": same matrices but non-zero biases and non-unit LayerNorm affine, so a
        kernel that drops a bias or a gamma/beta cannot pass parity.
    vocab_gain multiplies the tied word-embedding matrix (peaks the caption distribution,
        SURVEY.md section 7 "hard parts"); eos_bias is added to cls.predictions.bias[SEP] so
        early-exit / PAD-fill paths are exercised.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for key, shape, kind in state_dict_spec(cfg):
        n = int(np.prod(shape))
        if key == "module.cls.predictions.decoder.weight":
            continue
        if kind in ("w",):
            a = rng.standard_normal(n, dtype=np.float32) * 0.02
            np.clip(a, -2.0, 2.0, out=a)
        elif kind == "e":            # nn.Embedding default N(0,1)
            a = rng.standard_normal(n, dtype=np.float32)
        elif kind in ("u", "conv"):  # nn.Linear / Conv2d default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / np.sqrt(fan_in)
            a = (rng.random(n, dtype=np.float32) * 2 - 1) * bound
        elif kind == "ub":
            bound = 1.0 / np.sqrt(cfg.hidden if "pooler" in key else cfg.patch_dim)
            a = (rng.random(n, dtype=np.float32) * 2 - 1) * bound
        elif kind == "b":
            a = (rng.standard_normal(n, dtype=np.float32) * 0.02) if style == "stress" else np.zeros(n, np.float32)
        elif kind == "vb":
            a = (rng.standard_normal(n, dtype=np.float32) * 0.02) if style == "stress" else np.zeros(n, np.float32)
        elif kind == "ln_w":
            a = (1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32)) if style == "stress" else np.ones(n, np.float32)
        elif kind == "ln_b":
            a = (0.02 * rng.standard_normal(n, dtype=np.float32)) if style == "stress" else np.zeros(n, np.float32)
        else:
            raise ValueError(kind)
        out[key] = torch.from_numpy(a.astype(np.float32).reshape(shape))
    we = out["module.bert.embeddings.word_embeddings.weight"]
    if vocab_gain != 1.0:
        we.mul_(vocab_gain)
    # NB: nn.Embedding(padding_idx=0) zeroes row 0 at construction (modeling_bert.py:213) but
    # embeddings.apply(init_weights) (modeling_bert.py:1364) re-draws it, so row 0 stays random.
    out["module.cls.predictions.decoder.weight"] = we   # tied (modeling_bert.py:728-730)
    if eos_bias:
        out["module.cls.predictions.bias"][SEP_ID] += eos_bias
    if tag_bias:     # shifts every concept logit: moves topk_len (count of top-k probabilities >= 0.2, modeling_bert.py:1432)
        out["module.bert.tag_logit.predictions.bias"] += tag_bias
    # keep the reference's key order
    return {k: out[k] for k, _, _ in state_dict_spec(cfg)}


def make_images(cfg: VitCapConfig, batch, seed=1234):
    """``torch.randn`` stand-in for (x-0.5)/0.5-normalised images."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.standard_normal(batch * 3 * cfg.img_size * cfg.img_size, dtype=np.float32)
    return torch.from_numpy(a.reshape(batch, 3, cfg.img_size, cfg.img_size))


def make_text_inputs(cfg: VitCapConfig, batch, n_label=None, label_token=2000):
    """Test-time text tensors of the reference data layer: what
    ``CaptionTensorizer.tensorize_ab('', text_b=..., real_text_a_in_test=False)`` returns
    (dataset.py:206-417). ``n_label`` None / 0 (text_b == ''): ids [CLS, MASK x18, SEP, PAD x50], a 70x70 mask whose
    only non-zeros are the 20x20 caption triangle, zero segment ids, all-ones masked_pos.
    ``n_label`` = per-sample list of visible label slots (text_b with n-1 word pieces + [SEP]): label ids in slots
    20..20+n (their value never reaches the output: modeling_bert.py:1470 overwrites all 50 slot embeddings with the
    predicted tags), segment id 1 there, full attention L-L and C-L (dataset.py:405-408)."""
    a, s = cfg.max_seq_a, cfg.max_seq
    if n_label is None:
        n_label = [0] * batch
    assert len(n_label) == batch and all(0 <= int(n) <= s - a and int(n) != 1 for n in n_label)
    ids = torch.zeros(batch, s, dtype=torch.long)
    ids[:, 0] = CLS_ID
    ids[:, 1:a - 1] = MASK_ID
    ids[:, a - 1] = SEP_ID
    mask = torch.zeros(batch, s, s, dtype=torch.long)
    mask[:, :a, :a] = torch.tril(torch.ones(a, a, dtype=torch.long))
    typ = torch.zeros(batch, s, dtype=torch.long)
    for b, n in enumerate(n_label):
        n = int(n)
        if n:
            ids[b, a:a + n - 1] = label_token
            ids[b, a + n - 1] = SEP_ID
            typ[b, a:a + n] = 1
            mask[b, a:a + n, a:a + n] = 1
            mask[b, :a, a:a + n] = 1
    return {
        "input_ids": ids,
        "attention_mask": mask,
        "token_type_ids": typ,
        "masked_pos": torch.ones(batch, s, dtype=torch.int32),
    }


def default_test_extra_input(cfg: VitCapConfig, **overrides):
    """``test_extra_input`` of get_raw_model(is_train=False) (pipeline file lines 588-608)."""
    d = {
        "is_decode": True, "do_sample": False, "bos_token_id": CLS_ID, "pad_token_id": PAD_ID,
        "eos_token_ids": [SEP_ID], "mask_token_id": MASK_ID, "add_od_labels": True,
        "od_labels_start_posid": cfg.max_seq_a, "max_length": cfg.max_seq_a, "num_beams": 1,
        "temperature": 1, "top_k": 0, "top_p": 1, "repetition_penalty": 1, "length_penalty": 1,
        "num_return_sequences": 1, "num_keep_best": 1,
    }
    d.update(overrides)
    return d
