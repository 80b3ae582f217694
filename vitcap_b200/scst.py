"""SCST training-side use of the caption path.

The reference's self-critical recipe (legacy pipeline tagger_caption_uni_pipeline_expanding.py:447-468) runs, per training
batch, a greedy pass under ``no_grad`` (the baseline) and a sampling pass of ``scst_num_return`` captions per image whose
sequence log-probs must carry gradients (``assert sample_logprobs.requires_grad``) for
``ScstRewardCriterion`` (utils_caption_evaluate.py:162-242: ``loss = -(sample_logprobs * (reward - baseline)).mean()``).
It obtains those gradients by back-propagating through all 19 full-model calls of ``generate``.

Here the two GENERATION passes run on the B200 kernels (``FastImageCaptioning``: greedy + K sampled captions, no gradients),
and the gradient comes from ONE teacher-forced, differentiable re-evaluation of the sampled tokens' log-probs over the
module's fp32 master parameters -- the same function of the parameters (a caption row only ever sees the context and the
tokens before it, so the per-step [token, MASK] rows of the 19 calls are the rows of one masked pass), evaluated once
instead of nineteen times. That pass is plain PyTorch autograd: LIBRARY code, deliberately -- a hand-written backward of the
whole network is a training stack, out of this repository's scope. It is not a fallback of the
inference path: nothing in ``FastImageCaptioning.forward`` routes through it, and it is pinned by gradient parity against the
reference's own 19-call formulation (tests/test_scst_cpu.py) and against the kernels' log-probs (tests/test_scst_gpu.py).

Differences from the reference, by construction: dropout is off (the reference samples in ``.train()`` mode; the kernels
implement the eval-mode function), and the text mask is the eval pipeline's (no visible od/tag label region).
"""
import math

import torch
import torch.nn.functional as F

NEG = -10000.0          # the reference's additive mask value (modeling_bert.py:1498-1501)


def reference_params(model):
    """{reference state_dict key: Parameter} of a FastImageCaptioning module, tied vocabulary projection included."""
    P = dict(model.named_parameters())
    P.setdefault("module.cls.predictions.decoder.weight", P["module.bert.embeddings.word_embeddings.weight"])
    return P


def _gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))          # activations.py:16-23


def _ln(x, P, pre, eps):
    return F.layer_norm(x, (x.shape[-1],), P[pre + "weight"], P[pre + "bias"], eps)


def _heads(t, H, d):
    return t.view(t.shape[0], t.shape[1], H, d).permute(0, 2, 1, 3)


def _vit_block(P, pre, x, cfg):
    """Block / Attention / Mlp, vision_transformer.py:152-250 (pre-LN)."""
    B, N, C = x.shape
    H, d = cfg.heads, cfg.head_dim
    h = _ln(x, P, pre + "norm1.", cfg.vit_ln_eps)
    qkv = F.linear(h, P[pre + "attn.qkv.weight"], P[pre + "attn.qkv.bias"]).reshape(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
    a = torch.softmax((qkv[0] @ qkv[1].transpose(-2, -1)) * (d ** -0.5), dim=-1)
    o = (a @ qkv[2]).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(o, P[pre + "attn.proj.weight"], P[pre + "attn.proj.bias"])
    h = _ln(x, P, pre + "norm2.", cfg.vit_ln_eps)
    h = _gelu(F.linear(h, P[pre + "mlp.fc1.weight"], P[pre + "mlp.fc1.bias"]))
    return x + F.linear(h, P[pre + "mlp.fc2.weight"], P[pre + "mlp.fc2.bias"])


def encode_context(P, cfg, image):
    """Patch embed + cls + pos (vision_transformer.py:267-275, 411-427), split encoder (modeling_bert.py:458-478) and the
    decoder context [tag-CLS | caption-branch tokens] (modeling_bert.py:1493). image fp32 (B, 3, S, S)."""
    x = F.conv2d(image, P["image_encoder.module.patch_embed.proj.weight"], P["image_encoder.module.patch_embed.proj.bias"],
                 stride=cfg.patch).flatten(2).transpose(1, 2)
    x = torch.cat([P["image_encoder.module.cls_token"].expand(x.shape[0], -1, -1), x], dim=1) + P["image_encoder.module.pos_embed"]
    tag = None
    for i in range(cfg.enc_blocks):
        if i == cfg.enc_blocks - cfg.split_blocks:
            tag = x
        x = _vit_block(P, "module.bert.encoder.blocks.%d." % i, x, cfg)
    for j in range(cfg.split_blocks):
        tag = _vit_block(P, "module.bert.encoder.tag_blocks.%d." % j, tag, cfg)
    return torch.cat([tag[:, 0:1], x], dim=1)


def _bert_output(P, p, cfg, hq, attn_ctx):
    """BertSelfOutput -> BertIntermediate -> BertOutput (modeling_bert.py:343-419, post-LN) for query rows hq."""
    a1 = F.linear(attn_ctx, P[p + "attention.output.dense.weight"], P[p + "attention.output.dense.bias"])
    a1 = _ln(a1 + hq, P, p + "attention.output.LayerNorm.", cfg.bert_ln_eps)
    m = _gelu(F.linear(a1, P[p + "intermediate.dense.weight"], P[p + "intermediate.dense.bias"]))
    m = F.linear(m, P[p + "output.dense.weight"], P[p + "output.dense.bias"])
    return _ln(m + a1, P, p + "output.LayerNorm.", cfg.bert_ln_eps)


def _qkv(P, p, cfg, h):
    H, d = cfg.heads, cfg.head_dim
    return tuple(_heads(F.linear(h, P[p + "attention.self.%s.weight" % n], P[p + "attention.self.%s.bias" % n]), H, d)
                 for n in ("query", "key", "value"))


def sequence_logprobs(model, image, ids, num_return_sequences=1, temperature=1.0, top_k=0, top_p=1.0, extra=None):
    """Mean log-prob of every generated caption WITH gradients (what ``_generate_no_beam_search`` returns as ``logprobs``,
    modeling_utils.py:849-876: per step log_softmax of the (temperature-scaled, filtered) logits gathered at the chosen token,
    summed over the steps at which the sequence was still unfinished and divided by their number).

    image fp32 (B, 3, S, S); ids int64 (B * num_return_sequences, L) or (.., 1, L): [BOS, t_1, .., EOS, PAD ..], the tokens the
    search CHOSE at every step (``FastImageCaptioning.last_raw_ids``: the returned ids of a caption that ran to max_length have
    a forced EOS in the last position, modeling_utils.py:869-871, while its log-prob is the chosen token's).
    Returns fp32 (B * num_return_sequences,) attached to the module's parameters."""
    cfg = model.cfg
    P = reference_params(model)
    extra = dict(model.test_extra_input if extra is None else extra)
    mask_id, eos_ids = int(extra["mask_token_id"]), [int(e) for e in extra["eos_token_ids"]]
    if ids.dim() == 3:
        ids = ids[:, 0]
    ids = ids.to(image.device)
    R, L = ids.shape
    K = int(num_return_sequences)
    B = image.shape[0]
    assert R == B * K, "ids must hold num_return_sequences rows per image"
    T = L - 1                                               # decode steps; step t (1..T) predicts position t
    H, d = cfg.heads, cfg.head_dim
    dev = image.device

    hc = encode_context(P, cfg, image)                      # (B, C, H) context rows, once per image
    C = hc.shape[1]
    # text rows of every sequence: T token rows (positions 0..T-1: BOS, t_1 .. t_{T-1}) then T MASK rows (positions 1..T)
    tok = torch.cat([ids[:, :T], torch.full((R, T), mask_id, dtype=torch.long, device=dev)], dim=1)
    pos = torch.cat([torch.arange(0, T, device=dev), torch.arange(1, T + 1, device=dev)]).unsqueeze(0).expand(R, -1)
    pre = "module.bert.embeddings."
    e = P[pre + "word_embeddings.weight"][tok] + P[pre + "position_embeddings.weight"][pos] + P[pre + "token_type_embeddings.weight"][0]
    ht = _ln(e, P, pre + "LayerNorm.", cfg.bert_ln_eps)     # BertEmbeddings, modeling_bert.py:222-237 (segment 0)
    # who sees whom among the text rows (the context is visible to all of them; context rows see the context only):
    # token row i: token rows j <= i; MASK row of step t: token rows j < t and itself
    i_idx = torch.arange(T, device=dev)
    tok_tok = (i_idx.unsqueeze(1) >= i_idx.unsqueeze(0))
    msk_tok = (i_idx.unsqueeze(1) + 1 > i_idx.unsqueeze(0))
    eye = torch.eye(T, dtype=torch.bool, device=dev)
    vis = torch.cat([torch.cat([tok_tok, torch.zeros(T, T, dtype=torch.bool, device=dev)], 1), torch.cat([msk_tok, eye], 1)], 0)
    add = torch.cat([torch.zeros(2 * T, C, device=dev), (~vis).float() * NEG], dim=1).view(1, 1, 2 * T, C + 2 * T)
    scale = 1.0 / math.sqrt(d)
    for l in range(cfg.dec_layers):
        p = "module.bert.decoder.layer.%d." % l
        qc, kc, vc = _qkv(P, p, cfg, hc)
        qt, kt, vt = _qkv(P, p, cfg, ht)
        # context rows among themselves (the prefill)
        ac = torch.softmax((qc @ kc.transpose(-1, -2)) * scale, dim=-1) @ vc
        # text rows over [context of their image | own text rows]
        kk = torch.cat([kc.repeat_interleave(K, dim=0) if K > 1 else kc, kt], dim=2)
        vv = torch.cat([vc.repeat_interleave(K, dim=0) if K > 1 else vc, vt], dim=2)
        at = torch.softmax((qt @ kk.transpose(-1, -2)) * scale + add, dim=-1) @ vv
        hc = _bert_output(P, p, cfg, hc, ac.permute(0, 2, 1, 3).reshape(B, C, H * d))
        ht = _bert_output(P, p, cfg, ht, at.permute(0, 2, 1, 3).reshape(R, 2 * T, H * d))
    # BertLMPredictionHead on the MASK rows (modeling_bert.py:540-563, 809-812)
    hp = "module.cls.predictions."
    hm = _gelu(F.linear(ht[:, T:], P[hp + "transform.dense.weight"], P[hp + "transform.dense.bias"]))
    hm = _ln(hm, P, hp + "transform.LayerNorm.", cfg.bert_ln_eps)
    logits = F.linear(hm, P[hp + "decoder.weight"]) + P[hp + "bias"]           # (R, T, V)
    if temperature != 1.0:
        logits = logits / temperature
    if top_k > 0 or top_p < 1.0:
        logits = _filter(logits, top_k, top_p)
    chosen = ids[:, 1:]                                                        # token emitted by step t = position t
    lp = torch.gather(F.log_softmax(logits, dim=-1), -1, chosen.unsqueeze(-1)).squeeze(-1)   # (R, T)
    # unfinished BEFORE each step (modeling_utils.py:855, 861-863): 1 up to and including the step that emits the first EOS
    is_eos = torch.zeros_like(chosen, dtype=torch.bool)
    for eid in eos_ids:
        is_eos |= chosen == eid
    ended_before = torch.cat([torch.zeros(R, 1, dtype=torch.bool, device=dev), is_eos[:, :-1]], 1).cumsum(1) > 0
    unf = (~ended_before).float()
    lp = torch.where(ended_before, torch.zeros_like(lp), lp)        # (a filtered PAD would contribute -inf * 0)
    return lp.sum(1) / unf.sum(1)


def _filter(logits, top_k, top_p, min_tokens_to_keep=1):
    """top_k_top_p_filtering (modeling_utils.py:1103-1135) on (.., V) logits; the kept set is a function of the values, the
    gradient flows through the kept entries."""
    if top_k > 0:
        k = max(min(top_k, logits.shape[-1]), min_tokens_to_keep)
        kth = torch.topk(logits, k)[0][..., -1, None]
        logits = logits.masked_fill(logits < kth, float("-inf"))
    if top_p < 1.0:
        s, idx = torch.sort(logits, descending=True)
        cum = torch.cumsum(F.softmax(s, dim=-1), dim=-1)
        remove = cum > top_p
        if min_tokens_to_keep > 1:
            remove[..., :min_tokens_to_keep] = False
        remove[..., 1:] = remove[..., :-1].clone()
        remove[..., 0] = False
        logits = logits.masked_fill(remove.scatter(-1, idx, remove), float("-inf"))
    return logits


class ScstSampler:
    """The model side of the reference's SCST step (legacy pipeline :447-468):

        out = ScstSampler(model, num_return_sequences=5)(data)
        loss = scst_criterion(gt_captions, decode(out['greedy_ids']), decode(out['sample_ids']), out['sample_logprobs'])
        loss.backward()

    greedy_ids (B, L) / sample_ids (B*K, L): int64, generated by the CUDA kernels under no_grad; sample_logprobs (B*K,):
    ``sequence_logprobs`` of the sampled captions, requires_grad; sample_logprobs_kernels: the generation pass's own figure
    (no gradient; equals sample_logprobs.detach() to the kernels' precision)."""

    def __init__(self, model, num_return_sequences=5, temperature=1.0, top_k=0, top_p=1.0):
        self.model = model
        self.K, self.temperature, self.top_k, self.top_p = int(num_return_sequences), float(temperature), int(top_k), float(top_p)

    def _generate(self, data, **overrides):
        m = self.model
        saved = m.test_extra_input
        m.test_extra_input = dict(saved, **overrides)
        try:
            with torch.no_grad():
                return m(data)
        finally:
            m.test_extra_input = saved

    def __call__(self, data, image_for_grad=None):
        greedy_ids, greedy_lp = self._generate(data, do_sample=False, num_return_sequences=1, num_beams=1, num_keep_best=1)
        sample_ids, sample_lp = self._generate(data, do_sample=True, num_return_sequences=self.K, num_beams=1, num_keep_best=1,
                                               temperature=self.temperature, top_k=self.top_k, top_p=self.top_p)
        raw_ids = self.model.last_raw_ids                     # the chosen tokens (no forced EOS), see sequence_logprobs
        assert raw_ids is not None and raw_ids.shape[0] == sample_ids.shape[0]
        image = data["image"] if image_for_grad is None else image_for_grad
        if image.dtype == torch.uint8:
            raise NotImplementedError("ScstSampler needs the float image tensor of the reference contract for the gradient pass")
        with torch.enable_grad():
            lp = sequence_logprobs(self.model, image.float(), raw_ids, self.K, self.temperature, self.top_k, self.top_p)
        return {"greedy_ids": greedy_ids[:, 0], "greedy_logprobs": greedy_lp[:, 0], "sample_ids": sample_ids[:, 0],
                "sample_logprobs": lp, "sample_logprobs_kernels": sample_lp[:, 0]}
