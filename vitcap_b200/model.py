"""Drop-in replacement for the reference's eval model
``ImageCaptioning(ViTCAP(config), test_extra_input, image_encoder=InputAsDict(timm ViT))``
(tagger_caption_uni_pipeline_expanding_bertemb.py:23-189, 566-618; modeling_bert.py:695-1059, 1307-1516).

Same call contract:  ``model(data) -> (ids int64 (B*K, num_keep_best, max_length), logprobs fp32 (B*K, num_keep_best))``
with ``data`` = {'image', 'input_ids', 'attention_mask', 'token_type_ids', 'masked_pos', 'key'} and the decode flags
of ``test_extra_input``. Same parameter names, so ``state_dict`` / ``Checkpointer.load`` work unchanged.
All compute runs in the sm_100a kernels of libvitcap_b200.so; this file only holds parameters and orchestration.
"""
import os

import torch
import torch.nn as nn

from . import ops, synth
from .config import VitCapConfig
from .engine import CaptionEngine, PackedWeights


# --------------------------------------------------------------------------- parameter containers (never "called")
class _Mlp(nn.Module):
    def __init__(self, h, f):
        super().__init__()
        self.fc1 = nn.Linear(h, f)
        self.fc2 = nn.Linear(f, h)


class _Attn(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.qkv = nn.Linear(h, 3 * h)
        self.proj = nn.Linear(h, h)


class _Block(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.norm1 = nn.LayerNorm(cfg.hidden, eps=cfg.vit_ln_eps)
        self.attn = _Attn(cfg.hidden)
        self.norm2 = nn.LayerNorm(cfg.hidden, eps=cfg.vit_ln_eps)
        self.mlp = _Mlp(cfg.hidden, cfg.inter)


class _SplitEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.blocks = nn.ModuleList([_Block(cfg) for _ in range(cfg.enc_blocks)])
        self.tag_blocks = nn.ModuleList([_Block(cfg) for _ in range(cfg.split_blocks)])


class _Embeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab, cfg.hidden)
        self.position_embeddings = nn.Embedding(cfg.max_pos, cfg.hidden)
        self.token_type_embeddings = nn.Embedding(cfg.type_vocab, cfg.hidden)
        self.LayerNorm = nn.LayerNorm(cfg.hidden, eps=cfg.bert_ln_eps)


class _Dense(nn.Module):
    def __init__(self, i, o):
        super().__init__()
        self.dense = nn.Linear(i, o)


class _DenseLN(nn.Module):
    def __init__(self, i, o, eps):
        super().__init__()
        self.dense = nn.Linear(i, o)
        self.LayerNorm = nn.LayerNorm(o, eps=eps)


class _SelfAttn(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.query = nn.Linear(h, h)
        self.key = nn.Linear(h, h)
        self.value = nn.Linear(h, h)


class _BertAttention(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.self = _SelfAttn(cfg.hidden)
        self.output = _DenseLN(cfg.hidden, cfg.hidden, cfg.bert_ln_eps)


class _BertLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.attention = _BertAttention(cfg)
        self.intermediate = _Dense(cfg.hidden, cfg.inter)
        self.output = _DenseLN(cfg.inter, cfg.hidden, cfg.bert_ln_eps)


class _BertEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([_BertLayer(cfg) for _ in range(cfg.dec_layers)])


class _Predictions(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(cfg.vocab))
        self.transform = _DenseLN(cfg.hidden, cfg.hidden, cfg.bert_ln_eps)
        self.decoder = nn.Linear(cfg.hidden, cfg.vocab, bias=False)


class _Heads(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.predictions = _Predictions(cfg)


class _Bert(nn.Module):
    """Parameter layout of ViTSplitCLSEmbModel (modeling_bert.py:1307-1365)."""

    def __init__(self, cfg):
        super().__init__()
        self.embeddings = _Embeddings(cfg)
        self.extra_embeddings = _Embeddings(cfg)       # unused when tagemb == 'cls'; kept for checkpoint compatibility
        self.encoder = _SplitEncoder(cfg)
        self.caption_pooler = _Dense(cfg.hidden, cfg.hidden)   # dead in eval (modeling_bert.py:1513, result unused)
        self.pooler = _Dense(cfg.hidden, cfg.hidden)
        self.tag_logit = _Heads(cfg)
        self.decoder = _BertEncoder(cfg)


class _PatchEmbed(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.proj = nn.Conv2d(3, cfg.hidden, kernel_size=cfg.patch, stride=cfg.patch)


class _TimmVit(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, cfg.hidden))
        self.pos_embed = nn.Parameter(torch.zeros(1, cfg.n_tokens, cfg.hidden))
        self.patch_embed = _PatchEmbed(cfg)
        self.head = nn.Linear(cfg.hidden, 1000)        # unused classifier head of the timm model


_UNSUPPORTED = "vitcap_b200 does not implement %s (the reference path for it is host-side Python; see DESIGN.md)"


class FastImageEncoder(nn.Module):
    """``InputAsDict(timm ViT with blocks=[] and norm=Identity)`` (pipeline file lines 750-778): patch embed + cls + pos."""

    def __init__(self, cfg, owner):
        super().__init__()
        self.module = _TimmVit(cfg)
        self._owner = [owner]          # list: not registered as a sub-module

    def forward(self, data_dict):
        im = data_dict if isinstance(data_dict, torch.Tensor) else data_dict["image"]
        return self._owner[0]._patch_embed(im)


class FastViTCAP(nn.Module):
    """``ViTCAP`` (modeling_bert.py:695-1059) with the same ``forward``/``generate`` signatures."""

    def __init__(self, cfg, owner):
        super().__init__()
        self.cfg = cfg
        self.bert = _Bert(cfg)
        self.cls = _Heads(cfg)
        self.cls.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight    # tie_weights, :728-730
        self._owner = [owner]

    def forward(self, *args, **kwargs):
        if kwargs.get("inference_mode", ""):
            raise NotImplementedError(_UNSUPPORTED % "inference_mode='prod'/'prod_no_hidden' (batch-1 paths)")
        if kwargs.get("is_decode", False):
            return self.generate(*args, **kwargs)
        raise NotImplementedError(_UNSUPPORTED % "encode_forward (training / teacher-forced scoring)")

    def generate(self, img_feats, label=None, attention_mask=None, masked_pos=None, token_type_ids=None,
                 position_ids=None, head_mask=None, input_ids=None, max_length=None, do_sample=None, num_beams=None,
                 temperature=None, top_k=None, top_p=None, repetition_penalty=None, bos_token_id=None, pad_token_id=None,
                 eos_token_ids=None, mask_token_id=None, length_penalty=None, num_return_sequences=None, num_keep_best=1,
                 is_decode=None, add_od_labels=False, od_labels_start_posid=None, use_cbs=False, fsm=None,
                 num_constraints=None, min_constraints_to_satisfy=None, use_hypo=False, decoding_constraint_flag=None,
                 bad_ending_ids=None, gen_tag_ratio=None, seed=None):
        assert is_decode
        owner = self._owner[0]
        cfg = self.cfg
        if use_cbs:
            raise NotImplementedError(_UNSUPPORTED % "constrained beam search (use_cbs)")
        if repetition_penalty not in (None, 1, 1.0):
            raise NotImplementedError(_UNSUPPORTED % "repetition_penalty != 1")
        if head_mask is not None:
            raise NotImplementedError(_UNSUPPORTED % "head_mask")
        if position_ids is not None:
            raise NotImplementedError(_UNSUPPORTED % "caller-supplied position_ids")
        num_beams = 1 if num_beams is None else int(num_beams)
        nret = 1 if num_return_sequences is None else int(num_return_sequences)
        if num_beams > 1 and do_sample:
            raise NotImplementedError(_UNSUPPORTED % "beam sampling (num_beams > 1 with do_sample)")
        if num_beams > 1 and nret != 1:
            raise NotImplementedError(_UNSUPPORTED % "num_return_sequences > 1 with beam search")
        if num_beams == 1 and num_keep_best != 1:
            raise AssertionError("cannot generate >1 sentences in greedy search")      # modeling_utils.py:785
        B = img_feats.shape[0]
        assert input_ids is not None and input_ids.shape[0] == B                    # modeling_bert.py:962
        max_length = int(max_length)
        if max_length > cfg.max_seq or max_length < 2:
            raise ValueError("max_length must be in [2, %d]" % cfg.max_seq)
        # limits of the search kernels (search.cu VC_MAX_BEAMS / VC_MAX_LEN; decode_attention: cur_len + 1 <= 64), checked
        # before any kernel or graph capture starts
        if max_length > 64:
            raise ValueError("max_length must be <= 64 (caption rows per sequence the decode kernels hold)")
        if num_beams > 1 and (num_beams > 8 or not 1 <= int(num_keep_best) <= 64):
            raise ValueError("beam search supports num_beams <= 8 and 1 <= num_keep_best <= 64")
        n_label = owner._label_counts(attention_mask, input_ids, max_length)
        if n_label is not None and not add_od_labels:
            raise NotImplementedError(_UNSUPPORTED % "a visible label region with add_od_labels=False (the reference fails there)")
        return owner._generate(img_feats, n_label=n_label, max_length=max_length, do_sample=bool(do_sample), num_beams=num_beams,
                               temperature=float(1.0 if temperature is None else temperature),
                               top_k=int(top_k or 0), top_p=float(1.0 if top_p is None else top_p),
                               bos=int(bos_token_id), pad=int(pad_token_id), eos_ids=[int(e) for e in eos_token_ids],
                               mask_id=int(mask_token_id), length_penalty=float(1.0 if length_penalty is None else length_penalty),
                               nret=nret, keep=int(num_keep_best), seed=seed)


class FastImageCaptioning(nn.Module):
    """B200-native ``ImageCaptioning`` (eval branch). ``mode``: 'bf16' = tcgen05 tensor cores, bf16 operands, fp32
    accumulation; 'fp32' = exact mode on CUDA cores (token ids / tag indices match the fp32 reference)."""

    def __init__(self, cfg: VitCapConfig, test_extra_input=None, mode="bf16", tokenizer=None, max_batch=64,
                 use_cuda_graph=True, sample_seed=0, graph_forward=False, decode_precision=None):
        super().__init__()
        assert mode in ("bf16", "fp32")
        self.cfg = cfg
        self.mode = mode
        # fast mode only -- operands of the decode-step MLP and vocabulary-head GEMMs, the layers whose operand rounding flips
        # near-tie arg-max decisions against the fp32 reference:
        #   'fp16'   (default) ONE product on IEEE-half operands: 11-bit significands instead of bf16's 8 at bf16's tensor-core
        #            rate and bytes (LayerNorm / GELU outputs and weights sit well inside the half range; conversions saturate
        #            at +-65504). Token agreement with the fp32 reference 99.6 %, as bf16x3, without its two extra products.
        #   'bf16x3' split-bf16 operands, three tensor-core products per GEMM (~fp32 operand precision): 99.6-99.7 %
        #   'bf16'   plain bf16 operands everywhere: 98.5 %
        # Default from VITCAP_DECODE_PRECISION, else 'fp16'.
        if decode_precision is None:
            decode_precision = os.environ.get("VITCAP_DECODE_PRECISION", "fp16")
        assert decode_precision in ("bf16", "bf16x3", "fp16")
        if ops.HALF_STORE and mode == "bf16":
            # VITCAP_STORE=fp16: EVERY 16-bit operand of the fast mode is an IEEE half already (ops.STORE); the decode step then
            # runs its plain one-product path on them, which is what 'fp16' asks for and more
            decode_precision = "plain"
        self.operand_storage = ("fp16" if ops.HALF_STORE else "bf16") if mode == "bf16" else "fp32"
        self.decode_precision = decode_precision if mode == "bf16" else "fp32"
        self.module = FastViTCAP(cfg, self)
        self.image_encoder = FastImageEncoder(cfg, self)
        self.test_extra_input = dict(test_extra_input) if test_extra_input is not None else synth.default_test_extra_input(cfg)
        self.tokenizer = tokenizer
        self.max_batch = max_batch
        self.use_cuda_graph = use_cuda_graph
        self.sample_seed = sample_seed
        # small-batch serving: capture the WHOLE forward (patch embed .. token ids) in one CUDA graph per batch shape, so a call
        # costs one image copy + one graph launch instead of ~170 eager launches (see _forward_graphed)
        self.graph_forward = graph_forward
        self._skip_mask_check = False
        self.u8_channel_order = "bgr"       # what cv2 / the reference's TSV image decoder deliver (BGR2RGB is then fused)
        self._engine = None
        self.last_tags = None               # (topk idx int32 (B,K), prob fp32 (B,K)) of the most recent forward(), device tensors
        self._sample_calls = 0
        self._label_flip_hold = None        # None: decide per generate() call; False: forward() in progress, not decided yet
        self._tag_parts = None              # list while forward() runs: per-chunk concept top-k collected by _generate
        # greedy / sampling: the tokens the search actually chose at every step, int64 (B*K, L), BEFORE the forced EOS of rows
        # that ran to max_length (modeling_utils.py:869-871 overwrites the last position of the returned ids, while the
        # returned log-prob belongs to the token that was chosen there): what a teacher-forced re-scoring needs (scst.py)
        self.last_raw_ids = None
        self._raw_parts = None
        self.register_load_state_dict_post_hook(lambda m, keys: m._invalidate())
        self.eval()

    # ---- packing -----------------------------------------------------------------------------------------------------
    def _invalidate(self):
        self._engine = None

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() replace the parameter storage: the packed kernel copies (and an engine bound to the old
        # device) are stale
        self._invalidate()
        return super()._apply(fn, *args, **kwargs)

    def _param_version(self):
        return sum(p._version for p in self.parameters())

    def pack(self):
        """(Re)builds the kernel-side weight copies from the current parameters. Called lazily by forward()."""
        ops.load_library()
        dev = self.module.bert.embeddings.word_embeddings.weight.device
        if dev.type != "cuda":
            raise RuntimeError("vitcap_b200 runs on a CUDA device only (no CPU path): move the module with .cuda() first")
        sd = self.state_dict()
        w = PackedWeights(self.cfg, sd, self.mode, dev, decode_x3=self.decode_precision == "bf16x3",
                          decode_f16=self.decode_precision == "fp16")
        self._engine = CaptionEngine(self.cfg, w, dev, use_cuda_graph=self.use_cuda_graph)
        self._packed_version = self._param_version()
        return self

    @property
    def engine(self):
        # in-place parameter updates (an optimizer step, manual edits, re-tying) bump the tensors' version counters
        if self._engine is None or self._packed_version != self._param_version():
            self.pack()
        return self._engine

    # ---- pieces called by the sub-modules ----------------------------------------------------------------------------
    def _patch_embed(self, image):
        S = self.cfg.img_size
        if image.dtype == torch.uint8:
            # 8-bit pixels straight from the host pipeline (after resize / center-crop): uint8 (B, S, S, 3), HWC, channel order
            # self.u8_channel_order; ToTensor + Normalize(0.5, 0.5) run on the device (additive to the reference contract)
            if image.dim() != 4 or tuple(image.shape[1:]) != (S, S, 3):
                raise ValueError("uint8 images must be (B, %d, %d, 3) HWC" % (S, S))
            return self.engine.patch_embed(image.contiguous(), bgr=(self.u8_channel_order == "bgr"))
        image = image.to(dtype=torch.float32).contiguous()
        if image.shape[-1] != S or image.shape[-2] != S:
            raise NotImplementedError(_UNSUPPORTED % "position-embedding interpolation for a different image size")
        return self.engine.patch_embed(image)

    def _label_counts(self, attention_mask, input_ids, max_length):
        """Validates the text attention mask and returns the number of visible od/tag label slots per sample (int64 [B]), or
        None when there is none. Accepted: the seq2seq family of the reference data layer (dataset.py:395-408): the caption
        triangle; for the first n label slots full attention L-L and C-L; nothing else. n == 0 everywhere is what the eval
        pipeline passes (text_b == ''). The kernels encode exactly that structure; any other mask is refused
        instead of silently diverging."""
        if attention_mask is None or self._skip_mask_check:
            return None
        cfg = self.cfg
        m = attention_mask
        if m.dim() != 3 or m.shape[1] != m.shape[2]:
            raise NotImplementedError(_UNSUPPORTED % "attention_mask that is not (B, S, S)")
        S = m.shape[1]
        T = input_ids.shape[1]
        if S == T + cfg.n_tokens:                    # full mask built by construct_attn_mask: check its text block
            m = m[:, :T, :T]
            S = T
        a = max_length
        n = (m[:, 0, a:] != 0).sum(dim=1)
        ar = torch.arange(S, device=m.device)
        in_c = (ar < a)
        in_l = (ar.unsqueeze(0) >= a) & (ar.unsqueeze(0) < a + n.unsqueeze(1))            # (B, S) visible label slots
        tri = torch.tril(torch.ones(S, S, device=m.device, dtype=torch.bool)) & in_c.unsqueeze(0) & in_c.unsqueeze(1)
        ref = tri.unsqueeze(0) | ((in_c.view(1, S, 1) | in_l.unsqueeze(2)) & in_l.unsqueeze(1))
        if not bool(((m != 0) == ref).all()):
            raise NotImplementedError(_UNSUPPORTED % "a text attention mask outside the seq2seq family of dataset.py:395-408")
        if not bool((n > 0).any()):
            return None
        if cfg.tagemb != "cls":
            # the visible label rows would come from encode_tag_to_embedding(cls_emb=None) / extra_embeddings
            # (modeling_bert.py:1466, 1485); only the 'cls' recipes of the shipped configuration are built
            raise NotImplementedError(_UNSUPPORTED % ("a visible label region with config.tagemb=%r (only 'cls')" % cfg.tagemb))
        if T - a != cfg.topk:
            # modeling_bert.py:1470/1489 writes the topk tag embeddings over the LAST topk input slots
            raise NotImplementedError(_UNSUPPORTED % "a label region whose length differs from config.topk")
        return n

    def _generate(self, img_feats, max_length, do_sample, num_beams, temperature, top_k, top_p, bos, pad, eos_ids, mask_id,
                  length_penalty, nret, keep, seed, n_label=None):
        eng = self.engine
        B = img_feats.shape[0]
        outs_i, outs_l = [], []
        tags = self._tag_parts
        hold = self._label_flip_hold           # an int once forward() has decided it for a caller batch fed in several chunks
        flip = None if (hold is None or hold is False) else hold
        if do_sample:
            if seed is None:
                seed = self.sample_seed + self._sample_calls * 0x9E3779B97F4A7C15
                self._sample_calls += 1
        for s in range(0, B, self.max_batch):
            chunk = img_feats[s:s + self.max_batch]
            b = chunk.shape[0]
            if n_label is not None:
                eng.reserve(b, label_rows=True)    # before encode(): growing the workspace later would drop its outputs
            eng.encode(chunk)
            _, tag_idx, tag_prob, tag_len = eng.tag_head(b)
            if tags is not None:               # the concept head's top-k of this chunk (the workspace is reused by the next)
                tags.append((tag_idx.clone(), tag_prob.clone()))
            if n_label is None:
                eng.prefill(b)
            else:
                # The reference picks the label-embedding recipe per step from the FIRST sample's tag count
                # (modeling_bert.py:1435: topk_len[0] + 20 <= cur_len + 1 + label slots -> 'raw', else 'ln'); one host read,
                # where the reference itself synchronises. `flip` = first cur_len decoded under 'raw'.
                if flip is None:
                    flip = max(1, int(tag_len[0].item()) + 20 - 1 - self.cfg.topk)
                    if self._label_flip_hold is False:
                        self._label_flip_hold = flip
                eng.set_labels(b, n_label[s:s + b])
                eng.prefill(b, label_recipe="raw" if flip <= 1 else "ln")
            if num_beams > 1:
                ids, lp = eng.beam_search(b, num_beams, max_length, bos, pad, eos_ids, mask_id, length_penalty, keep,
                                          label_flip=flip if n_label is not None else None)
            else:
                ids, lp = eng.greedy_or_sample(b, nret, max_length, bos, pad, eos_ids, mask_id, do_sample, temperature, top_k,
                                               top_p, seed=(seed + s) if do_sample else 0,
                                               label_flip=flip if n_label is not None else None)
                if self._raw_parts is not None:
                    self._raw_parts.append(eng._decoder_ws(b, nret, max_length)["ids"].to(torch.int64))
            outs_i.append(ids)
            outs_l.append(lp)
        return torch.cat(outs_i, 0), torch.cat(outs_l, 0)

    # ---- public entry points -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, data):
        """Eval branch of ImageCaptioning.forward (pipeline file lines 87-112, 173-184)."""
        if self.training:
            raise NotImplementedError(_UNSUPPORTED % "training mode")
        data = dict(data.items())
        data.pop("key", None)
        image = data.pop("image")
        B = image.shape[0]
        extra = dict(self.test_extra_input)
        if self.graph_forward and self.use_cuda_graph and B <= self.max_batch and not extra.get("do_sample"):
            out = self._forward_graphed(image, data, extra)
            if out is not None:
                return out
        return self._forward_eager(image, data, extra)

    def _forward_graphed(self, image, data, extra):
        """Small-batch latency path (the role of the reference's batch-1 ``prod_generate``, modeling_bert.py:1075-1202, which
        fails on this model -- DESIGN.md section 8): the whole forward of one batch shape is one CUDA graph. The first call of
        a shape runs eagerly (and is the result), then the same call sequence is captured; later calls copy the image into the
        graph's input buffer and replay. Greedy and beam search without a visible label region (the label recipe needs a host
        read); anything else returns None and takes the eager path."""
        eng = self.engine
        if not image.is_cuda:
            return None
        ids_in, mask = data.get("input_ids"), data.get("attention_mask")
        if ids_in is None:
            return None
        if mask is not None and self._label_counts(mask, ids_in, int(extra["max_length"])) is not None:
            return None                                   # (the mask is validated here, outside the graph: one host read)
        key = (tuple(image.shape), image.dtype, tuple(ids_in.shape), self.u8_channel_order,
               tuple(sorted((k, tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in extra.items())))
        ent = eng.forward_graphs.get(key)
        if ent is None:
            eng.reserve(self.max_batch)                        # growing the image-side workspace later would drop every graph
            out = self._forward_eager(image, data, extra)      # allocates the workspaces, configures the kernels
            eager_tags = self.last_tags
            torch.cuda.current_stream().synchronize()
            static_img = image.detach().clone().contiguous()
            static_data = {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in data.items()}
            g = torch.cuda.CUDAGraph()
            eng.inline_graphs, self._skip_mask_check = True, True
            try:
                with torch.cuda.graph(g):
                    ids, lp = self._forward_eager(static_img, static_data, extra)
                    tags = self.last_tags
            finally:
                eng.inline_graphs, self._skip_mask_check = False, False
            # the entry keeps the decode workspaces alive (the engine's LRU may drop them; the graph holds their pointers)
            eng.forward_graphs[key] = {"graph": g, "image": static_img, "data": static_data, "out": (ids, lp), "tags": tags,
                                       "dec_ws": list(eng._dec_ws.values())}
            self.last_tags = eager_tags                    # (the capture pass set it to the graph's not yet written buffers)
            return out
        ent["image"].copy_(image, non_blocking=True)
        ent["graph"].replay()
        self.last_raw_ids = None
        eng.stats["forward_graph_replays"] = eng.stats.get("forward_graph_replays", 0) + 1
        self.last_tags = (ent["tags"][0].clone(), ent["tags"][1].clone()) if ent["tags"] is not None else None
        return ent["out"][0].clone(), ent["out"][1].clone()

    def _forward_eager(self, image, data, extra):
        B = image.shape[0]
        ids_all, lp_all = [], []
        self._label_flip_hold = False              # the label recipe follows the first sample of the WHOLE batch (see _generate)
        self._tag_parts = []
        self._raw_parts = [] if not self.engine.inline_graphs else None
        try:
            for s in range(0, B, self.max_batch):      # the image stream buffer holds max_batch images
                sub = {k: (v[s:s + self.max_batch] if torch.is_tensor(v) and v.shape[:1] == (B,) else v) for k, v in data.items()}
                sub["img_feats"] = self.image_encoder({"image": image[s:s + self.max_batch]})
                sub["gen_tag_ratio"] = 1
                sub.update(extra)
                ids, lp = self.module(**sub)
                ids_all.append(ids)
                lp_all.append(lp)
        finally:
            self._label_flip_hold = None
            parts, self._tag_parts = self._tag_parts, None
            raw, self._raw_parts = self._raw_parts, None
        self.last_raw_ids = torch.cat(raw, 0) if raw else None
        self.last_tags = (torch.cat([p[0] for p in parts], 0), torch.cat([p[1] for p in parts], 0)) if parts else None
        return torch.cat(ids_all, 0), torch.cat(lp_all, 0)

    @torch.no_grad()
    def forward_tags(self, image, caption_branch=True):
        """Additive API (BASELINE config 2): patch embed -> split encoder -> concept head -> top-k.
        Returns (tag_logits fp32 [B,V], topk_idx int64 [B,K] sorted by score, topk_prob fp32 [B,K], topk_len int64 [B])
        = what ``bert(...)`` computes at modeling_bert.py:1415-1432."""
        eng = self.engine
        outs = []
        for s in range(0, image.shape[0], self.max_batch):
            f = self._patch_embed(image[s:s + self.max_batch])
            eng.encode(f, caption_branch=caption_branch)
            lg, idx, pr, n = eng.tag_head(f.shape[0])
            outs.append((lg.clone(), idx.long(), pr.clone(), n.long()))
        return tuple(torch.cat([o[i] for o in outs], 0) for i in range(4))

    @torch.no_grad()
    def encode_features(self, image):
        """(caption feats, tag feats) fp32 [B,N,H] of the split encoder (modeling_bert.py:458-478), for parity checks."""
        eng = self.engine
        assert image.shape[0] <= self.max_batch
        f = self._patch_embed(image)
        cap, tag = eng.encode(f, full_tag_feats=True)
        return cap.clone(), tag.clone()


def build_from_state_dict(cfg, state_dict, device="cuda", **kw):
    m = FastImageCaptioning(cfg, **kw)
    m.load_state_dict(state_dict, strict=True)
    return m.to(device)
