"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32) of the reference's caption path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file. The product (``vitcap_b200``) never does.

Two restatements of jacobswan1/ViTCAP's eval hot path, both driven by the reference's own
``state_dict`` layout:

``faithful``  the algorithm exactly as shipped: every decode step re-runs the ViT trunk, the tag
              head, the 50 od/tag slots and the whole decoder over all T+578 rows with the dense
              647x647 mask (modeling_bert.py:825-923, 1408-1516; modeling_utils.py:768-1100).
              This is the CPU baseline that gets timed.
``cached``    the same mathematics with the shared trunk evaluated once, the 578 context rows
              prefetched through the decoder once with K/V cached, and two rows per step
              (SURVEY.md section 7.1). This is the executable spec of what the CUDA kernels compute.

Parity pinning: the reference ships no tests or golden vectors, so both
restatements are pinned against the reference *itself*, imported and run in the build container
by ``oracle/make_golden.py``; the outputs are committed under ``tests/golden/`` and checked by
``tests/test_oracle_golden.py`` (which needs no reference tree).
"""
import contextlib
import math

import numpy as np
import torch
import torch.nn.functional as F

from vitcap_b200.config import VitCapConfig

NEG_MASK = -10000.0   # modeling_bert.py:1501


def gelu_erf(x):
    """activations.py:16-23 (_gelu_python) and nn.GELU (vision_transformer.py:143)."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


class PortModel:
    def __init__(self, cfg: VitCapConfig, state_dict, dtype=torch.float32):
        self.cfg = cfg
        self.sd = {k: v.detach().to(dtype) for k, v in state_dict.items()}

    def p(self, key):
        return self.sd[key]

    # ---- image side ---------------------------------------------------------------------
    def patch_embed(self, image):
        """vision_transformer.py:267-275 (PatchEmbed) + 411-427 (cls token, pos embed); the
        image encoder has no blocks and norm=Identity (pipeline file lines 767-769)."""
        cfg = self.cfg
        w = self.p("image_encoder.module.patch_embed.proj.weight")
        b = self.p("image_encoder.module.patch_embed.proj.bias")
        x = F.conv2d(image, w, b, stride=cfg.patch).flatten(2).transpose(1, 2)
        cls = self.p("image_encoder.module.cls_token").expand(x.shape[0], -1, -1)
        x = torch.cat([cls, x], dim=1) + self.p("image_encoder.module.pos_embed")
        return x

    def vit_block(self, x, prefix):
        """vision_transformer.py:233-250 (Block), 174-210 (Attention), 152-158 (Mlp)."""
        cfg = self.cfg
        B, N, C = x.shape
        H, d = cfg.heads, cfg.head_dim
        h = F.layer_norm(x, (C,), self.p(prefix + "norm1.weight"), self.p(prefix + "norm1.bias"), cfg.vit_ln_eps)
        qkv = F.linear(h, self.p(prefix + "attn.qkv.weight"), self.p(prefix + "attn.qkv.bias"))
        qkv = qkv.reshape(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) * (d ** -0.5)
        attn = attn.softmax(dim=-1)
        o = (attn @ v).transpose(1, 2).reshape(B, N, C)
        x = x + F.linear(o, self.p(prefix + "attn.proj.weight"), self.p(prefix + "attn.proj.bias"))
        h = F.layer_norm(x, (C,), self.p(prefix + "norm2.weight"), self.p(prefix + "norm2.bias"), cfg.vit_ln_eps)
        h = F.linear(h, self.p(prefix + "mlp.fc1.weight"), self.p(prefix + "mlp.fc1.bias"))
        h = gelu_erf(h)
        h = F.linear(h, self.p(prefix + "mlp.fc2.weight"), self.p(prefix + "mlp.fc2.bias"))
        return x + h

    def split_encoder(self, x):
        """modeling_bert.py:458-478: 12 blocks, the tag branch forks from the input of block
        len(blocks)-split_blocks; the timm final norm is never applied."""
        cfg = self.cfg
        tag = None
        for i in range(cfg.enc_blocks):
            if i == cfg.enc_blocks - cfg.split_blocks:
                tag = x
            x = self.vit_block(x, "module.bert.encoder.blocks.%d." % i)
        for j in range(cfg.split_blocks):
            tag = self.vit_block(tag, "module.bert.encoder.tag_blocks.%d." % j)
        return x, tag

    def head(self, prefix, h):
        """BertLMPredictionHead, modeling_bert.py:540-563."""
        cfg = self.cfg
        h = F.linear(h, self.p(prefix + "transform.dense.weight"), self.p(prefix + "transform.dense.bias"))
        h = gelu_erf(h)
        h = F.layer_norm(h, (cfg.hidden,), self.p(prefix + "transform.LayerNorm.weight"),
                         self.p(prefix + "transform.LayerNorm.bias"), cfg.bert_ln_eps)
        return F.linear(h, self.p(prefix + "decoder.weight")) + self.p(prefix + "bias")

    def tag_head(self, tag_feats):
        """modeling_bert.py:1424-1432: pooler (tanh) -> tag_logit head -> sigmoid -> topk -> len."""
        cfg = self.cfg
        pooled = torch.tanh(F.linear(tag_feats[:, 0], self.p("module.bert.pooler.dense.weight"),
                                     self.p("module.bert.pooler.dense.bias")))
        logit = self.head("module.bert.tag_logit.predictions.", pooled)
        prob, idx = torch.sigmoid(logit).topk(cfg.topk, dim=1, largest=True)
        topk_len = (prob >= cfg.tag_thresh).sum(dim=1)
        return logit, prob, idx, topk_len

    # ---- text side ----------------------------------------------------------------------
    def embeddings(self, ids, pos, typ=None):
        """BertEmbeddings.forward, modeling_bert.py:222-237."""
        cfg = self.cfg
        pre = "module.bert.embeddings."
        e = self.p(pre + "word_embeddings.weight")[ids] + self.p(pre + "position_embeddings.weight")[pos]
        e = e + self.p(pre + "token_type_embeddings.weight")[torch.zeros_like(ids) if typ is None else typ]
        return F.layer_norm(e, (cfg.hidden,), self.p(pre + "LayerNorm.weight"), self.p(pre + "LayerNorm.bias"),
                            cfg.bert_ln_eps)

    def bert_layer(self, idx, hq, hkv, add_mask):
        """BertLayer (modeling_bert.py:303-437) for query rows ``hq`` attending over ``hkv`` rows;
        returns (layer output for the query rows, K, V of the hkv rows as (B,H,S,d))."""
        cfg = self.cfg
        H, d = cfg.heads, cfg.head_dim
        p = "module.bert.decoder.layer.%d." % idx
        B, Sq, C = hq.shape

        def heads(t):
            return t.view(t.shape[0], t.shape[1], H, d).permute(0, 2, 1, 3)
        q = heads(F.linear(hq, self.p(p + "attention.self.query.weight"), self.p(p + "attention.self.query.bias")))
        k = heads(F.linear(hkv, self.p(p + "attention.self.key.weight"), self.p(p + "attention.self.key.bias")))
        v = heads(F.linear(hkv, self.p(p + "attention.self.value.weight"), self.p(p + "attention.self.value.bias")))
        out = self.bert_layer_from_kv(idx, hq, q, k, v, add_mask)
        return out, k, v

    def bert_layer_from_kv(self, idx, hq, q, k, v, add_mask, step=False):
        """step: True when called for the two rows of a decode step (only the quantised subclass cares)."""
        cfg = self.cfg
        p = "module.bert.decoder.layer.%d." % idx
        B, Sq, C = hq.shape
        s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(cfg.head_dim)
        if add_mask is not None:
            s = s + add_mask
        a = torch.softmax(s, dim=-1)
        ctx = torch.matmul(a, v).permute(0, 2, 1, 3).contiguous().view(B, Sq, C)
        a1 = F.linear(ctx, self.p(p + "attention.output.dense.weight"), self.p(p + "attention.output.dense.bias"))
        a1 = F.layer_norm(a1 + hq, (C,), self.p(p + "attention.output.LayerNorm.weight"),
                          self.p(p + "attention.output.LayerNorm.bias"), cfg.bert_ln_eps)
        m = gelu_erf(F.linear(a1, self.p(p + "intermediate.dense.weight"), self.p(p + "intermediate.dense.bias")))
        m = F.linear(m, self.p(p + "output.dense.weight"), self.p(p + "output.dense.bias"))
        return F.layer_norm(m + a1, (C,), self.p(p + "output.LayerNorm.weight"), self.p(p + "output.LayerNorm.bias"),
                            cfg.bert_ln_eps)

    def qkv_rows(self, idx, h):
        cfg = self.cfg
        H, d = cfg.heads, cfg.head_dim
        p = "module.bert.decoder.layer.%d." % idx

        def heads(t):
            return t.view(t.shape[0], t.shape[1], H, d).permute(0, 2, 1, 3)
        return (heads(F.linear(h, self.p(p + "attention.self.query.weight"), self.p(p + "attention.self.query.bias"))),
                heads(F.linear(h, self.p(p + "attention.self.key.weight"), self.p(p + "attention.self.key.bias"))),
                heads(F.linear(h, self.p(p + "attention.self.value.weight"), self.p(p + "attention.self.value.bias"))))


# =========================================================================================
# quantisation-matched restatement of the fast (bf16) mode
# =========================================================================================
def q_bf16(x):
    """Round to bf16 (nearest even) and return as fp32: one kernel rounding point."""
    return x.to(torch.bfloat16).to(torch.float32)


def q_f16(x):
    """Round to IEEE half (nearest even, saturating at +-65504 as cvt.rn.satfinite.f16.f32) and back to fp32."""
    return x.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32)


@contextlib.contextmanager
def half_store():
    """Inside the block every operand rounding of QuantPortModel is an IEEE-half rounding (q_bf16 IS q_f16): the executable spec
    of a VITCAP_STORE=fp16 process, where the kernels store every 16-bit operand as a half (csrc/common.cuh VC_STORE_F16).
    Build AND run the model inside the block (rounded weights are cached per instance)."""
    global q_bf16
    saved = q_bf16
    q_bf16 = q_f16
    try:
        yield
    finally:
        q_bf16 = saved


def split_bf16(x):
    """x ~ hi + lo with hi = bf16(x), lo = bf16(x - hi): the split operand of the decode-step GEMMs (engine.py, bf16x3)."""
    hi = q_bf16(x)
    return hi, q_bf16(x - hi)


def gelu_fast(x):
    """The GELU of the tensor-core epilogues (common.cuh gelu_erf_tanh): x/2 (1 + tanh(z (c0 + c1 u + c2 u^2))), z = x/sqrt 2,
    u = min(z^2, 30); evaluated here with an exact tanh (the kernels use MUFU.TANH, 2^-11 relative)."""
    z = x * 0.70710678118654752440
    u = torch.clamp(z * z, max=30.0)
    p = (-1.988479253896676e-03 * u + 1.0466777301852825e-01) * u + 1.1278464660309704
    hx = 0.5 * x
    return hx * torch.tanh(z * p) + hx


class QuantPortModel(PortModel):
    """Executable spec of the fast mode's ARITHMETIC: the fp32 algorithm of PortModel with every operand rounded to bf16 exactly
    where the CUDA kernels round (vitcap_b200/engine.py), fp32 accumulation everywhere else. The difference between the CUDA
    path and THIS model is kernel error (summation order, MUFU approximations, rounding flips); the difference between this
    model and PortModel is the operand quantisation the north star allows the bf16 mode. Rounding points:

      patch embed      patches and conv weight bf16, fp32 out (engine.patch_embed)
      ViT block        qkv / attention output / GELU(fc1) stored bf16; the residual stream stays fp32. norm1 / norm2 are either
                       the LayerNorm kernel (h = bf16(LN(x)), W bf16) or FOLDED into the consuming GEMM (gemm_tc2.cu LN = 2):
                       rstd (bf16(x) bf16(gamma o W)^T - mean colsum) + (b + W beta), colsum over the rounded weight
      attention        S = Q K^T fp32, P = exp(S - max) rounded to bf16 for the P V product, row sum from the unrounded P
      decoder prefill  post-LN BertLayer: LayerNorm kernel output bf16 (operand) + fp32 (residual), qkv / attention / GELU bf16
      decode step      the same; with decode_x3 the MLP and the vocabulary head run on split operands (hi + lo, three products),
                       with decode_f16 on operands rounded to IEEE half (one product; decode_precision='fp16')
      heads            dense + fast GELU fp32 -> LayerNorm -> bf16 (tag head) or split (vocabulary head) -> decoder GEMM

    ln_fold: 0 = LayerNorm kernel everywhere, 1 = norm1 folded, 2 = norm1 and norm2 folded (the engine's default).
    cls_only_last: the last concept block is evaluated for the CLS row only with LayerNorm-kernel roundings and an fp32 softmax
    (engine._vit_block_cls_only); False = a full folded block (engine.encode(full_tag_feats=True))."""

    def __init__(self, cfg, state_dict, ln_fold=2, decode_x3=True, cls_only_last=True, acc64=False, prefill_fold=True,
                 decode_f16=False):
        super().__init__(cfg, state_dict, dtype=torch.float32)
        self.ln_fold = ln_fold
        # decode_f16 takes the split-operand route through the layers (lin_x3 is the GEMM of both forms)
        self.decode_f16 = decode_f16
        self.decode_x3 = decode_x3 or decode_f16
        self.cls_only_last = cls_only_last
        # prefill_fold (engine._prefill_folded): in the decoder prefill the intermediate GEMM folds the attention-output
        # LayerNorm and the q|k|v GEMM of layer i >= 1 folds the output LayerNorm of layer i - 1 (raw bf16 row, gamma-scaled
        # weights, statistics applied in the epilogue); the residuals are the fp32 LayerNorm outputs as before
        self.prefill_fold = prefill_fold
        self._pf_raw = None                    # (layer output tensor, raw pre-LayerNorm rows, gamma key, beta key)
        # acc64: every product is accumulated in fp64 and rounded to fp32 once -- the SAME quantised arithmetic with another
        # (better) summation. The distance between the acc64 and the plain model is the yardstick for how far two correct
        # implementations of this spec drift apart end to end: a bf16 rounding turns a relative perturbation e of its input
        # into ~0.04 sqrt(e) of its output (flipped roundings), so ~1e-7 grows to the quantisation-noise level within a few
        # stages whatever the implementation
        self.acc64 = acc64
        self._wq = {}

    def _mm(self, a, w, bias=None):
        if self.acc64:
            return F.linear(a.double(), w.double(), bias.double() if bias is not None else None).float()
        return F.linear(a, w, bias)

    def wq(self, key):
        if key not in self._wq:
            self._wq[key] = q_bf16(self.sd[key])
        return self._wq[key]

    # ---- building blocks ----------------------------------------------------------------
    def lin(self, h_q, wkey, bkey):
        """bf16 operands (h_q already rounded), fp32 accumulate + fp32 bias."""
        return self._mm(h_q, self.wq(wkey), self.sd[bkey] if bkey else None)

    def lin_x3(self, a, wkey, bkey):
        """Three-product split-bf16 GEMM (gemm_tc.cu X3 / the K-concatenated form): a_hi w_hi + a_lo w_hi + a_hi w_lo; with
        decode_f16 the one-product IEEE-half GEMM of gemm_dec.cu (operands rounded to nearest even, saturating)."""
        if self.decode_f16:
            key = ("f16", wkey)
            if key not in self._wq:
                self._wq[key] = q_f16(self.sd[wkey])
            out = self._mm(q_f16(a), self._wq[key])
            return out + self.sd[bkey] if bkey else out
        key = ("x3", wkey)
        if key not in self._wq:
            self._wq[key] = split_bf16(self.sd[wkey])
        w_hi, w_lo = self._wq[key]
        a_hi, a_lo = split_bf16(a)
        if self.acc64:
            out = (F.linear(a_hi.double(), w_hi.double()) + F.linear(a_lo.double(), w_hi.double()) +
                   F.linear(a_hi.double(), w_lo.double())).float()
        else:
            out = F.linear(a_hi, w_hi) + F.linear(a_lo, w_hi) + F.linear(a_hi, w_lo)
        return out + self.sd[bkey] if bkey else out

    def lin_ln(self, x, gkey, bekey, eps, wkey, bkey):
        """LayerNorm kernel (two-pass fp32 statistics, bf16 output) followed by a bf16 GEMM."""
        h = q_bf16(F.layer_norm(x, (x.shape[-1],), self.sd[gkey], self.sd[bekey], eps))
        return self.lin(h, wkey, bkey), h

    def lin_fold(self, x, gkey, bekey, eps, wkey, bkey):
        """Folded LayerNorm (gemm_tc2.cu LN = 1/3 producer + LN = 2 consumer; weights packed in engine.PackedWeights.folded)."""
        key = ("fold", wkey)
        if key not in self._wq:
            w, g, be = self.sd[wkey], self.sd[gkey], self.sd[bekey]
            wf = q_bf16(w * g.unsqueeze(0))
            self._wq[key] = (wf, wf.sum(1), self.sd[bkey] + w @ be)
        wf, colsum, bias_f = self._wq[key]
        K = x.shape[-1]
        mean = x.sum(-1, keepdim=True) / K
        var = torch.clamp((x * x).sum(-1, keepdim=True) / K - mean * mean, min=0.0)     # one-pass, as the GEMM epilogue
        rstd = torch.rsqrt(var + eps)
        acc = self._mm(q_bf16(x), wf)
        return rstd * acc + ((-mean * rstd) * colsum + bias_f)

    def attend(self, q, k, v, scale, add_mask=None, round_p=True):
        """q, k, v fp32 tensors holding bf16 values, (B, H, S, d). Returns the bf16-rounded attention output (B, Sq, H d)."""
        # exp2 domain with an INTEGER exponent reference, as the kernels (attention_tc.cu, decode_attention_mma.cu): bf16
        # rounding commutes with powers of two, so bf16(P) is the same whichever integer reference (lazy, per chunk, per warp)
        # a kernel happened to use
        mm = (lambda a, b: torch.matmul(a.double(), b.double()).float()) if self.acc64 else torch.matmul
        s = mm(q, k.transpose(-1, -2)) * (scale * 1.4426950408889634)
        if add_mask is not None:
            s = s + add_mask
        p = torch.exp2(s - torch.ceil(s.max(dim=-1, keepdim=True).values))
        l = p.sum(dim=-1, keepdim=True)
        o = mm(q_bf16(p) if round_p else p, v) / l
        B, H, Sq, d = o.shape
        return q_bf16(o.permute(0, 2, 1, 3).reshape(B, Sq, H * d))

    # ---- image side ---------------------------------------------------------------------
    def patch_embed(self, image):
        cfg = self.cfg
        w = self.wq("image_encoder.module.patch_embed.proj.weight")
        b = self.p("image_encoder.module.patch_embed.proj.bias")
        x = F.conv2d(q_bf16(image), w, b, stride=cfg.patch).flatten(2).transpose(1, 2)
        cls = self.p("image_encoder.module.cls_token").expand(x.shape[0], -1, -1)
        return torch.cat([cls, x], dim=1) + self.p("image_encoder.module.pos_embed")

    def vit_block(self, x, prefix, fold1=False, fold2=False):
        cfg = self.cfg
        B, N, C = x.shape
        H, d = cfg.heads, cfg.head_dim
        eps = cfg.vit_ln_eps
        n1 = (prefix + "norm1.weight", prefix + "norm1.bias", eps, prefix + "attn.qkv.weight", prefix + "attn.qkv.bias")
        qkv = self.lin_fold(x, *n1) if fold1 else self.lin_ln(x, *n1)[0]
        qkv = q_bf16(qkv).reshape(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
        o = self.attend(qkv[0], qkv[1], qkv[2], d ** -0.5)
        x = x + self.lin(o, prefix + "attn.proj.weight", prefix + "attn.proj.bias")
        n2 = (prefix + "norm2.weight", prefix + "norm2.bias", eps, prefix + "mlp.fc1.weight", prefix + "mlp.fc1.bias")
        h = self.lin_fold(x, *n2) if fold2 else self.lin_ln(x, *n2)[0]
        h = q_bf16(gelu_fast(h))
        return x + self.lin(h, prefix + "mlp.fc2.weight", prefix + "mlp.fc2.bias")

    def vit_block_cls_only(self, x, prefix):
        """engine._vit_block_cls_only: K | V of every row, everything else for row 0; returns x with row 0 replaced."""
        cfg = self.cfg
        B, N, C = x.shape
        H, d = cfg.heads, cfg.head_dim
        eps = cfg.vit_ln_eps
        qkv, h = self.lin_ln(x, prefix + "norm1.weight", prefix + "norm1.bias", eps, prefix + "attn.qkv.weight", prefix + "attn.qkv.bias")
        qkv = q_bf16(qkv).reshape(B, N, 3, H, d).permute(2, 0, 3, 1, 4)
        o = self.attend(qkv[0][:, :, 0:1], qkv[1], qkv[2], d ** -0.5, round_p=False)       # cls_attention_kernel: fp32 P
        xc = x[:, 0:1] + self.lin(o, prefix + "attn.proj.weight", prefix + "attn.proj.bias")
        hc, _ = self.lin_ln(xc, prefix + "norm2.weight", prefix + "norm2.bias", eps, prefix + "mlp.fc1.weight", prefix + "mlp.fc1.bias")
        hc = q_bf16(gelu_fast(hc))
        xc = xc + self.lin(hc, prefix + "mlp.fc2.weight", prefix + "mlp.fc2.bias")
        return torch.cat([xc, x[:, 1:]], dim=1)

    def split_encoder(self, x, taps=None):
        """engine.encode: norm1 of every block except the very first (and the CLS-only one) is folded when ln_fold >= 1, norm2
        of every full block when ln_fold >= 2. taps (list): receives (name, stream) after every block."""
        cfg = self.cfg
        f1, f2 = self.ln_fold >= 1, self.ln_fold >= 2
        split_at = cfg.enc_blocks - cfg.split_blocks
        tag = None
        for i in range(cfg.enc_blocks):
            if i == split_at:
                tag = x
            x = self.vit_block(x, "module.bert.encoder.blocks.%d." % i, fold1=f1 and i > 0, fold2=f2)
            if taps is not None:
                taps.append(("block%d" % i, x))
        for j in range(cfg.split_blocks):
            prefix = "module.bert.encoder.tag_blocks.%d." % j
            if j == cfg.split_blocks - 1 and self.cls_only_last:
                tag = self.vit_block_cls_only(tag, prefix)
            else:
                tag = self.vit_block(tag, prefix, fold1=f1 and (split_at + j) > 0, fold2=f2)
            if taps is not None:
                taps.append(("tag_block%d" % j, tag))
        return x, tag

    def head(self, prefix, h):
        """engine._head (tag head: bf16 operands) / the x3 branch of engine._decode_layers (vocabulary head)."""
        cfg = self.cfg
        x3 = self.decode_x3 and prefix.startswith("module.cls.")
        ln = (cfg.hidden,), self.p(prefix + "transform.LayerNorm.weight"), self.p(prefix + "transform.LayerNorm.bias"), cfg.bert_ln_eps
        if x3:
            t = gelu_fast(self.lin_x3(h, prefix + "transform.dense.weight", prefix + "transform.dense.bias"))
            t = F.layer_norm(t, *ln)
            return self.lin_x3(t, prefix + "decoder.weight", None) + self.p(prefix + "bias")
        t = gelu_fast(self.lin(q_bf16(h), prefix + "transform.dense.weight", prefix + "transform.dense.bias"))
        t = q_bf16(F.layer_norm(t, *ln))
        return self.lin(t, prefix + "decoder.weight", None) + self.p(prefix + "bias")

    def tag_head(self, tag_feats):
        cfg = self.cfg
        pooled = q_bf16(torch.tanh(self.lin(q_bf16(tag_feats[:, 0]), "module.bert.pooler.dense.weight", "module.bert.pooler.dense.bias")))
        logit = self.head("module.bert.tag_logit.predictions.", pooled)
        prob, idx = torch.sigmoid(logit).topk(cfg.topk, dim=1, largest=True)
        return logit, prob, idx, (prob >= cfg.tag_thresh).sum(dim=1)

    # ---- text side ----------------------------------------------------------------------
    def qkv_rows(self, idx, h):
        cfg = self.cfg
        H, d = cfg.heads, cfg.head_dim
        p = "module.bert.decoder.layer.%d.attention.self." % idx
        hq = q_bf16(h)

        def heads(t):
            return q_bf16(t).view(t.shape[0], t.shape[1], H, d).permute(0, 2, 1, 3)
        return tuple(heads(self.lin(hq, p + n + ".weight", p + n + ".bias")) for n in ("query", "key", "value"))

    def bert_layer(self, idx, hq, hkv, add_mask):
        """A prefill layer over the context rows (the decode steps call qkv_rows / bert_layer_from_kv directly)."""
        assert hq is hkv
        cfg = self.cfg
        if idx == 0:
            self._pf_raw = None
        if self.prefill_fold and self._pf_raw is not None and self._pf_raw[0] is hq:
            _, raw, gkey, bekey = self._pf_raw
            H, d = cfg.heads, cfg.head_dim
            p = "module.bert.decoder.layer.%d.attention.self." % idx

            def heads(t):
                return q_bf16(t).view(t.shape[0], t.shape[1], H, d).permute(0, 2, 1, 3)
            q, k, v = (heads(self.lin_fold(raw, gkey, bekey, cfg.bert_ln_eps, p + n + ".weight", p + n + ".bias"))
                       for n in ("query", "key", "value"))
        else:
            q, k, v = self.qkv_rows(idx, hq)
        return self.bert_layer_from_kv(idx, hq, q, k, v, add_mask, prefill=True), k, v

    def bert_layer_from_kv(self, idx, hq, q, k, v, add_mask, step=False, prefill=False):
        """hq: fp32 rows (residual); q / k / v hold bf16 values. step=True: a decode step (split-operand MLP when decode_x3);
        prefill=True with prefill_fold: the intermediate GEMM folds LayerNorm 1 and the raw output rows are kept for the next
        layer's folded q|k|v GEMM."""
        cfg = self.cfg
        p = "module.bert.decoder.layer.%d." % idx
        C = hq.shape[-1]
        ln1 = (p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias")
        ln2 = (p + "output.LayerNorm.weight", p + "output.LayerNorm.bias")
        ctx = self.attend(q, k, v, 1.0 / math.sqrt(cfg.head_dim), add_mask)
        tmp = self.lin(ctx, p + "attention.output.dense.weight", p + "attention.output.dense.bias") + hq
        a = F.layer_norm(tmp, (C,), self.p(ln1[0]), self.p(ln1[1]), cfg.bert_ln_eps)
        if step and self.decode_x3:
            m = gelu_fast(self.lin_x3(a, p + "intermediate.dense.weight", p + "intermediate.dense.bias"))
            m = self.lin_x3(m, p + "output.dense.weight", p + "output.dense.bias")
        else:
            if prefill and self.prefill_fold:
                pre = self.lin_fold(tmp, ln1[0], ln1[1], cfg.bert_ln_eps, p + "intermediate.dense.weight", p + "intermediate.dense.bias")
            else:
                pre = self.lin(q_bf16(a), p + "intermediate.dense.weight", p + "intermediate.dense.bias")
            m = self.lin(q_bf16(gelu_fast(pre)), p + "output.dense.weight", p + "output.dense.bias")
        raw = m + a
        out = F.layer_norm(raw, (C,), self.p(ln2[0]), self.p(ln2[1]), cfg.bert_ln_eps)
        if prefill:
            self._pf_raw = (out, raw, ln2[0], ln2[1])
        return out


# =========================================================================================
# faithful restatement: one full model call per decode step
# =========================================================================================
def construct_full_mask(text_mask, n_img):
    """ImageCaptioning.construct_attn_mask, mask_type='seq2seq' (pipeline file lines 57-85)."""
    B, T, _ = text_mask.shape
    top = torch.cat([text_mask.float(), torch.ones(B, T, n_img)], dim=2)
    bottom = torch.cat([torch.zeros(B, n_img, T), torch.ones(B, n_img, n_img)], dim=2)
    return torch.cat([top, bottom], dim=1)


class FaithfulStepper:
    """Holds what ViTCAP.generate stores on ``self`` (modeling_bert.py:937-1001) and evaluates
    one ``self(**prepare_inputs_for_generation(ids, past=None))`` call per step."""

    def __init__(self, model: PortModel, img_feats, full_mask, od_label_ids, max_len, num_expand,
                 od_labels_start_posid, mask_token_id, add_od_labels=True):
        self.m = model
        cfg = model.cfg

        def expand(x):   # _expand_for_beams, modeling_bert.py:1061-1070
            if num_expand == 1:
                return x
            return x.unsqueeze(1).expand(x.shape[0], num_expand, *x.shape[1:]).contiguous().view(
                x.shape[0] * num_expand, *x.shape[1:])
        self.img_feats = expand(img_feats)
        self.full_mask = expand(full_mask)
        self.od_label_ids = expand(od_label_ids)
        self.max_len = max_len
        self.od_len = od_label_ids.shape[1]
        self.add_od_labels = add_od_labels
        self.mask_token_id = mask_token_id
        start = max(od_labels_start_posid, max_len)                      # modeling_bert.py:959
        pos = torch.arange(max_len)
        if add_od_labels:
            pos = torch.cat([pos, torch.arange(start, start + self.od_len)])
        self.full_pos = pos
        self.last_tag = None
        self.n_calls = 0

    def __call__(self, cur_ids, beam_idx=None):
        m, cfg = self.m, self.m.cfg
        B, L = cur_ids.shape
        ids = torch.cat([cur_ids, torch.full((B, 1), self.mask_token_id, dtype=torch.long)], dim=1)
        cl = L + 1
        keep = torch.cat([torch.arange(0, cl), torch.arange(self.max_len, self.full_mask.shape[1])])
        mask = self.full_mask[:, keep][:, :, keep]                       # _remove_rows_cols
        pos = torch.cat([self.full_pos[:cl], self.full_pos[self.max_len:]]).unsqueeze(0).expand(B, -1)
        if self.add_od_labels:
            ids = torch.cat([ids, self.od_label_ids], dim=1)
        T = ids.shape[1]
        self.n_calls += 1

        # --- ViTSplitCLSEmbModel.forward (modeling_bert.py:1408-1516)
        cap, tag = m.split_encoder(self.img_feats)
        logit, prob, pred_topk, topk_len = m.tag_head(tag)
        self.last_tag = (logit, prob, pred_topk.clone(), topk_len)
        pred_topk = pred_topk.clone()
        word = m.p("module.cls.predictions.decoder.weight")              # cls_emb, tied to word embeddings
        if int(topk_len[0]) + 20 <= T:                                   # modeling_bert.py:1435
            pred_topk[:, -1] = 102
            emb = m.embeddings(ids, pos)
            emb[:, -pred_topk.shape[1]:] = word[pred_topk]               # raw F.embedding, 1456-1470
        else:
            ids = ids.clone()
            start_id = T - topk_len
            ids[:, start_id] = 102                                       # 1474-1476 (column set for all rows)
            pred_topk[:, -1] = 102
            k = pred_topk.shape[1]
            pre = "module.bert.embeddings."
            tpos = torch.arange(k) + 20                                  # encode_tag_to_embedding 1397
            te = word[pred_topk] + m.p(pre + "position_embeddings.weight")[tpos] + \
                m.p(pre + "token_type_embeddings.weight")[0]
            te = F.layer_norm(te, (cfg.hidden,), m.p(pre + "LayerNorm.weight"), m.p(pre + "LayerNorm.bias"),
                              cfg.bert_ln_eps)
            emb = m.embeddings(ids, pos)
            emb[:, -k:] = te
        ctx = torch.cat([tag[:, 0:1], cap], dim=1)                       # 1493
        mask = torch.cat([mask, mask[:, -1:].clone()], dim=1)            # 1494
        mask = torch.cat([mask, torch.ones(B, mask.shape[1], 1)], dim=2)  # 1495
        add = ((1.0 - mask) * NEG_MASK).unsqueeze(1)                     # 1498-1501
        h = torch.cat([emb, ctx], dim=1)
        for l in range(cfg.dec_layers):
            h, _, _ = m.bert_layer(l, h, h, add)
        logits = m.head("module.cls.predictions.", h[:, :T])             # 809-810: all T text rows
        return logits[:, L]                                              # next_token_idx = cur_len


# =========================================================================================
# cached restatement (the algorithm the kernels implement)
# =========================================================================================
class CachedStepper:
    """``n_label`` (int64 [B] or None): number of visible od/tag label slots per image, i.e. the mask family
    dataset.py:371-417 produces for a non-empty ``text_b`` (full L-L block, caption rows see L, image rows do not).
    The label rows then join the context: rows [tag-CLS | image | L_0..L_49]; image rows never see them, label rows
    see the image context and the first n_label[b] label rows, caption rows see all of that.

    L_i comes from one of two recipes, chosen PER STEP by the reference from the first sample's tag count
    (modeling_bert.py:1435, ``topk_len[0] + 20 <= input_ids.shape[1]`` with input length cur_len + 1 + 50):
      'raw' (1447-1470): word embedding of the i-th predicted tag, last slot forced to [SEP];
      'ln'  (1472-1489 -> encode_tag_to_embedding 1381-1406): the same + position 20+i + type 0 -> LayerNorm.
    The reference recomputes everything every step, so when the recipe flips at some step every cached row that saw the
    label rows is stale: the label rows are prefilled again and the caption rows replayed under the new recipe."""

    def __init__(self, model: PortModel, img_feats, num_expand, mask_token_id, keep_intermediates=False, n_label=None):
        self.m = model
        cfg = model.cfg
        self.E = num_expand
        self.mask_token_id = mask_token_id
        cap, tag = model.split_encoder(img_feats)
        self.cap, self.tag = cap, tag
        self.last_tag = model.tag_head(tag)
        self.ctx_img = torch.cat([tag[:, 0:1], cap], dim=1)
        self.n_label = None
        self.recipe = None
        if n_label is not None and int(n_label.max()) > 0:
            self.n_label = n_label.clone()
        else:
            self._prefill(None)
        self.n_calls = 0
        self.n_flips = 0

    def label_recipe(self, cur_len):
        """Which label embedding the reference uses at the step whose input holds cur_len tokens + [MASK]."""
        t0 = int(self.last_tag[3][0])
        return "raw" if t0 + 20 <= cur_len + 1 + self.m.cfg.topk else "ln"

    def _prefill(self, recipe):
        m, cfg = self.m, self.m.cfg
        ctx = self.ctx_img
        add = None
        if recipe is not None:
            B, C = ctx.shape[0], ctx.shape[1]
            pred_topk = self.last_tag[2].clone()
            pred_topk[:, -1] = 102                                         # modeling_bert.py:1447 / 1477
            lab = m.p("module.cls.predictions.decoder.weight")[pred_topk]   # cls_emb == tied word embeddings (:766)
            if recipe == "ln":
                pre = "module.bert.embeddings."
                tpos = torch.arange(cfg.topk) + 20                         # encode_tag_to_embedding, :1397
                lab = lab + m.p(pre + "position_embeddings.weight")[tpos] + m.p(pre + "token_type_embeddings.weight")[0]
                lab = F.layer_norm(lab, (cfg.hidden,), m.p(pre + "LayerNorm.weight"), m.p(pre + "LayerNorm.bias"),
                                   cfg.bert_ln_eps)
            ctx = torch.cat([ctx, lab], dim=1)
            S = ctx.shape[1]
            vis = torch.zeros(B, S, S)
            vis[:, :, :C] = 1.0
            for b in range(B):
                vis[b, C:, C:C + int(self.n_label[b])] = 1.0
            add = ((1.0 - vis) * NEG_MASK).unsqueeze(1)
            key_vis = torch.cat([torch.ones(B, C), (torch.arange(S - C).unsqueeze(0) < self.n_label.unsqueeze(1)).float()], 1)
            self.key_add = (1.0 - key_vis) * NEG_MASK
        self.recipe = recipe
        self.Kc, self.Vc = [], []
        for l in range(cfg.dec_layers):
            ctx_out, k, v = m.bert_layer(l, ctx, ctx, add)
            self.Kc.append(k)
            self.Vc.append(v)
            ctx = ctx_out
        self.Kt = [None] * cfg.dec_layers
        self.Vt = [None] * cfg.dec_layers

    def __call__(self, cur_ids, beam_idx=None):
        m, cfg = self.m, self.m.cfg
        R, L = cur_ids.shape                       # R = B * E rows
        self.n_calls += 1
        if self.n_label is not None:
            want = self.label_recipe(L)
            if want != self.recipe:
                self.n_flips += self.recipe is not None
                self._prefill(want)
                for s in range(1, L):              # replay the caption rows under the new label rows
                    self._step(cur_ids[:, :s])
                beam_idx = None                    # cur_ids already holds the re-parented histories
        return self._step(cur_ids, beam_idx)

    def _step(self, cur_ids, beam_idx=None):
        m, cfg = self.m, self.m.cfg
        R, L = cur_ids.shape
        E = self.E
        if beam_idx is not None:                   # reorder caption K/V rows (modeling_utils.py:1055-1065)
            for l in range(cfg.dec_layers):
                if self.Kt[l] is not None:
                    self.Kt[l] = self.Kt[l][beam_idx]
                    self.Vt[l] = self.Vt[l][beam_idx]
        inp = torch.stack([cur_ids[:, -1], torch.full((R,), self.mask_token_id, dtype=torch.long)], dim=1)
        pos = torch.tensor([L - 1, L]).unsqueeze(0).expand(R, -1)
        e = m.embeddings(inp, pos)
        for l in range(cfg.dec_layers):
            q, k, v = m.qkv_rows(l, e)
            Kc = self.Kc[l].repeat_interleave(E, dim=0) if E > 1 else self.Kc[l]
            Vc = self.Vc[l].repeat_interleave(E, dim=0) if E > 1 else self.Vc[l]
            parts_k = [Kc] + ([self.Kt[l]] if self.Kt[l] is not None else []) + [k]
            parts_v = [Vc] + ([self.Vt[l]] if self.Vt[l] is not None else []) + [v]
            K = torch.cat(parts_k, dim=2)
            V = torch.cat(parts_v, dim=2)
            add = torch.zeros(1, 1, 2, K.shape[2])
            add[..., 0, -1] = NEG_MASK             # the real-token row must not see the MASK row
            if self.recipe is not None:            # invisible label slots of the image's context
                ka = self.key_add.repeat_interleave(E, dim=0) if E > 1 else self.key_add
                add = add.repeat(R, 1, 1, 1)
                add[:, 0, :, :ka.shape[1]] += ka.unsqueeze(1)
            e = m.bert_layer_from_kv(l, e, q, K, V, add, step=True)
            self.Kt[l] = k[:, :, 0:1] if self.Kt[l] is None else torch.cat([self.Kt[l], k[:, :, 0:1]], dim=2)
            self.Vt[l] = v[:, :, 0:1] if self.Vt[l] is None else torch.cat([self.Vt[l], v[:, :, 0:1]], dim=2)
        return m.head("module.cls.predictions.", e[:, 1])


# =========================================================================================
# search loops (restating modeling_utils.py:768-1180)
# =========================================================================================
def greedy_or_sample(step, batch, max_length, bos, pad, eos_ids, do_sample=False, temperature=1.0,
                     top_k=0, top_p=1.0, sampler=None, trace=None):
    """_generate_no_beam_search, modeling_utils.py:768-886. ``sampler(logits, step_idx)`` returns the
    sampled token per row (default torch.multinomial, as the reference)."""
    ids = torch.full((batch, 1), bos, dtype=torch.long)
    unfinished = torch.ones(batch, dtype=torch.long)
    logprobs, unf = [], []
    cur_len = 1
    while cur_len < max_length:
        logits = step(ids)
        if do_sample:
            if temperature != 1.0:
                logits = logits / temperature
            logits = top_k_top_p_filtering(logits, top_k=top_k, top_p=top_p)
            if sampler is None:
                nxt = torch.multinomial(F.softmax(logits, dim=-1), num_samples=1).squeeze(1)
            else:
                nxt = sampler(logits, cur_len)
        else:
            nxt = torch.argmax(logits, dim=-1)
        sc = torch.gather(F.log_softmax(logits, dim=-1), -1, nxt.unsqueeze(-1))
        if trace is not None:
            trace.append(logits.clone())
        logprobs.append(sc)
        unf.append(unfinished)
        tok = nxt * unfinished + pad * (1 - unfinished)
        ids = torch.cat([ids, tok.unsqueeze(-1)], dim=-1)
        for e in eos_ids:
            unfinished = unfinished.mul(tok.ne(e).long())
        cur_len += 1
        if unfinished.max() == 0:
            break
    if cur_len == max_length:
        ids[:, -1].masked_fill_(unfinished.to(torch.bool), eos_ids[0])
    logprobs = torch.cat(logprobs, dim=1)
    unf = torch.stack(unf, dim=1).float()
    lp = (logprobs * unf).sum(dim=1) / unf.sum(dim=1)
    if max_length - ids.shape[1] > 0:
        ids = torch.cat([ids, ids.new_full((batch, max_length - ids.shape[1]), pad)], dim=1)
    return ids.unsqueeze(1), lp.unsqueeze(1)


def top_k_top_p_filtering(logits, top_k=0, top_p=1.0, filter_value=-float("inf"), min_tokens_to_keep=1):
    """modeling_utils.py:1103-1135."""
    if top_k > 0:
        top_k = min(max(top_k, min_tokens_to_keep), logits.size(-1))
        kth = torch.topk(logits, top_k)[0][..., -1, None]
        logits = logits.masked_fill(logits < kth, filter_value)
    if top_p < 1.0:
        sl, si = torch.sort(logits, descending=True)
        cp = torch.cumsum(F.softmax(sl, dim=-1), dim=-1)
        rm = cp > top_p
        if min_tokens_to_keep > 1:
            rm[..., :min_tokens_to_keep] = 0
        rm[..., 1:] = rm[..., :-1].clone()
        rm[..., 0] = 0
        rm = rm.scatter(1, si, rm)
        logits = logits.masked_fill(rm, filter_value)
    return logits


class _Hyps:
    """BeamHypotheses, modeling_utils.py:1138-1180 (early_stopping=False)."""

    def __init__(self, n_hyp, max_length, length_penalty):
        self.max_length = max_length - 1
        self.lp = length_penalty
        self.n = n_hyp
        self.hyp = []
        self.worst = 1e9

    def add(self, hyp, s):
        score = s / len(hyp) ** self.lp
        if len(self.hyp) < self.n or score > self.worst:
            self.hyp.append((score, hyp))
            if len(self.hyp) > self.n:
                ss = sorted([(sc, i) for i, (sc, _) in enumerate(self.hyp)])
                del self.hyp[ss[0][1]]
                self.worst = ss[1][0]
            else:
                self.worst = min(score, self.worst)

    def is_done(self, best):
        if len(self.hyp) < self.n:
            return False
        return self.worst >= best / self.max_length ** self.lp


def beam_search(step, batch, max_length, bos, pad, eos_ids, num_beams, vocab, length_penalty=1.0,
                num_keep_best=1, trace=None):
    """_generate_beam_search, modeling_utils.py:888-1100 (do_sample=False)."""
    ids = torch.full((batch * num_beams, 1), bos, dtype=torch.long)
    hyps = [_Hyps(num_keep_best, max_length, length_penalty) for _ in range(batch)]
    beam_scores = torch.zeros(batch, num_beams)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    done = [False] * batch
    cur_len = 1
    beam_idx = None
    while cur_len < max_length:
        scores = step(ids, beam_idx)
        if trace is not None:
            trace.append(scores.clone())
        scores = F.log_softmax(scores, dim=-1)
        _s = (scores + beam_scores[:, None]).view(batch, num_beams * vocab)
        nscores, nwords = torch.topk(_s, 2 * num_beams, dim=1, largest=True, sorted=True)
        nxt = []
        for b in range(batch):
            done[b] = done[b] or hyps[b].is_done(nscores[b].max().item())
            if done[b]:
                nxt.extend([(0, pad, 0)] * num_beams)
                continue
            sent = []
            for idx, sc in zip(nwords[b], nscores[b]):
                beam_id = int(idx) // vocab
                word_id = int(idx) % vocab
                if word_id in eos_ids or cur_len + 1 == max_length:
                    hyps[b].add(ids[b * num_beams + beam_id, :cur_len].clone(), sc.item())
                else:
                    sent.append((sc, word_id, b * num_beams + beam_id))
                if len(sent) == num_beams:
                    break
            if len(sent) == 0:
                sent = [(0, pad, 0)] * num_beams
            nxt.extend(sent)
        beam_scores = beam_scores.new_tensor([float(x[0]) for x in nxt])
        words = ids.new_tensor([x[1] for x in nxt])
        beam_idx = ids.new_tensor([x[2] for x in nxt])
        ids = torch.cat([ids[beam_idx, :], words.unsqueeze(1)], dim=-1)
        cur_len += 1
        if all(done):
            break
    logprobs = torch.full((batch, num_keep_best), -1e5)
    decoded = ids.new_full((batch, num_keep_best, max_length), pad)
    for i, h in enumerate(hyps):
        hs = torch.tensor([x[0] for x in h.hyp])
        _, best = torch.topk(hs, min(num_keep_best, len(hs)), largest=True)
        for j, hi in enumerate(best):
            conf, hyp = h.hyp[hi]
            logprobs[i, j] = conf
            decoded[i, j, :len(hyp)] = hyp
            decoded[i, j, len(hyp)] = eos_ids[0]
    return decoded, logprobs


# =========================================================================================
# entry points
# =========================================================================================
def _canonical(text_mask, max_seq_a):
    """True iff the 70x70 mask is the eval pipeline's: caption triangle only (dataset.py:371-390 with
    text_b == '')."""
    n = label_counts(text_mask, max_seq_a)
    return n is not None and int(n.max()) == 0


def label_counts(text_mask, max_seq_a):
    """Number of visible label slots per sample if the mask belongs to the seq2seq family of dataset.py:371-417
    (caption triangle; full attention L-L and C-L over the first n label slots; nothing else), else None."""
    B, S, _ = text_mask.shape
    n = text_mask[:, 0, max_seq_a:].sum(dim=1).long()
    ref = torch.zeros(B, S, S, dtype=text_mask.dtype)
    ref[:, :max_seq_a, :max_seq_a] = torch.tril(torch.ones(max_seq_a, max_seq_a, dtype=text_mask.dtype))
    for b in range(B):
        e = max_seq_a + int(n[b])
        ref[b, max_seq_a:e, max_seq_a:e] = 1
        ref[b, :max_seq_a, max_seq_a:e] = 1
    return n if bool((text_mask == ref).all()) else None


def caption(model: PortModel, data, extra, algorithm="cached", sampler=None, trace=None, info=None):
    """ImageCaptioning.forward eval branch (pipeline file lines 87-112, 173-184) ->
    ViTCAP.generate (modeling_bert.py:928-1059). Returns (ids (B*K, keep, max_len) int64,
    logprobs (B*K, keep) fp32)."""
    cfg = model.cfg
    image = data["image"]
    B = image.shape[0]
    max_length = extra["max_length"]
    nb, nret = extra["num_beams"], extra["num_return_sequences"]
    assert extra.get("repetition_penalty", 1) == 1
    img_feats = model.patch_embed(image)
    num_expand = nb * nret
    if algorithm == "faithful":
        full_mask = construct_full_mask(data["attention_mask"], img_feats.shape[1])
        od = data["input_ids"][:, max_length:]
        stepper = FaithfulStepper(model, img_feats, full_mask, od, max_length, num_expand,
                                  extra["od_labels_start_posid"], extra["mask_token_id"],
                                  add_od_labels=extra.get("add_od_labels", True))
    else:
        n_label = label_counts(data["attention_mask"], cfg.max_seq_a)
        assert n_label is not None, "cached algorithm needs a mask of the seq2seq family (dataset.py:371-417)"
        stepper = CachedStepper(model, img_feats, num_expand, extra["mask_token_id"], n_label=n_label)
    eff = B * nret
    if nb > 1:
        out = beam_search(stepper, eff, max_length, extra["bos_token_id"], extra["pad_token_id"],
                          extra["eos_token_ids"], nb, cfg.vocab, extra["length_penalty"], extra["num_keep_best"],
                          trace=trace)
    else:
        out = greedy_or_sample(stepper, eff, max_length, extra["bos_token_id"], extra["pad_token_id"],
                               extra["eos_token_ids"], extra["do_sample"], extra["temperature"], extra["top_k"],
                               extra["top_p"], sampler=sampler, trace=trace)
    if info is not None:
        info["n_calls"] = stepper.n_calls
        info["tag"] = stepper.last_tag
        info["img_feats"] = img_feats
        if algorithm == "cached":
            info["cap"], info["tag_feats"] = stepper.cap, stepper.tag
            info["Kc"], info["Vc"] = stepper.Kc, stepper.Vc
    return out


def encode_tags(model: PortModel, image):
    """BASELINE config 2: patch embed -> 8 shared + 4 caption + 4 tag blocks -> tag head -> top-k."""
    x = model.patch_embed(image)
    cap, tag = model.split_encoder(x)
    logit, prob, idx, n = model.tag_head(tag)
    return cap, tag, logit, prob, idx, n
