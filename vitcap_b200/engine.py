"""Device-side execution plan of the caption path: owns packed weights and workspaces in HBM and issues the
C-ABI kernels (vitcap_b200.ops) in the order of SURVEY.md section 7.1:

    patch embed -> 8 shared ViT blocks -> {4 caption blocks, 4 tag blocks} -> tag head + top-k
    -> context = [tag-CLS | caption features] prefilled ONCE through the decoder (K/V cached per layer)
    -> 19 decode steps of 2 rows per sequence (greedy / sampling / beam), captured in a CUDA graph.

What the reference recomputes every step (all 16 ViT blocks, the tag head, 50 dead od/tag slots and the decoder
over all ~648 rows; modeling_bert.py:845-876 with past=None) is computed exactly once here.

HBM layout (T = bf16 in fast mode, fp32 in exact mode; all row-major, rows = tokens):
    x, xt        fp32 [B*N, H]       residual streams of the caption / tag branches (never rounded to bf16)
    ln, att      T    [B*C, H]       LayerNorm output (GEMM A operand), attention output
    qkv          T    [B*N, 3H]      packed q|k|v of the current ViT block
    hid          T    [B*C, F]       MLP hidden
    ctx_qkv      T    [L, B*C, 3H]   prefill q|k|v of the 578 context rows per decoder layer == the KV cache
    step_qkv     T    [L, max_len, 2R, 3H]  q|k|v rows of every decode step == caption-token KV cache
    logits       fp32 [R, ldl]       vocabulary logits of the MASK rows (ldl = vocab rounded up to 64)
"""
import math
import os
import warnings
from collections import OrderedDict

import torch

from . import ops
from .config import VitCapConfig


def _round_up(a, b):
    return (a + b - 1) // b * b


HALF_MAX = 65504.0


def half_range_bounds(cfg: VitCapConfig, sd):
    """Worst-case magnitudes of everything decode_precision='fp16' stores as an IEEE half, from the weights alone:
      * the four weight matrices themselves (BertIntermediate / BertOutput of every decoder layer, head transform, vocabulary);
      * LayerNorm outputs (operands of fc1 and of the head GEMMs): |y_j| <= |gamma_j| sqrt(H - 1) + |beta_j|;
      * GELU outputs of fc1 (operand of fc2): |gelu(x)| <= |x| <= sum_j |W_ij| |y_j| + |b_i|  with the bound on y above.
    Returns {'weights': ..., 'layernorm': ..., 'gelu': ...}; a value below 65504 PROVES that the conversion never saturates for
    any input. (The conversions saturate instead of producing infinities in any case.)"""
    H = cfg.hidden
    root = math.sqrt(H - 1.0)

    def f(key):
        return sd[key].detach().float()

    def ln_bound(prefix):
        return f(prefix + "weight").abs() * root + f(prefix + "bias").abs()

    w_max, ln_max, gelu_max = 0.0, 0.0, 0.0
    for i in range(cfg.dec_layers):
        p = "module.bert.decoder.layer.%d." % i
        y1 = ln_bound(p + "attention.output.LayerNorm.")
        y2 = ln_bound(p + "output.LayerNorm.")
        wi, wo = f(p + "intermediate.dense.weight"), f(p + "output.dense.weight")
        w_max = max(w_max, float(wi.abs().max()), float(wo.abs().max()))
        ln_max = max(ln_max, float(y1.max()), float(y2.max()))
        gelu_max = max(gelu_max, float((wi.abs() @ y1 + f(p + "intermediate.dense.bias").abs()).max()))
    h = "module.cls.predictions."
    w_max = max(w_max, float(f(h + "transform.dense.weight").abs().max()), float(f(h + "decoder.weight").abs().max()))
    ln_max = max(ln_max, float(ln_bound(h + "transform.LayerNorm.").max()))
    return {"weights": w_max, "layernorm": ln_max, "gelu": gelu_max}


class PackedWeights:
    """Weights re-laid for the kernels from a reference-layout state_dict (fp32 masters stay in the nn.Module)."""

    def __init__(self, cfg: VitCapConfig, sd, mode, device, decode_x3=False, decode_f16=False):
        self.cfg = cfg
        self.mode = mode
        # split-bf16 copies [w_hi | w_hi | w_lo] of the decode-step MLP and vocabulary-head weights (see _decode_layers)
        self.decode_x3 = bool(decode_x3) and mode == "bf16"
        # IEEE-half copies of the same weights (decode_precision='fp16': one product on 11-bit significands)
        self.decode_f16 = bool(decode_f16) and mode == "bf16"
        assert not (self.decode_x3 and self.decode_f16)
        wt = ops.STORE if mode == "bf16" else torch.float32        # (IEEE halves in a VITCAP_STORE=fp16 process)
        self.wt = wt

        def W(key):
            return sd[key].detach().to(device=device, dtype=wt).contiguous()

        def Fp(key):
            return sd[key].detach().to(device=device, dtype=torch.float32).contiguous()

        def Hf(key):
            return Fp(key).clamp(-HALF_MAX, HALF_MAX).to(torch.float16).contiguous()      # (out of place: Fp may alias the master)

        H = cfg.hidden
        ie = "image_encoder.module."
        self.patch_w = sd[ie + "patch_embed.proj.weight"].detach().reshape(H, cfg.patch_dim).to(device=device, dtype=wt).contiguous()
        self.patch_b = Fp(ie + "patch_embed.proj.bias")
        self.cls_token = Fp(ie + "cls_token").reshape(H).contiguous()
        self.pos_embed = Fp(ie + "pos_embed").reshape(cfg.n_tokens, H).contiguous()

        def folded_t(w32, b32, g32, be32):
            """LayerNorm folded into the consuming Linear (vc_linear_ln_fold): Wf = bf16(gamma o W), colsum = fp32 row sums of
            the ROUNDED Wf (so that mean * colsum cancels what the tensor cores accumulate), bias_f = b + W beta."""
            wf = (w32 * g32.unsqueeze(0)).to(ops.STORE).contiguous()
            return wf, wf.float().sum(1).contiguous(), (b32 + w32 @ be32).contiguous()

        def folded(w_key, b_key, g_key, beta_key):
            return folded_t(Fp(w_key), Fp(b_key), Fp(g_key), Fp(beta_key))

        def block(prefix):
            d = {
                "n1w": Fp(prefix + "norm1.weight"), "n1b": Fp(prefix + "norm1.bias"),
                "qkv_w": W(prefix + "attn.qkv.weight"), "qkv_b": Fp(prefix + "attn.qkv.bias"),
                "proj_w": W(prefix + "attn.proj.weight"), "proj_b": Fp(prefix + "attn.proj.bias"),
                "n2w": Fp(prefix + "norm2.weight"), "n2b": Fp(prefix + "norm2.bias"),
                "fc1_w": W(prefix + "mlp.fc1.weight"), "fc1_b": Fp(prefix + "mlp.fc1.bias"),
                "fc2_w": W(prefix + "mlp.fc2.weight"), "fc2_b": Fp(prefix + "mlp.fc2.bias"),
            }
            if mode == "bf16":
                d["qkv_wf"], d["qkv_cf"], d["qkv_bf"] = folded(prefix + "attn.qkv.weight", prefix + "attn.qkv.bias",
                                                               prefix + "norm1.weight", prefix + "norm1.bias")
                d["fc1_wf"], d["fc1_cf"], d["fc1_bf"] = folded(prefix + "mlp.fc1.weight", prefix + "mlp.fc1.bias",
                                                               prefix + "norm2.weight", prefix + "norm2.bias")
            return d
        self.blocks = [block("module.bert.encoder.blocks.%d." % i) for i in range(cfg.enc_blocks)]
        self.tag_blocks = [block("module.bert.encoder.tag_blocks.%d." % i) for i in range(cfg.split_blocks)]

        self.pool_w, self.pool_b = W("module.bert.pooler.dense.weight"), Fp("module.bert.pooler.dense.bias")

        def head(prefix):
            return {
                "t_w": W(prefix + "transform.dense.weight"), "t_b": Fp(prefix + "transform.dense.bias"),
                "ln_w": Fp(prefix + "transform.LayerNorm.weight"), "ln_b": Fp(prefix + "transform.LayerNorm.bias"),
                "dec_w": W(prefix + "decoder.weight"), "bias": Fp(prefix + "bias"),
            }
        self.tag_head = head("module.bert.tag_logit.predictions.")
        self.cls_head = head("module.cls.predictions.")
        if self.decode_x3:
            self.cls_head["t_w3"] = ops.split_weight_bf16x3(Fp("module.cls.predictions.transform.dense.weight"))
            self.cls_head["dec_w3"] = ops.split_weight_bf16x3(Fp("module.cls.predictions.decoder.weight"))
        if self.decode_f16:
            # range proof of the half operands from the weights (half_range_bounds): weights beyond the half range are refused,
            # activations that COULD saturate (no checkpoint of this model family comes close) are reported once
            self.half_range = half_range_bounds(cfg, sd)
            if self.half_range["weights"] > HALF_MAX:
                raise ValueError("decode_precision='fp16': a decode-step weight of magnitude %.3g does not fit an IEEE half; use "
                                 "decode_precision='bf16x3'" % self.half_range["weights"])
            if max(self.half_range["layernorm"], self.half_range["gelu"]) > HALF_MAX:
                warnings.warn("decode_precision='fp16': the weights allow LayerNorm / GELU outputs up to %.3g / %.3g, beyond the IEEE-half "
                              "range (65504): such values would saturate. decode_precision='bf16x3' has bf16's range."
                              % (self.half_range["layernorm"], self.half_range["gelu"]))
            self.cls_head["t_wh"] = Hf("module.cls.predictions.transform.dense.weight")
            self.cls_head["dec_wh"] = Hf("module.cls.predictions.decoder.weight")

        e = "module.bert.embeddings."
        self.word = Fp(e + "word_embeddings.weight")
        self.pos = Fp(e + "position_embeddings.weight")
        self.type0 = Fp(e + "token_type_embeddings.weight")[0].contiguous()
        self.emb_ln_w, self.emb_ln_b = Fp(e + "LayerNorm.weight"), Fp(e + "LayerNorm.bias")

        self.dec = []
        for i in range(cfg.dec_layers):
            p = "module.bert.decoder.layer.%d." % i
            qkv_w = torch.cat([sd[p + "attention.self.%s.weight" % n].detach() for n in ("query", "key", "value")], 0)
            qkv_b = torch.cat([sd[p + "attention.self.%s.bias" % n].detach() for n in ("query", "key", "value")], 0)
            self.dec.append({
                "qkv_w": qkv_w.to(device=device, dtype=wt).contiguous(),
                "qkv_b": qkv_b.to(device=device, dtype=torch.float32).contiguous(),
                "o_w": W(p + "attention.output.dense.weight"), "o_b": Fp(p + "attention.output.dense.bias"),
                "ln1_w": Fp(p + "attention.output.LayerNorm.weight"), "ln1_b": Fp(p + "attention.output.LayerNorm.bias"),
                "i_w": W(p + "intermediate.dense.weight"), "i_b": Fp(p + "intermediate.dense.bias"),
                "f_w": W(p + "output.dense.weight"), "f_b": Fp(p + "output.dense.bias"),
                "ln2_w": Fp(p + "output.LayerNorm.weight"), "ln2_b": Fp(p + "output.LayerNorm.bias"),
            })
            if self.decode_x3:
                self.dec[-1]["i_w3"] = ops.split_weight_bf16x3(Fp(p + "intermediate.dense.weight"))
                self.dec[-1]["f_w3"] = ops.split_weight_bf16x3(Fp(p + "output.dense.weight"))
            if self.decode_f16:
                self.dec[-1]["i_wh"] = Hf(p + "intermediate.dense.weight")
                self.dec[-1]["f_wh"] = Hf(p + "output.dense.weight")
            if mode == "bf16":
                # prefill with folded LayerNorms (engine.prefill): the intermediate GEMM folds this layer's attention-output
                # LayerNorm, the q|k|v GEMM of layer i >= 1 folds the output LayerNorm of layer i - 1
                d = self.dec[-1]
                d["i_wf"], d["i_cf"], d["i_bf"] = folded(p + "intermediate.dense.weight", p + "intermediate.dense.bias",
                                                         p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias")
                if i > 0:
                    pp = "module.bert.decoder.layer.%d." % (i - 1)
                    d["qkv_wf"], d["qkv_cf"], d["qkv_bf"] = folded_t(qkv_w.to(device=device, dtype=torch.float32),
                                                                     qkv_b.to(device=device, dtype=torch.float32),
                                                                     Fp(pp + "output.LayerNorm.weight"), Fp(pp + "output.LayerNorm.bias"))


class CaptionEngine:
    def __init__(self, cfg: VitCapConfig, weights: PackedWeights, device, use_cuda_graph=True):
        self.cfg = cfg
        self.w = weights
        self.mode = weights.mode
        self.T = weights.wt
        self.dev = device
        self.use_cuda_graph = use_cuda_graph
        self.ldl = _round_up(cfg.vocab, 64)
        self._enc_ws = None
        self._dec_ws = OrderedDict()           # (B, E, max_len) -> decode workspace with its captured graphs, LRU
        self.max_decode_workspaces = 3         # e.g. SCST alternates a greedy (E = 1) and a sampling (E = 5) pass per batch
        self.attn_impl = "auto"
        self.stats = {}
        # norm1 of a ViT block folded into the fc2 GEMM before it (statistics + bf16 row copy) and the qkv GEMM after it
        # (gemm_tc2.cu, LN = 1 / 2): one 1.36 GB LayerNorm pass less per block. Fast mode only
        # (environment VITCAP_LN_FOLD = 0 / 1 / 2 selects none / norm1 / norm1 + norm2 for A/B measurements; default 2)
        level = int(os.environ.get("VITCAP_LN_FOLD", "2")) if self.mode == "bf16" else 0
        self.ln_fold = level >= 1
        self.ln_fold2 = level >= 2             # the same for norm2: the proj GEMM emits, the fc1 + GELU GEMM folds
        # decode-step MLP and vocabulary head on split-bf16 operands (three tensor-core products, ~fp32 operand precision)
        self.decode_x3 = weights.decode_x3
        # ... or on IEEE-half operands (one product, 11-bit significands; fused decode step only)
        self.decode_f16 = weights.decode_f16
        # fc2 of a decode step through vc_linear_x3 (distinct tiles loaded once) instead of the K-concatenated plain GEMM
        # (VITCAP_X3_DEDUP=0 for A/B measurements)
        self.x3_dedup = os.environ.get("VITCAP_X3_DEDUP", "1") != "0"
        # fused decode step (fast mode): the step's Linear layers on the CTA-pair decode kernels (gemm_dec.cu: split-K partial
        # planes, split-operand epilogues, vocabulary arg-max partials) with bias / residual / LayerNorm / operand split in the
        # row-wise finish kernel (decode_rowwise.cu). VITCAP_FUSED_DECODE=0 restores the round-1 kernel sequence (A/B runs)
        self.fused_decode = self.mode == "bf16" and os.environ.get("VITCAP_FUSED_DECODE", "1") != "0"
        if self.decode_f16 and not self.fused_decode:
            raise NotImplementedError("decode_precision='fp16' exists on the fused decode step only (VITCAP_FUSED_DECODE=0 is set)")
        # split-K factors of its partial-plane GEMMs (o-proj, fc2, head transform). CONSTANTS, not functions of the batch: the
        # summation order of a row must not depend on how many other rows are in flight (every image's result is bit-identical
        # in any batch, tests/test_fullsize_gpu.py). VITCAP_DEC_SPLITS="o,f,t" overrides them for tuning runs
        self.dec_splits = tuple(int(v) for v in os.environ.get("VITCAP_DEC_SPLITS", "3,6,6").split(","))
        # decoder prefill with every LayerNorm folded into the GEMMs around it (post-LN variant of the ViT fold; prefill())
        self.prefill_fold = self.mode == "bf16" and os.environ.get("VITCAP_PREFILL_FOLD", "1") != "0"
        # early exit of the captured decode loops (modeling_utils.py:865-867, 1071-1073): every step after the first is the body
        # of a conditional graph node keyed on "any caption unfinished", evaluated on the device (ops.graph_if_any)
        self.early_exit = os.environ.get("VITCAP_EARLY_EXIT", "1") != "0"
        self._body_stream = None
        self.forward_graphs = {}               # whole-forward CUDA graphs of the small-batch latency path (model.py); they hold
        self.inline_graphs = False             # raw workspace pointers. inline_graphs: an outer capture is running
        # parity instrumentation (tests / tools only): tap(name, index, tensor) is called with the stream after every ViT block
        # and with the vocabulary logits after every decode step; while it is set the decode loop runs eagerly (no graph)
        self.tap = None

    # ------------------------------------------------------------------ workspaces
    def _alloc(self, *shape, dtype=None):
        return torch.empty(*shape, device=self.dev, dtype=dtype or self.T)

    def _encoder_ws(self, B, ctx_rows=None):
        """ctx_rows: context rows per image (C = tag-CLS + image tokens; C + topk when the label rows are visible)."""
        ws = self._enc_ws
        cfg = self.cfg
        ctx_rows = cfg.n_ctx if ctx_rows is None else ctx_rows
        if ws is not None and ws["B"] >= B and ws["ctx_cap"] >= B * ctx_rows:
            return ws
        if ws is not None:                      # grow, never shrink (a captured graph keeps raw pointers)
            B = max(B, ws["B"])
            ctx_rows = max(ctx_rows, -(-ws["ctx_cap"] // B))
        N, C, H, F, L = cfg.n_tokens, ctx_rows, cfg.hidden, cfg.inter, cfg.dec_layers
        f32 = torch.float32
        ws = {"B": B, "ctx_cap": B * ctx_rows}
        ws["patches"] = self._alloc(B * cfg.n_patches, cfg.patch_dim)
        ws["patch_out"] = self._alloc(B * cfg.n_patches, H, dtype=f32)
        ws["x"] = self._alloc(B * N, H, dtype=f32)
        ws["xt"] = self._alloc(B * N, H, dtype=f32)
        ws["ln"] = self._alloc(B * C, H)
        ws["qkv"] = self._alloc(B * N, 3 * H)
        ws["att"] = self._alloc(B * C, H)
        ws["hid"] = self._alloc(B * C, F)
        ws["ctx_f"] = self._alloc(B * C, H, dtype=f32)
        ws["ctx_t"] = ws["ctx_f"] if self.T == f32 else self._alloc(B * C, H)
        ws["tmp_f"] = self._alloc(B * C, H, dtype=f32)
        ws["a_f"] = self._alloc(B * C, H, dtype=f32)
        ws["ctx_qkv"] = self._alloc(L, B * C, 3 * H)
        if self.mode == "bf16":
            st = (H + 255) // 256                  # per-256-column (sum, sum of squares) of the two raw prefill streams
            ws["pf_st"] = (self._alloc(B * C, st, 2, dtype=f32), self._alloc(B * C, st, 2, dtype=f32), st)
        # tag head
        ws["cls_t"] = self._alloc(B, H)
        ws["pooled"] = self._alloc(B, H, dtype=f32 if self.T == f32 else self.T)
        ws["th_f"] = self._alloc(B, H, dtype=f32)
        ws["th_t"] = self._alloc(B, H)
        ws["tag_logits"] = self._alloc(B, self.ldl, dtype=f32)
        ws["tag_idx"] = self._alloc(B, cfg.topk, dtype=torch.int32)
        ws["tag_prob"] = self._alloc(B, cfg.topk, dtype=f32)
        ws["tag_len"] = self._alloc(B, dtype=torch.int32)
        ws["n_label"] = torch.zeros(B, device=self.dev, dtype=torch.int32)
        ws["ctx_vis"] = torch.zeros(B, device=self.dev, dtype=torch.int32)
        # last concept-branch block, CLS row only
        ws["q_cls"] = self._alloc(B, H)
        ws["att_cls"] = self._alloc(B, H)
        ws["ln_cls"] = self._alloc(B, H)
        ws["hid_cls"] = self._alloc(B, F)
        if self.ln_fold:
            # raw bf16 copy + per-256-column (sum, sum of squares) of the caption / concept stream, written by the fc2 GEMM
            st = (H + 255) // 256
            ws["fold_x"] = (self._alloc(B * N, H), self._alloc(B * N, st, 2, dtype=f32), st)
            ws["fold_t"] = (self._alloc(B * N, H), self._alloc(B * N, st, 2, dtype=f32), st)
            ws["fold_2"] = (self._alloc(B * N, H), self._alloc(B * N, st, 2, dtype=f32), st)     # norm2, consumed inside the block
        self._enc_ws = ws
        self._dec_ws.clear()                   # captured graphs hold raw pointers into the old image-side workspace
        self.forward_graphs.clear()
        return ws

    def _new_decoder_ws(self, B, E, max_len):
        """Buffers of one decode loop over B images x E sequences."""
        cfg = self.cfg
        H, F, L = cfg.hidden, cfg.inter, cfg.dec_layers
        R = B * E
        f32, i32 = torch.float32, torch.int32
        ws = {"B": B, "E": E, "R": R, "max_len": max_len}
        ws["e_f"] = self._alloc(2 * R, H, dtype=f32)
        ws["e_t"] = ws["e_f"] if self.T == f32 else self._alloc(2 * R, H)
        ws["att"] = self._alloc(2 * R, H)
        ws["tmp"] = self._alloc(2 * R, H, dtype=f32)
        ws["a_f"] = self._alloc(2 * R, H, dtype=f32)
        ws["a_t"] = ws["a_f"] if self.T == f32 else self._alloc(2 * R, H)
        ws["hid"] = self._alloc(2 * R, F)
        ws["step_qkv"] = self._alloc(L, max_len, 2 * R, 3 * H)
        ws["head_f"] = self._alloc(R, H, dtype=f32)
        ws["head_t"] = ws["head_f"] if self.T == f32 else self._alloc(R, H)
        ws["logits"] = self._alloc(R, self.ldl, dtype=f32)
        if self.decode_x3:
            ws["a_t3"] = self._alloc(2 * R, 3 * H)          # LayerNorm 1 output, [hi | lo | hi]
            ws["hid_f"] = self._alloc(2 * R, F, dtype=f32)  # GELU output in fp32 ...
            ws["hid3"] = self._alloc(2 * R, 3 * F)          # ... and split
            ws["e_t3"] = self._alloc(2 * R, 3 * H)          # last layer's LayerNorm 2 output (feeds the head)
            ws["head_t3"] = self._alloc(R, 3 * H)
        if self.decode_f16:
            ws["a_th"] = self._alloc(2 * R, H, dtype=torch.float16)       # LayerNorm 1 output as halves
            ws["hid_h"] = self._alloc(2 * R, F, dtype=torch.float16)      # GELU output as halves
            ws["e_t2"] = self._alloc(2 * R, 2 * H)                        # LayerNorm 2 output: [bf16 | bit patterns of the halves]
            ws["head_th"] = self._alloc(R, H, dtype=torch.float16)
        if self.fused_decode:
            x3 = self.decode_x3
            # split-K factors of the partial-plane GEMMs (o-proj, fc2, head transform) and their fp32 planes
            ws["splits"] = dict(zip("oft", self.dec_splits))
            assert (H // 64) % ws["splits"]["o"] == 0 and (F // 64) % ws["splits"]["f"] == 0 and (H // 64) % ws["splits"]["t"] == 0
            ws["m_pad"], ws["m_pad_h"] = _round_up(2 * R, 128), _round_up(R, 128)
            ws["part"] = self._alloc(max(ws["splits"]["o"], ws["splits"]["f"]), ws["m_pad"], H, dtype=f32)
            ws["part_h"] = self._alloc(ws["splits"]["t"], ws["m_pad_h"], H, dtype=f32)
            ws["vpart"] = self._alloc(R, ops.vocab_partials(cfg.vocab), 4, dtype=f32)     # vocabulary arg-max partials
        ws["ids"] = torch.zeros(R, max_len, device=self.dev, dtype=i32)
        ws["unfinished"] = torch.ones(R, device=self.dev, dtype=i32)
        ws["sum_lp"] = torch.zeros(R, device=self.dev, dtype=f32)
        ws["n_steps"] = torch.zeros(R, device=self.dev, dtype=i32)
        ws["ident_rows"] = torch.arange(R, device=self.dev, dtype=i32)
        return ws

    def _decoder_ws(self, B, E, max_len):
        """Decode workspace of a (B, E, max_len) call, cached with its captured graphs."""
        key = (B, E, max_len)
        if key in self._dec_ws:
            self._dec_ws.move_to_end(key)
            return self._dec_ws[key]
        R = B * E
        ws = self._new_decoder_ws(B, E, max_len)
        ws["out_ids"] = torch.zeros(R, max_len, device=self.dev, dtype=torch.int64)
        ws["out_lp"] = torch.zeros(R, device=self.dev, dtype=torch.float32)
        ws["seed"] = torch.zeros(1, device=self.dev, dtype=torch.int64)     # sampling seed, rewritten before every replay
        ws["graphs"] = {}                  # captured decode loops over THIS workspace; they die with it
        self._dec_ws[key] = ws
        while len(self._dec_ws) > self.max_decode_workspaces:
            self._dec_ws.popitem(last=False)
        return ws

    def reserve(self, B, label_rows=False):
        """Sizes the image-side workspace before a batch starts (growing it later would drop the encoder outputs)."""
        return self._encoder_ws(B, self.cfg.n_ctx + (self.cfg.topk if label_rows else 0))

    # ------------------------------------------------------------------ building blocks
    def _t(self, name, idx, tensor):
        """Parity tap (tests / tools): hands a named intermediate to self.tap; free when no tap is set."""
        if self.tap is not None:
            self.tap(name, idx, tensor)

    def _ln(self, x, g, b, eps, out_t, out_f=None, rows=None):
        """LayerNorm of fp32 rows -> operand copy (and optional fp32 copy). In exact mode both are the same buffer."""
        if self.T == torch.float32:
            tgt = out_f if out_f is not None else out_t
            ops.layernorm(x, g, b, eps, out_t=None, out_f=tgt, rows=rows)
            return tgt
        ops.layernorm(x, g, b, eps, out_t=out_t, out_f=out_f, rows=rows)
        return out_t

    def _vit_block(self, p, x, rows, B, N, ws, out=None, pre=None, emit=None, tid=("block", -1)):
        """Pre-LN ViT block on the fp32 stream x (in place). vision_transformer.py:233-250.
        out: another stream buffer that receives the block's result while x stays untouched (the fork of the split encoder:
        the first concept block reads the shared trunk's output and starts its own stream without a copy).
        pre: (bf16 copy, statistics, tiles) of x left by the GEMM that produced it -> norm1 is folded into the qkv GEMM.
        emit: the same triple to be filled for the block's RESULT by its fc2 GEMM (for the next block's norm1)."""
        cfg = self.cfg
        H = cfg.hidden
        ln, qkv, att, hid = ws["ln"][:rows], ws["qkv"][:rows], ws["att"][:rows], ws["hid"][:rows]
        if pre is not None:
            ops.linear_ln_fold(pre[0][:rows], p["qkv_wf"], p["qkv_bf"], p["qkv_cf"], pre[1], pre[2], cfg.vit_ln_eps, qkv, M=rows)
        else:
            h = self._ln(x, p["n1w"], p["n1b"], cfg.vit_ln_eps, ln, rows=rows)
            ops.linear(h, p["qkv_w"], p["qkv_b"], qkv, M=rows)
        self._t(tid[0] + ".qkv", tid[1], qkv)
        ops.attention(qkv, att, B, N, cfg.heads, cfg.head_dim ** -0.5, impl=self.attn_impl)
        self._t(tid[0] + ".att", tid[1], att)
        dst = x if out is None else out
        if self.ln_fold and self.ln_fold2:
            f2 = ws["fold_2"]
            ops.linear_ln_emit(att, p["proj_w"], p["proj_b"], dst, x, f2[0][:rows], f2[1], M=rows)
            x = dst
            ops.linear_ln_fold(f2[0][:rows], p["fc1_wf"], p["fc1_bf"], p["fc1_cf"], f2[1], f2[2], cfg.vit_ln_eps, hid,
                               act=ops.ACT_GELU, M=rows)
        else:
            ops.linear(att, p["proj_w"], p["proj_b"], dst, resid=x, M=rows)
            x = dst
            h = self._ln(x, p["n2w"], p["n2b"], cfg.vit_ln_eps, ln, rows=rows)
            ops.linear(h, p["fc1_w"], p["fc1_b"], hid, act=ops.ACT_GELU, M=rows)
        self._t(tid[0] + ".mid", tid[1], x)            # stream after the attention branch (fp32; both taps precede fc2)
        self._t(tid[0] + ".hid", tid[1], hid)
        if emit is not None:
            ops.linear_ln_emit(hid, p["fc2_w"], p["fc2_b"], x, x, emit[0][:rows], emit[1], M=rows)
        else:
            ops.linear(hid, p["fc2_w"], p["fc2_b"], x, resid=x, M=rows)

    def _vit_block_cls_only(self, p, x, rows, B, N, ws):
        """The same block, evaluated for the CLS row of every image only (K and V still come from all rows). Used for the last
        block of the concept branch: its output is consumed at row 0 alone (pooler -> tag head, modeling_bert.py:1424-1425;
        tag token of the context, modeling_bert.py:1493), so the other 576 rows of Q / attention / proj / MLP are dead work."""
        cfg = self.cfg
        H = cfg.hidden
        ln, qkv = ws["ln"][:rows], ws["qkv"][:rows]
        q_cls, att_cls, ln_cls, hid_cls = ws["q_cls"][:B], ws["att_cls"][:B], ws["ln_cls"][:B], ws["hid_cls"][:B]
        h = self._ln(x, p["n1w"], p["n1b"], cfg.vit_ln_eps, ln, rows=rows)
        ops.linear(h, p["qkv_w"][H:], p["qkv_b"][H:], qkv[:, H:], M=rows, ldo=3 * H)          # K | V of every row
        h_cls = h.view(B, N * H)[:, :H]                                                        # row 0 of every image
        ops.linear(h_cls, p["qkv_w"][:H], p["qkv_b"][:H], q_cls, M=B)
        self._t("cls.kv", 0, qkv[:, H:])
        self._t("cls.q", 0, q_cls)
        ops.cls_attention(q_cls, qkv, att_cls, B, N, cfg.heads, cfg.head_dim ** -0.5)
        self._t("cls.att", 0, att_cls)
        x_cls = x.view(B, N * H)[:, :H]                                                        # fp32 stream, row 0 of every image
        ops.linear(att_cls, p["proj_w"], p["proj_b"], x_cls, resid=x_cls, M=B)
        self._t("cls.mid", 0, x_cls)
        h2 = self._ln(x_cls, p["n2w"], p["n2b"], cfg.vit_ln_eps, ln_cls, rows=B)
        ops.linear(h2, p["fc1_w"], p["fc1_b"], hid_cls, act=ops.ACT_GELU, M=B)
        self._t("cls.hid", 0, hid_cls)
        ops.linear(hid_cls, p["fc2_w"], p["fc2_b"], x_cls, resid=x_cls, M=B)

    # ------------------------------------------------------------------ stages
    def patch_embed(self, image, bgr=True):
        """image fp32 [B,3,S,S] normalised, or uint8 [B,S,S,3] HWC pixels (ToTensor + Normalize fused on the device; ``bgr``
        gives their channel order) -> img_feats fp32 [B,N,H] (a view of the trunk stream buffer)."""
        cfg, w = self.cfg, self.w
        B = image.shape[0]
        ws = self._encoder_ws(B)
        P, H, N = cfg.n_patches, cfg.hidden, cfg.n_tokens
        patches = ws["patches"][:B * P]
        po = ws["patch_out"][:B * P]
        if image.dtype == torch.uint8:
            ops.patchify_u8(image, patches, cfg.patch, bgr=bgr)
        else:
            ops.patchify(image, patches, cfg.patch)
        ops.linear(patches, w.patch_w, w.patch_b, po, M=B * P)
        x = ws["x"][:B * N]
        ops.assemble_tokens(po, w.cls_token, w.pos_embed, x, B, P, H)
        self._t("patch", 0, x.view(B, N, H))
        return x.view(B, N, H)

    def encode(self, img_feats, caption_branch=True, full_tag_feats=False):
        """TIMMVitSplitEncoder.forward (modeling_bert.py:458-478): returns (caption feats, tag feats) fp32 [B,N,H]. Unless
        ``full_tag_feats`` is set only row 0 (CLS) of the tag features is valid -- all the caption path consumes."""
        cfg, w = self.cfg, self.w
        B, N, H = img_feats.shape
        ws = self._encoder_ws(B)
        rows = B * N
        x = ws["x"][:rows]
        if img_feats.data_ptr() != x.data_ptr():
            x.copy_(img_feats.reshape(rows, H))
        xt = ws["xt"][:rows]
        split_at = cfg.enc_blocks - cfg.split_blocks
        # folded norm1: a block's fc2 GEMM leaves the bf16 copy + statistics of its result for the NEXT block's qkv GEMM
        # (fold_x: trunk / caption stream, fold_t: concept stream). Not for the very first block (its input comes from the
        # patch embedding) nor for the CLS-only last concept block (it normalises with the LayerNorm kernel)
        fx = ws["fold_x"] if self.ln_fold else None
        ft = ws["fold_t"] if self.ln_fold else None
        n_cap = cfg.enc_blocks - split_at if caption_branch else 0
        cls_only_last = not full_tag_feats
        n_tag_fold = cfg.split_blocks - (1 if cls_only_last else 0)     # concept blocks that run as full blocks
        pre = None
        for i in range(split_at):
            last_trunk = (i == split_at - 1)
            wanted = (not last_trunk) or n_cap > 0 or n_tag_fold > 0
            self._vit_block(w.blocks[i], x, rows, B, N, ws, pre=pre, emit=fx if wanted else None, tid=("block", i))
            pre = fx if wanted else None
            if self.tap is not None:
                self.tap("block", i, x.view(B, N, H))
        pre_trunk = pre
        # fork: both branches start from the block-8 input (modeling_bert.py:464-474). The concept branch runs first; its first
        # block reads x and writes xt, so the 0.9 GB stream is never copied
        forked = False
        pre_t = pre_trunk
        for j in range(cfg.split_blocks):
            if j == cfg.split_blocks - 1 and cls_only_last:
                if not forked:
                    xt.copy_(x)
                    forked = True
                self._vit_block_cls_only(w.tag_blocks[j], xt, rows, B, N, ws)
                if self.tap is not None:
                    self.tap("tag_block", j, xt.view(B, N, H))
                continue
            nxt_full = (j + 1 < n_tag_fold)                              # the next concept block consumes the folded form
            emit_t = ft if (self.ln_fold and nxt_full) else None
            if not forked:
                # reads the trunk's stream x (and its copy/statistics), writes the concept stream xt out of place
                self._vit_block_fork(w.tag_blocks[j], x, xt, rows, B, N, ws, pre=pre_t, emit=emit_t, tid=("tag_block", j))
                forked = True
            else:
                self._vit_block(w.tag_blocks[j], xt, rows, B, N, ws, pre=pre_t, emit=emit_t, tid=("tag_block", j))
            pre_t = emit_t
            if self.tap is not None:
                self.tap("tag_block", j, xt.view(B, N, H))
        if not forked:
            xt.copy_(x)
        if caption_branch:
            pre = pre_trunk
            for i in range(split_at, cfg.enc_blocks):
                emit = fx if (self.ln_fold and i + 1 < cfg.enc_blocks) else None
                self._vit_block(w.blocks[i], x, rows, B, N, ws, pre=pre, emit=emit, tid=("block", i))
                pre = emit
                if self.tap is not None:
                    self.tap("block", i, x.view(B, N, H))
        return x.view(B, N, H), xt.view(B, N, H)

    def _vit_block_fork(self, p, x, xt, rows, B, N, ws, pre=None, emit=None, tid=("tag_block", 0)):
        """First block of the concept branch: input = the shared trunk's stream x (left untouched), result -> xt."""
        self._vit_block(p, x, rows, B, N, ws, out=xt, pre=pre, emit=emit, tid=tid)

    def _head(self, hp, a_t, rows, th_f, th_t, logits):
        """BertLMPredictionHead (modeling_bert.py:540-563): dense + gelu -> LN(1e-12) -> tied/untied decoder + bias."""
        cfg = self.cfg
        ops.linear(a_t, hp["t_w"], hp["t_b"], th_f, act=ops.ACT_GELU, M=rows, lda=a_t.stride(0))
        t = self._ln(th_f, hp["ln_w"], hp["ln_b"], cfg.bert_ln_eps, th_t, rows=rows)
        ops.linear(t, hp["dec_w"], hp["bias"], logits[:, :cfg.vocab], M=rows, ldo=logits.stride(0))

    def tag_head(self, B):
        """pooler -> tag_logit -> sigmoid -> topk -> len (modeling_bert.py:1424-1432) on the tag stream."""
        cfg, w = self.cfg, self.w
        ws = self._encoder_ws(B)
        N, H = cfg.n_tokens, cfg.hidden
        cls_t, pooled = ws["cls_t"][:B], ws["pooled"][:B]
        ops.gather_rows(ws["xt"], N * H, cls_t, B, H)
        ops.linear(cls_t, w.pool_w, w.pool_b, pooled, act=ops.ACT_TANH, M=B)
        self._t("tag.pooled", 0, pooled)
        logits = ws["tag_logits"][:B]
        self._head(w.tag_head, pooled, B, ws["th_f"][:B], ws["th_t"][:B], logits)
        self._t("tag.logits", 0, logits[:, :cfg.vocab])
        ops.tag_topk(logits, cfg.vocab, cfg.topk, cfg.tag_thresh, ws["tag_idx"], ws["tag_prob"], ws["tag_len"], rows=B)
        return logits[:, :cfg.vocab], ws["tag_idx"][:B], ws["tag_prob"][:B], ws["tag_len"][:B]

    def set_labels(self, B, n_label):
        """Per-image count of visible od/tag label slots (int tensor [B]) for the label-region path; the device copies live in
        the workspace so that captured graphs see the current values."""
        cfg = self.cfg
        ws = self._encoder_ws(B, cfg.n_ctx + cfg.topk)
        ws["n_label"][:B].copy_(n_label.to(device=self.dev, dtype=torch.int32))
        ws["ctx_vis"][:B].copy_(ws["n_label"][:B] + cfg.n_ctx)

    def prefill(self, B, label_recipe=None):
        """Context rows [tag-CLS | caption feats] through the decoder once; per-layer q|k|v stay in ctx_qkv (KV cache).
        BertLayer, modeling_bert.py:303-437, bidirectional over the context (image rows only see image columns,
        pipeline file lines 57-85).
        label_recipe 'raw' / 'ln' (after set_labels): the topk od/tag label rows follow the C context rows of every image
        (modeling_bert.py:1447-1489); they see the context and the visible labels, the context rows do not see them."""
        cfg, w = self.cfg, self.w
        N, C, H = cfg.n_tokens, cfg.n_ctx, cfg.hidden
        Cp = C if label_recipe is None else C + cfg.topk
        ws = self._encoder_ws(B, Cp)
        rows = B * Cp
        ctx_f, ctx_t = ws["ctx_f"][:rows], ws["ctx_t"][:rows]
        ops.assemble_ctx(ws["x"], ws["xt"], ctx_f, ctx_t, B, N, H, rows_per_image=Cp)
        n_extra = None
        if label_recipe is not None:
            assert label_recipe in ("raw", "ln")
            n_extra = ws["n_label"]
            ops.label_rows(ws["tag_idx"], cfg.sep_id, label_recipe == "ln", cfg.max_seq_a, w.word, w.pos, w.type0, w.emb_ln_w,
                           w.emb_ln_b, cfg.bert_ln_eps, ctx_f, ctx_t, B, Cp, C)
        att, hid, tmp, a_f, ln = ws["att"][:rows], ws["hid"][:rows], ws["tmp_f"][:rows], ws["a_f"][:rows], ws["ln"][:rows]
        self._t("prefill.in", 0, ctx_f)
        if self.prefill_fold:
            return self._prefill_folded(ws, B, Cp, C, rows, n_extra)
        for l, p in enumerate(w.dec):
            qkv = ws["ctx_qkv"][l][:rows]
            last = (l == cfg.dec_layers - 1)
            if last:
                # only K and V of the last layer are ever used: project the k|v two thirds of the fused weight
                ops.linear(ctx_t, p["qkv_w"][H:], p["qkv_b"][H:], qkv[:, H:], M=rows, ldo=3 * H)
                self._t("prefill.qkv", l, qkv)
                break
            ops.linear(ctx_t, p["qkv_w"], p["qkv_b"], qkv, M=rows)
            self._t("prefill.qkv", l, qkv)
            ops.attention(qkv, att, B, Cp, cfg.heads, 1.0 / math.sqrt(cfg.head_dim), impl=self.attn_impl, n_base=C, n_extra=n_extra)
            self._t("prefill.att", l, att)
            ops.linear(att, p["o_w"], p["o_b"], tmp, resid=ctx_f, M=rows)
            a_t = self._ln(tmp, p["ln1_w"], p["ln1_b"], cfg.bert_ln_eps, ln, out_f=a_f, rows=rows)
            self._t("prefill.a", l, a_f)
            ops.linear(a_t, p["i_w"], p["i_b"], hid, act=ops.ACT_GELU, M=rows)
            self._t("prefill.hid", l, hid)
            ops.linear(hid, p["f_w"], p["f_b"], tmp, resid=a_f, M=rows)
            self._ln(tmp, p["ln2_w"], p["ln2_b"], cfg.bert_ln_eps, ctx_t, out_f=ctx_f, rows=rows)
            self._t("prefill.out", l, ctx_f)

    def _prefill_folded(self, ws, B, Cp, C, rows, n_extra):
        """The same layers without a stand-alone LayerNorm pass (each one re-read and re-wrote 2.3 GB at 512 images). BertLayer
        is post-LN: a = LN1(o(att) + x), y = LN2(f(gelu(i(a))) + a). Both normalised streams only ever feed (1) a GEMM -- which
        folds the normalisation (vc_linear_ln_fold on the raw bf16 copy + row statistics the producing GEMM emitted) -- and
        (2) the residual input of the next residual GEMM, which re-applies it to the raw fp32 tile in its epilogue
        (vc_linear_ln_emit_postln). raw1 / raw2 are the pre-LayerNorm rows o(att) + x and f(.) + a."""
        cfg, w = self.cfg, self.w
        H, eps = cfg.hidden, cfg.bert_ln_eps
        ctx_f, ctx_t = ws["ctx_f"][:rows], ws["ctx_t"][:rows]
        att, hid = ws["att"][:rows], ws["hid"][:rows]
        raw1, raw2 = ws["tmp_f"][:rows], ws["a_f"][:rows]
        xb1, xb2 = ws["ln"][:rows], ctx_t                  # ctx_t is free once layer 0 has projected it
        st1, st2, nst = ws["pf_st"]
        L = cfg.dec_layers
        for l, p in enumerate(w.dec):
            qkv = ws["ctx_qkv"][l][:rows]
            lo = H if l == L - 1 else 0                    # the last layer only ever serves K and V
            if l == 0:
                ops.linear(ctx_t, p["qkv_w"][lo:], p["qkv_b"][lo:], qkv[:, lo:], M=rows, ldo=3 * H)
            else:
                ops.linear_ln_fold(xb2, p["qkv_wf"][lo:], p["qkv_bf"][lo:], p["qkv_cf"][lo:], st2, nst, eps, qkv[:, lo:], M=rows, ldo=3 * H)
            self._t("prefill.qkv", l, qkv)
            if l == L - 1:
                break
            ops.attention(qkv, att, B, Cp, cfg.heads, 1.0 / math.sqrt(cfg.head_dim), impl=self.attn_impl, n_base=C, n_extra=n_extra)
            self._t("prefill.att", l, att)
            if l == 0:
                ops.linear_ln_emit(att, p["o_w"], p["o_b"], raw1, ctx_f, xb1, st1, M=rows)
            else:
                q = w.dec[l - 1]
                ops.linear_ln_emit(att, p["o_w"], p["o_b"], raw1, raw2, xb1, st1, M=rows, resid_ln=(st2, nst, q["ln2_w"], q["ln2_b"], eps))
            self._t("prefill.raw1", l, raw1)
            ops.linear_ln_fold(xb1, p["i_wf"], p["i_bf"], p["i_cf"], st1, nst, eps, hid, act=ops.ACT_GELU, M=rows)
            self._t("prefill.hid", l, hid)
            ops.linear_ln_emit(hid, p["f_w"], p["f_b"], raw2, raw1, xb2, st2, M=rows, resid_ln=(st1, nst, p["ln1_w"], p["ln1_b"], eps))
            self._t("prefill.raw2", l, raw2)

    def _decode_layers(self, ws, B, E, cur_len, anc, mask_id, labels=False, head=True, vocab="logits", live=(None, None)):
        """One decode step up to the vocabulary logits of the MASK rows. labels: the context holds C + topk rows per image of
        which ctx_vis[b] are visible. head=False stops after the decoder layers (caption-row replay after a recipe flip).
        vocab (fused path): 'logits' = ws['logits'] is filled; 'argmax' = only the per-tile (max, arg max, sum exp) partials in
        ws['vpart'] (greedy decoding); 'both' for parity taps. Returns True when ws['vpart'] holds this step's partials.
        live = (seq_unfinished int32 [R] or None, img_done int32 [B] or None): the attention skips finished captions."""
        if self.fused_decode:
            return self._decode_layers_fused(ws, B, E, cur_len, anc, mask_id, labels, head, vocab, live)
        cfg, w = self.cfg, self.w
        R, H = ws["R"], cfg.hidden
        C = cfg.n_ctx + (cfg.topk if labels else 0)
        enc = self._enc_ws
        ctx_vis = enc["ctx_vis"] if labels else None
        e_f, e_t = ws["e_f"], ws["e_t"]
        ops.embed_ln(ws["ids"], cur_len, mask_id, w.word, w.pos, w.type0, w.emb_ln_w, w.emb_ln_b, cfg.bert_ln_eps, e_f, e_t, R)
        scale = 1.0 / math.sqrt(cfg.head_dim)
        x3 = self.decode_x3
        n_layers = len(w.dec)
        for l, p in enumerate(w.dec):
            sq = ws["step_qkv"][l]
            ops.linear(e_t, p["qkv_w"], p["qkv_b"], sq[cur_len - 1], M=2 * R)
            ops.decode_attention(enc["ctx_qkv"][l], sq, anc, ws["att"], B, C, cfg.heads, E, cur_len, scale, ctx_vis=ctx_vis)
            ops.linear(ws["att"], p["o_w"], p["o_b"], ws["tmp"], resid=e_f, M=2 * R)
            if x3:
                # BertIntermediate / BertOutput (modeling_bert.py:395-419) on split-bf16 operands: the operand rounding of
                # these two layers and of the vocabulary head is what flips near-tie argmax decisions against the fp32
                # reference; the attention projections stay plain bf16
                ops.layernorm(ws["tmp"], p["ln1_w"], p["ln1_b"], cfg.bert_ln_eps, out_t=ws["a_t3"], out_f=ws["a_f"], rows=2 * R,
                              x3=True)
                # (K' = 2304: the CTA-pair kernel's 256 x 256 tiles halve the operand bytes each SM pulls from L2 per flop;
                # 23.4 us against 31.9 us on the heuristic's 128-wide tiles at 1024 rows, profiles/r01_x3probe_s10.log)
                ops.linear(ws["a_t3"], p["i_w3"], p["i_b"], ws["hid_f"], act=ops.ACT_GELU, M=2 * R,
                           impl="tc" if R >= 512 else "auto", tile_n=512 if R >= 512 else 0)
                ops.split_bf16x3(ws["hid_f"], ws["hid3"], rows=2 * R)
                if self.x3_dedup:
                    ops.linear_x3(ws["hid3"], p["f_w3"], p["f_b"], ws["tmp"], ws["a_f"], M=2 * R)
                else:
                    ops.linear(ws["hid3"], p["f_w3"], p["f_b"], ws["tmp"], resid=ws["a_f"], M=2 * R)
                if head and l == n_layers - 1:
                    ops.layernorm(ws["tmp"], p["ln2_w"], p["ln2_b"], cfg.bert_ln_eps, out_t=ws["e_t3"], out_f=e_f, rows=2 * R,
                                  x3=True)
                else:
                    self._ln(ws["tmp"], p["ln2_w"], p["ln2_b"], cfg.bert_ln_eps, e_t, out_f=e_f, rows=2 * R)
                continue
            a_t = self._ln(ws["tmp"], p["ln1_w"], p["ln1_b"], cfg.bert_ln_eps, ws["a_t"], out_f=ws["a_f"], rows=2 * R)
            ops.linear(a_t, p["i_w"], p["i_b"], ws["hid"], act=ops.ACT_GELU, M=2 * R)
            ops.linear(ws["hid"], p["f_w"], p["f_b"], ws["tmp"], resid=ws["a_f"], M=2 * R)
            self._ln(ws["tmp"], p["ln2_w"], p["ln2_b"], cfg.bert_ln_eps, e_t, out_f=e_f, rows=2 * R)
        # vocabulary head on the MASK rows only (rows 1::2); the reference runs it over all T text rows
        # (modeling_bert.py:809-810) and keeps one
        if head and x3:
            hp = w.cls_head
            ops.linear(ws["e_t3"][1::2], hp["t_w3"], hp["t_b"], ws["head_f"], act=ops.ACT_GELU, M=R)
            ops.layernorm(ws["head_f"], hp["ln_w"], hp["ln_b"], cfg.bert_ln_eps, out_t=ws["head_t3"], rows=R, x3=True)
            ops.linear(ws["head_t3"], hp["dec_w3"], hp["bias"], ws["logits"][:, :cfg.vocab], M=R, ldo=ws["logits"].stride(0))
        elif head:
            mask_rows = e_t[1::2]
            self._head(w.cls_head, mask_rows, R, ws["head_f"], ws["head_t"], ws["logits"])

    def _decode_layers_fused(self, ws, B, E, cur_len, anc, mask_id, labels, head, vocab, live=(None, None)):
        """The same decode step (BertLayer x L + BertLMPredictionHead, modeling_bert.py:303-437, 540-563) on the decode-step
        kernels. Per layer: q|k|v GEMM -> attention over the K/V cache -> o-proj as split-K partial planes -> finish (bias +
        residual + LayerNorm 1 -> fp32 row + operand) -> fc1 (+ GELU, written as the split pair) -> fc2 partial planes ->
        finish (LayerNorm 2). With decode_x3 the MLP and the head run on split operands [hi | lo] (three tensor-core products)."""
        cfg, w = self.cfg, self.w
        R, H, F = ws["R"], cfg.hidden, cfg.inter
        M = 2 * R
        C = cfg.n_ctx + (cfg.topk if labels else 0)
        enc = self._enc_ws
        ctx_vis = enc["ctx_vis"] if labels else None
        x3, f16 = self.decode_x3, self.decode_f16
        eps = cfg.bert_ln_eps
        scale = 1.0 / math.sqrt(cfg.head_dim)
        sp, part, m_pad = ws["splits"], ws["part"], ws["m_pad"]
        e_f, a_f = ws["e_f"], ws["a_f"]
        # operand copies of the two streams: bf16 rows; (x3) [hi | lo | .] rows of pitch 3H whose first H columns ARE the bf16
        # copy (the q|k|v GEMM reads them with lda = 3H); (f16) IEEE halves -- the LayerNorm 2 stream as [bf16 | half] rows of
        # pitch 2H, because the q|k|v projection stays a bf16 product
        e_op = ws["e_t3"] if x3 else (ws["e_t2"] if f16 else ws["e_t"])
        a_op = ws["a_t3"] if x3 else (ws["a_th"] if f16 else ws["a_t"])
        hid = ws["hid3"] if x3 else (ws["hid_h"] if f16 else ws["hid"])
        split1 = "f16" if f16 else x3
        split2 = "bf16+f16" if f16 else x3
        wi, wf = ("i_w3", "f_w3") if x3 else (("i_wh", "f_wh") if f16 else ("i_w", "f_w"))
        ops.embed_ln(ws["ids"], cur_len, mask_id, w.word, w.pos, w.type0, w.emb_ln_w, w.emb_ln_b, eps, e_f, ws["e_t"], R)
        cur = ws["e_t"]                                   # layer 0 reads the embedding rows (bf16, pitch H)
        for l, p in enumerate(w.dec):
            sq = ws["step_qkv"][l]
            ops.dec_linear(ops.DEC_BF16, cur[:, :H], p["qkv_w"], p["qkv_b"], sq[cur_len - 1], M=M)
            ops.decode_attention(enc["ctx_qkv"][l], sq, anc, ws["att"], B, C, cfg.heads, E, cur_len, scale, ctx_vis=ctx_vis,
                                 seq_unfinished=live[0], img_done=live[1])
            ops.dec_linear(ops.DEC_PARTIAL, ws["att"], p["o_w"], None, part, M=M, splits=sp["o"], m_pad=m_pad)
            ops.finish_ln(part, sp["o"], p["o_b"], p["ln1_w"], p["ln1_b"], eps, M, resid=e_f, out_f=a_f, out_t=a_op, split=split1)
            if x3:
                ops.dec_linear(ops.DEC_GELU_SPLIT, a_op, p[wi], p["i_b"], hid, M=M, x3=True)
            else:
                ops.dec_linear(ops.DEC_GELU_BF16, a_op, p[wi], p["i_b"], hid, M=M)
            ops.dec_linear(ops.DEC_PARTIAL, hid, p[wf], None, part, M=M, x3=x3, splits=sp["f"], m_pad=m_pad)
            ops.finish_ln(part, sp["f"], p["f_b"], p["ln2_w"], p["ln2_b"], eps, M, resid=a_f, out_f=e_f, out_t=e_op, split=split2)
            cur = e_op
        if not head:
            return False
        # vocabulary head on the MASK rows only (rows 1::2); the reference runs it over all T text rows
        # (modeling_bert.py:809-810) and keeps one
        hp = w.cls_head
        if f16:
            mask_rows = e_op.view(torch.float16)[1::2, H:2 * H]          # the half columns of the [bf16 | half] rows
            head_op, t_w, dec_w = ws["head_th"], hp["t_wh"], hp["dec_wh"]
        else:
            mask_rows = e_op[1::2]
            head_op = ws["head_t3"] if x3 else ws["head_t"]
            t_w, dec_w = (hp["t_w3"], hp["dec_w3"]) if x3 else (hp["t_w"], hp["dec_w"])
        ops.dec_linear(ops.DEC_PARTIAL, mask_rows, t_w, None, ws["part_h"], M=R, x3=x3, splits=sp["t"], m_pad=ws["m_pad_h"])
        # (the K-concatenated GEMM that materialises the logits reads [hi | lo | hi]; the arg-max kernel only [hi | lo])
        ops.finish_ln(ws["part_h"], sp["t"], hp["t_b"], hp["ln_w"], hp["ln_b"], eps, R, gelu=True, out_t=head_op,
                      split="f16" if f16 else ((3 if vocab != "argmax" else True) if x3 else False))
        if vocab != "argmax":
            if f16:       # fp32 logits from the decode-step kernel itself: one plane with the bias (no half form of vc_linear)
                ops.dec_linear(ops.DEC_PARTIAL, head_op, dec_w, hp["bias"], ws["logits"][:, :cfg.vocab], M=R)
            else:
                ops.linear(head_op, dec_w, hp["bias"], ws["logits"][:, :cfg.vocab], M=R, ldo=ws["logits"].stride(0))
        if vocab != "logits":
            ops.dec_vocab_argmax(head_op, dec_w, hp["bias"], ws["vpart"], M=R, x3=x3)
            return True
        return False

    def _flip_labels(self, ws, B, E, cur_len, mask_id, anc_table=None):
        """The reference switches the label embedding to the 'raw' recipe at this step (modeling_bert.py:1435) and, having no
        cache, recomputes every row under it: prefill the context again and replay the caption rows of steps 1..cur_len-1
        (each sequence's own token history is in ws['ids']; beams were re-parented there, so the replay is row-local)."""
        self.prefill(B, label_recipe="raw")
        for s in range(1, cur_len):
            self._decode_layers(ws, B, E, s, None, mask_id, labels=True, head=False)
        if anc_table is not None and cur_len > 1:
            anc_table[:cur_len - 1].copy_(ws["ident_rows"].unsqueeze(0).expand(cur_len - 1, -1))

    # ------------------------------------------------------------------ search drivers
    @staticmethod
    def _eos_tensor(ws, eos_ids):
        """Device copy of the EOS id list, created ONCE per workspace: a captured CUDA graph keeps its raw pointer, so it must
        outlive every replay (a per-call tensor would be freed and its block reused under the graph's feet)."""
        key = ("eos", tuple(int(e) for e in eos_ids))
        if key not in ws:
            ws[key] = torch.tensor(list(key[1]), device=ws["out_ids"].device, dtype=torch.int32)
        return ws[key]

    def _if_stream(self):
        """The stream the bodies of the conditional decode steps are captured on (ops.body_stream: one per device)."""
        if self._body_stream is None:
            self._body_stream = ops.body_stream(self.dev)
        return self._body_stream

    def _exit_every(self, n):
        """Decode steps per early-exit condition. A conditional node costs ~15 us per replay (condition kernel + node), and the
        chance that ALL n captions have finished at a given step falls quickly with n: one node per step for small batches
        (serving: the call returns at the first step after its last EOS), one per 2 / 4 steps for larger ones. Steps that run
        after every caption has finished change nothing (finished rows receive PAD, their attention is skipped).
        VITCAP_EARLY_EXIT_EVERY overrides."""
        env = os.environ.get("VITCAP_EARLY_EXIT_EVERY")
        if env:
            return max(1, int(env))
        return 1 if n <= 16 else (2 if n <= 128 else 4)

    def _step_groups(self, max_len, n, conditional):
        """[(first step, last step + 1, behind a condition?)] covering steps 1 .. max_len - 1: step 1 always runs."""
        if not conditional:
            return [(1, max_len, False)]
        every = self._exit_every(n)
        return [(1, 2, False)] + [(f, min(f + every, max_len), True) for f in range(2, max_len, every)]

    def _maybe_graph(self, ws, key, fn):
        """Runs fn() eagerly once (warm-up: lazy kernel attribute setup, descriptor cache), then captures and replays it."""
        if not self.use_cuda_graph or self.inline_graphs or self.tap is not None:
            fn()
            return
        graphs = ws["graphs"]
        g = graphs.get(key)
        if g is None:
            fn()                                         # warm-up / first execution is the real one
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            before = ops.launch_count()
            with torch.cuda.graph(g):
                fn()
            self.stats["graph_kernels"] = ops.launch_count() - before
            graphs[key] = g
            return "captured_after_eager"
        g.replay()
        self.stats["graph_replays"] = self.stats.get("graph_replays", 0) + 1

    def greedy_or_sample(self, B, E, max_len, bos, pad, eos_ids, mask_id, do_sample=False, temperature=1.0, top_k=0,
                         top_p=1.0, seed=0, label_flip=None):
        """_generate_no_beam_search (modeling_utils.py:768-886) for R = B*E sequences; returns (ids int64 [R,1,max_len],
        logprob fp32 [R,1]). label_flip: None = no visible label region; else the first cur_len decoded under the 'raw' label
        recipe (the context must have been prefilled with 'ln' if label_flip > 1, else with 'raw')."""
        labels = label_flip is not None
        cfg = self.cfg
        ws = self._decoder_ws(B, E, max_len)
        R = ws["R"]
        eos = self._eos_tensor(ws, eos_ids)
        filt = do_sample and (top_k > 0 or top_p < 1.0)

        def step(cur_len):
            if labels and cur_len == label_flip and cur_len > 1:
                self._flip_labels(ws, B, E, cur_len, mask_id)
            # greedy decoding on the fused path never materialises the logits (a parity tap asks for both)
            want = "logits" if do_sample else ("both" if self.tap is not None else "argmax")
            partials = self._decode_layers(ws, B, E, cur_len, None, mask_id, labels=labels, vocab=want,
                                           live=(ws["unfinished"], None))
            if want != "argmax" or not partials:
                self._t("logits", cur_len, ws["logits"][:, :cfg.vocab])
            if partials:
                ops.token_step_partials(ws["vpart"], R, cur_len, pad, eos, ws["ids"], ws["unfinished"], ws["sum_lp"],
                                        ws["n_steps"])
                return
            t = temperature
            if filt:
                ops.filter_logits(ws["logits"], cfg.vocab, R, 1.0 / temperature, top_k, top_p)
                t = 1.0
            ops.token_step(ws["logits"], cfg.vocab, R, do_sample, t, 0, cur_len, pad, eos, ws["ids"], ws["unfinished"],
                           ws["sum_lp"], ws["n_steps"], seed_dev=ws["seed"] if do_sample else None)

        def run():
            # positions a finished batch never reaches stay PAD (the reference pads after its break, modeling_utils.py:879-883)
            ws["ids"].fill_(pad)
            ws["ids"][:, 0] = bos
            ws["unfinished"].fill_(1)
            ws["sum_lp"].zero_()
            ws["n_steps"].zero_()
            # `if cur_unfinished.max() == 0: break` (modeling_utils.py:865-867) on the device: in a captured loop the steps
            # after the first are bodies of conditional nodes, `every` steps per node (_exit_every; the label-recipe flip
            # re-runs the prefill with torch copies in it: those loops keep the plain sequence)
            for first, last, cond in self._step_groups(max_len, R, self.early_exit and not labels):
                with ops.graph_if_any(ws["unfinished"], self._if_stream(), enabled=cond):
                    for cur_len in range(first, last):
                        step(cur_len)
            ops.greedy_finalize(ws["ids"], ws["unfinished"], ws["sum_lp"], ws["n_steps"], int(eos_ids[0]), R, ws["out_ids"],
                                ws["out_lp"])

        if do_sample:
            # the seed lives in device memory, outside the captured loop: one graph serves every call
            s64 = int(seed) & 0xFFFFFFFFFFFFFFFF
            ws["seed"].fill_(s64 - (1 << 64) if s64 >= (1 << 63) else s64)
        self._maybe_graph(ws, ("tok", B, E, max_len, do_sample, temperature, top_k, top_p, bos, pad, tuple(eos_ids), mask_id,
                           label_flip), run)
        return ws["out_ids"].view(R, 1, max_len).clone(), ws["out_lp"].view(R, 1).clone()

    def beam_search(self, B, nb, max_len, bos, pad, eos_ids, mask_id, length_penalty=1.0, keep=1, label_flip=None):
        """_generate_beam_search (modeling_utils.py:888-1100), do_sample=False. Returns (ids int64 [B,keep,max_len],
        logprob fp32 [B,keep]). label_flip as in greedy_or_sample."""
        labels = label_flip is not None
        cfg = self.cfg
        ws = self._decoder_ws(B, nb, max_len)
        R, K = ws["R"], 2 * nb
        f32, i32 = torch.float32, torch.int32
        # one state per `keep`: a graph captured for another keep holds raw pointers into ITS state, which must stay alive
        if ("beam", keep) not in ws:
            ws[("beam", keep)] = {
                "keep": keep,
                "ids": ws["ids"], "beam_scores": torch.zeros(R, device=self.dev, dtype=f32),
                "done": torch.zeros(B, device=self.dev, dtype=i32), "anc": torch.zeros(max_len, R, device=self.dev, dtype=i32),
                "hyp_score": torch.zeros(B, keep, device=self.dev, dtype=torch.float64),
                "hyp_len": torch.zeros(B, keep, device=self.dev, dtype=i32),
                "hyp_ids": torch.zeros(B, keep, max_len, device=self.dev, dtype=i32),
                "hyp_count": torch.zeros(B, device=self.dev, dtype=i32),
                "worst": torch.zeros(B, device=self.dev, dtype=torch.float64),
                "cand_val": torch.zeros(R, K, device=self.dev, dtype=f32), "cand_idx": torch.zeros(R, K, device=self.dev, dtype=i32),
                "row_max": torch.zeros(R, device=self.dev, dtype=f32), "row_logsum": torch.zeros(R, device=self.dev, dtype=f32),
                "out_ids": torch.zeros(B, keep, max_len, device=self.dev, dtype=torch.int64),
                "out_lp": torch.zeros(B, keep, device=self.dev, dtype=f32),
                "init_scores": torch.tensor(([0.0] + [-1e9] * (nb - 1)) * B, device=self.dev, dtype=f32),
            }
        st = ws[("beam", keep)]
        eos = self._eos_tensor(ws, eos_ids)

        def run():
            st["ids"].zero_()
            st["ids"][:, 0] = bos
            st["beam_scores"].copy_(st["init_scores"])       # [0, -1e9, ...] per image (modeling_utils.py:922-924)
            st["done"].zero_()
            st["anc"].zero_()
            st["hyp_count"].zero_()
            st["worst"].fill_(1e9)
            # `if all(done): break` (modeling_utils.py:1071-1073) on the device, as in greedy_or_sample (B images decide)
            for first, last, cond in self._step_groups(max_len, B, self.early_exit and not labels):
                with ops.graph_if_any(st["done"], self._if_stream(), invert=True, enabled=cond):
                    for cur_len in range(first, last):
                        if labels and cur_len == label_flip and cur_len > 1:
                            self._flip_labels(ws, B, nb, cur_len, mask_id, anc_table=st["anc"])
                        self._decode_layers(ws, B, nb, cur_len, st["anc"], mask_id, labels=labels, live=(None, st["done"]))
                        ops.beam_row_topk(ws["logits"], cfg.vocab, R, K, st["cand_val"], st["cand_idx"], st["row_max"], st["row_logsum"])
                        ops.beam_advance(st, st["cand_val"], st["cand_idx"], st["row_max"], st["row_logsum"], B, nb, cfg.vocab,
                                         cur_len, keep, length_penalty, pad, eos)
            ops.beam_finalize(st, B, keep, pad, int(eos_ids[0]), st["out_ids"], st["out_lp"])

        self._maybe_graph(ws, ("beam", B, nb, max_len, keep, float(length_penalty), bos, pad, tuple(eos_ids), mask_id, label_flip), run)
        return st["out_ids"].clone(), st["out_lp"].clone()
