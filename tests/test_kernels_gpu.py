"""Per-kernel parity on the GPU: every C-ABI entry point against a plain torch fp32 expression of the same
reference operator (tolerances stated per test; integer outputs are compared bit-exactly)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from vitcap_b200 import ops  # noqa: E402


def dev():
    return torch.device("cuda:0")


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(dev())


def ref_linear(a, w, bias, act, resid):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == ops.ACT_GELU:
        y = y * 0.5 * (1 + torch.erf(y / math.sqrt(2)))
    elif act == ops.ACT_TANH:
        y = torch.tanh(y)
    if resid is not None:
        y = y + resid
    return y


@pytest.mark.parametrize("M,N,K", [(300, 200, 64), (128, 768, 768), (577 * 2, 2304, 768), (77, 30522 // 16, 128)])
@pytest.mark.parametrize("act", [ops.ACT_NONE, ops.ACT_GELU, ops.ACT_TANH])
def test_linear_exact_fp32(M, N, K, act):
    a, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3)
    resid = rnd(M, N, seed=4)
    ldo = (N + 7) // 8 * 8
    out = torch.zeros(M, ldo, device=dev())
    ops.linear(a, w, b, out[:, :N], act=act, resid=resid, ldo=ldo)
    ref = ref_linear(a, w, b, act, resid)
    torch.testing.assert_close(out[:, :N], ref, rtol=1e-4, atol=1e-4)


TC_SHAPES = [(128, 256, 64), (128, 256, 768), (256, 512, 128), (300, 200, 64), (577 * 2, 2304, 768), (1154, 768, 3072),
             (64, 1000, 768), (1024, 768, 768)]


@pytest.mark.parametrize("tile_n", [512, 256, 128, 64])
@pytest.mark.parametrize("M,N,K", TC_SHAPES)
def test_linear_tc_bf16_plain(M, N, K, tile_n):
    """tcgen05 GEMM, bf16 out, bias only. Tolerance: bf16 output rounding (2^-8 relative) on fp32-accumulated sums."""
    a, w, b = rnd(M, K, seed=1, dtype=torch.bfloat16), rnd(N, K, seed=2, scale=0.05, dtype=torch.bfloat16), rnd(N, seed=3)
    ldo = (N + 7) // 8 * 8
    out = torch.zeros(M, ldo, device=dev(), dtype=torch.bfloat16)
    ops.linear(a, w, b, out[:, :N], ldo=ldo, impl="tc", tile_n=tile_n)
    ref = ref_linear(a, w, b, ops.ACT_NONE, None)
    torch.testing.assert_close(out[:, :N].float(), ref, rtol=1e-2, atol=2e-2)
    Nr = (N + 7) // 8 * 8
    if ldo > Nr:
        assert float(out[:, Nr:].abs().max()) == 0.0         # never writes past the 16-byte boundary after N


@pytest.mark.parametrize("act", [ops.ACT_NONE, ops.ACT_GELU, ops.ACT_TANH])
@pytest.mark.parametrize("out_f32,with_resid", [(True, True), (True, False), (False, False)])
def test_linear_tc_epilogues(act, out_f32, with_resid):
    M, N, K = 577 * 3, 768, 768
    a, w, b = rnd(M, K, seed=5, dtype=torch.bfloat16), rnd(N, K, seed=6, scale=0.03, dtype=torch.bfloat16), rnd(N, seed=7)
    resid = rnd(M, N, seed=8) if with_resid else None
    out = torch.zeros(M, N, device=dev(), dtype=torch.float32 if out_f32 else torch.bfloat16)
    ops.linear(a, w, b, out, act=act, resid=resid)
    ref = ref_linear(a, w, b, act, resid)
    if out_f32:
        torch.testing.assert_close(out, ref, rtol=2e-4, atol=2e-4)   # fp32 accumulate of exact bf16 products
    else:
        torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("M,N,K", [(577 * 3, 768, 768), (20000, 768, 768), (9000, 1024, 3072), (4111, 2304, 768)])
@pytest.mark.parametrize("act", [ops.ACT_NONE, ops.ACT_GELU, ops.ACT_TANH])
@pytest.mark.parametrize("out_f32,with_resid", [(True, True), (True, False), (False, False)])
def test_linear_tc2_cta_pair_epilogues(M, N, K, act, out_f32, with_resid):
    """CTA-pair kernel (cta_group::2, forced with tile_n=512): several rounds of tiles per cluster (accumulator / stage phase
    wrap), ragged M, every epilogue. GELU here is the tanh-form fit of erf (|err| <= 3e-5 + MUFU.TANH's 2^-11 relative)."""
    a, w, b = rnd(M, K, seed=5, dtype=torch.bfloat16), rnd(N, K, seed=6, scale=0.03, dtype=torch.bfloat16), rnd(N, seed=7)
    resid = rnd(M, N, seed=8) if with_resid else None
    out = torch.zeros(M, N, device=dev(), dtype=torch.float32 if out_f32 else torch.bfloat16)
    ops.linear(a, w, b, out, act=act, resid=resid, impl="tc", tile_n=512)
    ref = ref_linear(a, w, b, act, resid)
    if out_f32:
        tol = 2e-3 if act == ops.ACT_GELU else 3e-4
        torch.testing.assert_close(out, ref, rtol=tol, atol=tol)
    else:
        torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)


def test_linear_tc2_inplace_residual_and_heuristic():
    """x = x + proj(h) in place on the CTA-pair kernel; the auto heuristic must pick it for the encoder shape and agree."""
    M, N, K = 577 * 64, 768, 768
    a, w, b = rnd(M, K, seed=5, dtype=torch.bfloat16), rnd(N, K, seed=6, scale=0.03, dtype=torch.bfloat16), rnd(N, seed=7)
    x = rnd(M, N, seed=9)
    x2 = x.clone()
    ref = ref_linear(a, w, b, ops.ACT_NONE, x.clone())
    ops.linear(a, w, b, x, resid=x, impl="tc", tile_n=512)
    torch.testing.assert_close(x, ref, rtol=3e-4, atol=3e-4)
    ops.linear(a, w, b, x2, resid=x2)
    torch.testing.assert_close(x, x2, rtol=0, atol=1e-6)
    y1 = torch.empty(M, N, device=dev(), dtype=torch.bfloat16)
    y2 = torch.empty_like(y1)
    ops.linear(a, w, b, y1, impl="tc", tile_n=512)
    ops.linear(a, w, b, y2, impl="tc", tile_n=256)
    assert torch.equal(y1, y2)                       # same products, same accumulation order per output element


def test_linear_tc_inplace_residual_stream():
    """x = x + proj(h): out aliases resid (vision_transformer.py:246)."""
    M, N, K = 1154, 768, 768
    a, w, b = rnd(M, K, seed=5, dtype=torch.bfloat16), rnd(N, K, seed=6, scale=0.03, dtype=torch.bfloat16), rnd(N, seed=7)
    x = rnd(M, N, seed=9)
    ref = ref_linear(a, w, b, ops.ACT_NONE, x.clone())
    ops.linear(a, w, b, x, resid=x)
    torch.testing.assert_close(x, ref, rtol=2e-4, atol=2e-4)


def test_linear_tc_strided_rows_and_vocab_tail():
    """A rows taken with a pitch (the MASK rows 1::2 of the decode buffer) and the ragged 30522-wide vocabulary."""
    R, K, V = 96, 768, 30522
    buf = rnd(2 * R, K, seed=11, dtype=torch.bfloat16)
    w = rnd(V, K, seed=12, scale=0.02, dtype=torch.bfloat16)
    bias = rnd(V, seed=13)
    ldl = (V + 63) // 64 * 64
    logits = torch.full((R, ldl), -7.0, device=dev())
    ops.linear(buf[1::2], w, bias, logits[:, :V], M=R, lda=2 * K, ldo=ldl)
    ref = ref_linear(buf[1::2], w, bias, ops.ACT_NONE, None)
    torch.testing.assert_close(logits[:, :V], ref, rtol=2e-4, atol=2e-4)
    # the TMA store clips at 16-byte granularity: only the pad columns up to the next 16-byte boundary may be touched
    Vr = (V + 3) // 4 * 4
    assert bool((logits[:, Vr:] == -7.0).all())


def test_linear_tc_matches_simt_bf16_bitwise_inputs():
    """Independent on-device cross-check: CUDA-core kernel on the same bf16 operands."""
    M, N, K = 640, 512, 1024
    a, w = rnd(M, K, seed=21, dtype=torch.bfloat16), rnd(N, K, seed=22, scale=0.05, dtype=torch.bfloat16)
    o1 = torch.zeros(M, N, device=dev())
    o2 = torch.zeros(M, N, device=dev())
    ops.linear(a, w, None, o1, impl="tc")
    ops.linear(a, w, None, o2, impl="simt")
    torch.testing.assert_close(o1, o2, rtol=1e-4, atol=1e-4)


def ref_attention(qkv, heads, scale):
    B, N, H3 = qkv.shape
    H = H3 // 3
    q, k, v = qkv.float().view(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    a = torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1)
    return (a @ v).transpose(1, 2).reshape(B, N, H)


@pytest.mark.parametrize("B,N,heads", [(2, 577, 12), (3, 197, 12), (1, 578, 12), (2, 145, 12), (2, 17, 2), (1, 128, 1), (1, 129, 1)])
def test_attention_exact_fp32(B, N, heads):
    qkv = rnd(B, N, 3 * heads * 64, seed=3)
    out = torch.zeros(B, N, heads * 64, device=dev())
    ops.attention(qkv, out, B, N, heads, 0.125)
    torch.testing.assert_close(out, ref_attention(qkv, heads, 0.125), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("impl", ["auto", "simt"])
@pytest.mark.parametrize("B,N,heads", [(2, 577, 12), (3, 197, 12), (1, 578, 12), (2, 145, 12), (2, 17, 2), (1, 128, 1), (1, 129, 1), (1, 256, 3),
                                       (2, 288, 2), (1, 289, 3), (2, 290, 1), (1, 385, 2), (1, 674, 2), (1, 576, 1)])
def test_attention_bf16(B, N, heads, impl):
    """bf16 operands; P is rounded to bf16 before the PV product (as any flash kernel does): 2e-2 abs on O(1) outputs."""
    qkv = rnd(B, N, 3 * heads * 64, seed=3, dtype=torch.bfloat16)
    out = torch.zeros(B, N, heads * 64, device=dev(), dtype=torch.bfloat16)
    ops.attention(qkv, out, B, N, heads, 0.125, impl=impl)
    ref = ref_attention(qkv, heads, 0.125)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)


def test_attention_bf16_many_ctas():
    """Enough (image, head) work items that CTAs are co-resident and SMs run several waves (pipeline hand-over,
    TMEM re-allocation, barrier phase bookkeeping across tiles)."""
    B, N, heads = 48, 577, 12
    qkv = rnd(B, N, 3 * heads * 64, seed=9, dtype=torch.bfloat16)
    out = torch.zeros(B, N, heads * 64, device=dev(), dtype=torch.bfloat16)
    for _ in range(3):
        ops.attention(qkv, out, B, N, heads, 0.125)
    torch.cuda.synchronize()
    ref = ref_attention(qkv, heads, 0.125)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)


def _ref_attention_rounded_p(qkv, heads, scale):
    """The kernels' own statement of the arithmetic (oracle/port.py QuantPortModel.attend): exp2 domain, integer exponent
    reference, P rounded to bf16 for the V product, row sum of the unrounded P, bf16 output."""
    B, N, H3 = qkv.shape
    q, k, v = qkv.double().view(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * (scale * 1.4426950408889634)
    p = torch.exp2(s - torch.ceil(s.max(-1, keepdim=True).values))
    o = (p.float().to(torch.bfloat16).double() @ v) / p.sum(-1, keepdim=True)
    return o.transpose(1, 2).reshape(B, N, heads * 64).float()


@pytest.mark.parametrize("B,N,heads", [(3, 577, 12), (2, 578, 12), (2, 288, 2), (2, 197, 3), (1, 674, 1)])
@pytest.mark.parametrize("gain", [1.0, 2.5])
def test_attention_bf16_vs_rounded_p_statement(B, N, heads, gain):
    """Both tensor-core attention kernels (64-key chunks; 96-key chunks with the leading keys peeled off onto the CUDA cores for
    N = p + 96 m) against the quantisation-matched statement: what is left is the final bf16 rounding of the output (2^-8
    relative) and the rare P whose bf16 rounding flips under MUFU.EX2's 2^-22 error. A mishandled peeled key (a missing score
    in the row maximum / row sum, an unrounded P) shows as a 1e-2 .. 1e-1 error here."""
    qkv = rnd(B, N, 3 * heads * 64, seed=11, scale=gain, dtype=torch.bfloat16)
    out = torch.zeros(B, N, heads * 64, device=dev(), dtype=torch.bfloat16)
    ops.attention(qkv, out, B, N, heads, 0.125)
    ref = _ref_attention_rounded_p(qkv, heads, 0.125)
    err = (out.float() - ref).abs()
    # element-wise: one bf16 ulp of the output plus one flipped P rounding of a dominating key (2^-9 of the row's scale)
    assert float((err - (2.0 ** -7) * ref.abs()).max()) <= 2.0 ** -8 * float(ref.abs().max()), float(err.max())
    assert float(err.norm() / ref.norm()) < 3e-3


def test_attention_96_key_kernel_equals_64_key_kernel():
    """VITCAP_ATTN96=0 routes N = 577 / 578 to the 64-key kernel: same P roundings (integer exponent reference), different
    summation order only."""
    import os
    for N in (577, 578):
        qkv = rnd(4, N, 3 * 12 * 64, seed=13, dtype=torch.bfloat16)
        a = torch.zeros(4, N, 768, device=dev(), dtype=torch.bfloat16)
        b = torch.zeros_like(a)
        ops.attention(qkv, a, 4, N, 12, 0.125)
        os.environ["VITCAP_ATTN96"] = "0"
        try:
            ops.attention(qkv, b, 4, N, 12, 0.125)
        finally:
            del os.environ["VITCAP_ATTN96"]
        torch.cuda.synchronize()
        d = (a.float() - b.float()).abs()
        assert float(d.max()) <= 2.0 ** -7 * float(b.float().abs().max())      # at most one bf16 ulp of the largest output
        assert float((d > 0).float().mean()) < 0.05


def test_attention_bf16_peaky_scores():
    """Large-magnitude scores (online-softmax rescaling across chunks must stay exact)."""
    B, N, heads = 1, 577, 2
    qkv = rnd(B, N, 3 * heads * 64, seed=4, scale=3.0, dtype=torch.bfloat16)
    out = torch.zeros(B, N, heads * 64, device=dev(), dtype=torch.bfloat16)
    ops.attention(qkv, out, B, N, heads, 0.125)
    torch.testing.assert_close(out.float(), ref_attention(qkv, heads, 0.125), rtol=3e-2, atol=6e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("eps", [1e-6, 1e-12])
def test_layernorm(dtype, eps):
    rows, H = 1000, 768
    x, g, b = rnd(rows, H, seed=1, scale=2.0), rnd(H, seed=2), rnd(H, seed=3)
    o_t = torch.zeros(rows, H, device=dev(), dtype=dtype)
    o_f = torch.zeros(rows, H, device=dev())
    ops.layernorm(x, g, b, eps, out_t=o_t, out_f=o_f)
    ref = F.layer_norm(x, (H,), g, b, eps)
    torch.testing.assert_close(o_f, ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(o_t.float(), ref, rtol=1e-2 if dtype == torch.bfloat16 else 1e-5, atol=1e-2 if dtype == torch.bfloat16 else 1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_patchify_and_tokens(dtype):
    B, S, p, H = 2, 64, 16, 768
    img = rnd(B, 3, S, S, seed=1)
    P = (S // p) ** 2
    a = torch.zeros(B * P, 3 * p * p, device=dev(), dtype=dtype)
    ops.patchify(img, a, p)
    ref = F.unfold(img, kernel_size=p, stride=p).transpose(1, 2).reshape(B * P, 3 * p * p)
    assert torch.equal(a.float(), ref.to(dtype).float())
    po, cls, pos = rnd(B * P, H, seed=2), rnd(H, seed=3), rnd(P + 1, H, seed=4)
    x = torch.zeros(B, P + 1, H, device=dev())
    ops.assemble_tokens(po, cls, pos, x, B, P, H)
    ref = torch.cat([cls.view(1, 1, H).expand(B, 1, H), po.view(B, P, H)], 1) + pos
    assert torch.equal(x, ref)


def test_gather_rows_and_ctx():
    B, N, H = 3, 17, 768
    cap, tag = rnd(B, N, H, seed=1), rnd(B, N, H, seed=2)
    o = torch.zeros(B, H, device=dev(), dtype=torch.bfloat16)
    ops.gather_rows(tag, N * H, o, B, H)
    assert torch.equal(o, tag[:, 0].to(torch.bfloat16))
    cf = torch.zeros(B, N + 1, H, device=dev())
    ct = torch.zeros(B, N + 1, H, device=dev(), dtype=torch.bfloat16)
    ops.assemble_ctx(cap, tag, cf, ct, B, N, H)
    ref = torch.cat([tag[:, 0:1], cap], 1)
    assert torch.equal(cf, ref) and torch.equal(ct, ref.to(torch.bfloat16))


@pytest.mark.parametrize("scale", [1.0, 4.0])
def test_tag_topk_matches_torch(scale):
    """sigmoid -> topk(50) -> count(prob >= 0.2) (modeling_bert.py:1429-1432). Selection runs on the logits: where
    sigmoid saturates in fp32 (scale 4: logits ~ 17) torch's topk-on-probabilities sees exact ties whose order is
    implementation-defined, so there the index comparison is against topk of the logits and the probabilities are
    compared as sorted values."""
    B, V, K = 37, 30522, 50
    logits = rnd(B, V, seed=5, scale=scale)
    ld = 30528
    buf = torch.zeros(B, ld, device=dev())
    buf[:, :V] = logits
    idx = torch.zeros(B, K, device=dev(), dtype=torch.int32)
    prob = torch.zeros(B, K, device=dev())
    n = torch.zeros(B, device=dev(), dtype=torch.int32)
    ops.tag_topk(buf, V, K, 0.2, idx, prob, n)
    rp, ri = torch.sigmoid(logits).topk(K, dim=1)
    lv, li = logits.topk(K, dim=1)
    assert torch.equal(idx.long(), li)
    torch.testing.assert_close(prob, rp, rtol=1e-6, atol=1e-7)
    assert torch.equal(n.long(), (rp >= 0.2).sum(1))
    if scale == 1.0:
        assert torch.equal(idx.long(), ri)


def test_tag_topk_ties_and_small_rows():
    V, K = 100, 50
    logits = torch.zeros(2, 104, device=dev())
    logits[0, :V] = torch.arange(V, device=dev()).float() % 7          # heavy ties
    logits[1, :V] = -torch.arange(V, device=dev()).float()
    idx = torch.zeros(2, K, device=dev(), dtype=torch.int32)
    prob = torch.zeros(2, K, device=dev())
    n = torch.zeros(2, device=dev(), dtype=torch.int32)
    ops.tag_topk(logits, V, K, 0.2, idx, prob, n)
    v0 = logits[0, :V]
    # sorted by value desc, ties by lowest index first
    order = sorted(range(V), key=lambda i: (-float(v0[i]), i))[:K]
    assert idx[0].tolist() == order
    assert idx[1].tolist() == list(range(K))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,N,heads", [(3, 577, 12), (2, 197, 12), (1, 17, 2), (5, 64, 1)])
def test_cls_attention(B, N, heads, dtype):
    """Single-query attention of the last concept-branch block: row 0 of the full attention output."""
    H = heads * 64
    qkv = rnd(B, N, 3 * H, seed=31, dtype=dtype)
    ref = ref_attention(qkv, heads, 0.125)[:, 0]                     # [B, H]
    q = qkv[:, 0, :H]                                                # strided view: pitch N*3H
    out = torch.full((B, 2 * H), float("nan"), device=dev(), dtype=dtype)
    ops.cls_attention(q, qkv, out[:, :H], B, N, heads, 0.125)
    tol = 2e-5 if dtype == torch.float32 else 1e-2
    torch.testing.assert_close(out[:, :H].float(), ref, rtol=tol, atol=tol)
    assert bool(torch.isnan(out[:, H:]).all())                       # pitch respected


@pytest.mark.parametrize("M", [577 * 3, 20000, 4111])
@pytest.mark.parametrize("act", [ops.ACT_NONE, ops.ACT_GELU])
@pytest.mark.parametrize("row_mean", [0.0, 3.0])
@pytest.mark.parametrize("K1", [3072, 768])              # 3072: copy stored from registers; 768: staged in smem, TMA store
def test_linear_ln_emit_and_fold(M, act, row_mean, K1):
    """Folded LayerNorm (vc_linear_ln_emit / vc_linear_ln_fold): the producer GEMM's fp32 result is bit-identical to the plain
    residual GEMM, its bf16 copy is the rounded result, its statistics are the fp32 row sums; the consumer GEMM on the raw copy
    reproduces act(LayerNorm(x) W^T + b) as closely as the two-kernel path (LayerNorm kernel + plain GEMM) does."""
    H, N2 = 768, 2304
    a, w, b = rnd(M, K1, seed=5, dtype=torch.bfloat16), rnd(H, K1, seed=6, scale=0.03, dtype=torch.bfloat16), rnd(H, seed=7)
    x = rnd(M, H, seed=8) + row_mean
    plain = torch.empty(M, H, device=dev())
    ops.linear(a, w, b, plain, resid=x, impl="tc", tile_n=512)
    out = x.clone()                                              # in place, as the residual stream is updated
    xb = torch.zeros(M, H, device=dev(), dtype=torch.bfloat16)
    stats = torch.zeros(M, 3, 2, device=dev())
    ops.linear_ln_emit(a, w, b, out, out, xb, stats)
    assert torch.equal(out, plain)
    assert torch.equal(xb, out.to(torch.bfloat16))
    s = stats.sum(1)
    torch.testing.assert_close(s[:, 0], out.double().sum(1).float(), rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(s[:, 1], (out.double() ** 2).sum(1).float(), rtol=1e-5, atol=1e-3)
    # consumer
    eps = 1e-6
    gamma, beta = 1.0 + 0.1 * rnd(H, seed=11), 0.05 * rnd(H, seed=12)
    w2, b2 = rnd(N2, H, seed=13, scale=0.03), rnd(N2, seed=14)
    wf = (w2 * gamma).to(torch.bfloat16)
    colsum = wf.float().sum(1).contiguous()
    bias_f = (b2 + w2 @ beta).contiguous()
    got = torch.empty(M, N2, device=dev(), dtype=torch.bfloat16)
    ops.linear_ln_fold(xb, wf, bias_f, colsum, stats, 3, eps, got, act=act)
    ln_ref = torch.nn.functional.layer_norm(out.double(), (H,), gamma.double(), beta.double(), eps)
    ref = ln_ref @ w2.double().t() + b2.double()
    if act == ops.ACT_GELU:
        ref = torch.nn.functional.gelu(ref)
    ref = ref.float()
    # the two-kernel path it replaces
    ln_t = torch.empty(M, H, device=dev(), dtype=torch.bfloat16)
    ops.layernorm(out, gamma, beta, eps, out_t=ln_t)
    two = torch.empty(M, N2, device=dev(), dtype=torch.bfloat16)
    ops.linear(ln_t, w2.to(torch.bfloat16), b2, two, act=act, impl="tc", tile_n=512)
    e_fold = float((got.float() - ref).norm() / ref.norm())
    e_two = float((two.float() - ref).norm() / ref.norm())
    print("folded LN rel err %.3g, LayerNorm kernel + GEMM rel err %.3g" % (e_fold, e_two))
    assert e_fold < 6e-3 and e_fold < 1.5 * e_two + 1e-4
    torch.testing.assert_close(got.float(), ref, rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("M", [578 * 2, 9000])
@pytest.mark.parametrize("K1", [3072, 768])
def test_linear_ln_emit_postln_residual(M, K1):
    """vc_linear_ln_emit_postln: the residual added is LayerNorm(raw rows) evaluated in the epilogue from the raw fp32 tile and
    the partial sums a previous emit left -- equal to materialising the LayerNorm first (post-LN BertLayer, the decoder prefill)."""
    H = 768
    a, w, b = rnd(M, K1, seed=5, dtype=torch.bfloat16), rnd(H, K1, seed=6, scale=0.03, dtype=torch.bfloat16), rnd(H, seed=7)
    raw = rnd(M, H, seed=8) * 1.7 + 0.4
    gamma, beta, eps = 1.0 + 0.1 * rnd(H, seed=11), 0.05 * rnd(H, seed=12), 1e-12
    # the statistics exactly as a producer would have emitted them: per-256-column partial (sum, sum of squares)
    rstats = torch.stack([raw.view(M, 3, 256).sum(2), (raw.view(M, 3, 256) ** 2).sum(2)], dim=2).contiguous()
    resid = torch.nn.functional.layer_norm(raw, (H,), gamma, beta, eps)
    plain = torch.empty(M, H, device=dev())
    ops.linear(a, w, b, plain, resid=resid, impl="tc", tile_n=512)
    out = torch.empty(M, H, device=dev())
    xb = torch.zeros(M, H, device=dev(), dtype=torch.bfloat16)
    stats = torch.zeros(M, 3, 2, device=dev())
    ops.linear_ln_emit(a, w, b, out, raw, xb, stats, resid_ln=(rstats, 3, gamma, beta, eps))
    scale = float(plain.abs().max())
    assert float((out - plain).abs().max()) <= 2e-5 * scale
    assert torch.equal(xb, out.to(torch.bfloat16))
    s = stats.sum(1)
    torch.testing.assert_close(s[:, 0], out.double().sum(1).float(), rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(s[:, 1], (out.double() ** 2).sum(1).float(), rtol=1e-5, atol=1e-3)


# ---- split-bf16 operands of the decode-step GEMMs (include/vitcap_b200.h, VC_OPERAND_BF16X3) ---------------------------------
def _split_ref(y):
    hi = y.to(torch.bfloat16)
    lo = (y - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo, hi], dim=-1)


@pytest.mark.parametrize("rows,K", [(1, 8), (1024, 3072), (77, 768)])
def test_split_bf16x3_bit_exact(rows, K):
    x = rnd(rows, K + 8, seed=5, scale=3.0)[:, :K]                # pitched input view
    out = torch.zeros(rows, 3 * K + 16, device=dev(), dtype=torch.bfloat16)
    ops.split_bf16x3(x, out[:, :3 * K])
    assert torch.equal(out[:, :3 * K].view(torch.int16), _split_ref(x).view(torch.int16))
    assert int(out[:, 3 * K:].abs().sum()) == 0                   # nothing written past the operand


def test_split_bf16x3_refuses_narrow_output():
    x = rnd(4, 64, seed=1)
    out = torch.zeros(4, 128, device=dev(), dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops._check(ops.load_library().vc_split_bf16x3(ops._ptr(x), 64, ops._ptr(out), 128, 4, 64, ops._stream()), "vc_split_bf16x3")


def test_layernorm_x3_is_the_split_of_the_fp32_output():
    rows, H = 1000, 768
    x, g, b = rnd(rows, H, seed=1, scale=2.0), rnd(H, seed=2), rnd(H, seed=3)
    o_t = torch.zeros(2 * rows, 3 * H, device=dev(), dtype=torch.bfloat16)[1::2]     # strided rows, as the MASK rows are
    o_f = torch.zeros(rows, H, device=dev())
    ops.layernorm(x, g, b, 1e-12, out_t=o_t, out_f=o_f, x3=True)
    torch.testing.assert_close(o_f, F.layer_norm(x, (H,), g, b, 1e-12), rtol=1e-5, atol=1e-5)
    assert torch.equal(o_t.contiguous().view(torch.int16), _split_ref(o_f).view(torch.int16))
    # the first H columns are what the plain bf16 flavour writes
    o_b = torch.zeros(rows, H, device=dev(), dtype=torch.bfloat16)
    ops.layernorm(x, g, b, 1e-12, out_t=o_b)
    assert torch.equal(o_t[:, :H].contiguous().view(torch.int16), o_b.view(torch.int16))


@pytest.mark.parametrize("M,N,K,act,resid", [(1024, 3072, 768, ops.ACT_GELU, False), (1024, 768, 3072, ops.ACT_NONE, True),
                                             (512, 30522, 768, ops.ACT_NONE, False), (100, 768, 768, ops.ACT_GELU, False)])
def test_linear_bf16x3_reaches_fp32_operand_precision(M, N, K, act, resid):
    """Three tensor-core products on [hi | lo | hi] x [w_hi | w_hi | w_lo]: error against the fp64 product of the UNROUNDED
    operands must be ~fp32-accumulation sized (< 2e-5 of the output scale), two orders below the plain bf16 GEMM's."""
    a, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3)
    r = rnd(M, N, seed=4) if resid else None
    a3 = torch.zeros(M, 3 * K, device=dev(), dtype=torch.bfloat16)
    ops.split_bf16x3(a, a3)
    w3 = ops.split_weight_bf16x3(w)
    assert w3.shape == (N, 3 * K)
    ldo = (N + 63) // 64 * 64
    out = torch.zeros(M, ldo, device=dev())
    ops.linear(a3, w3, b, out[:, :N], act=act, resid=r, ldo=ldo)
    y = a.double() @ w.double().t() + b.double()
    if act == ops.ACT_GELU:
        y = y * 0.5 * (1 + torch.erf(y / math.sqrt(2)))
    if resid:
        y = y + r.double()
    scale = float(y.abs().max())
    err3 = float((out[:, :N].double() - y).abs().max()) / scale
    out1 = torch.zeros(M, ldo, device=dev())
    ops.linear(a.to(torch.bfloat16), w.to(torch.bfloat16), b, out1[:, :N], act=act, resid=r, ldo=ldo)
    err1 = float((out1[:, :N].double() - y).abs().max()) / scale
    print("bf16x3 max error / scale %.3g, plain bf16 %.3g" % (err3, err1))
    assert err3 < 2e-5 and err3 < err1 / 50


@pytest.mark.parametrize("M,N,K", [(1024, 768, 3072), (2, 768, 3072), (300, 200, 64), (129, 64, 192), (16, 768, 768)])
def test_linear_x3_equals_the_k_concatenated_gemm(M, N, K):
    """vc_linear_x3 (four distinct tiles per k-block, three products) against vc_linear on the same split operands walked as one
    K' = 3K product: identical mathematics, fp32 summation order differs (measured 4e-6 of the output scale over K' = 9216 terms; tolerance 2e-5), and against the
    fp64 product of the unrounded operands."""
    a, w, b, r = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3), rnd(M, N + 8, seed=4)[:, :N]
    a3 = torch.zeros(M, 3 * K, device=dev(), dtype=torch.bfloat16)
    ops.split_bf16x3(a, a3)
    w3 = ops.split_weight_bf16x3(w)
    ldo = (N + 63) // 64 * 64
    o1 = torch.zeros(M, ldo, device=dev())
    o2 = torch.full((M, ldo), 7.0, device=dev())
    ops.linear(a3, w3, b, o1[:, :N], resid=r, ldo=ldo)
    ops.linear_x3(a3, w3, b, o2[:, :N], r)
    y = a.double() @ w.double().t() + b.double() + r.double()
    scale = float(y.abs().max())
    assert float((o1[:, :N] - o2[:, :N]).abs().max()) / scale < 2e-5
    assert float((o2[:, :N].double() - y).abs().max()) / scale < 2e-5
    assert bool((o2[:, N:] == 7.0).all())                          # columns past N untouched
    o3 = torch.zeros(M, ldo, device=dev())
    ops.linear_x3(a3, w3, b, o3[:, :N], r)
    assert torch.equal(o2[:, :N], o3[:, :N])                       # deterministic
