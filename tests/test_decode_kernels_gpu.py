"""Kernel-level parity of the decode-step entry points (through the C ABI) against torch expressions / the CPU oracle's
search loops (oracle/port.py, restating modeling_utils.py:768-1180) driven by identical synthetic logits."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import philox, port  # noqa: E402
from vitcap_b200 import ops  # noqa: E402

DEV = "cuda:0"


def _rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ------------------------------------------------------------------------------------------------ decode attention
def _decode_attention_ref(ctx, stepq, anc, B, C, heads, E, cur_len, scale):
    """fp32 torch restatement: rows (2r, 2r+1) = (token, MASK) queries of sequence r = b*E+e (modeling_bert.py:303-340 with the
    structural mask of modeling_bert.py:1494-1501): context keys of image b, cached caption keys through the ancestor table,
    this step's token key, and -- for the MASK query only -- this step's MASK key."""
    H = heads * 64
    R = B * E
    step = cur_len - 1
    ctx = ctx.float().view(B, C, 3, heads, 64)
    sq = stepq.float().view(stepq.shape[0], 2 * R, 3, heads, 64)
    out = torch.zeros(2 * R, heads, 64)
    for r in range(R):
        b = r // E
        ks = [ctx[b, :, 1]]
        vs = [ctx[b, :, 2]]
        for j in range(step):
            src = int(anc[j, r]) if anc is not None else r
            ks.append(sq[j, 2 * src, 1][None])
            vs.append(sq[j, 2 * src, 2][None])
        ks.append(sq[step, 2 * r, 1][None]); vs.append(sq[step, 2 * r, 2][None])
        ks.append(sq[step, 2 * r + 1, 1][None]); vs.append(sq[step, 2 * r + 1, 2][None])
        K = torch.cat(ks, 0)            # [keys, heads, 64]
        V = torch.cat(vs, 0)
        for w in range(2):
            q = sq[step, 2 * r + w, 0]  # [heads, 64]
            s = torch.einsum("hd,khd->hk", q, K) * scale
            if w == 0:
                s[:, -1] = -1e30
            p = torch.softmax(s, dim=-1)
            out[2 * r + w] = torch.einsum("hk,khd->hd", p, V)
    return out.view(2 * R, H)


DA_CASES = [(2, 578, 12, 1, 1), (2, 578, 12, 1, 19), (3, 198, 12, 5, 7), (2, 578, 12, 4, 10), (1, 18, 2, 1, 3),
            (2, 146, 12, 8, 19), (1, 578, 12, 10, 5), (2, 50, 3, 3, 2), (1, 578, 12, 2, 40)]


def _da_inputs(B, C, heads, E, cur_len, dtype, seed=0):
    H = heads * 64
    R = B * E
    max_len = max(20, cur_len + 1)
    ctx = _rnd(B, C, 3 * H, seed=seed).to(dtype)
    stepq = _rnd(max_len, 2 * R, 3 * H, seed=seed + 1).to(dtype)
    g = torch.Generator().manual_seed(seed + 2)
    anc = torch.zeros(max_len, R, dtype=torch.int32)
    for j in range(max_len):
        for r in range(R):
            b = r // E
            anc[j, r] = b * E + int(torch.randint(0, E, (1,), generator=g))
    return ctx, stepq, anc


@pytest.mark.parametrize("use_anc", [False, True])
@pytest.mark.parametrize("B,C,heads,E,cur_len", DA_CASES)
def test_decode_attention_bf16_mma(B, C, heads, E, cur_len, use_anc):
    scale = 0.125
    ctx, stepq, anc = _da_inputs(B, C, heads, E, cur_len, torch.bfloat16)
    ref = _decode_attention_ref(ctx, stepq, anc if use_anc else None, B, C, heads, E, cur_len, scale)
    H = heads * 64
    out = torch.full((2 * B * E, H), float("nan"), dtype=torch.bfloat16, device=DEV)
    d_anc = anc.to(DEV) if use_anc else None
    ops.decode_attention(ctx.to(DEV), stepq.to(DEV), d_anc, out, B, C, heads, E, cur_len, scale)
    torch.cuda.synchronize()
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    # bf16 probabilities and bf16 output rounding: |err| <~ 2^-8 of the value scale (outputs are O(0.1..1))
    assert float((got - ref).abs().max()) < 2.5e-2, float((got - ref).abs().max())
    assert float((got - ref).norm() / ref.norm()) < 6e-3
    # cross-check against the CUDA-core kernel on the same bf16 inputs
    out2 = torch.empty_like(out)
    ops.decode_attention(ctx.to(DEV), stepq.to(DEV), d_anc, out2, B, C, heads, E, cur_len, scale, impl="simt")
    assert float((out2.float().cpu() - ref).abs().max()) < 2.5e-2


@pytest.mark.parametrize("B,C,heads,E,cur_len", [(2, 578, 12, 1, 5), (2, 198, 12, 4, 19), (1, 18, 2, 5, 3)])
def test_decode_attention_fp32_exact(B, C, heads, E, cur_len):
    scale = 0.125
    ctx, stepq, anc = _da_inputs(B, C, heads, E, cur_len, torch.float32, seed=5)
    ref = _decode_attention_ref(ctx, stepq, anc, B, C, heads, E, cur_len, scale)
    out = torch.empty(2 * B * E, heads * 64, dtype=torch.float32, device=DEV)
    ops.decode_attention(ctx.to(DEV), stepq.to(DEV), anc.to(DEV), out, B, C, heads, E, cur_len, scale)
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), atol=2e-5, rtol=1e-4)


def test_decode_attention_peaky_scores_bf16():
    """Large score range: the lazy rescale path (running reference moves by more than 2^8) must stay exact."""
    B, C, heads, E, cur_len = 1, 578, 12, 2, 6
    ctx, stepq, anc = _da_inputs(B, C, heads, E, cur_len, torch.bfloat16, seed=9)
    ctx = (ctx.float() * 3.0).to(torch.bfloat16)
    stepq = (stepq.float() * 3.0).to(torch.bfloat16)
    ref = _decode_attention_ref(ctx, stepq, anc, B, C, heads, E, cur_len, 0.125)
    out = torch.empty(2 * B * E, heads * 64, dtype=torch.bfloat16, device=DEV)
    ops.decode_attention(ctx.to(DEV), stepq.to(DEV), anc.to(DEV), out, B, C, heads, E, cur_len, 0.125)
    got = out.float().cpu()
    assert float((got - ref).abs().max()) < 8e-2 and float((got - ref).norm() / ref.norm()) < 8e-3


# ------------------------------------------------------------------------------------------------ embedding
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_embed_ln(dtype):
    R, H, V, L, cur_len, mask_id = 37, 768, 500, 20, 7, 103
    word, pos, typ = _rnd(V, H, seed=1, scale=0.05), _rnd(64, H, seed=2, scale=0.05), _rnd(H, seed=3, scale=0.05)
    gam, bet = 1 + _rnd(H, seed=4, scale=0.1), _rnd(H, seed=5, scale=0.1)
    ids = torch.randint(0, V, (R, L), generator=torch.Generator().manual_seed(6), dtype=torch.int32)
    out_f = torch.empty(2 * R, H, device=DEV)
    out_t = out_f if dtype == torch.float32 else torch.empty(2 * R, H, device=DEV, dtype=dtype)
    ops.embed_ln(ids.to(DEV), cur_len, mask_id, word.to(DEV), pos.to(DEV), typ.to(DEV), gam.to(DEV), bet.to(DEV), 1e-12, out_f, out_t, R)
    tok = torch.stack([ids[:, cur_len - 1].long(), torch.full((R,), mask_id)], 1).reshape(-1)
    p = torch.tensor([cur_len - 1, cur_len]).repeat(R)
    ref = torch.nn.functional.layer_norm(word[tok] + pos[p] + typ, (H,), gam, bet, 1e-12)
    np.testing.assert_allclose(out_f.cpu().numpy(), ref.numpy(), atol=2e-5)
    if dtype == torch.bfloat16:
        np.testing.assert_allclose(out_t.float().cpu().numpy(), ref.numpy(), atol=2e-2)


# ------------------------------------------------------------------------------------------------ search loops
class _TableLogits:
    """logits(row) = table[hash(prefix ids of the row)]: a pure gather, so the CPU oracle loop and the GPU kernels see
    bit-identical logits as long as their token prefixes agree -- any divergence in ids / reordering shows up at once."""

    def __init__(self, V, T=509, seed=0, eos=102, eos_boost=6.0, scale=3.0):
        g = torch.Generator().manual_seed(seed)
        self.table = torch.randn(T, V, generator=g) * scale
        boost = torch.rand(T, generator=g) < 0.12
        self.table[boost, eos] += eos_boost
        self.T = T
        self.V = V
        self._dev = {}

    def index(self, ids):
        L = ids.shape[1]
        w = torch.arange(1, L + 1, device=ids.device, dtype=torch.int64) * 7919
        return ((ids.to(torch.int64) * w).sum(1) + 31 * L) % self.T

    def __call__(self, ids, beam_idx=None):
        key = str(ids.device)
        if key not in self._dev:
            self._dev[key] = self.table.to(ids.device)
        return self._dev[key][self.index(ids)]


def _greedy_kernels(tl, R, max_len, bos, pad, eos_ids, do_sample=False, temperature=1.0, top_k=0, top_p=1.0, seed=0,
                    seed_on_device=False):
    V = tl.V
    seed_dev = None
    if seed_on_device:                      # the seed word is read by the kernel; the scalar argument is then ignored
        s64 = seed & 0xFFFFFFFFFFFFFFFF
        seed_dev = torch.tensor([s64 - (1 << 64) if s64 >= (1 << 63) else s64], dtype=torch.int64, device=DEV)
        seed = 12345
    ldl = (V + 63) // 64 * 64
    ids = torch.zeros(R, max_len, dtype=torch.int32, device=DEV)
    ids[:, 0] = bos
    unf = torch.ones(R, dtype=torch.int32, device=DEV)
    sum_lp = torch.zeros(R, device=DEV)
    n_steps = torch.zeros(R, dtype=torch.int32, device=DEV)
    logits = torch.zeros(R, ldl, device=DEV)
    eos = torch.tensor(eos_ids, dtype=torch.int32, device=DEV)
    for cur_len in range(1, max_len):
        logits[:, :V] = tl(ids[:, :cur_len])
        t = temperature
        if do_sample and (top_k > 0 or top_p < 1.0):
            ops.filter_logits(logits, V, R, 1.0 / temperature, top_k, top_p)
            t = 1.0
        ops.token_step(logits, V, R, do_sample, t, seed, cur_len, pad, eos, ids, unf, sum_lp, n_steps, seed_dev=seed_dev)
    out_ids = torch.zeros(R, max_len, dtype=torch.int64, device=DEV)
    out_lp = torch.zeros(R, device=DEV)
    ops.greedy_finalize(ids, unf, sum_lp, n_steps, eos_ids[0], R, out_ids, out_lp)
    return out_ids.cpu(), out_lp.cpu()


@pytest.mark.parametrize("V,R", [(3000, 64), (30522, 16)])
def test_greedy_kernels_vs_oracle_loop(V, R):
    tl = _RowTable(_TableLogits(V, seed=V, eos_boost=6.0 if V < 10000 else 14.0), 1)
    rids, rlp = port.greedy_or_sample(tl, R, 20, 101, 0, [102])
    ids, lp = _greedy_kernels(tl, R, 20, 101, 0, [102])
    assert torch.equal(ids, rids[:, 0])
    np.testing.assert_allclose(lp.numpy(), rlp[:, 0].numpy(), atol=1e-5)
    assert (ids == 0).any() and (ids[:, -1] == 102).any()      # both early EOS (PAD fill) and forced EOS occur


def test_sampling_seed_from_device_memory_equals_scalar_seed():
    """vc_token_step with seed_dev (what a captured decode loop uses) draws the same noise as the scalar seed."""
    tl = _RowTable(_TableLogits(3000, seed=5, eos_boost=6.0), 1)
    for seed in (7, 0x9E3779B97F4A7C15, (1 << 64) - 3):
        a = _greedy_kernels(tl, 32, 20, 101, 0, [102], do_sample=True, seed=seed)
        b = _greedy_kernels(tl, 32, 20, 101, 0, [102], do_sample=True, seed=seed, seed_on_device=True)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    c = _greedy_kernels(tl, 32, 20, 101, 0, [102], do_sample=True, seed=8, seed_on_device=True)
    assert not torch.equal(a[0], c[0])


class _RowTable:
    """Makes the hash depend on the image (row // rows_per_image) too, so different images get different captions."""

    def __init__(self, tl, rows_per_image):
        self.tl, self.n, self.V = tl, rows_per_image, tl.V

    def __call__(self, ids, beam_idx=None):
        img = (torch.arange(ids.shape[0], device=ids.device, dtype=ids.dtype).unsqueeze(1) // self.n) % 97
        return self.tl(torch.cat([img, ids], 1))


def test_sampling_kernels_vs_oracle_loop_same_noise():
    V, R, seed = 3000, 48, 4321
    tl = _RowTable(_TableLogits(V, seed=3, scale=2.0), 1)
    for (temp, top_k, top_p) in [(1.0, 0, 1.0), (0.7, 40, 0.9), (1.3, 0, 0.8), (1.0, 5, 1.0)]:
        rids, rlp = port.greedy_or_sample(tl, R, 20, 101, 0, [102], do_sample=True, temperature=temp, top_k=top_k, top_p=top_p,
                                          sampler=philox.make_sampler(seed))
        ids, lp = _greedy_kernels(tl, R, 20, 101, 0, [102], do_sample=True, temperature=temp, top_k=top_k, top_p=top_p, seed=seed)
        same = (ids == rids[:, 0]).all(1)
        # Gumbel noise is computed with different log implementations on CPU and GPU: a sample may flip at a near-tie
        assert float(same.float().mean()) >= 0.9, (temp, top_k, top_p, float(same.float().mean()))
        np.testing.assert_allclose(lp[same].numpy(), rlp[:, 0][same].numpy(), atol=2e-5)


@pytest.mark.parametrize("top_k,top_p,temp", [(50, 1.0, 1.0), (0, 0.9, 1.0), (20, 0.5, 0.7), (1, 1.0, 1.0), (0, 0.05, 2.0)])
def test_filter_logits_vs_reference_filter(top_k, top_p, temp):
    R, V = 33, 30522
    x = _rnd(R, V, seed=11, scale=2.5)
    ldl = (V + 63) // 64 * 64
    d = torch.zeros(R, ldl, device=DEV)
    d[:, :V] = x.to(DEV)
    ops.filter_logits(d, V, R, 1.0 / temp, top_k, top_p)
    ref = port.top_k_top_p_filtering(x / temp if temp != 1.0 else x.clone(), top_k=top_k, top_p=top_p)
    got = d[:, :V].cpu()
    kept_ref, kept_got = torch.isfinite(ref), torch.isfinite(got)
    # the top-p boundary is a cumulative fp32 sum: allow the kept set to differ by at most one boundary token per row
    diff = (kept_ref != kept_got).sum(1)
    assert int(diff.max()) <= 1, diff
    both = kept_ref & kept_got
    np.testing.assert_allclose(got[both].numpy(), ref[both].numpy(), rtol=1e-6, atol=1e-6)
    if top_p >= 1.0:
        assert int(diff.max()) == 0


def _beam_kernels(tl, B, nb, max_len, bos, pad, eos_ids, keep, length_penalty):
    V = tl.V
    R, K = B * nb, 2 * nb
    ldl = (V + 63) // 64 * 64
    f32, i32 = torch.float32, torch.int32
    st = {
        "ids": torch.zeros(R, max_len, dtype=i32, device=DEV),
        "beam_scores": torch.tensor(([0.0] + [-1e9] * (nb - 1)) * B, device=DEV, dtype=f32),
        "done": torch.zeros(B, device=DEV, dtype=i32), "anc": torch.zeros(max_len, R, device=DEV, dtype=i32),
        "hyp_score": torch.zeros(B, keep, device=DEV, dtype=torch.float64), "hyp_len": torch.zeros(B, keep, device=DEV, dtype=i32),
        "hyp_ids": torch.zeros(B, keep, max_len, device=DEV, dtype=i32), "hyp_count": torch.zeros(B, device=DEV, dtype=i32),
        "worst": torch.full((B,), 1e9, device=DEV, dtype=torch.float64),
    }
    st["ids"][:, 0] = bos
    cand_val = torch.zeros(R, K, device=DEV, dtype=f32)
    cand_idx = torch.zeros(R, K, device=DEV, dtype=i32)
    row_max, row_logsum = torch.zeros(R, device=DEV), torch.zeros(R, device=DEV)
    logits = torch.zeros(R, ldl, device=DEV)
    eos = torch.tensor(eos_ids, dtype=i32, device=DEV)
    anc_hist = []
    for cur_len in range(1, max_len):
        logits[:, :V] = tl(st["ids"][:, :cur_len])
        ops.beam_row_topk(logits, V, R, K, cand_val, cand_idx, row_max, row_logsum)
        ops.beam_advance(st, cand_val, cand_idx, row_max, row_logsum, B, nb, V, cur_len, keep, length_penalty, pad, eos)
        anc_hist.append(st["anc"].clone())
    out_ids = torch.zeros(B, keep, max_len, dtype=torch.int64, device=DEV)
    out_lp = torch.zeros(B, keep, device=DEV)
    ops.beam_finalize(st, B, keep, pad, eos_ids[0], out_ids, out_lp)
    return out_ids.cpu(), out_lp.cpu(), st


@pytest.mark.parametrize("B,nb,keep,lp,V", [(24, 4, 1, 1.0, 3000), (16, 3, 3, 0.6, 3000), (8, 8, 4, 1.4, 3000), (6, 4, 2, 1.0, 30522),
                                            (12, 2, 1, 0.0, 997)])
def test_beam_kernels_vs_oracle_loop(B, nb, keep, lp, V):
    tl = _RowTable(_TableLogits(V, seed=100 + nb, eos_boost=5.0), nb)
    rids, rlp = port.beam_search(tl, B, 20, 101, 0, [102], nb, V, length_penalty=lp, num_keep_best=keep)
    ids, glp, st = _beam_kernels(tl, B, nb, 20, 101, 0, [102], keep, lp)
    assert torch.equal(ids, rids), (ids[0], rids[0])
    np.testing.assert_allclose(glp.numpy(), rlp.numpy(), atol=2e-5, rtol=1e-5)
    # the ancestor table must reproduce every live beam's prefix: ids[r, j+1] was generated by row anc[j, r] at step j
    anc = st["anc"].cpu()[:19]                                        # one row per decode step (max_len - 1 steps)
    assert int(anc.min()) >= 0 and int(anc.max()) < B * nb
    rows = torch.arange(B * nb)
    assert torch.equal(anc // nb, (rows // nb).expand_as(anc))        # beams never cross images


# ------------------------------------------------------------------------------------------------ fused decode step kernels
def _bf(x):
    return x.to(torch.bfloat16)


def _split_pair(x):
    hi = _bf(x)
    return hi, _bf(x - hi.float())


def _x3_operands(M, N, Kt, seed, wscale=0.05):
    """(a fp32, w fp32, A3 = [a_hi | a_lo | junk], W3 = [w_hi | w_hi | w_lo], fp32 reference of the three products)."""
    a, w = _rnd(M, Kt, seed=seed), _rnd(N, Kt, seed=seed + 1, scale=wscale)
    a_hi, a_lo = _split_pair(a)
    w_hi, w_lo = _split_pair(w)
    A3 = torch.cat([a_hi, a_lo, _bf(torch.full((M, Kt), 7.0))], 1).contiguous()       # the third block must never be read
    W3 = ops.split_weight_bf16x3(w)
    ref = a_hi.float() @ w_hi.float().t() + a_lo.float() @ w_hi.float().t() + a_hi.float() @ w_lo.float().t()
    return a, w, A3, W3, ref


@pytest.mark.parametrize("M,N,Kt,x3,splits", [(1024, 768, 768, False, 1), (1024, 768, 768, False, 6), (1000, 768, 3072, True, 6),
                                              (512, 768, 768, True, 12), (130, 2304, 768, False, 3), (1, 768, 768, True, 2),
                                              (5000, 768, 768, False, 2)])
def test_dec_linear_partial_planes(M, N, Kt, x3, splits):
    """vc_dec_linear VC_DEC_PARTIAL: plane s = the product over the s-th K slice (fp32, no bias); their sum = the full product.
    Covers ragged M, a single row, and more tiles than clusters (the producer then waits for the epilogue's staging tiles)."""
    if x3:
        a, w, A, W, ref = _x3_operands(M, N, Kt, seed=M + N)
    else:
        A, W = _bf(_rnd(M, Kt, seed=M)), _bf(_rnd(N, Kt, seed=N, scale=0.05))
        ref = A.float() @ W.float().t()
    m_pad = (M + 127) // 128 * 128
    out = torch.full((splits, m_pad, N), float("nan"), device=DEV)
    ops.dec_linear(ops.DEC_PARTIAL, A.to(DEV), W.to(DEV), None, out, M=M, x3=x3, splits=splits, m_pad=m_pad)
    got = out[:, :M].cpu()
    assert torch.isfinite(got).all()
    scale = float(ref.abs().max())
    assert float((got.sum(0) - ref).abs().max()) <= 3e-5 * scale
    ks = Kt // splits
    for s in range(splits):
        sl = slice(s * ks, (s + 1) * ks)
        if x3:
            a_hi, a_lo = A[:, :Kt].float(), A[:, Kt:2 * Kt].float()
            w_hi, w_lo = W[:, :Kt].float(), W[:, 2 * Kt:].float()
            part = a_hi[:, sl] @ w_hi[:, sl].t() + a_lo[:, sl] @ w_hi[:, sl].t() + a_hi[:, sl] @ w_lo[:, sl].t()
        else:
            part = A[:, sl].float() @ W[:, sl].float().t()
        assert float((got[s] - part).abs().max()) <= 3e-5 * scale, s


@pytest.mark.parametrize("M,N,Kt,x3,mode", [(1024, 2304, 768, False, "bf16"), (1024, 3072, 768, True, "split"), (777, 3072, 768, False, "gelu"),
                                            (5120, 3072, 768, True, "split"), (200, 768, 768, True, "bf16"), (64, 128, 64, False, "gelu")])
def test_dec_linear_bf16_and_split_epilogues(M, N, Kt, x3, mode):
    """VC_DEC_BF16 / VC_DEC_GELU_BF16 / VC_DEC_GELU_SPLIT against fp32 torch on the same operands. The split pair (hi, lo) must
    carry the GELU output to ~2^-16 where the plain bf16 output stops at 2^-9, and hi must be the plain bf16 rounding."""
    if x3:
        a, w, A, W, ref = _x3_operands(M, N, Kt, seed=M + N)
    else:
        A, W = _bf(_rnd(M, Kt, seed=M)), _bf(_rnd(N, Kt, seed=N, scale=0.05))
        ref = A.float() @ W.float().t()
    bias = _rnd(N, seed=9)
    ref = ref + bias
    if mode != "bf16":
        ref = port.gelu_fast(ref)
    scale = float(ref.abs().max())
    if mode == "split":
        out = torch.full((M, 3 * N), 9.0, device=DEV, dtype=torch.bfloat16)
        ops.dec_linear(ops.DEC_GELU_SPLIT, A.to(DEV), W.to(DEV), bias.to(DEV), out, M=M, x3=x3)
        o = out.cpu()
        hi, lo = o[:, :N].float(), o[:, N:2 * N].float()
        assert bool((o[:, 2 * N:] == 9.0).all())                               # the third block is left alone
        assert float((hi + lo - ref).abs().max()) <= 1.2e-3 * scale            # MUFU.TANH (2^-11) dominates
        assert float((hi - ref).abs().max()) <= 6e-3 * scale
        assert float((hi + lo - ref).abs().mean()) < 0.2 * float((hi - ref).abs().mean())
        assert float((lo.abs() > hi.abs() * 2.0 ** -7 + 1e-30).float().mean()) == 0.0      # lo is a rounding remainder
    else:
        out = torch.full((M, N), 9.0, device=DEV, dtype=torch.bfloat16)
        ops.dec_linear(ops.DEC_BF16 if mode == "bf16" else ops.DEC_GELU_BF16, A.to(DEV), W.to(DEV), bias.to(DEV), out, M=M, x3=x3)
        assert float((out.cpu().float() - ref).abs().max()) <= 6e-3 * scale


@pytest.mark.parametrize("rows,splits,gelu,resid,mode", [(1024, 6, False, True, "split"), (1000, 1, False, True, "bf16"),
                                                         (512, 12, True, False, "split"), (3, 2, False, True, "none")])
def test_finish_ln(rows, splits, gelu, resid, mode):
    H = 768
    m_pad = (rows + 127) // 128 * 128
    part = _rnd(splits, m_pad, H, seed=rows, scale=0.5)
    bias, gamma, beta = _rnd(H, seed=1), 1.0 + 0.1 * _rnd(H, seed=2), 0.1 * _rnd(H, seed=3)
    res = _rnd(rows, H, seed=4) if resid else None
    x = part[:, :rows].sum(0) + bias
    if gelu:
        x = port.gelu_fast(x)
    if resid:
        x = x + res
    ref = torch.nn.functional.layer_norm(x, (H,), gamma, beta, 1e-12)
    out_f = torch.empty(rows, H, device=DEV)
    out_t = None if mode == "none" else torch.full((rows, 3 * H if mode == "split" else H), 5.0, device=DEV, dtype=torch.bfloat16)
    ops.finish_ln(part.to(DEV), splits, bias.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-12, rows, resid=res.to(DEV) if resid else None,
                  gelu=gelu, out_f=out_f, out_t=out_t, split=(mode == "split"))
    of = out_f.cpu()
    assert float((of - ref).abs().max()) <= (2e-3 if gelu else 2e-5) * float(ref.abs().max())
    if mode == "bf16":
        assert torch.equal(out_t.cpu(), of.to(torch.bfloat16))
    elif mode == "split":
        hi, lo = _split_pair(of)
        o = out_t.cpu()
        assert torch.equal(o[:, :H], hi) and torch.equal(o[:, H:2 * H], lo) and bool((o[:, 2 * H:] == 5.0).all())


@pytest.mark.parametrize("R,V,x3", [(512, 30522, True), (40, 30522, False), (130, 3000, True), (7, 500, False)])
def test_vocab_argmax_partials_and_token_step(R, V, x3):
    """Greedy decoding without materialised logits: vc_dec_vocab_argmax + vc_token_step_partials against the explicit
    logits -> argmax -> log_softmax -> gather (modeling_utils.py:849-853) and the state update of vc_token_step."""
    Kt = 768
    if x3:
        a, w, A, W, ref = _x3_operands(R, V, Kt, seed=R + V, wscale=0.03)
    else:
        A, W = _bf(_rnd(R, Kt, seed=R)), _bf(_rnd(V, Kt, seed=V, scale=0.03))
        ref = A.float() @ W.float().t()
    bias = 0.5 * _rnd(V, seed=5)
    bias[102] += 3.0                                          # a popular [SEP] so that rows finish
    ref = ref + bias
    n_part = ops.vocab_partials(V)
    part = torch.full((R, n_part, 4), float("nan"), device=DEV)
    ops.dec_vocab_argmax(A.to(DEV), W.to(DEV), bias.to(DEV), part, M=R, x3=x3)
    p = part.cpu()
    # every partial: max / arg max / sum over its own columns
    for tile in (0, n_part // 2 - 1):
        for g in range(2):
            T = ops.VOCAB_TILE
            cols = [c for ch in range(g, 7, 2) for c in range(tile * T + ch * 32, min(tile * T + ch * 32 + 32, (tile + 1) * T)) if c < V]
            if not cols:
                assert bool((p[:, 2 * tile + g, 2] == 0).all())
                continue
            sub = ref[:, cols]
            np.testing.assert_allclose(p[:, 2 * tile + g, 0].numpy(), sub.max(1).values.numpy(), atol=2e-4)
            np.testing.assert_allclose(p[:, 2 * tile + g, 2].numpy(), torch.exp(sub - sub.max(1, keepdim=True).values).sum(1).numpy(), rtol=2e-4)
    eos = torch.tensor([102], dtype=torch.int32, device=DEV)
    L = 6
    ids = torch.zeros(R, L, dtype=torch.int32, device=DEV)
    unf = torch.ones(R, dtype=torch.int32, device=DEV)
    unf[::5] = 0
    slp, nst = torch.zeros(R, device=DEV), torch.zeros(R, dtype=torch.int32, device=DEV)
    ops.token_step_partials(part, R, 2, 0, eos, ids, unf, slp, nst)
    top2 = ref.topk(2, dim=1)
    lsm = torch.log_softmax(ref, dim=1)
    tok = ids[:, 2].cpu()
    was_unf = torch.ones(R, dtype=torch.bool)
    was_unf[::5] = False
    for r in range(R):
        if not was_unf[r]:
            assert int(tok[r]) == 0 and float(slp[r]) == 0.0 and int(nst[r]) == 0 and int(unf[r]) == 0
            continue
        if int(tok[r]) != int(top2.indices[r, 0]):
            assert float(top2.values[r, 0] - top2.values[r, 1]) < 2e-4, r          # only a near-tie may flip
        assert abs(float(slp[r]) - float(lsm[r, int(tok[r])])) < 3e-4
        assert int(nst[r]) == 1 and int(unf[r]) == int(int(tok[r]) != 102)
    assert int((tok == 102).sum()) > 0


@pytest.mark.parametrize("B,C,heads,E,cur_len", [(6, 578, 12, 1, 7), (3, 198, 12, 5, 4), (2, 146, 12, 10, 9)])
def test_decode_attention_skips_finished_sequences(B, C, heads, E, cur_len):
    """vc_decode_attention_skip: a CTA (image, head, group of <= 8 sequences) whose sequences have all finished -- or whose image's
    beam search is done -- reads no K/V and leaves its output rows untouched; every other row is bit-identical to the plain call."""
    scale = 0.125
    ctx, stepq, anc = _da_inputs(B, C, heads, E, cur_len, torch.bfloat16, seed=11)
    H, R = heads * 64, B * E
    ctx, stepq, anc = ctx.to(DEV), stepq.to(DEV), anc.to(DEV)
    full = torch.empty(2 * R, H, dtype=torch.bfloat16, device=DEV)
    ops.decode_attention(ctx, stepq, anc, full, B, C, heads, E, cur_len, scale)
    g = torch.Generator().manual_seed(3)
    unf = (torch.rand(R, generator=g) < 0.5).to(torch.int32)
    unf[:E] = 0                                               # image 0: every sequence finished
    done = (torch.rand(B, generator=g) < 0.5).to(torch.int32)
    done[0], done[-1] = 1, 0
    groups = (E + 7) // 8
    for kw, live_seq in ((dict(seq_unfinished=unf.to(DEV)), None), (dict(img_done=done.to(DEV)), None)):
        out = torch.full((2 * R, H), 123.0, dtype=torch.bfloat16, device=DEV)
        ops.decode_attention(ctx, stepq, anc, out, B, C, heads, E, cur_len, scale, **kw)
        o, f = out.cpu(), full.cpu()
        n_skipped = 0
        for b in range(B):
            for grp in range(groups):
                seqs = range(b * E + grp * 8, b * E + min(E, grp * 8 + 8))
                live = (int(done[b]) == 0) if "img_done" in kw else any(int(unf[r]) for r in seqs)
                for r in seqs:
                    rows = slice(2 * r, 2 * r + 2)
                    if live:
                        assert torch.equal(o[rows], f[rows]), (b, r)
                    else:
                        assert bool((o[rows] == 123.0).all()), (b, r)
                        n_skipped += 1
        assert n_skipped > 0


# ------------------------------------------------------------------------------------------------ IEEE-half operand format
def _h(x):
    return x.clamp(-65504.0, 65504.0).to(torch.float16)


@pytest.mark.parametrize("M,N,K,splits", [(1024, 768, 3072, 6), (777, 768, 768, 3), (1, 768, 768, 6), (512, 768, 768, 1)])
def test_dec_linear_half_partial_planes(M, N, K, splits):
    """VC_DEC_FMT_F16 (decode_precision='fp16'): the product of IEEE-half operands, fp32 accumulation -- against fp32 torch on
    the SAME half values, so the only difference is the summation order; and it must be the half product, not a bf16 one."""
    a, w = _rnd(M, K, seed=M + 1), _rnd(N, K, seed=N + 2, scale=0.05)
    A, W = _h(a), _h(w)
    ref = A.float() @ W.float().t()
    m_pad = (M + 127) // 128 * 128
    out = torch.full((splits, m_pad, N), float("nan"), device=DEV)
    ops.dec_linear(ops.DEC_PARTIAL, A.to(DEV), W.to(DEV), None, out, M=M, splits=splits, m_pad=m_pad)
    got = out[:, :M].cpu().sum(0)
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 3e-5 * scale
    # the same bits read as bf16 would be another number altogether; the operand rounding itself is 8 x finer than bf16's
    e_half = float((ref - a @ w.t()).abs().mean())
    e_bf16 = float((_bf(a).float() @ _bf(w).float().t() - a @ w.t()).abs().mean())
    assert e_half < 0.2 * e_bf16


@pytest.mark.parametrize("M,N,K", [(1024, 3072, 768), (130, 3072, 768), (64, 128, 64)])
def test_dec_linear_half_gelu_epilogue(M, N, K):
    """VC_DEC_GELU_BF16 with half operands writes GELU(A W^T + bias) as IEEE halves (the operand of the next half GEMM)."""
    A, W = _h(_rnd(M, K, seed=M)), _h(_rnd(N, K, seed=N, scale=0.05))
    bias = _rnd(N, seed=9)
    ref = port.gelu_fast(A.float() @ W.float().t() + bias)
    out = torch.full((M, N), 9.0, device=DEV, dtype=torch.float16)
    ops.dec_linear(ops.DEC_GELU_BF16, A.to(DEV), W.to(DEV), bias.to(DEV), out, M=M)
    got = out.cpu().float()
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 1.2e-3 * scale                  # MUFU.TANH (2^-11) + the half rounding (2^-12)
    assert float((got - ref).abs().mean()) <= 1.5e-4 * scale


@pytest.mark.parametrize("R,V", [(512, 30522), (5, 30522), (130, 3000)])
def test_dec_linear_half_logits_plane_with_bias(R, V):
    """VC_DEC_PARTIAL with splits = 1 and a bias = the materialised fp32 vocabulary logits (beam search / sampling with
    decode_precision='fp16'): any N (30522 is not a multiple of 4), rows clipped at M. The TMA store works in 16-byte units: the
    pad columns up to the next multiple of 4 may receive zeros (the pitch is a multiple of 4, so they exist), the rest is
    untouched."""
    K = 768
    A, W = _h(_rnd(R, K, seed=R)), _h(_rnd(V, K, seed=V, scale=0.03))
    bias = 0.5 * _rnd(V, seed=5)
    ref = A.float() @ W.float().t() + bias
    ldl = (V + 63) // 64 * 64
    logits = torch.full((R + 3, ldl), 7.0, device=DEV)
    ops.dec_linear(ops.DEC_PARTIAL, A.to(DEV), W.to(DEV), bias.to(DEV), logits[:, :V], M=R)
    got = logits.cpu()
    assert float((got[:R, :V] - ref).abs().max()) <= 3e-5 * float(ref.abs().max())
    V4 = (V + 3) // 4 * 4
    assert bool((got[R:] == 7.0).all()) and bool((got[:, V4:] == 7.0).all())
    assert bool(((got[:R, V:V4] == 7.0) | (got[:R, V:V4] == 0.0)).all())


@pytest.mark.parametrize("rows,splits,gelu,mode", [(1024, 3, False, "f16"), (1000, 6, False, "bf16+f16"), (512, 6, True, "f16"),
                                                   (3, 2, False, "bf16+f16")])
def test_finish_ln_half_operand_copies(rows, splits, gelu, mode):
    H = 768
    m_pad = (rows + 127) // 128 * 128
    part = _rnd(splits, m_pad, H, seed=rows, scale=0.5)
    bias, gamma, beta = _rnd(H, seed=1), 1.0 + 0.1 * _rnd(H, seed=2), 0.1 * _rnd(H, seed=3)
    res = None if gelu else _rnd(rows, H, seed=4)
    out_f = torch.empty(rows, H, device=DEV)
    if mode == "f16":
        out_t = torch.full((rows, H), 5.0, device=DEV, dtype=torch.float16)
    else:
        out_t = torch.full((rows, 3 * H), 5.0, device=DEV, dtype=torch.bfloat16)
    ops.finish_ln(part.to(DEV), splits, bias.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-12, rows, resid=None if gelu else res.to(DEV),
                  gelu=gelu, out_f=out_f, out_t=out_t, split=mode)
    of = out_f.cpu()
    if mode == "f16":
        assert torch.equal(out_t.cpu(), _h(of))
    else:
        o = out_t.cpu()
        assert torch.equal(o[:, :H], _bf(of))
        assert torch.equal(o.view(torch.float16)[:, H:2 * H], _h(of))
        assert bool((o[:, 2 * H:] == 5.0).all())


@pytest.mark.parametrize("R,V", [(512, 30522), (40, 3000)])
def test_vocab_argmax_half_operands(R, V):
    K = 768
    A, W = _h(_rnd(R, K, seed=R)), _h(_rnd(V, K, seed=V, scale=0.03))
    bias = 0.5 * _rnd(V, seed=5)
    ref = A.float() @ W.float().t() + bias
    n_part = ops.vocab_partials(V)
    part = torch.full((R, n_part, 4), float("nan"), device=DEV)
    ops.dec_vocab_argmax(A.to(DEV), W.to(DEV), bias.to(DEV), part, M=R)
    p = part.cpu()
    mx = p[:, :, 0].max(1).values
    np.testing.assert_allclose(mx.numpy(), ref.max(1).values.numpy(), atol=2e-4)
    lse = torch.log((p[:, :, 2] * torch.exp(p[:, :, 0] - mx[:, None])).nan_to_num(0.0).sum(1)) + mx
    np.testing.assert_allclose(lse.numpy(), torch.logsumexp(ref, 1).numpy(), atol=3e-4)
    best = p[:, :, 0].argmax(1)
    idx = p[torch.arange(R), best, 1].contiguous().view(torch.int32)
    top2 = ref.topk(2, dim=1)
    for r in range(R):
        if int(idx[r]) != int(top2.indices[r, 0]):
            assert float(top2.values[r, 0] - top2.values[r, 1]) < 2e-4, r


def test_half_conversions_saturate_instead_of_overflowing():
    """Values beyond the IEEE-half range become +-65504, never infinities (cvt.rn.satfinite): LayerNorm outputs under an absurd
    gain through vc_finish_ln, GELU outputs of a GEMM with huge weights through vc_dec_linear."""
    rows, H = 64, 768
    part = _rnd(1, 128, H, seed=1)
    gamma, beta = torch.full((H,), 1.0e5), torch.zeros(H)
    out_f = torch.empty(rows, H, device=DEV)
    out_t = torch.empty(rows, H, device=DEV, dtype=torch.float16)
    ops.finish_ln(part.to(DEV), 1, None, gamma.to(DEV), beta.to(DEV), 1e-12, rows, out_f=out_f, out_t=out_t, split="f16")
    of, ot = out_f.cpu(), out_t.cpu()
    assert float(of.abs().max()) > 65504.0 and bool(torch.isfinite(ot.float()).all())
    assert torch.equal(ot, _h(of)) and float(ot.float().abs().max()) == 65504.0
    M, N, K = 128, 128, 64
    A, W = _h(_rnd(M, K, seed=2) * 100.0), _h(_rnd(N, K, seed=3) * 100.0)
    ref = port.gelu_fast(A.float() @ W.float().t())
    out = torch.empty(M, N, device=DEV, dtype=torch.float16)
    ops.dec_linear(ops.DEC_GELU_BF16, A.to(DEV), W.to(DEV), None, out, M=M)
    got = out.cpu().float()
    assert float(ref.max()) > 65504.0 and bool(torch.isfinite(got).all()) and float(got.max()) == 65504.0
    big = ref.abs() > 70000.0
    assert bool((got[big] == 65504.0).all())
