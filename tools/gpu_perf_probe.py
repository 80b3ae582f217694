"""Per-kernel and per-stage timing on the GPU (CUDA events, warm, inputs larger than L2). Not a bench value: it is the
map of where a step's time goes, used to pick the next kernel to optimise."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import ops, synth  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm_probe(B):
    M = B * 577
    res = []
    for (N, K, act, resid, outf32, name) in [(2304, 768, 0, False, False, "qkv"), (768, 768, 0, True, True, "proj+res"),
                                             (3072, 768, 1, False, False, "fc1+gelu"), (768, 3072, 0, True, True, "fc2+res")]:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
        b = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.float32 if outf32 else torch.bfloat16)
        r = out if resid else None
        if resid:
            out.normal_()
        for tn in (512, 256):
            ms = timeit(lambda: ops.linear(a, w, b, out, act=act, resid=r, impl="tc", tile_n=tn))
            tf = 2.0 * M * N * K / ms / 1e9
            res.append((name, M, N, K, tn, ms, tf))
            print("gemm %-9s M=%d N=%d K=%d tile_n=%d: %.3f ms  %.1f TFLOP/s" % (name, M, N, K, tn, ms, tf), flush=True)
        # torch (cuBLAS) comparator for the plain product
        ms = timeit(lambda: torch.matmul(a, w.t()))
        print("     cublas bf16 matmul same shape: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9), flush=True)
        del a, w, out
    return res


def attn_probe(B):
    N, heads = 577, 12
    qkv = torch.randn(B, N, 3 * 768, device=dev).to(torch.bfloat16)
    out = torch.empty(B, N, 768, device=dev, dtype=torch.bfloat16)
    fl = 4.0 * B * heads * N * N * 64
    for rnd in range(2):
        for knob, name in (("1", "attention_tc96 (96-key chunks, CLS key peeled)"), ("0", "attention_tc (64-key chunks)")):
            os.environ["VITCAP_ATTN96"] = knob
            ms = timeit(lambda: ops.attention(qkv, out, B, N, heads, 0.125), iters=20, warm=3)
            print("%s B=%d N=%d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (name, B, N, ms, fl / ms / 1e9), flush=True)
    del os.environ["VITCAP_ATTN96"]
    qkv2 = torch.randn(B, 578, 3 * 768, device=dev).to(torch.bfloat16)
    out2 = torch.empty(B, 578, 768, device=dev, dtype=torch.bfloat16)
    for knob in ("1", "0"):
        os.environ["VITCAP_ATTN96"] = knob
        ms = timeit(lambda: ops.attention(qkv2, out2, B, 578, heads, 0.125), iters=20, warm=3)
        print("N=578 (decoder context) VITCAP_ATTN96=%s: %.3f ms" % (knob, ms), flush=True)
    del os.environ["VITCAP_ATTN96"]
    del qkv2, out2
    ms = timeit(lambda: ops.attention(qkv, out, B, N, heads, 0.125))
    q, k, v = qkv.view(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ms2 = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    print("     torch SDPA same shape: %.3f ms  %.1f TFLOP/s" % (ms2, fl / ms2 / 1e9), flush=True)


def ln_probe(B):
    rows = B * 577
    x = torch.randn(rows, 768, device=dev)
    g, b = torch.randn(768, device=dev), torch.randn(768, device=dev)
    o = torch.empty(rows, 768, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.layernorm(x, g, b, 1e-6, out_t=o))
    print("layernorm rows=%d: %.3f ms  %.0f GB/s" % (rows, ms, rows * 768 * 6 / ms / 1e6), flush=True)


def decode_attn_probe(B, E=1, C=578, L=4):
    """One decode-step attention call per decoder layer over distinct caches (so L2 cannot help): algorithmic bytes =
    K+V of the visible keys, context rows counted once per image."""
    heads, H = 12, 768
    R = B * E
    ctx = [torch.randn(B, C, 3 * H, device=dev).to(torch.bfloat16) for _ in range(L)]
    stepq = [torch.randn(20, 2 * R, 3 * H, device=dev).to(torch.bfloat16) for _ in range(L)]
    out = torch.empty(2 * R, H, device=dev, dtype=torch.bfloat16)
    for impl in ("auto", "simt"):
        for cur_len in (1, 10, 19):
            def run():
                for l in range(L):
                    ops.decode_attention(ctx[l], stepq[l], None, out, B, C, heads, E, cur_len, 0.125, impl=impl)
            ms = timeit(run) / L
            byt = B * (C + E * (cur_len + 1)) * 2 * H * 2
            print("decode_attention[%s] B=%d E=%d cur_len=%d: %.1f us  %.0f GB/s (algorithmic)" % (impl, B, E, cur_len, ms * 1e3, byt / ms / 1e6),
                  flush=True)


def stage_probe(B, variant="16_384"):
    cfg = vcfg.variant(variant)
    sd = synth.make_state_dict(cfg, seed=0)
    m = FastImageCaptioning(cfg, mode="bf16", max_batch=B)
    m.load_state_dict(sd)
    m = m.to(dev)
    img = synth.make_images(cfg, B, seed=1).to(dev)
    eng = m.engine
    data = {k: v.to(dev) for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = img
    m(data)
    m(data)
    torch.cuda.synchronize()
    t = {}
    t["patch_embed"] = timeit(lambda: eng.patch_embed(img), iters=3, warm=1)
    f = eng.patch_embed(img)
    t["encode(16 blocks)"] = timeit(lambda: eng.encode(f), iters=3, warm=1)
    t["tag_head"] = timeit(lambda: eng.tag_head(B), iters=3, warm=1)
    t["prefill"] = timeit(lambda: eng.prefill(B), iters=3, warm=1)
    ex = m.test_extra_input
    t["decode(19 steps, graph)"] = timeit(lambda: eng.greedy_or_sample(B, 1, 20, 101, 0, [102], 103), iters=3, warm=1)
    t["full forward"] = timeit(lambda: m(data), iters=3, warm=1)
    for k, v in t.items():
        print("stage %-26s %.2f ms" % (k, v), flush=True)
    print("images/s (full forward): %.1f" % (B / t["full forward"] * 1e3))
    print("graph kernels per decode loop:", eng.stats.get("graph_kernels"))
    return t


def fold_probe(B):
    """Folded LayerNorm: producer (fc2 + residual, emitting the bf16 copy and statistics) and consumer (qkv from the raw copy)
    against the plain GEMMs and the LayerNorm kernel they replace. Long loops (steady state under the power cap)."""
    M, H = B * 577, 768
    hid = torch.randn(M, 3072, device=dev).to(torch.bfloat16)
    w2 = (torch.randn(H, 3072, device=dev) * 0.02).to(torch.bfloat16)
    b2 = torch.randn(H, device=dev)
    x = torch.randn(M, H, device=dev)
    xb = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
    stats = torch.empty(M, 3, 2, device=dev)
    wq = (torch.randn(2304, H, device=dev) * 0.02).to(torch.bfloat16)
    bq, cq = torch.randn(2304, device=dev), torch.randn(2304, device=dev)
    qkv = torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
    ln = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
    g, b = torch.randn(H, device=dev), torch.randn(H, device=dev)
    for name, fn in [("fc2+res plain", lambda: ops.linear(hid, w2, b2, x, resid=x)),
                     ("fc2+res emit", lambda: ops.linear_ln_emit(hid, w2, b2, x, x, xb, stats)),
                     ("layernorm", lambda: ops.layernorm(x, g, b, 1e-6, out_t=ln)),
                     ("qkv plain", lambda: ops.linear(ln, wq, bq, qkv)),
                     ("qkv fold", lambda: ops.linear_ln_fold(xb, wq, bq, cq, stats, 3, 1e-6, qkv))]:
        x.normal_()
        ms_burst = timeit(fn, iters=5, warm=2)
        ms = timeit(fn, iters=400, warm=20)
        print("%-16s burst %.3f ms   steady %.3f ms" % (name, ms_burst, ms), flush=True)


def prefill_probe(B, variant="16_384"):
    """The decoder prefill alone (steady state), for old / new library comparisons (VITCAP_LIB)."""
    cfg = vcfg.variant(variant)
    m = FastImageCaptioning(cfg, mode="bf16", max_batch=B)
    m.load_state_dict(synth.make_state_dict(cfg, seed=0))
    m = m.to(dev)
    data = {k: v.to(dev) for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = synth.make_images(cfg, B, seed=1).to(dev)
    m(data)
    f = m.engine.patch_embed(data["image"])
    for rnd in range(3):
        t_enc = timeit(lambda: m.engine.encode(f), iters=20, warm=3)
        t_pre = timeit(lambda: m.engine.prefill(B), iters=40, warm=5)
        t_all = timeit(lambda: m(data), iters=20, warm=3)
        print("round %d: encode %.3f ms  prefill %.3f ms  full forward %.2f ms (%.1f images/s)" % (rnd, t_enc, t_pre, t_all, B / t_all * 1e3),
              flush=True)


def fold_consumers_probe(B):
    """The two LayerNorm-folding consumer GEMMs of a ViT block at the bench shape: q|k|v (N = 2304) and fc1 + GELU (N = 3072)."""
    M, H = B * 577, 768
    xb = torch.randn(M, H, device=dev).to(torch.bfloat16)
    stats = torch.stack([xb.float().view(M, 3, 256).sum(2), (xb.float().view(M, 3, 256) ** 2).sum(2)], dim=2).contiguous()
    for name, N, act in (("qkv fold", 2304, ops.ACT_NONE), ("fc1+GELU fold", 3072, ops.ACT_GELU)):
        w = (torch.randn(N, H, device=dev) * 0.02).to(torch.bfloat16)
        bq, cq = torch.randn(N, device=dev), torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        fn = lambda: ops.linear_ln_fold(xb, w, bq, cq, stats, 3, 1e-6, out, act=act)   # noqa: E731
        for rnd in range(2):
            ms_burst = timeit(fn, iters=5, warm=2)
            ms = timeit(fn, iters=300, warm=20)
            print("%-14s round %d: burst %.3f ms   steady %.3f ms  (%.0f TFLOP/s steady)" % (name, rnd, ms_burst, ms, 2.0 * M * N * H / ms / 1e9),
                  flush=True)


def fold_ab_probe(B, variant="16_384"):
    """A/B of the folded norm1 (engine.ln_fold) in one process, alternating, steady state (20 forwards per sample)."""
    cfg = vcfg.variant(variant)
    sd = synth.make_state_dict(cfg, seed=0)
    m = FastImageCaptioning(cfg, mode="bf16", max_batch=B)
    m.load_state_dict(sd)
    m = m.to(dev)
    eng = m.engine
    data = {k: v.to(dev) for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = synth.make_images(cfg, B, seed=1).to(dev)
    for rnd in range(3):
        for fold in (2, 1, 0):
            eng.ln_fold, eng.ln_fold2 = fold >= 1, fold >= 2
            f = eng.patch_embed(data["image"])
            t_enc = timeit(lambda: eng.encode(f), iters=20, warm=3)
            t_all = timeit(lambda: m(data), iters=20, warm=3)
            print("ln_fold=%d round %d: encode %.2f ms  full forward %.2f ms  (%.1f images/s)" % (fold, rnd, t_enc, t_all, B / t_all * 1e3),
                  flush=True)
    eng.ln_fold = eng.ln_fold2 = True


def pdl_probe(B, variant="16_384"):
    """A/B of programmatic dependent launch (ops.set_pdl) in one process, alternating, same model and inputs."""
    cfg = vcfg.variant(variant)
    sd = synth.make_state_dict(cfg, seed=0)
    m = FastImageCaptioning(cfg, mode="bf16", max_batch=B)
    m.load_state_dict(sd)
    m = m.to(dev)
    eng = m.engine
    data = {k: v.to(dev) for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = synth.make_images(cfg, B, seed=1).to(dev)
    for rnd in range(2):
        for pdl in (2, 1, 0):
            ops.set_pdl(pdl)
            eng._dec_ws.clear()                      # captured decode loops keep the setting they were captured with
            m(data)
            m(data)
            torch.cuda.synchronize()
            f = eng.patch_embed(data["image"])
            t_enc = timeit(lambda: eng.encode(f), iters=3, warm=1)
            t_pre = timeit(lambda: eng.prefill(B), iters=3, warm=1)
            t_dec = timeit(lambda: eng.greedy_or_sample(B, 1, 20, 101, 0, [102], 103), iters=5, warm=1)
            t_all = timeit(lambda: m(data), iters=5, warm=1)
            print("pdl=%d round %d: encode %.2f  prefill %.2f  decode(graph) %.2f  full forward %.2f ms  (%.1f images/s)"
                  % (pdl, rnd, t_enc, t_pre, t_dec, t_all, B / t_all * 1e3), flush=True)
    ops.set_pdl(1)


def x3_probe(B):
    """Decode-step GEMMs at their shapes (2B or B rows): plain bf16 against the split-bf16 three-product form (K' = 3K), per
    tile width. Eight weight copies are rotated so that no launch finds its weights in L2 (as inside a decode step, where
    3.7 GB of K/V pass between two uses of a weight)."""
    R = B
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for (name, M, N, K, act, resid) in [("fc1+gelu", 2 * R, 3072, 768, 1, False), ("fc2+res", 2 * R, 768, 3072, 0, True),
                                        ("vocab", R, 30522, 768, 0, False), ("head_t", R, 768, 768, 1, False)]:
        for kk, tag in ((K, "bf16"), (3 * K, "bf16x3")):
            a = torch.randn(M, kk, device=dev).to(torch.bfloat16)
            ncopy = 2 if N > 10000 else 8
            ws = [(torch.randn(N, kk, device=dev) * 0.02).to(torch.bfloat16) for _ in range(ncopy)]
            b = torch.randn(N, device=dev)
            ldo = (N + 63) // 64 * 64
            out = torch.empty(M, ldo, device=dev)
            r = torch.randn(M, ldo, device=dev) if resid else None
            for tile in (0, 64, 128, 256, 512):
                if tile == 512 and N < 256:
                    continue
                i = [0]

                def fn():
                    i[0] += 1
                    ops.linear(a, ws[i[0] % ncopy], b, out[:, :N], act=act, resid=r[:, :N] if resid else None, ldo=ldo,
                               impl="tc" if tile else "auto", tile_n=tile)
                try:
                    ms = timeit(fn, iters=16, warm=4)
                except RuntimeError as e:
                    print("x3probe %-9s %-7s tile %3d: %s" % (name, tag, tile, str(e)[:60]))
                    continue
                print("x3probe %-9s %-7s M=%d N=%d K=%d tile %3d: %7.1f us  %6.1f TFLOP/s" %
                      (name, tag, M, N, kk, tile, ms * 1e3, 2.0 * M * N * kk / ms / 1e9), flush=True)
    x = torch.randn(2 * R, 3072, device=dev)
    o = torch.empty(2 * R, 3 * 3072, device=dev, dtype=torch.bfloat16)
    print("x3probe split_bf16x3 [%d, 3072]: %.1f us" % (2 * R, timeit(lambda: ops.split_bf16x3(x, o), iters=20, warm=3) * 1e3))


def prefill_fold_ab_probe(B, variant="16_384"):
    """A/B of the folded decoder prefill (engine.prefill_fold) in one process, alternating, steady state (20 iterations)."""
    cfg = vcfg.variant(variant)
    sd = synth.make_state_dict(cfg, seed=0)
    m = FastImageCaptioning(cfg, mode="bf16", max_batch=B)
    m.load_state_dict(sd)
    m = m.to(dev)
    eng = m.engine
    data = {k: v.to(dev) for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = synth.make_images(cfg, B, seed=1).to(dev)
    m(data)
    for rnd in range(3):
        for fold in (True, False):
            eng.prefill_fold = fold
            t_pre = timeit(lambda: eng.prefill(B), iters=20, warm=3)
            t_all = timeit(lambda: m(data), iters=20, warm=3)
            print("prefill_fold=%d round %d: prefill %.2f ms  full forward %.2f ms  (%.1f images/s)" % (fold, rnd, t_pre, t_all, B / t_all * 1e3),
                  flush=True)
    eng.prefill_fold = True


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    what = sys.argv[2] if len(sys.argv) > 2 else "all"
    print(torch.cuda.get_device_name(0), "B =", B, flush=True)
    if what in ("all", "gemm"):
        gemm_probe(B)
    if what in ("all", "attn"):
        attn_probe(B)
        ln_probe(B)
    if what in ("all", "dattn"):
        decode_attn_probe(B)
        decode_attn_probe(min(B, 256), E=4)
    if what in ("all", "stage"):
        stage_probe(B)
    if what == "x3":
        x3_probe(B)
    if what == "pdl":
        pdl_probe(B)
    if what == "fold":
        fold_probe(B)
    if what == "prefill":
        prefill_probe(B)
    if what == "foldc":
        fold_consumers_probe(B)
    if what == "foldab":
        fold_ab_probe(B)
    if what == "prefillab":
        prefill_fold_ab_probe(B)
