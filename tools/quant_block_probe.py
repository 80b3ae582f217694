"""Stage-by-stage comparison of the first ViT blocks of the fast path with oracle/port.py QuantPortModel: every oracle stage is
fed the KERNEL's own input of that stage, so a mismatch is local to the stage that prints it (test infrastructure; not shipped)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port  # noqa: E402
from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import synth  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def neq(a, b):
    return float((a.float() != b.float()).float().mean())


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    cfg = vcfg.variant("16_384")
    sd = synth.make_state_dict(cfg, seed=0, eos_bias=1.0)
    m = FastImageCaptioning(cfg, mode="bf16", max_batch=B)
    m.load_state_dict(sd)
    m = m.to(dev)
    eng = m.engine
    w = eng.w
    img = synth.make_images(cfg, B, seed=2024).to(dev)
    qm = port.QuantPortModel(cfg, {k: v.to(dev) for k, v in sd.items()})
    N, H, heads, d = cfg.n_tokens, cfg.hidden, cfg.heads, cfg.head_dim
    rows = B * N
    with torch.no_grad():
        x = eng.patch_embed(img)
        ws = eng._enc_ws
        print("patch embed: rel %.3g" % rel(x, qm.patch_embed(img)))
        xs = ws["x"][:rows]
        fx = ws["fold_x"]
        pre = None
        for blk in range(3):
            p = w.blocks[blk]
            prefix = "module.bert.encoder.blocks.%d." % blk
            x_in = xs.clone().view(B, N, H)
            eng._vit_block(p, xs, rows, B, N, ws, pre=pre, emit=fx)
            torch.cuda.synchronize()
            k_qkv = ws["qkv"][:rows].float().view(B, N, 3 * H)
            k_att = ws["att"][:rows].float().view(B, N, H)
            k_xb2 = ws["fold_2"][0][:rows].float().view(B, N, H)
            k_hid = ws["hid"][:rows].float().view(B, N, cfg.inter)
            k_out = xs.clone().view(B, N, H)
            k_xb = fx[0][:rows].float().view(B, N, H)
            n1 = (prefix + "norm1.weight", prefix + "norm1.bias", cfg.vit_ln_eps, prefix + "attn.qkv.weight", prefix + "attn.qkv.bias")
            if pre is None:
                o_qkv, o_ln = qm.lin_ln(x_in, *n1)
                print("block %d ln1 (kernel): rel %.3g, elements differing %.4f" % (blk, rel(ws["ln"][:rows].float().view(B, N, H), o_ln),
                                                                                     neq(ws["ln"][:rows].view(B, N, H), o_ln)))
            else:
                o_qkv = qm.lin_fold(x_in, *n1)
            o_qkv = port.q_bf16(o_qkv)
            print("block %d qkv (%s): rel %.3g, elements differing %.4f" % (blk, "ln kernel" if pre is None else "fold", rel(k_qkv, o_qkv), neq(k_qkv, o_qkv)))
            kq = k_qkv.reshape(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
            o_att = qm.attend(kq[0], kq[1], kq[2], d ** -0.5)
            print("block %d attention (from the kernel's qkv): rel %.3g, elements differing %.4f" % (blk, rel(k_att, o_att), neq(k_att, o_att)))
            o_mid = x_in + qm.lin(k_att, prefix + "attn.proj.weight", prefix + "attn.proj.bias")
            print("block %d xb after proj (bf16 copy of the stream): rel %.3g differing %.4f" % (blk, rel(k_xb2, port.q_bf16(o_mid)), neq(k_xb2, port.q_bf16(o_mid))))
            n2 = (prefix + "norm2.weight", prefix + "norm2.bias", cfg.vit_ln_eps, prefix + "mlp.fc1.weight", prefix + "mlp.fc1.bias")
            pre_act = qm.lin_fold(o_mid, *n2)
            o_hid = port.q_bf16(port.gelu_fast(pre_act))
            o_hid_exact = port.q_bf16(port.gelu_erf(pre_act))
            print("block %d hid = GELU(fc1 fold): rel %.3g differing %.4f  (vs exact-erf GELU: rel %.3g)" % (blk, rel(k_hid, o_hid), neq(k_hid, o_hid), rel(k_hid, o_hid_exact)))
            o_out = o_mid + qm.lin(k_hid, prefix + "mlp.fc2.weight", prefix + "mlp.fc2.bias")
            print("block %d out (from the kernel's hid): rel %.3g; emitted copy differing %.4f" % (blk, rel(k_out, o_out), neq(k_xb, port.q_bf16(o_out))))
            # statistics emitted for the next block
            st = fx[1][:rows * fx[2] * 2].view(rows, fx[2], 2).sum(1)
            mean_k = st[:, 0] / H
            var_k = st[:, 1] / H - mean_k * mean_k
            xo = k_out.view(rows, H)
            print("block %d emitted stats: mean rel %.3g var rel %.3g; |mean|/std median %.3g" % (
                blk, rel(mean_k, xo.mean(1)), rel(var_k, xo.var(1, unbiased=False)), float((xo.mean(1).abs() / xo.std(1)).median())))
            pre = fx
        # the whole-block chain error for reference
        q_taps = []
        qm.split_encoder(qm.patch_embed(img), taps=q_taps)
        print("chained oracle block2 vs kernel stream: rel %.3g" % rel(xs.view(B, N, H), q_taps[2][1]))


if __name__ == "__main__":
    main()
