"""TEST INFRASTRUCTURE ONLY. numpy restatement of the sampling noise of vc_token_step (search.cu):
Philox4x32-10 keyed by the 64-bit seed, counter = (vocab_index // 4, row, cur_len, 0); lane j of the output block is
the draw for vocab index 4*(idx//4)+j; u = ((x >> 8) + 0.5) / 2^24; Gumbel = -log(-log(u)) in fp32."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = [np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3)]
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return c0, c1, c2, c3


def gumbel_noise(seed, rows, V, cur_len):
    """fp32 [rows, V] Gumbel(0,1) noise identical to the kernel's."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    n4 = (V + 3) // 4
    i4 = np.broadcast_to(np.arange(n4, dtype=np.uint64)[None, :], (rows, n4))
    r = np.broadcast_to(np.arange(rows, dtype=np.uint64)[:, None], (rows, n4))
    out = philox4x32_10(i4, r, np.full((rows, n4), cur_len, np.uint64), np.zeros((rows, n4), np.uint64),
                        seed & 0xFFFFFFFF, seed >> 32)
    x = np.stack(out, axis=-1).reshape(rows, n4 * 4)[:, :V]
    u = ((x >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    return -np.log(-np.log(u, dtype=np.float32), dtype=np.float32)


def make_sampler(seed):
    import torch

    def sampler(logits, cur_len):
        g = torch.from_numpy(gumbel_noise(seed, logits.shape[0], logits.shape[1], cur_len))
        return torch.argmax(logits + g, dim=-1)
    return sampler
