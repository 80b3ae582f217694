"""Decode lanes A/B (engine.decode_lanes and the vc_set_tuning knobs that let kernels of different lanes share an SM), one
process, alternating configurations: decode loop alone (19 steps, one CUDA graph) and the full forward at B images."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gpu_perf_probe import timeit  # noqa: E402
from vitcap_b200 import config as vcfg  # noqa: E402
from vitcap_b200 import ops, synth  # noqa: E402
from vitcap_b200.model import FastImageCaptioning  # noqa: E402

dev = torch.device("cuda:0")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    # (lanes, decode-attention CTAs per SM [0 = one CTA per item], GEMM smem cap KiB [0 = none], GEMM launch priority)
    configs = [(1, 0, 0, 0), (2, 0, 0, 0), (3, 0, 0, 0), (4, 0, 0, 0),
               (1, 2, 0, 0), (1, 3, 0, 0),
               (2, 2, 96, 0), (2, 3, 80, 0), (3, 2, 96, 0), (4, 2, 96, 0),
               (2, 2, 96, -2), (3, 2, 96, -2), (2, 0, 0, -2), (2, 2, 0, 0), (1, 0, 96, 0)]
    if len(sys.argv) > 2:
        configs = [tuple(int(v) for v in c.split(",")) for c in sys.argv[2:]]
    cfg = vcfg.variant("16_384")
    sd = synth.make_state_dict(cfg, seed=0)
    m = FastImageCaptioning(cfg, mode="bf16", max_batch=B)
    m.load_state_dict(sd)
    m = m.to(dev)
    eng = m.engine
    data = {k: v.to(dev) for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = synth.make_images(cfg, B, seed=1).to(dev)
    ref_ids, ref_lp = m(data)
    print(torch.cuda.get_device_name(0), "B =", B, flush=True)
    for rnd in range(2):
        for lanes, per_sm, smem_kb, prio in configs:
            eng.decode_lanes, eng.lane_gemm_priority = lanes, prio
            ops.set_tuning(ops.TUNE_DATTN_CTAS_PER_SM, per_sm)
            ops.set_tuning(ops.TUNE_GEMM_SMEM_KB, smem_kb)
            eng._dec_ws.clear()                      # captured loops keep the settings they were captured with
            try:
                ids, lp = m(data)
                m(data)
                torch.cuda.synchronize()
                same = bool(torch.equal(ids, ref_ids))
                t_dec = timeit(lambda: eng.greedy_or_sample(B, 1, 20, 101, 0, [102], 103), iters=5, warm=1)
                t_all = timeit(lambda: m(data), iters=5, warm=1)
                print("lanes=%d dattn_ctas/sm=%d gemm_smem_kb=%d prio=%d round %d: decode(graph) %.2f ms  full forward %.2f ms "
                      "(%.1f images/s)  ids_equal=%s kernels=%s" % (lanes, per_sm, smem_kb, prio, rnd, t_dec, t_all, B / t_all * 1e3, same,
                                                                    eng.stats.get("graph_kernels")), flush=True)
            except Exception as e:  # noqa: BLE001
                print("lanes=%d dattn_ctas/sm=%d gemm_smem_kb=%d prio=%d: FAILED %s" % (lanes, per_sm, smem_kb, prio, str(e)[:200]), flush=True)
    ops.set_tuning(ops.TUNE_DATTN_CTAS_PER_SM, 0)
    ops.set_tuning(ops.TUNE_GEMM_SMEM_KB, 0)


if __name__ == "__main__":
    main()
