#!/usr/bin/env python
"""Benchmark of the ViTCAP caption hot path on B200 (driver contract: one JSON line on stdout from rank 0).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (bf16, tcgen05)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

A step = one pass of the hot path (patch embed -> split ViT encoder -> concept head/top-50 -> decoder prefill ->
19 greedy decode steps -> token ids) over one batch of 512 synthetic 384x384 images per GPU (BASELINE.json configs[2],
the configuration the metric "images/sec captioned" is quoted on). Weak scaling: every rank captions its own 512 images
and the packed results are all-gathered (the path's only exchange step).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec captioned (ViT-B/16-384, 20-tok greedy)"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="images per GPU per step")
    ap.add_argument("--variant", default="16_384")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--dec-layers", type=int, default=4)
    ap.add_argument("--decode-precision", default=None, choices=["bf16", "bf16x3", "fp16"],
                    help="operands of the decode-step MLP / vocabulary-head GEMMs (default: the model's, fp16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-images", type=int, default=8, help="cpu_baseline leg: BASELINE.json configs[0] is B = 8")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra keys (12-layer decoder, EOS-planted step)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only. cpu (the contract's reference arm): the reference algorithm on the host cores. "
                         "cuda: the same oracle port executed by stock PyTorch library kernels (cuBLAS / ATen) on cuda:0 -- the "
                         "'library-kernel comparator' of BASELINE.md section 4 item 5; its line says impl = reference-on-gpu")
    ap.add_argument("--ref-dtype", default="f32", choices=["f32", "bf16"], help="with --ref-device cuda: fp32 (TF32 off) or "
                    "torch.autocast(bfloat16)")
    ap.add_argument("--ref-algorithm", default="faithful", choices=["faithful", "cached"],
                    help="with --ref-device cuda: the algorithm as shipped (no KV cache) or the cached re-formulation")
    ap.add_argument("--ref-batch", type=int, default=32, help="with --ref-device cuda: images per step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch images on EVERY GPU (the contract's default); strong: --batch images in total, "
                         "split over the GPUs (SURVEY.md section 8d asks for both)")
    return ap.parse_args()


def per_gpu_batch(args):
    world = max(1, int(os.environ.get("WORLD_SIZE", str(max(1, args.gpus)))))
    if args.scaling == "strong":
        if args.batch % world:
            raise SystemExit("bench.py: --scaling strong needs --batch divisible by the number of GPUs")
        return args.batch // world
    return args.batch


def workload_config(args, cfg):
    b = per_gpu_batch(args)
    return {
        "workload": "BASELINE.json configs[2]: full ViTCAP greedy captioning, ViT-B/16-%d, %d-layer decoder, max_len 20, "
                    "batch %d per GPU, data-parallel" % (cfg.img_size, cfg.dec_layers, b),
        "variant": args.variant, "batch_per_gpu": b, "global_batch": b * max(1, args.gpus), "batch_per_step": b * max(1, args.gpus),
        "max_length": 20, "num_beams": 1, "decoder_layers": cfg.dec_layers, "parallelism": "dp%d" % max(1, args.gpus),
        "weights": "random-init (synth.make_state_dict seed 0, reference layout)",
        "decode_precision": getattr(args, "decode_precision", None) or "fp16",
        "operand_storage": "fp16 (VITCAP_STORE=fp16: every 16-bit operand an IEEE half)" if os.environ.get("VITCAP_STORE", "bf16").lower()
                           in ("fp16", "f16", "half") else "bf16",
        "precision": "bf16 operands / fp32 accumulation throughout (fp32 residual stream); decode_precision names the operands of "
                     "the decode-step MLP and vocabulary-head GEMMs: fp16 = IEEE half (11-bit significand, one product), "
                     "bf16x3 = split bf16 (three products), bf16 = plain",
        "l2": "inputs larger than L2 (906 MB of images and >10 GB of activations per step vs 126 MB L2)",
    }


def algorithmic_flops_per_image(cfg, max_len=20):
    """BASELINE.md section 3 / SURVEY.md section 8(d): FLOPs of the cached algorithm per image (187.95 G for ViT-B/16-384 with the
    4-layer decoder): every block in full, although the path skips the dead rows of the last concept block."""
    N, C, H, F, V, L = cfg.n_tokens, cfg.n_ctx, cfg.hidden, cfg.inter, cfg.vocab, cfg.dec_layers

    def block(n):
        return 2 * n * H * 3 * H + 4 * n * n * H + 2 * n * H * H + 4 * n * H * F
    patch = 2 * cfg.n_patches * cfg.patch_dim * H
    enc = (cfg.enc_blocks + cfg.split_blocks) * block(N)
    tag = 2 * H * H * 2 + 2 * H * V
    prefill = L * block(C)
    steps = 0
    for cl in range(1, max_len):
        steps += L * (2 * (2 * H * 3 * H + 2 * H * H + 4 * H * F) + 4 * 2 * (C + cl + 1) * H) + 2 * H * H + 2 * H * V
    return patch + enc + tag + prefill + steps


# ------------------------------------------------------------------------------------------------ clocks sampling
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Samples that arrived inside [t0, t1] (the timed region); the sampler is started before the warm-up steps because
        nvidia-smi needs a few hundred ms to deliver its first line. A timed region shorter than the sampling period (strong
        scaling at 8 GPUs: 0.2 s) falls back to the samples of the warm-up steps right before it and says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.2)]
        window = "timed region"
        if not rows:
            rows, window = [r for t, r in self.rows], "warm-up steps + timed region (timed region shorter than the sampling period)"
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        load = [s for s, p in zip(sm, pw) if p > 300] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_captioner(cfg, sd, extra, variant):
    """The CPU implementation both CPU legs time: the UNMODIFIED reference when it can be imported here -- from /root/reference
    in the build container, from the bytecode tree oracle/build_ref.py staged under oracle/_ref on the GPU box -- through its
    own public entry (ImageCaptioning.forward under no_grad, uni_pipeline.py:745-746); otherwise the restatement
    oracle/port.py 'faithful' (pinned to the reference's outputs by tests/golden). Returns (fn(data) -> (ids, logprobs),
    kind, what)."""
    import torch
    from vitcap_b200 import config as vcfg
    try:
        from oracle import ref_loader
        if ref_loader.reference_available() and vcfg.variant(variant, dec_layers=cfg.dec_layers) == cfg:
            ref, _ = ref_loader.build_reference(variant, decoder_layer=cfg.dec_layers)
            ref.load_state_dict(sd, strict=True)
            ref.test_extra_input = extra

            def run_ref(data):
                d = dict(data)
                d["key"] = ["k%d" % i for i in range(d["image"].shape[0])]
                with torch.no_grad(), ref_loader.on_cpu():
                    return ref(d)
            return run_ref, "reference", "the unmodified reference (ImageCaptioning.forward, imported from %s: %s)" \
                % (ref_loader.REF_ROOT, ref_loader.REF_KIND)
    except Exception as e:  # noqa: BLE001 -- a reference tree that does not import falls back to the pinned restatement
        print("reference import failed (%s: %s): timing oracle/port.py instead" % (type(e).__name__, e), file=sys.stderr)
    from oracle import port
    pm = port.PortModel(cfg, sd)

    def run_port(data):
        with torch.no_grad():
            return port.caption(pm, data, extra, algorithm="faithful")
    return run_port, "port", "oracle/port.py 'faithful' (restatement of the reference algorithm, pinned by tests/golden)"


def cpu_baseline(cfg, sd, n_images, extra, variant):
    """The reference as shipped (every decode step re-runs the ViT trunk, the tag head and the whole decoder) on the host
    cores, as BASELINE.md section 4 prescribes: all host threads, one untimed B = 1 warm-up, one timed run of B = n_images
    (configs[0]: 8)."""
    import torch
    from vitcap_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    fn, kind, what = cpu_captioner(cfg, sd, extra, variant)
    warm = synth.make_text_inputs(cfg, 1)
    warm["image"] = synth.make_images(cfg, 1, seed=1233)
    t0 = time.time()
    fn(warm)
    t_warm = time.time() - t0
    data = synth.make_text_inputs(cfg, n_images)
    data["image"] = synth.make_images(cfg, n_images, seed=1234)
    t0 = time.time()
    ids, lp = fn(data)
    dt = time.time() - t0
    return {"value": n_images / dt, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "BASELINE.json configs[0]: one batch of %d image(s) of the same workload (ViT-B/16-%d, 20-token greedy, fp32, "
                      "no KV cache: 19 full-model calls) through %s, %.1f s, after one untimed B=1 warm-up (%.1f s)"
                      % (n_images, cfg.img_size, what, dt, t_warm)}, ids


def decode_roofline(torch, ops, cfg, B, dev, peaks, model=None, extra=None):
    """Algorithmic bytes (K+V rows a step must read, context rows counted once per image) / measured
    duration of decode_attention_mma_kernel at the middle decode step, against the measured copy bandwidth."""
    H, heads, C, L = cfg.hidden, cfg.heads, cfg.n_ctx, cfg.dec_layers
    cur_len = 10
    ctx = [torch.randn(B, C, 3 * H, device=dev).to(ops.STORE) for _ in range(L)]       # (ops.STORE: bf16, or halves under VITCAP_STORE=fp16)
    stepq = [torch.randn(20, 2 * B, 3 * H, device=dev).to(ops.STORE) for _ in range(L)]
    out = torch.empty(2 * B, H, device=dev, dtype=ops.STORE)

    def run():
        for l in range(L):
            ops.decode_attention(ctx[l], stepq[l], None, out, B, C, heads, 1, cur_len, cfg.head_dim ** -0.5)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (iters * L)
    byt = B * (C + cur_len + 1) * 2 * H * 2
    peak = peaks.get("hbm_gbs")
    src = "MEASURED_PEAKS.json hbm_gbs (copy bandwidth, read+write)"
    if not peak:
        peak, src = 6500.0, "fallback (B200_PROFILING.md: ~6.5 TB/s measured copy)"
    ach = byt / (ms * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        traffic = next(v["dram_bytes"] for k, v in tj.items() if "decode_attention_mma_kernel" in k)
    except Exception:
        pass
    # the WHOLE decode loop (19 steps, one CUDA graph: attention + the small GEMMs / LayerNorms / search kernels around it)
    # against its HBM floor: SURVEY.md section 8(d) bytes = K+V of the visible keys per layer and step + one pass over the
    # decoder / vocabulary-head weights (bf16) per step
    loop = None
    if model is not None:
        eng = model.engine
        max_len = int(extra["max_length"])
        call = lambda: eng.greedy_or_sample(B, 1, max_len, int(extra["bos_token_id"]), int(extra["pad_token_id"]),   # noqa: E731
                                            [int(e) for e in extra["eos_token_ids"]], int(extra["mask_token_id"]))
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(5):
            call()
        l1.record()
        torch.cuda.synchronize()
        loop_ms = l0.elapsed_time(l1) / 5
        F = cfg.inter
        w_bytes = L * (3 * H * H + H * H + 2 * H * F) * 2 + (H * H + cfg.vocab * H) * 2
        kv_bytes = sum(L * B * (C + cl + 1) * 2 * H * 2 for cl in range(1, max_len))
        floor_ms = (kv_bytes + (max_len - 1) * w_bytes) / (peak * 1e9) * 1e3
        loop = {"ms": loop_ms, "floor_ms": floor_ms, "frac": floor_ms / loop_ms, "steps": max_len - 1,
                "kernels": eng.stats.get("graph_kernels"),
                "algorithmic_bytes": kv_bytes + (max_len - 1) * w_bytes,
                "how": "%d-step greedy decode loop replayed 5 times as its CUDA graph right after the timed region (context K/V of "
                       "the last batch), CUDA events; floor = algorithmic bytes / copy bandwidth" % (max_len - 1)}
    return {"kernel": "decode_attention_mma_kernel (2 query rows per sequence over 578 context + %d caption keys, bf16 K/V, "
                      "%d launches per decode step, 19 steps per batch)" % (cur_len + 1, L), "loop": loop,
            "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src,
            "algorithmic_bytes_per_launch": byt, "us_per_launch": ms * 1e3, "traffic": traffic,
            "traffic_note": "dram__bytes_read+write of one launch at B=512 from the committed ncu --set full capture "
                            "(profiles/r02_ncu_summary.md)",
            "how": "eager launches after the timed region (the decode loop itself is one CUDA graph): one launch per decoder "
                   "layer over distinct %0.2f GB caches, CUDA events, %d iterations" % (byt / 1e9, iters)}


def run_reference_on_gpu(args):
    """the reference's modules through stock PyTorch on one B200. /root/reference does not exist
    on the GPU box, so the thing executed is oracle/port.py (pinned to the reference's outputs by tests/golden) with every
    tensor on cuda:0: ATen / cuBLAS kernels, no kernel of this repo. A comparator line, never the product and never the
    contract's reference arm."""
    import contextlib
    import torch
    from vitcap_b200 import config as vcfg
    from vitcap_b200 import synth
    from oracle import port
    cfg = vcfg.variant(args.variant, dec_layers=args.dec_layers)
    sd = synth.make_state_dict(cfg, seed=0)
    extra = synth.default_test_extra_input(cfg)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B = args.ref_batch
    data = {k: v.to("cuda:0") for k, v in synth.make_text_inputs(cfg, B).items()}
    data["image"] = synth.make_images(cfg, B, seed=1234).to("cuda:0")
    pm = port.PortModel(cfg, {k: v.to("cuda:0") for k, v in sd.items()})
    torch.set_default_device("cuda:0")          # the port creates its masks / position ids with bare factory calls
    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if args.ref_dtype == "bf16" else contextlib.nullcontext

    def step():
        with torch.no_grad(), ctx():
            return port.caption(pm, data, extra, algorithm=args.ref_algorithm)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    wl = workload_config(args, cfg)
    wl.update(batch_per_gpu=B, global_batch=B,
              workload="library-kernel comparator: oracle/port.py '%s' algorithm (%s) executed by stock PyTorch on cuda:0, "
                       "%s, batch %d, ViT-B/16-%d, %d-layer decoder, greedy 20 tokens"
                       % (args.ref_algorithm, "the reference as shipped: every decode step re-runs the whole model"
                          if args.ref_algorithm == "faithful" else "trunk once, K/V cached, two rows per step",
                          "fp32, TF32 off" if args.ref_dtype == "f32" else "torch.autocast(bfloat16)", B, cfg.img_size,
                          cfg.dec_layers))
    print(json.dumps({"impl": "reference-on-gpu", "metric": METRIC, "value": B / ms * 1e3, "unit": UNIT, "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "dtype": args.ref_dtype, "data": "synthetic", "config": wl, "gpu_launches": 0}), flush=True)


def run_reference_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_device == "cuda":
        return run_reference_on_gpu(args)
    from vitcap_b200 import config as vcfg
    from vitcap_b200 import synth
    cfg = vcfg.variant(args.variant, dec_layers=args.dec_layers)
    sd = synth.make_state_dict(cfg, seed=0)
    extra = synth.default_test_extra_input(cfg)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    fn, kind, what = cpu_captioner(cfg, sd, extra, args.variant)

    def run(n, seed):
        data = synth.make_text_inputs(cfg, n)
        data["image"] = synth.make_images(cfg, n, seed=seed)
        t0 = time.time()
        fn(data)
        return time.time() - t0

    # Warm-up: W untimed B = 1 captions ("one untimed B=1 warm-up"); their median also sizes the timed
    # steps: EXACTLY --steps steps are timed, each one batch of `per_step` images, per_step = the largest batch <= 8
    # (BASELINE.json configs[0]) for which the whole arm still ends within ~4 minutes on this host.
    warm = [run(1, 1000 + i) for i in range(max(1, args.warmup))]
    t_img = statistics.median(warm)
    budget = 240.0
    per_step = int(budget / (max(1, args.steps) * t_img * 1.25))
    per_step = max(1, min(8, per_step))
    times = [run(per_step, 1234 + i) for i in range(args.steps)]
    total = sum(times)
    value = per_step * len(times) / total
    rates = sorted(per_step / t for t in times)
    spread = {"min": rates[0], "median": statistics.median(rates), "max": rates[-1], "unit": UNIT, "timed_steps": len(times)}
    sample = "%d timed step(s) of %d image(s) each (BASELINE.json configs[0] is B=8; the per-step batch is the largest <= 8 that " \
             "keeps %d steps within ~4 min at this host's %.2f s per B=1 caption) through %s: no KV cache, 19 full-model " \
             "calls per caption, fp32, %d host threads; per-step rate min/median/max %.3f/%.3f/%.3f images/s" \
             % (len(times), per_step, args.steps, t_img, what, threads, spread["min"], spread["median"], spread["max"])
    wl = workload_config(args, cfg)
    wl["batch_per_step"] = per_step
    wl["reference_sample"] = "the reference arm captions %d image(s) per step, not %d: a bounded sample of the workload" \
                             % (per_step, wl["global_batch"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": wl,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, "spread": spread},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    from vitcap_b200 import config as vcfg
    from vitcap_b200 import ops, parallel, synth
    from vitcap_b200.model import FastImageCaptioning

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the caption path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cores = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        # one process per GPU: keep the process and its pinned image buffers on the GPU's own socket
        numa_cores = parallel.bind_to_gpu_numa(local_rank)

    cfg = vcfg.variant(args.variant, dec_layers=args.dec_layers)
    sd = synth.make_state_dict(cfg, seed=0)
    extra = synth.default_test_extra_input(cfg)
    B = per_gpu_batch(args)
    model = FastImageCaptioning(cfg, test_extra_input=extra, mode=args.mode, max_batch=B, decode_precision=args.decode_precision)
    args.decode_precision = model.decode_precision
    model.load_state_dict(sd)
    model = model.to(dev)
    host = synth.make_text_inputs(cfg, B)
    # every rank captions different images (seed by rank)
    img_host = synth.make_images(cfg, B, seed=1234 + rank).pin_memory()
    text_dev = {k: v.to(dev) for k, v in host.items()}
    img_dev = img_host.to(dev)
    data_dev = dict(text_dev, image=img_dev)

    # the path's one exchange step (gather of the caption records) runs on a side stream: the compute stream of a rank never
    # waits for the other ranks inside the timed region, only the side streams meet (vitcap_b200/parallel.py SideStreamGather)
    side_gather = parallel.SideStreamGather(dev)
    last_gather = [None]

    def step_device():
        ids, lp = model(data_dev)
        rec = parallel.pack_records(ids, lp, *model.last_tags)
        full, done = side_gather(rec)
        last_gather[0] = done
        return full

    def join_gather():
        if last_gather[0] is not None:
            torch.cuda.current_stream().wait_event(last_gather[0])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- GEMM profiler: CUDA events around every eager tensor-core GEMM launch of the timed steps
    prof = []
    ops.GEMM_PROFILE = None
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    sync_all()
    ops.reset_launch_count()
    replays0 = model.engine.stats.get("graph_replays", 0)
    ops.GEMM_PROFILE = prof
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_region0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    join_gather()                                         # the timed region ends when the last batch's records are gathered
    ev1.record()
    sync_all()
    t_region1 = time.time()
    ops.GEMM_PROFILE = None
    clocks = sampler.stop(t_region0, t_region1)
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    eager = ops.launch_count()
    replays = model.engine.stats.get("graph_replays", 0) - replays0
    launches = eager + replays * model.engine.stats.get("graph_kernels", 0)
    value = world * B * args.steps / (ms_max / 1e3)

    # ---- roofline of the dominant kernel (gemm_tc_kernel): algorithmic FLOPs / measured duration
    flops = sum(p[0] for p in prof)
    gemm_ms = sum(p[1].elapsed_time(p[2]) for p in prof)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    if not peak:
        peak, peak_src = 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"
    achieved = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # DRAM traffic of the largest GEMM launch (MLP fc1 + GELU, M = 512 x 577) from the committed ncu --set full capture
    traffic, traffic_note = None, "no ncu capture committed"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        k = next(k for k in tj if "gemm_tc2_kernel<1, 0, 0, 2, 2" in k)
        traffic = tj[k]["dram_bytes"]
        traffic_note = "dram__bytes_read+write of one fc1+GELU launch with the folded LayerNorm (M=295424, N=3072, K=768; " \
                       "algorithmic 2.27 GB), profiles/r02_ncu_summary.md"
    except Exception:
        pass
    # the same launches shape by shape, each against the roof that bounds IT: max(flops / tensor peak, bytes / copy bandwidth).
    # (The o-proj GEMMs carry the fp32 residual stream -- 2.7 GB per launch for 0.35 TFLOP -- and are HBM bound; folded into the
    # family's tensor-roof figure above they read as a low tensor fraction.)
    hbm_peak = peaks.get("hbm_gbs") or 6500.0
    by_shape = {}
    for pr in prof:
        if len(pr) < 5:
            continue
        d = by_shape.setdefault(pr[4], {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["launches"] += 1
        d["ms"] += pr[1].elapsed_time(pr[2])
        d["flops"] += pr[0]
        d["bytes"] += pr[3]
    shapes, floor_ms = [], 0.0
    for tag, d in sorted(by_shape.items(), key=lambda kv: -kv[1]["ms"]):
        t_tensor = d["flops"] / (peak * 1e12) * 1e3
        t_hbm = d["bytes"] / (hbm_peak * 1e9) * 1e3
        floor_ms += max(t_tensor, t_hbm)
        if d["ms"] / max(gemm_ms, 1e-9) < 0.01:
            continue
        shapes.append({"gemm": tag, "launches_per_step": d["launches"] / max(1, args.steps), "us_per_launch": 1e3 * d["ms"] / d["launches"],
                       "tflops": d["flops"] / (d["ms"] * 1e-3) / 1e12, "gbs": d["bytes"] / (d["ms"] * 1e-3) / 1e9,
                       "bound": "tensor" if t_tensor >= t_hbm else "hbm", "frac_of_its_roof": max(t_tensor, t_hbm) / d["ms"]})
    roofline = {
        "kernel": "gemm_tc2_kernel / gemm_tc_kernel (tcgen05 bf16 GEMM, cta_group::2 pairs for the large shapes; %d eager "
                  "launches/step: ViT qkv/proj/fc1/fc2, patch embed, decoder prefill, heads)" % (len(prof) // max(1, args.steps)),
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
        "peak_source": peak_src, "traffic": traffic, "traffic_note": traffic_note,
        "share_of_step": gemm_ms / ms if ms > 0 else None,
        "algorithmic_flops_per_step": flops / max(1, args.steps),
        "by_shape": shapes,
        "frac_per_kernel_roof": floor_ms / gemm_ms if gemm_ms > 0 else None,
        "by_shape_note": "every eager GEMM launch of the timed region grouped by epilogue / N / K, against max(algorithmic flops / "
                         "%.0f TFLOP/s, algorithmic bytes / %.0f GB/s); frac_per_kernel_roof = sum of those floors / measured time "
                         "(shapes below 1 %% of the GEMM time are counted but not listed)" % (peak, hbm_peak),
    }

    # ---- second roofline object: the HBM-bound kernel of the decode steps (single-token attention over the KV cache). It runs
    # inside the CUDA graph of the decode loop, where events cannot bracket it, so the same launches (one per decoder layer,
    # distinct caches of the bench's own size: 3.6 GB, far beyond L2) are timed eagerly right after the timed region.
    roofline_decode = None
    if args.mode == "bf16" and rank == 0:
        roofline_decode = decode_roofline(torch, ops, cfg, B, dev, peaks, model=model, extra=extra)
    sync_all()

    # ---- e2e: the same steps through the public host loop (vitcap_b200.stream.OverlappedCaptioner, the drop-in for the
    # reference's predict_iter) with HOST buffers: every step uploads its own 512 images from pinned memory and reads its
    # result records back; the upload of step i+1 overlaps the captioning of step i (double-buffered device staging)
    e2e = None
    if not args.no_e2e:
        from vitcap_b200.stream import OverlappedCaptioner
        host_batch = {k: v.pin_memory() for k, v in host.items()}
        host_batch["image"] = img_host
        oc = OverlappedCaptioner(model, dev, depth=2, gather=True, with_tags=True)

        def batches(n):
            for _ in range(n):
                yield host_batch

        for _ in oc.run(batches(2)):                      # warm-up: staging buffers, pinned result buffers
            pass
        sync_all()
        oc.h2d_bytes = oc.d2h_bytes = 0
        # the first upload of a run cannot overlap anything (pipeline fill: 0.9 GB over PCIe = 3 % of a 5-step run): time at
        # least 10 steps so that the figure is the steady state a long evaluation sees, fill included
        n_e2e = max(10, min(args.steps, 20))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        n_out = 0
        for res in oc.run(batches(n_e2e)):
            n_out += res[0].shape[0]
        e1.record()
        sync_all()
        wall = time.time() - t0
        assert n_out == world * B * n_e2e
        ems = max(e0.elapsed_time(e1), wall * 1e3)
        te = torch.tensor([ems], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * n_e2e / (float(te.item()) / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(oc.h2d_bytes // n_e2e),
               "d2h_bytes_per_step": int(oc.d2h_bytes // n_e2e), "steps": n_e2e,
               "api": "vitcap_b200.stream.OverlappedCaptioner(FastImageCaptioning).run(host batches): upload of step i+1 "
                      "overlaps captioning of step i"}

    # ---- the same host loop fed with 8-bit pixels (an additive input format: uint8 HWC after
    # resize / crop; ToTensor + Normalize + BGR2RGB run inside patch extraction on the device): a quarter of the upload
    e2e_u8 = None
    if e2e is not None:
        u8 = ((img_host * 0.5 + 0.5).clamp_(0, 1) * 255.0).round_().to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
        host_batch_u8 = dict(host_batch, image=u8)

        def batches_u8(n):
            for _ in range(n):
                yield host_batch_u8

        for _ in oc.run(batches_u8(2)):
            pass
        sync_all()
        oc.h2d_bytes = oc.d2h_bytes = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        n_out = 0
        for res in oc.run(batches_u8(n_e2e)):
            n_out += res[0].shape[0]
        e1.record()
        sync_all()
        wall = time.time() - t0
        ems = max(e0.elapsed_time(e1), wall * 1e3)
        te = torch.tensor([ems], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_u8 = {"value": world * B * n_e2e / (float(te.item()) / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(oc.h2d_bytes // n_e2e),
                  "d2h_bytes_per_step": int(oc.d2h_bytes // n_e2e), "steps": n_e2e,
                  "api": "the same loop with uint8 HWC images (8-bit pixels; ToTensor/Normalize/BGR2RGB fused into patch extraction)"}

    # ---- extra keys: the whole step against the tensor roofline, and the BASELINE.json wording of the decoder (12 layers)
    extras = {}
    fl_img = algorithmic_flops_per_image(cfg)
    step_tflops = fl_img * B * world * args.steps / (ms_max * 1e-3) / 1e12 / world
    extras["whole_step_frac"] = {"algorithmic_gflop_per_image": fl_img / 1e9, "achieved_tflops_per_gpu": step_tflops, "peak": peak,
                                 "frac": step_tflops / peak if peak else None,
                                 "note": "SURVEY.md section 8(d) FLOPs of the whole step (HBM-bound decode loop included) / step "
                                         "time, against the sustained bf16 matmul peak"}
    if not args.no_extras and world == 1 and args.mode == "bf16" and cfg.dec_layers != 12:
        del out
        cfg12 = vcfg.variant(args.variant, dec_layers=12)
        m12 = FastImageCaptioning(cfg12, test_extra_input=extra, mode=args.mode, max_batch=B, decode_precision=args.decode_precision)
        m12.load_state_dict(synth.make_state_dict(cfg12, seed=0))
        m12 = m12.to(dev)
        for _ in range(3):
            m12(data_dev)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            m12(data_dev)
        a1.record()
        torch.cuda.synchronize()
        ms12 = a0.elapsed_time(a1) / 5
        extras["dec12"] = {"value": B / ms12 * 1e3, "unit": UNIT, "ms_per_step": ms12, "steps": 5, "warmup": 3,
                           "note": "the same step with a 12-layer decoder (BASELINE.json's wording; the shipped model builds 4, "
                                   "modeling_bert.py:1342-1346), device-resident inputs, decode_precision %s" % args.decode_precision}
        del m12
        torch.cuda.empty_cache()

    if not args.no_extras and world == 1 and args.mode == "bf16":
        # the same step on EOS-planted weights (cls.predictions.bias[SEP] += 1.825: captions end after 11 tokens on average, as
        # with a trained checkpoint; with the bench's plain random weights no caption ever ends). Finished captions are skipped
        # by the decode-step attention, and the captured loop leaves through its conditional nodes once all have ended
        # (modeling_utils.py:865-867)
        eos_bias = 1.825
        model.load_state_dict(synth.make_state_dict(cfg, seed=0, eos_bias=eos_bias))
        for _ in range(3):
            ids_e, _lp = model(data_dev)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            ids_e, _lp = model(data_dev)
        a1.record()
        torch.cuda.synchronize()
        ms_e = a0.elapsed_time(a1) / 5
        eos_id = int(extra["eos_token_ids"][0])
        hit = ids_e[:, 0] == eos_id
        length = torch.where(hit.any(1), hit.float().argmax(1) + 1, torch.full_like(ids_e[:, 0, 0], ids_e.shape[-1])).float()
        eng = model.engine
        max_len = int(extra["max_length"])
        loop = lambda: eng.greedy_or_sample(B, 1, max_len, int(extra["bos_token_id"]), int(extra["pad_token_id"]),   # noqa: E731
                                            [int(e) for e in extra["eos_token_ids"]], int(extra["mask_token_id"]))
        loop()
        torch.cuda.synchronize()
        a0.record()
        for _ in range(5):
            loop()
        a1.record()
        torch.cuda.synchronize()
        extras["eos_planted"] = {"value": B / ms_e * 1e3, "unit": UNIT, "ms_per_step": ms_e, "steps": 5, "warmup": 3,
                                 "eos_bias": eos_bias, "mean_caption_tokens": float(length.mean()),
                                 "captions_reaching_max_length": int((length >= max_len).sum()),
                                 "decode_loop_ms": a0.elapsed_time(a1) / 5,
                                 "decode_loop_ms_no_eos": (roofline_decode or {}).get("loop", {}).get("ms"),
                                 "note": "device-resident inputs; BOS and the closing EOS count as caption tokens"}
        model.load_state_dict(sd)
        del ids_e

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic", "config": workload_config(args, cfg),
        "roofline": roofline, "roofline_decode": roofline_decode, "e2e": e2e, "e2e_u8": e2e_u8, "gpu_launches": int(launches), "clocks": clocks,
        "extra": extras,
    }
    if world > 1:
        line["config"]["host_binding"] = ("rank bound to the %d CPU cores NVML reports as local to its GPU" % len(numa_cores)
                                          if numa_cores else "none (NVML affinity unavailable)")
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_baseline(cfg, sd, args.cpu_sample_images, extra, args.variant)
        line["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
