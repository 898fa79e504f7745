#!/usr/bin/env python
"""bench.py — images/s image->CXSMILES token ids (batch 32 per GPU, greedy, <=512 tok) on 1..8 B200.

One "step" = one generate() pass (encode + 511 greedy decode steps) over one batch of 32 synthetic 512x512
images per GPU (BASELINE.json configs[1]).  `value` times the C-ABI call with inputs resident in HBM,
`e2e` times the host-buffer call (pinned host inputs -> H2D -> generate -> D2H ids) a user makes.
`--impl reference` times the CPU oracle (the stock-transformers restatement of the reference path, see
oracle/mg_oracle.py) on the host cores on a bounded sample of the same workload.

Prints ONE JSON line (rank 0).
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
sys.path.insert(0, ROOT)

METRIC = "images/sec image->CXSMILES (batch32, <=512 tok)"
UNIT = "images/s"
BATCH = 32
TEXT_LEN = 64
MAX_LENGTH = 512


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [s for s, p in zip(sm, power) if p > 300] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ inputs
def synth_inputs(image_size, batch, text_len, seed, vocab):
    """Seeded synthetic batch of the SURVEY §8d shape, generated with torch on the host (pinned)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    px = torch.ones(batch, 3, image_size, image_size)
    for b in range(batch):
        for _ in range(40):
            x0, y0 = [int(v) for v in torch.randint(0, image_size - 2, (2,), generator=g)]
            ln = int(torch.randint(4, image_size // 4, (1,), generator=g))
            if torch.rand(1, generator=g).item() < 0.5:
                px[b, :, y0:y0 + 2, x0:x0 + ln] = 0.0
            else:
                px[b, :, y0:y0 + ln, x0:x0 + 2] = 0.0
        for _ in range(12):
            x0, y0 = [int(v) for v in torch.randint(0, image_size - 8, (2,), generator=g)]
            w = int(torch.randint(3, image_size // 12, (1,), generator=g))
            h = int(torch.randint(2, image_size // 32, (1,), generator=g))
            px[b, :, y0:y0 + h, x0:x0 + w] = 0.25
    px = (px - 0.5) / 0.5
    n_prefix = 14
    ids = torch.randint(3, min(32000, vocab), (batch, text_len), generator=g)
    ids[:, n_prefix] = 1
    ids[:, -1] = 1
    box = torch.zeros(batch, text_len, 4)
    n_ocr = text_len - n_prefix - 2
    x0 = torch.rand(batch, n_ocr, generator=g) * 0.88 + 0.02
    y0 = torch.rand(batch, n_ocr, generator=g) * 0.88 + 0.02
    w = torch.rand(batch, n_ocr, generator=g) * 0.07 + 0.01
    h = torch.rand(batch, n_ocr, generator=g) * 0.02 + 0.01
    box[:, n_prefix + 1:-1] = torch.stack([x0, y0, (x0 + w).clamp(max=1), (y0 + h).clamp(max=1)], -1)
    box[:, n_prefix] = 1.0
    box[:, -1] = 1.0
    mask = torch.ones(batch, text_len, dtype=torch.long)
    return {"input_ids": ids, "bbox": box, "pixel_values": px, "attention_mask": mask}


# ------------------------------------------------------------------------------------------------ CPU oracle arm
_ORACLE = {}


def cpu_reference_sample(sample_batch, sample_len, full_len, threads):
    """Times the CPU oracle (kind="port": stock-transformers restatement of the reference path) on a bounded
    sample and extrapolates the per-image time to a full `full_len`-token greedy decode of the same batch.
    This is the ONLY place bench.py touches oracle/."""
    import torch
    from oracle import mg_oracle as O

    torch.set_num_threads(threads)
    if "model" not in _ORACLE:  # full-size random-init model, built once per process
        cfg = O.MGConfig.full()
        _ORACLE["cfg"] = cfg
        _ORACLE["model"] = O.build(cfg, seed=0)
    cfg, model = _ORACLE["cfg"], _ORACLE["model"]
    inp = O.make_inputs(cfg, sample_batch, TEXT_LEN, seed=1234)
    t0 = time.perf_counter()
    mem, mask = model.encode(**inp)
    t1 = time.perf_counter()
    model.generate_greedy(None, None, None, memory=mem, mask=mask, max_length=sample_len)
    t2 = time.perf_counter()
    t_enc, t_step = t1 - t0, (t2 - t1) / (sample_len - 1)
    full = t_enc + t_step * (full_len - 1)
    return {"images_per_s": sample_batch / full, "t_encode_s": t_enc, "t_step_s": t_step, "wall_s": t2 - t0,
            "sample": (f"B={sample_batch} images of the batch-{BATCH} workload, full-size random-init model, fp32, "
                       f"{threads} torch threads: encode timed in full + {sample_len - 1} greedy steps with KV cache; "
                       f"per-step time extrapolated to {full_len - 1} steps")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_reference_sample(4, 9, MAX_LENGTH, threads)
        if i >= args.warmup:
            vals.append(info["images_per_s"])
    v = statistics.mean(vals)
    out = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1000.0 * BATCH / v if v else None, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"configs[1]: batch-{BATCH} synthetic 512x512, greedy <={MAX_LENGTH} tok, "
                                  "random-init MarkushGrapher-2 dims", "text_len": TEXT_LEN,
                      "note": "reference path = CPU oracle (stock transformers UDOP+Swin restatement; the reference's "
                              "own model code lives in un-vendored forks and cannot be installed offline) on host cores"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": info["sample"]},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from markushgrapher_b200.configuration import MarkushgrapherConfig, random_state
    from markushgrapher_b200.engine import MGEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = MarkushgrapherConfig()  # MarkushGrapher-2 dims (UDOP-large + Swin-B + MLP projector)
    if args.small:
        cfg = MarkushgrapherConfig(vocab_size=2051, d_model=256, d_ff=512, num_layers=3, num_heads=4, image_size=128,
                                   swin_image=192, swin_embed=32, swin_depths=(2, 2, 2), swin_heads=(1, 2, 4),
                                   proj_hidden=256)
    state = random_state(cfg, seed=0, device=dev)
    eng = MGEngine(cfg, state, precision=0, device=dev)
    del state
    torch.cuda.empty_cache()

    if world > 1:
        eng.comm_init_from_torch()  # NCCL communicator owned by the library: per-step all-gather of token ids
    B = args.batch
    host = synth_inputs(cfg.image_size, B, TEXT_LEN, seed=1234 + rank, vocab=cfg.vocab_size)
    host = {k: v.pin_memory() for k, v in host.items()}
    devin = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = B * args.max_length * 8 * world

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gen_dev():
        if world > 1:
            return eng.generate_dist(**devin, max_length=args.max_length)
        return eng.generate(**devin, max_length=args.max_length, trim=False)

    def gen_host():
        if world > 1:  # host shard -> device, sharded generate with per-step id exchange, all ids back to the host
            d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            return eng.generate_dist(**d, max_length=args.max_length).cpu()
        return eng.generate_host(**host, max_length=args.max_length, trim=False)

    stream = torch.cuda.Stream(device=dev)
    sampler = ClockSampler(local)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            gen_dev()
        barrier()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        enc_ms, dec_ms, loop_ms, p50_ms, kernels, fused = [], [], [], [], 0, False
        for _ in range(args.steps):
            gen_dev()
            s = eng.last_stats()
            enc_ms.append(s["encode_ms"])
            dec_ms.append(s["decode_ms"])
            kernels = s["kernels"]
            lp = eng.last_decode_loop()
            loop_ms.append(lp["loop_ms"] / max(1, lp["steps"]))
            p50_ms.append(lp["step_p50_ms"])
            fused = lp["fused"]
        ev1.record(stream)
        barrier()
        ms_dev = ev0.elapsed_time(ev1)
        # ---- e2e: host buffers through the public host entry, copies inside the timed region
        gen_host()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            ids = gen_host()
        e1.record(stream)
        barrier()
        ms_e2e_dev = e0.elapsed_time(e1)
        ms_e2e_wall = (time.perf_counter() - t0) * 1000.0
        clocks = sampler.stop()
        # ---- dominant kernel, timed alone with CUDA events on its launch stream
        prof = eng.profile_cross_attn(reps=3)

    t = torch.tensor([ms_dev, max(ms_e2e_dev, ms_e2e_wall)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    steps_run = eng.last_steps
    if rank == 0:
        peak, peak_src = read_peaks()
        images = B * world * args.steps
        value = images / (ms_dev / 1000.0)
        e2e = images / (ms_e2e / 1000.0)
        ach = prof["bytes_per_launch"] / (prof["ms_per_launch"] * 1e-3) / 1e9
        # decode-step HBM model: weights (hi+lo planes = 4 B/param) + B*(cross KV + self KV), fp32
        d, L, M = cfg.d_model, cfg.num_decoder_layers, cfg.swin_tokens + TEXT_LEN + cfg.n_patches
        # weights read by one step: q,k,v,o + cross q,o + wi,wo per layer (cross k,v run once per image, not per step)
        w_bytes = 4 * (L * (6 * d * d + 2 * d * cfg.d_ff) + cfg.vocab_size * d)
        # cross K/V: kv24 = 3 bytes per element (fp32 rounded to 24 significant bits); self K/V (mean cached length
        # = half the decode) and weights: 4 bytes
        step_bytes = w_bytes + B * (L * 2 * M * d * 3 + L * 2 * (steps_run // 2) * d * 4)
        step_ms = statistics.mean(loop_ms)  # the step loop alone, CUDA events on the launch stream
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (split-bf16 tcgen05 GEMMs, fp32 accumulate / softmax, cross K/V stored with 24 significant bits)",
            "data": "synthetic",
            "config": {"workload": f"configs[1]: batch-{B} synthetic 512x512 images per GPU, random-init "
                                   f"MarkushGrapher-2 dims (831M params), greedy <={args.max_length} tok",
                       "images_per_gpu": B, "text_len": TEXT_LEN, "max_length": args.max_length,
                       "decode_steps_run": steps_run,
                       "parallelism": f"image-batch sharding x{world}" + (
                           ", ncclAllGather of token ids per decode step" if world > 1 else ""),
                       "l2": "inputs larger than L2 (cross-KV working set {:.1f} GB per step)".format(
                           B * L * 2 * M * d * 3 / 1e9)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(kernels) * args.steps,
            "phases": {"encode_ms": statistics.mean(enc_ms), "decode_ms": statistics.mean(dec_ms),
                       "decode_step_ms_mean": step_ms,
                       "decode_step_ms_p50": (statistics.median(p50_ms) if fused and p50_ms else step_ms),
                       "decode_step_algorithmic_GB": step_bytes / 1e9,
                       "decode_step_frac_of_hbm_peak": step_bytes / (step_ms * 1e-3) / 1e9 / peak},
        }
        full = B == 32 and not args.small
        cross = {"kernel": "cross_attn_stream24_kernel (decoder cross-attention over the kv24 encoder memory), timed "
                           "alone back to back over all layers",
                 "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                 "traffic": 246.97e6 if full else None,
                 "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full, "
                                   "profiles/r1_ncu_cross24.txt (batch 32, M = 1232)",
                 "algorithmic_bytes_per_launch": prof["bytes_per_launch"],
                 "ms_per_launch": prof["ms_per_launch"], "launches_timed": prof["launches"]}
        if fused:
            # the dominant kernel IS the decode step: one persistent kernel per generated token
            a2 = step_bytes / (step_ms * 1e-3) / 1e9
            out["roofline"] = {"kernel": "decode_step_kernel (fused persistent decode step: 24 layers + LM head, one "
                                         "launch per generated token)",
                               "bound": "hbm", "achieved": a2, "peak": peak, "unit": "GB/s", "frac": a2 / peak,
                               "traffic": 9.133e9 if (full and args.max_length == 512) else None,
                               "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the launch at cached "
                                                 "length 255, ncu --set full, profiles/r1c_ncu_decode_step.txt",
                               "peak_source": peak_src,
                               "algorithmic_bytes_per_launch": step_bytes, "ms_per_launch": step_ms,
                               "launches_timed": steps_run * args.steps,
                               "note": "algorithmic bytes = decoder weights (4 B/param) + B x (kv24 cross K/V + fp32 self "
                                       "K/V at the mean cached length); the kernel alternates HBM-bound attention "
                                       "phases with latency-bound linears separated by grid barriers",
                               "cross_attention_alone": cross}
        else:
            cross["peak_source"] = peak_src
            out["roofline"] = cross
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            c = cpu_reference_sample(4, 9, args.max_length, threads)
            out["cpu_baseline"] = {"value": c["images_per_s"], "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": c["sample"], "t_encode_s": c["t_encode_s"], "t_step_s": c["t_step_s"]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--max-length", dest="max_length", type=int, default=MAX_LENGTH)
    ap.add_argument("--small", action="store_true", help="debug: small dims (NOT the benchmark config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
