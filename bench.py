#!/usr/bin/env python
"""bench.py — images/s image->CXSMILES token ids (batch 32 per GPU, greedy, <=512 tok) on 1..8 B200.

One "step" = one generate() pass (encode + 511 greedy decode steps) over one batch of 32 synthetic 512x512
images per GPU (BASELINE.json configs[1]).  `value` times the C-ABI call with inputs resident in HBM,
`e2e` times the host-buffer call (pinned host inputs -> H2D -> generate -> D2H ids) a user makes.
`--impl reference` times the CPU oracle (the stock-transformers restatement of the reference path, see
oracle/mg_oracle.py) on the host cores on THE SAME workload, in full: batch 32, full encode, all 511 greedy steps
(about two minutes per pass on 16 cores, so it runs as many passes as fit a wall-clock budget and reports that count
as `steps`).  `--workload enc256` measures BASELINE.json configs[2] instead (batch-256 encoder only).

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


def profile_traffic(kernel_prefix, name_glob):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel_prefix`, parsed from the newest committed
    ncu summary profiles/<name_glob> (written by tools/ncu_summary.py from an `ncu --set full` capture).
    Returns (bytes or None, file or None)."""
    import glob
    import re

    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for fn in sorted(glob.glob(os.path.join(ROOT, "profiles", name_glob)), reverse=True):
        tot, on = {}, False
        for line in open(fn):
            if not line.startswith(" "):
                on = line.startswith(kernel_prefix) and not tot
                continue
            m = re.match(r"\s+(dram__bytes_(?:read|write)\.sum)\s+([0-9.eE+-]+)\s+(\w+)", line)
            if on and m and m.group(3) in unit:
                tot[m.group(1)] = float(m.group(2)) * unit[m.group(3)]
        if len(tot) == 2:
            return sum(tot.values()), os.path.relpath(fn, ROOT)
    return None, None


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


def _oracle_model(threads):
    """full-size random-init oracle, built once per process.  This file touches oracle/ ONLY in this section."""
    import torch
    from oracle import mg_oracle as O

    torch.set_num_threads(threads)
    if "model" not in _ORACLE:
        cfg = O.MGConfig.full()
        _ORACLE["cfg"] = cfg
        _ORACLE["model"] = O.build(cfg, seed=0)
    return _ORACLE["cfg"], _ORACLE["model"]


def _greedy_steps(model, memory, mask, n_steps, prefill=0):
    """n_steps greedy decode steps with the KV cache (the oracle's generate_greedy loop body); seconds per step list.
    prefill > 0: the cache is first filled with `prefill` positions in ONE teacher-forced decoder pass, so that the timed
    steps run at that cached length (the stock cache re-concatenates K/V every step: a step at length 255 costs about
    twice a step at length 10)"""
    import torch

    B = memory.shape[0]
    cur = torch.zeros((B, 1), dtype=torch.long)
    past, ts = None, []
    with torch.no_grad():
        if prefill > 0:
            out = model.decoder(input_ids=torch.zeros((B, prefill), dtype=torch.long), encoder_hidden_states=memory,
                                encoder_attention_mask=mask, use_cache=True, return_dict=True)
            past = out.past_key_values
        for _ in range(n_steps):
            t0 = time.perf_counter()
            out = model.decoder(input_ids=cur, encoder_hidden_states=memory, encoder_attention_mask=mask,
                                past_key_values=past, use_cache=True, return_dict=True)
            past = out.past_key_values
            h = out.last_hidden_state[:, -1, :] * (model.cfg.d_model ** -0.5)
            cur = torch.argmax(model.lm_head(h).float(), dim=-1)[:, None]
            ts.append(time.perf_counter() - t0)
    return ts


def cpu_full_pass(threads, batch, max_length):
    """ONE full pass of the bench workload on the host cores: same inputs as the GPU arm, batch `batch`, full encode,
    greedy decode to max_length (all max_length-1 steps unless every row emits EOS).  Nothing is extrapolated."""
    cfg, model = _oracle_model(threads)
    inp = synth_inputs(cfg.image_size, batch, TEXT_LEN, seed=1234, vocab=cfg.vocab_size)
    t0 = time.perf_counter()
    mem, mask = model.encode(**inp)
    t1 = time.perf_counter()
    ids = model.generate_greedy(None, None, None, memory=mem, mask=mask, max_length=max_length)
    t2 = time.perf_counter()
    return {"t_encode_s": t1 - t0, "t_decode_s": t2 - t1, "wall_s": t2 - t0, "decode_steps": int(ids.shape[1]) - 1}


def cpu_bounded_sample(threads, batch, max_length, enc_images=8, steps=12):
    """cpu_baseline of the GPU arm's line: a BOUNDED sample (~15 s) of the same workload.  Encode is timed on
    `enc_images` of the batch (compute-bound, linear in the image count); the decode step is timed at the TRUE batch
    size -- the step is bound by the batch's cross-K/V and weight reads, so a smaller batch would understate the CPU --
    over `steps` real greedy steps on the encoder memory tiled to `batch` rows.  Both extrapolations are stated."""
    cfg, model = _oracle_model(threads)
    inp = synth_inputs(cfg.image_size, enc_images, TEXT_LEN, seed=1234, vocab=cfg.vocab_size)
    t0 = time.perf_counter()
    mem, mask = model.encode(**inp)
    t_enc = (time.perf_counter() - t0) * batch / enc_images
    rep = (batch + enc_images - 1) // enc_images
    mem, mask = mem.repeat(rep, 1, 1)[:batch].contiguous(), mask.repeat(rep, 1)[:batch].contiguous()
    mid = (max_length - 1) // 2
    ts = _greedy_steps(model, mem, mask, steps, prefill=mid)  # steps at the MEAN cached length of the workload
    t_step = statistics.mean(ts[2:])
    full = t_enc + t_step * (max_length - 1)
    return {"images_per_s": batch / full, "t_encode_s": t_enc, "t_step_s": t_step,
            "sample": (f"encode of {enc_images} of the {batch} images timed and scaled x{batch / enc_images:g}; "
                       f"{steps} real greedy steps at batch {batch} and cached length {mid} (cache prefilled by one "
                       f"teacher-forced pass; mean of the last {steps - 2}) scaled to "
                       f"{max_length - 1} steps; full-size random-init model, fp32, {threads} torch threads; the "
                       "un-extrapolated full pass is what `bench.py --impl reference` times")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    budget = float(os.environ.get("MG_REF_BUDGET_S", "200"))
    t_start = time.perf_counter()
    # warm-up: thread pool, allocator and code paths (one image, a few steps) -- a full pass costs minutes
    cfg, model = _oracle_model(threads)
    w = synth_inputs(cfg.image_size, 1, TEXT_LEN, seed=99, vocab=cfg.vocab_size)
    mem, mask = model.encode(**w)
    _greedy_steps(model, mem, mask, 3)
    passes = []
    while len(passes) < max(1, args.steps):
        passes.append(cpu_full_pass(threads, args.batch, args.max_length))
        if time.perf_counter() - t_start + passes[-1]["wall_s"] > budget:
            break
    wall = statistics.mean(p["wall_s"] for p in passes)
    v = args.batch / wall
    out = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": len(passes),
           "warmup": 1, "steps_requested": args.steps, "warmup_requested": args.warmup,
           "ms_per_step": 1000.0 * wall, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "extrapolated": False,
           "config": {"workload": f"configs[1]: batch-{args.batch} synthetic 512x512 images, random-init "
                                  f"MarkushGrapher-2 dims (831M params), greedy <={args.max_length} tok",
                      "images_per_gpu": args.batch, "text_len": TEXT_LEN, "max_length": args.max_length,
                      "decode_steps_run": passes[-1]["decode_steps"],
                      "note": "reference path = CPU oracle (stock transformers UDOP+Swin restatement; the reference's "
                              "own model code lives in un-vendored forks and cannot be installed offline) on the host "
                              "cores; every timed step is one FULL pass of the workload (full encode + every greedy "
                              f"step at batch {args.batch}); a pass takes minutes, so `steps` is the number of passes "
                              f"that fit the {budget:.0f} s budget (MG_REF_BUDGET_S), not the requested count; warm-up is "
                              "a 1-image encode + 3 steps"},
           "phases": {"encode_s": statistics.mean(p["t_encode_s"] for p in passes),
                      "decode_s": statistics.mean(p["t_decode_s"] for p in passes),
                      "decode_step_s_mean": statistics.mean(p["t_decode_s"] / max(1, p["decode_steps"]) for p in passes)},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"{len(passes)} full pass(es) of the workload, nothing extrapolated"},
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
    nbeams = args.num_beams
    text_len = TEXT_LEN
    cfg_name = "configs[1]"
    if args.workload == "gen128":      # BASELINE.json configs[3]: 128 images per GPU, greedy <= 512 tok, id all-gather per step
        B, cfg_name = (args.batch if args.batch != BATCH else 128), "configs[3]"
    elif args.workload == "beam4":     # BASELINE.json configs[4]: IP5-M shape (1000 images / 8 GPUs), beam 4, <= 768 tok
        B, cfg_name = (args.batch if args.batch != BATCH else 125), "configs[4]"
        nbeams = args.num_beams if args.num_beams > 1 else 4
        args.max_length = args.max_length if args.max_length != MAX_LENGTH else 768
        text_len = 512
    if text_len == TEXT_LEN:
        host = synth_inputs(cfg.image_size, B, TEXT_LEN, seed=1234 + rank, vocab=cfg.vocab_size)
    else:                              # IP5-M shape: ragged OCR text 32..512 tokens
        host = synth_inputs_ragged(cfg.image_size, B, text_len, seed=1239 + rank, vocab=cfg.vocab_size)
    host = {k: v.pin_memory() for k, v in host.items()}
    # e2e alternates between two host batches so that the H2D copy of batch i+1 (mg_prefetch_host, copy stream) can
    # overlap the decode of batch i -- both copies stay inside the timed region
    host_b = {k: v.clone().pin_memory() for k, v in host.items()}
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
            return eng.generate_dist(**devin, max_length=args.max_length, num_beams=nbeams)
        return eng.generate(**devin, max_length=args.max_length, num_beams=nbeams, trim=False)

    def gen_host(i=0):
        cur, nxt = (host, host_b) if i % 2 == 0 else (host_b, host)
        if world > 1:  # host shard -> device, sharded generate with per-step id exchange, all ids back to the host
            d = {k: v.to(dev, non_blocking=True) for k, v in cur.items()}
            return eng.generate_dist(**d, max_length=args.max_length, num_beams=nbeams).cpu()
        eng.prefetch_host(**nxt, max_length=args.max_length)  # returns at once; overlaps this batch's decode
        return eng.generate_host(**cur, max_length=args.max_length, num_beams=nbeams, trim=False)

    stream = torch.cuda.Stream(device=dev)
    sampler = ClockSampler(local)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            gen_dev()
        barrier()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        enc_ms, dec_ms, loop_ms, p50_ms, p99_ms, kernels, fused = [], [], [], [], [], 0, False
        for _ in range(args.steps):
            gen_dev()
            s = eng.last_stats()
            enc_ms.append(s["encode_ms"])
            dec_ms.append(s["decode_ms"])
            kernels = s["kernels"]
            lp = eng.last_decode_loop()
            loop_ms.append(lp["loop_ms"] / max(1, lp["steps"]))
            p50_ms.append(lp["step_p50_ms"])
            p99_ms.append(lp["step_p99_ms"])
            fused = lp["fused"]
        ev1.record(stream)
        barrier()
        ms_dev = ev0.elapsed_time(ev1)
        # ---- e2e: host buffers through the public host entry, copies inside the timed region
        gen_host(1)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            ids = gen_host(i)
        e1.record(stream)
        barrier()
        ms_e2e_dev = e0.elapsed_time(e1)
        ms_e2e_wall = (time.perf_counter() - t0) * 1000.0
        clocks = sampler.stop()
        # ---- opt-in throughput mode (not the headline): the encoder of batch i+1 runs ahead on its own SM partition while
        # batch i decodes (mg_encode_ahead); K batches, the fill outside the timed region like a warm-up step,
        # the last run-ahead encoder (nobody takes it) drained inside it
        ahead = None
        if world == 1 and nbeams == 1 and args.workload == "generate" and not args.no_ahead:
            sets = [devin, {k: v.clone() for k, v in devin.items()}]
            if eng.encode_ahead(**sets[0]):
                barrier()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                a_step, a_enc = [], []
                for i in range(args.steps):
                    eng.encode_ahead(**sets[(i + 1) % 2])
                    eng.generate(**sets[i % 2], max_length=args.max_length, trim=False)
                    lp = eng.last_decode_loop()
                    a_step.append(lp["loop_ms"] / max(1, lp["steps"]))
                    a_enc.append(eng.last_ahead()["encoder_ms"])
                eng.ahead_reset()
                a1.record(stream)
                barrier()
                info = eng.last_ahead()
                ahead = {"value": B * args.steps / (a0.elapsed_time(a1) / 1000.0), "unit": UNIT,
                         "decode_step_ms_mean": statistics.mean(a_step), "encoder_ms_on_its_partition": statistics.mean(a_enc),
                         "sms_encoder": info["sms_encoder"], "sms_decoder": info["sms_decoder"],
                         "note": "opt-in Engine.encode_ahead: encoder of the next batch on an SM partition of its own (CUDA "
                                 "green contexts) under the decode loop of the current one; device-resident inputs; not "
                                 "the headline `value`, which runs one batch at a time"}
            else:
                eng.ahead_reset()
        # ---- dominant kernel, timed alone with CUDA events on its launch stream
        prof = eng.profile_cross_attn(reps=3) if nbeams == 1 else None

    t = torch.tensor([ms_dev, max(ms_e2e_dev, ms_e2e_wall)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    steps_run = eng.last_steps
    m_enc, m_dec = eng.last_memory_len()
    if rank == 0:
        peak, peak_src = read_peaks()
        images = B * world * args.steps
        value = images / (ms_dev / 1000.0)
        e2e = images / (ms_e2e / 1000.0)
        ach = prof["bytes_per_launch"] / (prof["ms_per_launch"] * 1e-3) / 1e9 if prof else 0.0
        # decode-step HBM model: weights (hi+lo planes = 4 B/param) + B*(cross KV + self KV), fp32
        # M = memory positions the decoder holds K/V for: masked positions (patches absorbed by OCR tokens leave a padded
        # tail, text padding) are dropped before the cross K/V projection -- they contribute exactly 0 -- so the
        # ALGORITHMIC bytes of a step count the kept positions only (declared in config.memory_positions)
        d, L, M = cfg.d_model, cfg.num_decoder_layers, m_dec
        # weights read by one step: q,k,v,o + cross q,o + wi,wo per layer (cross k,v run once per image, not per step)
        w_bytes = 4 * (L * (6 * d * d + 2 * d * cfg.d_ff) + cfg.vocab_size * d)
        # cross K/V: kv24 = 3 bytes per element (fp32 rounded to 24 significant bits); self K/V (mean cached length
        # = half the decode) and weights: 4 bytes
        # (beam search: the cross K/V of an image are streamed once for all its beams, the self K/V once per beam)
        step_bytes = w_bytes + B * (L * 2 * M * d * 3 + nbeams * L * 2 * (steps_run // 2) * d * 4)
        step_ms = statistics.mean(loop_ms)  # the step loop alone, CUDA events on the launch stream
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ("f32 (GEMM operands as hi+lo bf16 planes on tcgen05 = ~16 significant bits per product, fp32 accumulate / "
                                      "residual / softmax; cross K/V stored as kv24 = 24 stored bits, 16 significant)"),
            "data": "synthetic",
            "config": {"workload": f"{cfg_name}: batch-{B} synthetic 512x512 images per GPU, random-init "
                                   f"MarkushGrapher-2 dims (831M params), "
                                   f"{'greedy' if nbeams == 1 else 'beam=' + str(nbeams)} <={args.max_length} tok",
                       "images_per_gpu": B, "text_len": text_len, "max_length": args.max_length, "num_beams": nbeams,
                       "decode_steps_run": steps_run,
                       "memory_positions": {"encoder": m_enc, "decoder_after_dropping_masked": m_dec},
                       "parallelism": f"image-batch sharding x{world}" + (
                           (", token ids of every decode step exchanged by NVLink peer stores fused into the selection kernel"
                            if eng.dist_mode() == 2 and nbeams == 1 else ", ncclAllGather of token ids per decode step")
                           if world > 1 else ""),
                       "l2": "inputs larger than L2 (cross-KV working set {:.1f} GB per step)".format(
                           B * L * 2 * M * d * 3 / 1e9)},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(kernels) * args.steps,
            "phases": {"encode_ms": statistics.mean(enc_ms), "decode_ms": statistics.mean(dec_ms),
                       "decode_step_ms_mean": step_ms,
                       "decode_step_ms_p50": statistics.median(p50_ms), "decode_step_ms_p99": statistics.median(p99_ms),
                       "decode_step_latency_source": "per-step %globaltimer stamps written by the kernel that ends each "
                                                     "step; median / 99th percentile over the steps of a generate call",
                       "decode_step_algorithmic_GB": step_bytes / 1e9,
                       "decode_step_frac_of_hbm_peak": step_bytes / (step_ms * 1e-3) / 1e9 / peak},
        }
        full = B == 32 and not args.small
        tr_cross, tr_cross_src = profile_traffic("cross_attn_stream24_kernel", "r*_ncu_cross24.txt")
        tr_step, tr_step_src = profile_traffic("decode_step_kernel", "r*_ncu_decode_step.txt")
        if prof is None:
            prof = {"bytes_per_launch": None, "ms_per_launch": None, "launches": 0}
        cross = {"kernel": "cross_attn_stream24_kernel (decoder cross-attention over the kv24 encoder memory), timed "
                           "alone back to back over all layers",
                 "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                 "traffic": tr_cross if full else None,
                 "traffic_source": f"dram__bytes_read.sum + dram__bytes_write.sum of one launch (batch 32, M = 1232), "
                                   f"parsed from the committed ncu --set full summary {tr_cross_src}",
                 "algorithmic_bytes_per_launch": prof["bytes_per_launch"],
                 "ms_per_launch": prof["ms_per_launch"], "launches_timed": prof["launches"]}
        if fused:
            # the dominant kernel IS the decode step: one persistent kernel per generated token
            a2 = step_bytes / (step_ms * 1e-3) / 1e9
            out["roofline"] = {"kernel": "decode_step_kernel (fused persistent decode step: 24 layers + LM head, one "
                                         "launch per generated token)",
                               "bound": "hbm", "achieved": a2, "peak": peak, "unit": "GB/s", "frac": a2 / peak,
                               "traffic": tr_step if (full and args.max_length == 512) else None,
                               "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the launch at cached "
                                                 "length 255 (= the mean length of this run), parsed from the committed "
                                                 f"ncu --set full summary {tr_step_src}",
                               "peak_source": peak_src,
                               "algorithmic_bytes_per_launch": step_bytes, "ms_per_launch": step_ms,
                               "launches_timed": steps_run * args.steps,
                               "note": "algorithmic bytes = decoder weights (4 B/param) + B x (kv24 cross K/V + fp32 self "
                                       "K/V at the mean cached length); the kernel alternates HBM-bound attention "
                                       "phases with latency-bound linears separated by grid barriers",
                               "cross_attention_alone": cross}
        elif nbeams > 1 or B != 32:
            a2 = step_bytes / (step_ms * 1e-3) / 1e9
            out["roofline"] = {"kernel": "one decode step of the per-operation kernel chain (skinny_tc_kernel linears, "
                                         "self-attention, kv24 streaming cross-attention; CUDA-graph replay)",
                               "bound": "hbm", "achieved": a2, "peak": peak, "unit": "GB/s", "frac": a2 / peak,
                               "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": step_bytes,
                               "ms_per_launch": step_ms, "launches_timed": steps_run * args.steps}
        else:
            cross["peak_source"] = peak_src
            out["roofline"] = cross
        if ahead is not None:
            out["run_ahead"] = ahead
        if world == 1 and not args.no_cpu_baseline and args.workload == "generate":
            threads = os.cpu_count() or 1
            c = cpu_bounded_sample(threads, B, args.max_length)
            out["cpu_baseline"] = {"value": c["images_per_s"], "unit": UNIT, "cores": threads, "kind": "port",
                                   "sample": c["sample"], "t_encode_s": c["t_encode_s"], "t_step_s": c["t_step_s"]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ configs[2]: encoder
def synth_inputs_ragged(image_size, batch, max_text, seed, vocab):
    """USPTO-shape batch: per-image OCR text length ~ U[32, max_text], padded to max_text with mask 0"""
    import torch

    base = synth_inputs(image_size, batch, max_text, seed, vocab)
    g = torch.Generator().manual_seed(seed + 7)
    lens = torch.randint(32, max_text + 1, (batch,), generator=g)
    lens[0] = max_text
    for b in range(batch):
        n = int(lens[b])
        base["input_ids"][b, n - 1] = 1
        base["bbox"][b, n - 1] = 1.0
        base["input_ids"][b, n:] = 0
        base["bbox"][b, n:] = 0.0
        base["attention_mask"][b, n:] = 0
    return base


def encoder_work(cfg, S):
    """algorithmic work of ONE image through mg_encode (Swin-B + projector + VTL encoder at S = text + patches positions):
    flops = 2 x multiply-adds of the reference's dense contractions (TF/models/swin/modeling_swin.py, modeling_udop.py);
    bytes = weights excluded, the fp32 activations a layer must read and write once (residual stream in/out of both
    sub-blocks, normalised copy, q/k/v, context, FF hidden) -- the S x S score / probability matrices are NOT counted,
    a fused attention never moves them"""
    d, dff = cfg.d_model, cfg.d_ff
    fl = 0.0
    res, C = cfg.swin_image // cfg.swin_patch, cfg.swin_embed
    fl += 2.0 * res * res * (3 * cfg.swin_patch ** 2) * C
    by = 0.0
    for s, (depth, heads) in enumerate(zip(cfg.swin_depths, cfg.swin_heads)):
        T, ws = res * res, min(cfg.swin_window, res)
        per_block = 2.0 * T * (3 * C * C + C * C + 8 * C * C) + 2.0 * (T // (ws * ws)) * heads * 2 * (ws * ws) ** 2 * (C // heads)
        fl += depth * per_block
        by += depth * T * C * 4.0 * (4 + 2 + 3 + 1 + 4 + 4)
        if s + 1 < len(cfg.swin_depths):
            fl += 2.0 * (T // 4) * (4 * C) * (2 * C)
            res //= 2
            C *= 2
    n_sw = res * res
    fl += 2.0 * n_sw * (C * cfg.proj_hidden + cfg.proj_hidden * d)
    fl += 2.0 * cfg.n_patches * (3 * cfg.patch_size ** 2) * d
    fl += cfg.num_layers * (8.0 * S * d * d + 4.0 * S * d * dff + 4.0 * S * S * d)
    by += cfg.num_layers * S * 4.0 * (4 * d + 2 * d + 3 * d + d + 2 * dff)
    return fl, by


def run_enc256(args):
    """BASELINE.json configs[2]: batch-256 USPTO-shape synthetic images + OCR boxes, VTL encoder on, 1 GPU: images/s of
    mg_encode with achieved TFLOP/s (the binding roof: tensor cores) and algorithmic GB/s, each against its measured peak"""
    import torch

    from markushgrapher_b200.configuration import MarkushgrapherConfig, random_state
    from markushgrapher_b200.engine import MGEngine

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    cfg = MarkushgrapherConfig()
    eng = MGEngine(cfg, random_state(cfg, seed=0, device=dev), precision=0, device=dev)
    torch.cuda.empty_cache()
    B, Lt = (args.batch if args.batch != BATCH else 256), 256
    host = {k: v.pin_memory() for k, v in synth_inputs_ragged(cfg.image_size, B, Lt, 1237, cfg.vocab_size).items()}
    devin = {k: v.to(dev) for k, v in host.items()}
    stream = torch.cuda.Stream(device=dev)
    sampler = ClockSampler(dev.index)
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            eng.encode(**devin)
        torch.cuda.synchronize()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        ev0.record(stream)
        for _ in range(args.steps):
            mem, mask = eng.encode(**devin)
        ev1.record(stream)
        torch.cuda.synchronize()
        launches = eng.launch_count() - l0
        ms = ev0.elapsed_time(ev1) / args.steps
        mask_h = torch.empty(mask.shape, dtype=mask.dtype).pin_memory()
        t0 = time.perf_counter()
        for _ in range(args.steps):  # e2e: pinned host inputs -> H2D -> encode -> D2H of the memory mask
            d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            mem, mask = eng.encode(**d)
            mask_h.copy_(mask, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        ms_e2e = (time.perf_counter() - t0) * 1000.0 / args.steps
        clocks = sampler.stop()
    S = Lt + cfg.n_patches
    fl, by = encoder_work(cfg, S)
    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        peaks = json.load(open(pth))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1407.0))
    hbm_peak, hbm_src = read_peaks()
    w_bytes = 4.0 * (cfg.num_layers * (4 * cfg.d_model ** 2 + 2 * cfg.d_model * cfg.d_ff) + 87e6)
    tfs = fl * B / (ms * 1e-3) / 1e12
    gbs = (by * B + w_bytes) / (ms * 1e-3) / 1e9
    out = {"metric": "images/sec OCSR+VTL encoder (batch256 USPTO-shape, text<=256)", "value": B / (ms * 1e-3),
           "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32 (GEMM operands as hi+lo bf16 planes on tcgen05: 3 bf16 MMA products per fp32 product)",
           "data": "synthetic",
           "config": {"workload": f"configs[2]: batch-{B} USPTO-shape synthetic 512x512 images + OCR boxes (ragged text "
                                  f"32..{Lt} tokens padded to {Lt}), Swin-B + projector + 24-layer VTL encoder, 1 GPU",
                      "images_per_gpu": B, "text_len": Lt, "S": S,
                      "l2": "activations per layer far larger than L2 (residual stream alone {:.1f} GB)".format(
                          min(B, 64) * S * cfg.d_model * 4 / 1e9)},
           "clocks": clocks, "gpu_launches": int(launches),
           "e2e": {"value": B / (ms_e2e * 1e-3), "unit": UNIT,
                   "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host.values()),
                   "d2h_bytes_per_step": mask_h.numel() * mask_h.element_size()},
           "roofline": {"kernel": "mg_encode (tcgen05 GEMMs + fused encoder attention; the whole encoder pass)",
                        "bound": "tensor", "achieved": tfs, "peak": tf_peak, "unit": "TFLOP/s", "frac": tfs / tf_peak,
                        "traffic": None,
                        "note": "achieved = fp32-equivalent algorithmic flops (2 x MACs of the reference's contractions at "
                                "the padded S) / time; the split-bf16 policy issues 3 bf16 MMA products per fp32 product, "
                                "so the tensor pipe carries `issued_bf16_tflops`",
                        "issued_bf16_tflops": 3.0 * tfs, "issued_frac_of_peak": 3.0 * tfs / tf_peak,
                        "algorithmic_flops_per_image": fl,
                        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1407 TF/s"},
           "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                   "algorithmic_bytes_per_image": by, "weights_bytes": w_bytes, "peak_source": hbm_src}}
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cfg_o, model = _oracle_model(threads)
        n = 2
        sub = {k: v[:n] for k, v in host.items()}
        t0 = time.perf_counter()
        model.encode(**sub)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"oracle.encode of {n} images of the batch ({dt:.1f} s), fp32, {threads} torch threads"}
    print(json.dumps(out))


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
    ap.add_argument("--no-ahead", dest="no_ahead", action="store_true", help="skip the opt-in encoder run-ahead measurement")
    ap.add_argument("--workload", default="generate", choices=["generate", "enc256", "gen128", "beam4"],
                    help="generate = BASELINE.json configs[1] (the bench line); enc256 = configs[2] (encoder only); "
                         "gen128 = configs[3] per-GPU shape (128 images, greedy); beam4 = configs[4] (125 images, beam 4, "
                         "<=768 tok, text <=512)")
    ap.add_argument("--num-beams", dest="num_beams", type=int, default=1)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "enc256":
        run_enc256(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
