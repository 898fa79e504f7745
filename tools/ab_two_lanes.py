"""Feasibility measurement for a two-lane decode: does running the bench batch as TWO independent
half-batches, each on its own persistent decode_step_kernel instance that owns half of the SMs, overlap one lane's
latency-bound linear phases with the other lane's HBM-bound attention streams?

Two full-size engines live in one process (each with its own weights -> the weight stream is paid twice, as a real
two-lane kernel without L2 sharing would), each decodes half of the batch from its own thread on its own stream with
MG_MEGA_CTAS=<half the SMs>.  Reported: per-lane ms/step (device events / per-step stamps) alone and together, and
the wall time of the concurrent pair; compare with `tools/ab_env.py` (one lane, all SMs, whole batch).

    MG_MEGA_CTAS=74 python tools/ab_two_lanes.py [--max-length 512] [--batch 32]
"""
import argparse
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import BATCH, TEXT_LEN, synth_inputs
from markushgrapher_b200.configuration import MarkushgrapherConfig, random_state
from markushgrapher_b200.engine import MGEngine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-length", type=int, default=512)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--stagger-ms", type=float, default=0.0, help="start lane 1 this much after lane 0")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = MarkushgrapherConfig()
    state = random_state(cfg, seed=0, device=dev)
    engs = [MGEngine(cfg, state, precision=0, device=dev) for _ in range(2)]
    del state
    inp = {k: v.to(dev) for k, v in synth_inputs(cfg.image_size, a.batch, TEXT_LEN, seed=1234, vocab=cfg.vocab_size).items()}
    half = a.batch // 2
    parts = [{k: v[:half].contiguous() for k, v in inp.items()}, {k: v[half:].contiguous() for k, v in inp.items()}]
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    print(f"MG_MEGA_CTAS={os.environ.get('MG_MEGA_CTAS', '(all)')} MG_MEGA_NOCOOP={os.environ.get('MG_MEGA_NOCOOP', '0')} "
          f"batch {a.batch} -> 2 x {half}", flush=True)

    def run(i, out, delay=0.0):
        torch.cuda.set_device(dev)
        if delay:
            time.sleep(delay)
        with torch.cuda.stream(streams[i]):
            ids = engs[i].generate(**parts[i], max_length=a.max_length, trim=False)
            streams[i].synchronize()
        lp = engs[i].last_decode_loop()
        out[i] = (ids, lp)

    # each lane alone (the other half of the chip idle)
    alone = {}
    for i in range(2):
        for _ in range(a.reps):
            run(i, alone)
        lp = alone[i][1]
        print(f"lane {i} alone   : {lp['loop_ms'] / max(1, lp['steps']):.4f} ms/step mean, {lp['step_p50_ms']:.4f} p50, "
              f"steps {lp['steps']}, fused {lp['fused']}", flush=True)
    # both lanes concurrently
    for rep in range(a.reps):
        both = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=run, args=(i, both, a.stagger_ms * 1e-3 * i)) for i in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        for i in range(2):
            lp = both[i][1]
            same = bool(torch.equal(both[i][0], alone[i][0]))
            print(f"lane {i} together: {lp['loop_ms'] / max(1, lp['steps']):.4f} ms/step mean, {lp['step_p50_ms']:.4f} p50, "
                  f"p99 {lp['step_p99_ms']:.4f}, steps {lp['steps']}, ids identical to alone {same}", flush=True)
        print(f"pair wall (encode + decode of both halves): {wall:.1f} ms", flush=True)
    for e in engs:
        e.close()


if __name__ == "__main__":
    main()
