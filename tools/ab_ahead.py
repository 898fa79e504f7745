"""Encoder run-ahead A/B on the bench workload: K batches decoded one after the other, (a) plain generate calls,
(b) with the encoder of batch i+1 queued on its own SM partition before batch i is decoded (Engine.encode_ahead).
Prints ms per batch (device events around the K calls), per-step decode times, and checks the ids are identical.
    python tools/ab_ahead.py [--batches 4] [--max-length 512]      (MG_AHEAD_SMS=16 sizes the encoder partition)"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import BATCH, TEXT_LEN, synth_inputs
from markushgrapher_b200.configuration import MarkushgrapherConfig, random_state
from markushgrapher_b200.engine import MGEngine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-length", type=int, default=512)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--batches", type=int, default=4)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = MarkushgrapherConfig()
    state = random_state(cfg, seed=0, device=dev)
    eng = MGEngine(cfg, state, precision=0, device=dev)
    del state
    sets = [{k: v.to(dev) for k, v in synth_inputs(cfg.image_size, a.batch, TEXT_LEN, seed=1234 + i, vocab=cfg.vocab_size).items()}
            for i in range(2)]
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        ref = [eng.generate(**sets[i], max_length=a.max_length, trim=False).clone() for i in range(2)]  # warm-up + reference ids
        for mode in ("plain", "ahead", "plain", "ahead"):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if mode == "ahead":
                armed = eng.encode_ahead(**sets[0])  # pipeline fill (outside the timed region, like a warm-up step)
            torch.cuda.synchronize()
            e0.record(st)
            steps, same, enc = [], True, []
            for i in range(a.batches):
                cur, nxt = sets[i % 2], sets[(i + 1) % 2]
                if mode == "ahead":
                    eng.encode_ahead(**nxt)
                ids = eng.generate(**cur, max_length=a.max_length, trim=False)
                lp = eng.last_decode_loop()
                steps.append(lp["loop_ms"] / max(1, lp["steps"]))
                enc.append(eng.last_ahead()["encoder_ms"] if mode == "ahead" else eng.last_stats()["encode_ms"])
                same = same and bool(torch.equal(ids, ref[i % 2]))
            if mode == "ahead":
                eng.ahead_reset()  # the last run-ahead encoder (nobody takes it) still counts
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.batches
            extra = f", partitions {eng.last_ahead()}" if mode == "ahead" else ""
            print(f"[{mode:5s}] {ms:8.1f} ms / batch = {a.batch / ms * 1e3:6.2f} img/s; decode step {sum(steps) / len(steps):.4f} ms; "
                  f"encoder {sum(enc) / len(enc):.1f} ms; ids identical {same}{extra}", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
