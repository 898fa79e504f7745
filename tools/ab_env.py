"""A/B of runtime switches of the decode path on the bench workload: one full-size engine, every setting run back to
back (the library reads its MG_* switches at each generate call); prints ms/step (mean over the loop and p50 over
16-step windows) and checks the token ids against the first setting's.
    python tools/ab_env.py [--max-length 512] [--settings "" "MG_MEGA_GATE=1" "MG_MEGA_GATE=1,MG_MEGA_INFLIGHT=3" ...]
A/B of two builds: run it twice with MG_B200_LIB=<path to the other libmg_b200.so>."""
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
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--settings", nargs="*", default=["", ""])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = MarkushgrapherConfig()
    state = random_state(cfg, seed=0, device=dev)
    eng = MGEngine(cfg, state, precision=0, device=dev)
    del state
    inp = {k: v.to(dev) for k, v in synth_inputs(cfg.image_size, a.batch, TEXT_LEN, seed=1234, vocab=cfg.vocab_size).items()}
    ref = None
    touched = set()
    for s in a.settings:
        for k in touched:
            os.environ.pop(k, None)
        for kv in filter(None, s.split(",")):
            k, _, v = kv.partition("=")
            os.environ[k] = v
            touched.add(k)
        best = None
        for _ in range(a.reps):
            ids = eng.generate(**inp, max_length=a.max_length, trim=False)
            torch.cuda.synchronize()
            lp = eng.last_decode_loop()
            ms = lp["loop_ms"] / max(1, lp["steps"])
            if best is None or ms < best[0]:
                best = (ms, lp["step_p50_ms"], lp["steps"], lp["fused"])
        if ref is None:
            ref = ids.clone()
        print(f"[{s or 'default':40s}] {best[0]:.4f} ms/step mean, {best[1]:.4f} p50, steps {best[2]}, fused {best[3]}, "
              f"ids identical to first {bool(torch.equal(ids, ref))}", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
