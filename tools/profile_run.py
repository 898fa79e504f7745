"""Short run of the hot path for ncu: full-size model, batch 32, a handful of decode steps."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from markushgrapher_b200.configuration import MarkushgrapherConfig, random_state
from markushgrapher_b200.engine import MGEngine

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--max-length", type=int, default=6)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--precision", type=int, default=0)
ap.add_argument("--num-beams", type=int, default=1)
ap.add_argument("--text-len", type=int, default=bench.TEXT_LEN)
a = ap.parse_args()
cfg = MarkushgrapherConfig()
dev = torch.device("cuda", 0)
eng = MGEngine(cfg, random_state(cfg, 0, dev), precision=a.precision, device=dev)
inp = {k: v.to(dev) for k, v in bench.synth_inputs(cfg.image_size, a.batch, a.text_len, 1234, cfg.vocab_size).items()}
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(a.reps):
        eng.generate(**inp, max_length=a.max_length, num_beams=a.num_beams, trim=False)
torch.cuda.synchronize()
print(eng.last_stats())
