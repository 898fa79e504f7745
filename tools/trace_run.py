"""Runs a few decode steps with the MG_TRACE build and prints per-phase times (ns, %globaltimer) of CTA (0,0)
of the last 64 skinny-linear launches."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from markushgrapher_b200 import _lib
from markushgrapher_b200.configuration import MarkushgrapherConfig, random_state
from markushgrapher_b200.engine import MGEngine

cfg = MarkushgrapherConfig()
dev = torch.device("cuda", 0)
eng = MGEngine(cfg, random_state(cfg, 0, dev), device=dev)
inp = {k: v.to(dev) for k, v in bench.synth_inputs(512, 32, 64, 1234, cfg.vocab_size).items()}
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    eng.generate(**inp, max_length=40, trim=False)
torch.cuda.synchronize()
buf = (ctypes.c_uint64 * (64 * 16))()
n = ctypes.c_int(0)
_lib.lib().mg_trace_dump(buf, ctypes.byref(n))
print("launches traced:", n.value)
names = ["start", "wait_done", "x_staged", "zero_done", "stats_reduced", "stats_bar", "pre_epi", "mma_done", "atomics_issued"]
rows = []
for slot in range(64):
    t = [buf[slot * 16 + i] for i in range(10)]
    if t[0] == 0:
        continue
    meta = t[9]
    rows.append((t[0], (meta >> 32, (meta >> 16) & 0xffff, meta & 0xffff), [(t[i] - t[0]) if t[i] else 0 for i in range(9)]))
rows.sort()
base = rows[-20][0]
print("absolute timeline (ns) of the last launches: start | dep_wait_done | atomics_issued | kernel")
for t0, (gx, gy, pro), d in rows[-20:]:
    print(f"  {t0 - base:8d} {t0 - base + d[1]:8d} {t0 - base + d[8]:8d}  grid({gx},{gy}) pro{pro}  [work after wait: {d[8] - d[1]}]")
prev_end = None
for t0, (gx, gy, pro), d in rows[-16:]:
    gap = (t0 - prev_end) if prev_end else 0
    print(f"grid({gx:3d},{gy:2d}) pro{pro} gap_from_prev_atomics {gap:6d} | " + " ".join(f"{names[i]}={d[i] - d[1]:5d}" for i in range(2, 9)))
    prev_end = t0 + d[8]
