"""Per-phase timing of the fused decode step from the in-kernel globaltimer stamps (MG_MEGA_PROF=<file>).
usage: MG_MEGA_PROF=gpurun_out/mega_prof.bin python tools/profile_run.py --max-length 260 ; python tools/mega_phase_profile.py gpurun_out/mega_prof.bin
For every phase (7 per layer + LM head): work = time from the previous barrier exit to this CTA's arrival,
wait = time spent in the grid barrier; reported as max / mean over CTAs, averaged over layers."""
import sys

import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 512, 2).astype(np.int64)
G = a.shape[0]
NL = int(sys.argv[2]) if len(sys.argv) > 2 else 24
names = ["qkv|cq", "self", "o|cq", "cross", "co", "wi", "wo"]
NPH = len(names)
nph = NL * NPH
arrive, leave = a[:, :nph + 1, 0], a[:, :nph, 1]
start = np.concatenate([np.full((G, 1), arrive[:, 0].min() - 1), leave], axis=1)  # phase start per CTA (phase 0: unknown)
work = arrive[:, :nph + 1] - start[:, :nph + 1]
wait = leave - arrive[:, :nph]
t0 = leave[:, 0].min()
print(f"CTAs {G}; step span (first barrier exit -> last LM-head done): {(arrive[:, nph].max() - t0) / 1e3:.1f} us")
print(f"{'phase':8s} {'dur(us)':>8s} {'work max':>9s} {'work mean':>9s} {'wait mean':>9s}")
tot = 0.0
for k, nm in enumerate(names):
    idx = np.arange(k, nph, NPH)
    idx = idx[idx > 0]
    # phase duration = last barrier exit of this phase - last barrier exit of the previous phase
    dur = (leave[:, idx].max(0) - leave[:, idx - 1].max(0)).mean() / 1e3
    tot += dur
    print(f"{nm:8s} {dur:8.2f} {work[:, idx].max(0).mean() / 1e3:9.2f} {work[:, idx].mean() / 1e3:9.2f} {wait[:, idx].mean() / 1e3:9.2f}")
lm = (arrive[:, nph].max() - leave[:, nph - 1].max()) / 1e3
print(f"{'lm_head':8s} {lm:8.2f}")
print(f"per layer {tot:.1f} us -> {tot * NL / 1e3 + lm / 1e3:.3f} ms/step")

# ---- fine stamps (profiling build: MG_B200_CFLAGS=-DMK_FINE): linear phases of layer 1, per CTA
raw = a
lin = [(0, "qkv|cq"), (2, "o|cq"), (4, "co"), (5, "wi"), (6, "wo")]
if raw[:, 256:384].any():
    print("\nlinear phases of layer 1 (us since the worker's phase start; mean over CTAs with work)")
    print(f"{'phase':6s} {'x loaded':>9s} {'staged':>9s} {'rs sync':>9s} {'tmem full':>9s} {'epi done':>9s} {'stat done':>9s} | {'mma:xrdy':>9s} {'W full':>9s} {'commit':>9s} {'xrdy last':>9s} {'W last':>9s} | start-exit")
    for j, (k, nm) in enumerate(lin):
        f = raw[:, 256 + k * 16:256 + k * 16 + 16, :].reshape(G, 32)
        fm = raw[:, 400 + j * 4:400 + j * 4 + 4, :].reshape(G, 8)
        ok = (f[:, 5] > 0) & (fm[:, 3] > 0)
        base = f[ok, 0]
        ex = raw[ok, NPH + k - 1, 1]
        cols = [(f[ok, i] - base).mean() / 1e3 for i in (1, 2, 3, 4, 5, 10)] + [(fm[ok, i] - base).mean() / 1e3 for i in (1, 2, 3, 4, 5)]
        print(f"{nm:6s} " + " ".join(f"{c:9.2f}" for c in cols[:6]) + " | " + " ".join(f"{c:9.2f}" for c in cols[6:]) + f" | {(base - ex).mean() / 1e3:6.2f}")
