"""Reader of the -DMK_XPROF cycle accounts of the fused decode step's cross-attention phase (decode_mega.cu).
usage: tools/build_variant.sh xprof -DMK_XPROF
       MG_B200_LIB=.../libmg_b200_xprof.so MG_MEGA_PROF=gpurun_out/xprof.bin python tools/profile_run.py --max-length 260
       python tools/cross_phase_cycles.py gpurun_out/xprof.bin [sm_mhz]"""
import sys

import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.uint64)
mhz = float(sys.argv[2]) if len(sys.argv) > 2 else 1965.0
n_cta = a.size // 1024
a = a.reshape(n_cta, 512, 2).reshape(n_cta, 1024)
prod = a[:, 880:883].astype(np.float64)   # slot 440: wait, issue, chunks
cons = a[:, 896:904].astype(np.float64)   # slot 448: wait, math, sync, head, soft, tail, chunks, total
act = cons[:, 7] > 0
us = lambda c: c / mhz
print(f"CTAs with cross work: {int(act.sum())} of {n_cta}; chunks per CTA: mean {cons[act, 6].mean():.1f}")
names = ["wait for data", "arithmetic", "barrier+release", "item head (q/mask loads)", "softmax", "item tail (reduce, ctx)"]
tot = cons[act, 7]
print(f"consumer thread 0, microseconds per phase (mean / max over CTAs); phase total {us(tot).mean():.2f} / {us(tot).max():.2f}")
for i, nm in enumerate(names):
    print(f"  {nm:28s} {us(cons[act, i]).mean():7.2f} / {us(cons[act, i]).max():7.2f}   per chunk {us(cons[act, i] / np.maximum(cons[act, 6], 1)).mean():.3f}")
rest = tot - cons[act, :6].sum(axis=1)
print(f"  {'unaccounted':28s} {us(rest).mean():7.2f}")
pa = prod[:, 2] > 0
print(f"producer, microseconds per phase (mean): wait for a free stage {us(prod[pa, 0]).mean():.2f}, issue {us(prod[pa, 1]).mean():.2f} "
      f"({us(prod[pa, 1] / prod[pa, 2]).mean() * 1e3:.0f} ns per chunk), chunks {prod[pa, 2].mean():.1f}")

sf = a[:, 928:936].astype(np.float64)   # slot 464: self-attention phase, group 0 thread 0
sa = sf[:, 7] > 0
if sa.any():
    print(f"self-attention phase (group 0 of {int(sa.sum())} CTAs), microseconds per phase (mean / max); total {us(sf[sa, 7]).mean():.2f} / {us(sf[sa, 7]).max():.2f}")
    for i, nm in enumerate(["wait for data", "K arithmetic", "V arithmetic", "group barrier+release", "item head (q/k/v, bias loads)", "softmax", "item tail (combine, ctx)"]):
        print(f"  {nm:30s} {us(sf[sa, i]).mean():7.2f} / {us(sf[sa, i]).max():7.2f}")
    print(f"  {'unaccounted':30s} {us(sf[sa, 7] - sf[sa, :7].sum(axis=1)).mean():7.2f}")
