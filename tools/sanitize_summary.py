"""Condenses gpurun_out/<tag>_sanitize_*.log into one text file for profiles/: per run the command, the tool's
summary lines, the pytest verdict and the first reported hazards (if any).
usage: python tools/sanitize_summary.py r2 > profiles/r2_compute_sanitizer.txt"""
import glob
import re
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
for fn in sorted(glob.glob(f"gpurun_out/{tag}_sanitize_*.log")):
    lines = open(fn, errors="replace").read().splitlines()
    print("=" * 100)
    print(fn)
    keep = [l for l in lines if re.search(r"ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|kernels|COMPUTE-SANITIZER$", l)]
    for l in keep[-8:]:
        print("   ", l.strip())
    haz = [l for l in lines if re.search(r"========= (Error|Warning|Invalid|Race|Barrier|Hazard|Program hit)", l)]
    print(f"    reported records: {len(haz)}")
    for l in haz[:12]:
        print("       ", l.strip())
