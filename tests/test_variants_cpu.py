"""CPU: the flagged experiment / profiling builds of the fused decode step keep compiling for sm_100a
(-DMK_FINE in-kernel phase stamps, -DMK_PF_LANE dedicated L2-prefetch lane, -DMK_SELF_ALL self-attention over all
CTAs, -DMK_FOLD_FF folded FF input projection, -DMK_FLAGBAR flag-array grid barrier, -DMK_RACECHECK progress flags
through atomics, -DMK_XPROF cycle accounts of the attention phases, -DMK_WARPREL per-warp stage release, -DMK_OCC=2
-DMK_NST_N=2 two instances per SM). nvcc cross-compiles without a GPU; nothing is executed."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "markushgrapher_b200", "csrc")


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
@pytest.mark.parametrize("flags", [("-DMK_FINE", "-DMK_PF_LANE", "-DMK_SELF_ALL"),
                                   ("-DMK_FOLD_FF", "-DMK_FLAGBAR", "-DMK_RACECHECK"),
                                   ("-DMK_XPROF", "-DMK_WARPREL", "-DMK_OCC=2", "-DMK_NST_N=2", "-DMK_XUNROLL=4")])
def test_flagged_builds_of_the_fused_decode_step_compile(tmp_path, flags):
    out = tmp_path / "decode_mega_variants.o"
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
           "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, *flags,
           "-c", os.path.join(CSRC, "decode_mega.cu"), "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert out.stat().st_size > 0
