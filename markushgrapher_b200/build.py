"""Builds libmg_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = ["gemm_tc.cu", "ops.cu", "enc_flash.cu", "swin.cu", "decode.cu", "decode_mega.cu", "beam.cu", "model.cu", "api_ops.cu", "pack.cu", "detok.cu"]
OUT = os.path.join(HERE, "lib", "libmg_b200.so")


def build(force: bool = False, verbose: bool = False) -> str:
    csrc = os.path.join(HERE, "csrc")
    srcs = [os.path.join(csrc, s) for s in SRC]
    deps = srcs + [os.path.join(csrc, h) for h in ("ptx.cuh", "mg_internal.h", "kernels.h")] + [
        os.path.join(ROOT, "include", "mg_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objdir = os.path.join(HERE, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    common = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + csrc]
    common += os.environ.get("MG_B200_CFLAGS", "").split()  # e.g. -DMK_FINE for the in-kernel phase profile
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        if not force and os.path.exists(o) and all(os.path.getmtime(o) >= os.path.getmtime(d) for d in deps if d.endswith(('.h', '.cuh')) or d == s):
            continue
        procs.append((s, subprocess.Popen(common + ["-c", s, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError(f"nvcc failed on {s}")
        if verbose and out:
            print(out.decode())
    subprocess.check_call(["nvcc", "-shared", "-o", OUT] + objs + ["-lcudart"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
