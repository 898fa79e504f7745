"""images/s of the input packer: GPU mg_pack_pixels (uint8 host batch -> H2D -> pixel_values on the device) vs the
reference's per-sample CPU path (PIL LANCZOS resize + image processor + H2D of the fp32 result)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image

from markushgrapher_b200 import packing
from markushgrapher_b200.processing import MarkushgrapherImageProcessor

B = 32
dev = torch.device("cuda", 0)
proc = MarkushgrapherImageProcessor()
for hw in [(1024, 768), (2200, 1700)]:
    rng = np.random.default_rng(1)
    arr = rng.integers(0, 256, size=(B, hw[0], hw[1], 3), dtype=np.uint8)
    host = torch.from_numpy(arr).pin_memory()
    for _ in range(3):
        packing.pack_pixels(host.to(dev, non_blocking=True))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        out = packing.pack_pixels(host.to(dev, non_blocking=True))
    torch.cuda.synchronize()
    gpu = B * n / (time.perf_counter() - t0)
    ims = [Image.fromarray(a) for a in arr[:8]]
    t0 = time.perf_counter()
    ref = proc([im.resize((512, 512), resample=Image.LANCZOS) for im in ims])["pixel_values"].to(dev)
    torch.cuda.synchronize()
    cpu = len(ims) / (time.perf_counter() - t0)
    same = torch.equal(out[:8].cpu(), ref.cpu())
    print(f"{hw[0]}x{hw[1]} -> 512x512 LANCZOS + normalise: GPU {gpu:9.1f} img/s (incl. H2D of uint8), "
          f"CPU PIL path {cpu:7.1f} img/s (1 core), bit-identical: {same}")
