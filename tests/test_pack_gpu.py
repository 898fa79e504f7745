"""GPU: mg_pack_pixels (Pillow-exact resize + image-processor normalisation on the device) against PIL + the host
image processor on the same bytes: bit exact (uint8 / fixed-point work, and explicitly rounded fp32 normalisation)."""
import numpy as np
import pytest
import torch
from PIL import Image

from markushgrapher_b200 import packing
from markushgrapher_b200.processing import MarkushgrapherImageProcessor

pytestmark = pytest.mark.gpu


def reference_pixel_values(arrs, pil_filter):
    proc = MarkushgrapherImageProcessor()
    ims = [Image.fromarray(a).resize((512, 512), resample=pil_filter) for a in arrs]  # mdu_dataset.py:118
    return proc(ims)["pixel_values"]                                                  # utils/common.py:34-42


@pytest.mark.parametrize("hw", [(300, 420), (1024, 768), (512, 700), (640, 512), (512, 512), (2200, 1700), (64, 48)])
@pytest.mark.parametrize("filt,pil", [(packing.LANCZOS, Image.LANCZOS), (packing.BILINEAR, Image.BILINEAR)])
def test_pack_pixels_bit_exact_vs_pil(hw, filt, pil):
    rng = np.random.default_rng(hw[0] + 31 * hw[1])
    arrs = [rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8) for _ in range(3)]
    arrs[1][: hw[0] // 2] = 255
    arrs[2][:, hw[1] // 4: hw[1] // 4 + 3] = 0
    ref = reference_pixel_values(arrs, pil)
    got = packing.pack_pixels(torch.from_numpy(np.stack(arrs)).cuda(), (512, 512), filt)
    assert got.shape == (3, 3, 512, 512)
    assert torch.equal(got.cpu(), ref)


def test_pack_pil_images_mixed_sizes_keeps_order():
    rng = np.random.default_rng(5)
    sizes = [(300, 400), (512, 512), (300, 400), (900, 650)]
    ims = [Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)) for h, w in sizes]
    got = packing.pack_pil_images(ims, torch.device("cuda"))
    ref = reference_pixel_values([np.asarray(im) for im in ims], Image.LANCZOS)
    assert torch.equal(got.cpu(), ref)
