"""CPU: the host half of the GPU input packer -- Pillow's resampling coefficient tables -- must reproduce PIL's
resize bit for bit when applied with plain integer arithmetic (numpy), for up- and down-scaling, both filters."""
import numpy as np
import pytest
from PIL import Image

from markushgrapher_b200 import packing


def numpy_resize(arr, out_hw, filt):
    """Pillow's two-pass fixed-point resampler with the library's coefficient tables (horizontal pass first)"""
    H, W, _ = arr.shape
    Ho, Wo = out_hw
    a = arr.astype(np.int64)
    if W != Wo:
        ks, bounds, kk = packing.resample_coeffs(W, Wo, filt)
        tmp = np.zeros((H, Wo, 3), dtype=np.int64)
        for xx in range(Wo):
            x0, n = bounds[xx]
            acc = (1 << 21) + (a[:, x0:x0 + n, :] * kk[xx, :n].astype(np.int64)[None, :, None]).sum(1)
            tmp[:, xx, :] = np.clip(acc >> 22, 0, 255)
        a = tmp
    if H != Ho:
        ks, bounds, kk = packing.resample_coeffs(H, Ho, filt)
        out = np.zeros((Ho, a.shape[1], 3), dtype=np.int64)
        for yy in range(Ho):
            y0, n = bounds[yy]
            acc = (1 << 21) + (a[y0:y0 + n, :, :] * kk[yy, :n].astype(np.int64)[:, None, None]).sum(0)
            out[yy] = np.clip(acc >> 22, 0, 255)
        a = out
    return a.astype(np.uint8)


@pytest.mark.parametrize("hw", [(300, 420), (640, 512), (512, 700), (1024, 768), (77, 1300), (512, 512), (2200, 1700)])
@pytest.mark.parametrize("filt,pil", [(packing.LANCZOS, Image.LANCZOS), (packing.BILINEAR, Image.BILINEAR)])
def test_coefficient_tables_reproduce_pil_resize(hw, filt, pil):
    rng = np.random.default_rng(hw[0] * 7 + hw[1])
    arr = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    arr[: hw[0] // 3, : hw[1] // 2] = 255  # flat white regions + hard edges (ringing / clipping of the Lanczos lobes)
    arr[hw[0] // 2:, hw[1] // 3: hw[1] // 3 + 2] = 0
    ref = np.asarray(Image.fromarray(arr).resize((512, 512), resample=pil))
    got = numpy_resize(arr, (512, 512), filt)
    assert np.array_equal(got, ref)


def test_pad_batch_shapes_and_mask():
    import torch

    encs = [{"input_ids": torch.arange(5)[None], "bbox": torch.rand(1, 5, 4)},
            {"input_ids": torch.arange(9)[None], "bbox": torch.rand(1, 9, 4), "attention_mask": torch.ones(1, 9, dtype=torch.long)}]
    out = packing.pad_batch(encs)
    assert out["input_ids"].shape == (2, 9) and out["bbox"].shape == (2, 9, 4)
    assert out["attention_mask"].sum(1).tolist() == [5, 9]
    assert (out["input_ids"][0, 5:] == 0).all() and (out["bbox"][0, 5:] == 0).all()
