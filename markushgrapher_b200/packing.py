"""GPU input packing: page images -> `pixel_values`, ragged (ids, boxes) -> padded batch.

`pack_pixels` is the reference's per-sample CPU preprocessing
    row["page_image"].resize((512, 512), resample=Image.LANCZOS)          (core/datasets/mdu_dataset.py:118)
    processor(images=image.convert("RGB"), ...)["pixel_values"]           (utils/common.py:34-42)
done on the device for a whole batch; the resize is bit-identical to Pillow (csrc/pack.cu).  No CPU fallback:
without the CUDA library the call raises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib

BILINEAR, LANCZOS = 0, 1


def resample_coeffs(in_size: int, out_size: int, filt: int = LANCZOS):
    """Pillow's fixed-point coefficient table of one axis (host only; needs no GPU): (ksize, bounds[out,2], kk[out,ksize])"""
    L = _lib.lib()
    L.mg_resample_coeffs.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_int]
    ks = ctypes.c_int32(0)
    _lib.check(L.mg_resample_coeffs(in_size, out_size, filt, ctypes.addressof(ks), None, None, 0), "mg_resample_coeffs")
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ks.value), dtype=np.int32)
    _lib.check(L.mg_resample_coeffs(in_size, out_size, filt, ctypes.addressof(ks), bounds.ctypes.data, kk.ctypes.data,
                                    kk.size), "mg_resample_coeffs")
    return ks.value, bounds, kk


def pack_pixels(images: torch.Tensor, size=(512, 512), resample: int = LANCZOS, image_mean=(0.5, 0.5, 0.5),
                image_std=(0.5, 0.5, 0.5)) -> torch.Tensor:
    """images: (B, H, W, 3) uint8 RGB on a CUDA device -> (B, 3, size[0], size[1]) float32 pixel_values"""
    if not images.is_cuda:
        raise _lib.MgError("pack_pixels needs the images on a CUDA device; there is no CPU fallback")
    assert images.dtype == torch.uint8 and images.dim() == 4 and images.shape[-1] == 3
    images = images.contiguous()
    B, H, W, _ = images.shape
    out = torch.empty(B, 3, size[0], size[1], dtype=torch.float32, device=images.device)
    mean = (ctypes.c_float * 3)(*image_mean)
    std = (ctypes.c_float * 3)(*image_std)
    L = _lib.lib()
    L.mg_pack_pixels.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p]
    with torch.cuda.device(images.device):
        _lib.check(L.mg_pack_pixels(_lib.cur_stream(), B, H, W, _lib.ptr(images), size[0], size[1], resample, mean, std,
                                    _lib.ptr(out)), "mg_pack_pixels")
    return out


def pack_pil_images(images: Sequence, device, size=(512, 512), resample: int = LANCZOS, **kw) -> torch.Tensor:
    """a list of PIL images of arbitrary sizes: images of equal size share one launch; returns (N,3,H,W) in order"""
    arrs = [np.asarray(im.convert("RGB"), dtype=np.uint8) for im in images]
    out = torch.empty(len(arrs), 3, size[0], size[1], dtype=torch.float32, device=device)
    groups: Dict[tuple, List[int]] = {}
    for i, a in enumerate(arrs):
        groups.setdefault(a.shape[:2], []).append(i)
    for idx in groups.values():
        batch = torch.from_numpy(np.stack([arrs[i] for i in idx])).to(device, non_blocking=True)
        out[torch.tensor(idx, device=device)] = pack_pixels(batch, size, resample, **kw)
    return out


def pad_batch(encodings: Sequence[Dict[str, torch.Tensor]], pad_token_id: int = 0) -> Dict[str, torch.Tensor]:
    """batching / padding of per-sample processor outputs (leading dim 1, reference utils/common.py:68-71) into the
    (B, Lt) tensors generate() takes: ids padded with pad_token_id, boxes with zeros, mask 0 on the padding"""
    Lt = max(int(e["input_ids"].shape[1]) for e in encodings)
    B = len(encodings)
    ids = torch.full((B, Lt), pad_token_id, dtype=torch.long)
    box = torch.zeros(B, Lt, 4, dtype=torch.float32)
    mask = torch.zeros(B, Lt, dtype=torch.long)
    for i, e in enumerate(encodings):
        n = int(e["input_ids"].shape[1])
        ids[i, :n] = e["input_ids"][0]
        box[i, :n] = e["bbox"][0]
        mask[i, :n] = e.get("attention_mask", torch.ones(1, n, dtype=torch.long))[0]
    out = {"input_ids": ids, "bbox": box, "attention_mask": mask}
    if all("pixel_values" in e for e in encodings):
        out["pixel_values"] = torch.cat([e["pixel_values"] for e in encodings])
    return out
