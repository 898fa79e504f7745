"""Batched detokeniser: generated ids -> CXSMILES / substituent-table strings on the device.

Mirrors `MarkushTokenizer.decode_plus_decode_other_tokens` (reference core/common/markush_tokenizer.py:615-670).
`build_table` evaluates each of the reference's string predicates ONCE per vocabulary entry (with the same Python
expressions) and packs the result into a byte table + flags; the CUDA kernels (csrc/detok.cu) follow the control
flow for a whole (B, T) id matrix.  No CPU fallback: without the CUDA library / a device the calls raise.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib

I_OPEN, I_CLOSE_EQ, I_CLOSE_IN, LOC, OTHER, NEXT_SPACE = 1, 2, 4, 8, 16, 32


def build_table(pieces: Sequence[str], vocabulary: Dict[str, str], vocabulary_inverse: Dict[str, str],
                encode_index: bool):
    """pieces[id] = tokenizer.convert_ids_to_tokens(id); vocabulary / vocabulary_inverse = MarkushTokenizer's
    (markush_tokenizer.py:237-285).  Returns (text bytes, text_off int32[V+1], flags uint8[V])."""
    i_open = vocabulary["<i>"] if encode_index else None
    i_close = vocabulary["</i>"] if encode_index else None
    flags = np.zeros(len(pieces), dtype=np.uint8)
    chunks: List[bytes] = []
    off = np.zeros(len(pieces) + 1, dtype=np.int32)
    for i, tok in enumerate(pieces):
        f = 0
        if encode_index and i_open in tok:                                   # :637
            f |= I_OPEN
        if encode_index and tok == i_close:                                  # :633 (token != vocabulary["</i>"])
            f |= I_CLOSE_EQ
        if encode_index and i_close in tok:                                  # :641
            f |= I_CLOSE_IN
        if "loc" in tok and "<" in tok and ">" in tok:                       # :645
            f |= LOC
        if "▁" in tok or "other" in tok:                                     # :661-663 (seen from the previous token)
            f |= NEXT_SPACE
        if "other" in tok and "<" in tok and ">" in tok:                     # :649
            f |= OTHER
            text = vocabulary_inverse[tok] + " " if tok in vocabulary_inverse else tok
        else:
            text = tok[1:] if tok[:1] == "▁" else tok                        # :658-659
        b = text.encode("utf-8")
        chunks.append(b)
        off[i + 1] = off[i] + len(b)
        flags[i] = f
    return b"".join(chunks), off, flags


class BatchedDetokenizer:
    def __init__(self, pieces: Sequence[str], vocabulary: Dict[str, str], vocabulary_inverse: Dict[str, str],
                 encode_index: bool = False, device: Optional[torch.device] = None):
        L = _lib.lib()
        if not torch.cuda.is_available() or not L.mg_device_available():
            raise _lib.MgError("BatchedDetokenizer needs a CUDA device; there is no CPU fallback")
        self.device = torch.device(device or "cuda")
        text, off, flags = build_table(pieces, vocabulary, vocabulary_inverse, encode_index)
        self._h = ctypes.c_void_p()
        L.mg_detok_create.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.POINTER(ctypes.c_void_p)]
        L.mg_detok_measure.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p]
        L.mg_detok_write.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p]
        L.mg_detok_destroy.argtypes = [ctypes.c_void_p]
        L.mg_detok_destroy.restype = None
        with torch.cuda.device(self.device):
            _lib.check(L.mg_detok_create(len(pieces), text, off.ctypes.data, flags.ctypes.data, ctypes.byref(self._h)),
                       "mg_detok_create")

    def decode(self, ids: torch.Tensor, lens: Optional[torch.Tensor] = None) -> List[str]:
        """ids (B, T) int64 on the device (e.g. generate()'s output); lens (B) = tokens to decode per row (default T)"""
        if not ids.is_cuda:
            raise _lib.MgError("BatchedDetokenizer.decode needs the ids on a CUDA device")
        ids = ids.to(torch.int64).contiguous()
        B, T = ids.shape
        if B == 0 or T == 0:
            return [""] * B
        lens32 = None if lens is None else lens.to(device=ids.device, dtype=torch.int32).contiguous()
        row_off = np.zeros(B + 1, dtype=np.int64)
        L = _lib.lib()
        with torch.cuda.device(ids.device):
            _lib.check(L.mg_detok_measure(self._h, _lib.cur_stream(), B, T, _lib.ptr(ids), _lib.ptr(lens32),
                                          row_off.ctypes.data), "mg_detok_measure")
            out = torch.empty(max(int(row_off[B]), 1), dtype=torch.uint8, device=ids.device)
            _lib.check(L.mg_detok_write(self._h, _lib.cur_stream(), B, T, _lib.ptr(ids), _lib.ptr(out)), "mg_detok_write")
        raw = out.cpu().numpy().tobytes()
        return [raw[row_off[b]:row_off[b + 1]].decode("utf-8") for b in range(B)]

    def close(self):
        if self._h:
            _lib.lib().mg_detok_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
