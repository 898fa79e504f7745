"""Thin ctypes layer over the C ABI (include/mg_b200.h): PyTorch tensors in, PyTorch tensors out.

No arithmetic happens here and there is no fallback: every call lands in libmg_b200.so or raises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib

MG_MAX_SWIN_STAGES = 4


class mg_config(ctypes.Structure):
    _fields_ = [
        ("vocab_size", ctypes.c_int32), ("d_model", ctypes.c_int32), ("d_kv", ctypes.c_int32),
        ("d_ff", ctypes.c_int32), ("num_layers", ctypes.c_int32), ("num_decoder_layers", ctypes.c_int32),
        ("num_heads", ctypes.c_int32),
        ("rel_buckets", ctypes.c_int32), ("rel_max_distance", ctypes.c_int32), ("max_2d", ctypes.c_int32),
        ("image_size", ctypes.c_int32), ("patch_size", ctypes.c_int32),
        ("ln_eps", ctypes.c_float),
        ("swin_image", ctypes.c_int32), ("swin_patch", ctypes.c_int32), ("swin_embed", ctypes.c_int32),
        ("swin_num_stages", ctypes.c_int32),
        ("swin_depths", ctypes.c_int32 * MG_MAX_SWIN_STAGES), ("swin_heads", ctypes.c_int32 * MG_MAX_SWIN_STAGES),
        ("swin_window", ctypes.c_int32),
        ("swin_ln_eps", ctypes.c_float),
        ("proj_hidden", ctypes.c_int32),
        ("precision", ctypes.c_int32),
        ("logit_scale", ctypes.c_float),
        ("decoder_start_token_id", ctypes.c_int32), ("eos_token_id", ctypes.c_int32), ("pad_token_id", ctypes.c_int32),
        ("enc_chunk", ctypes.c_int32),
    ]


def make_c_config(cfg, precision: int = 0, enc_chunk: int = 0) -> mg_config:
    """cfg: any object with the MGConfig attribute names (oracle.mg_oracle.MGConfig or MarkushgrapherConfig)."""
    c = mg_config()
    for k in ("vocab_size", "d_model", "d_kv", "d_ff", "num_layers", "num_decoder_layers", "num_heads", "max_2d",
              "image_size", "patch_size", "swin_image", "swin_patch", "swin_embed", "swin_window", "proj_hidden"):
        setattr(c, k, int(getattr(cfg, k)))
    c.rel_buckets = int(cfg.rel_buckets)
    c.rel_max_distance = int(cfg.rel_max_distance)
    c.ln_eps = float(cfg.ln_eps)
    c.swin_ln_eps = float(cfg.swin_ln_eps)
    depths, heads = list(cfg.swin_depths), list(cfg.swin_heads)
    assert len(depths) == len(heads) <= MG_MAX_SWIN_STAGES
    c.swin_num_stages = len(depths)
    for i, (dp, hd) in enumerate(zip(depths, heads)):
        c.swin_depths[i] = int(dp)
        c.swin_heads[i] = int(hd)
    c.precision = int(precision)
    c.logit_scale = float(getattr(cfg, "logit_scale", cfg.d_model ** -0.5))
    c.decoder_start_token_id = int(getattr(cfg, "decoder_start_token_id", 0))
    c.eos_token_id = int(getattr(cfg, "eos_token_id", 1))
    c.pad_token_id = int(getattr(cfg, "pad_token_id", 0))
    c.enc_chunk = int(enc_chunk)
    return c


class MGEngine:
    """One model instance on the current CUDA device."""

    def __init__(self, cfg, state: Dict[str, torch.Tensor], precision: int = 0, enc_chunk: int = 0,
                 device: Optional[torch.device] = None):
        L = _lib.lib()
        if not torch.cuda.is_available() or not L.mg_device_available():
            raise _lib.MgError("markushgrapher_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.device = torch.device(device or "cuda")
        self._c = make_c_config(cfg, precision, enc_chunk)
        self._h = ctypes.c_void_p()
        L.mg_create.argtypes = [ctypes.POINTER(mg_config), ctypes.POINTER(ctypes.c_void_p)]
        _lib.check(L.mg_create(ctypes.byref(self._c), ctypes.byref(self._h)), "mg_create")
        L.mg_load_weight.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_int64), ctypes.c_int]
        keep = []
        with torch.cuda.device(self.device):
            for name, t in state.items():
                t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
                keep.append(t)
                shp = (ctypes.c_int64 * max(1, t.dim()))(*t.shape)
                rc = L.mg_load_weight(self._h, name.encode(), ctypes.c_void_p(t.data_ptr()), 0, shp, t.dim())
                if rc < 0:
                    _lib.check(rc, f"mg_load_weight({name})")
            torch.cuda.synchronize()
            L.mg_finalize.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            _lib.check(L.mg_finalize(self._h, _lib.cur_stream()), "mg_finalize")
        del keep
        L.mg_encode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7
        L.mg_generate.argtypes = ([ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4 +
                                  [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4)
        L.mg_generate_host.argtypes = ([ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] +
                                       [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3)
        L.mg_last_stats.argtypes = [ctypes.c_void_p] * 4
        L.mg_last_decode_loop.argtypes = [ctypes.c_void_p] * 4
        L.mg_last_decode_p50.argtypes = [ctypes.c_void_p] * 2
        L.mg_last_decode_latency.argtypes = [ctypes.c_void_p] * 3
        L.mg_prefetch_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4 + [ctypes.c_int]
        L.mg_destroy.argtypes = [ctypes.c_void_p]
        L.mg_destroy.restype = None
        self.n_patches = (cfg.image_size // cfg.patch_size) ** 2
        g = cfg.swin_image // cfg.swin_patch // 2 ** (len(list(cfg.swin_depths)) - 1)
        self.swin_tokens = g * g

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().mg_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _prep(self, input_ids, bbox, pixel_values, attention_mask, dev):
        ids = input_ids.to(device=dev, dtype=torch.int64).contiguous()
        box = bbox.to(device=dev, dtype=torch.float32).contiguous()
        px = pixel_values.to(device=dev, dtype=torch.float32).contiguous()
        am = None if attention_mask is None else attention_mask.to(device=dev, dtype=torch.int64).contiguous()
        B, Lt = ids.shape
        assert box.shape == (B, Lt, 4), box.shape
        assert px.shape == (B, 3, self.cfg.image_size, self.cfg.image_size), px.shape
        return ids, box, px, am, B, Lt

    def encode(self, input_ids, bbox, pixel_values, attention_mask=None):
        """-> (memory (B, M, d) f32, mask (B, M) i32) on the device"""
        ids, box, px, am, B, Lt = self._prep(input_ids, bbox, pixel_values, attention_mask, self.device)
        M = self.swin_tokens + Lt + self.n_patches
        out = torch.empty((B, M, self.cfg.d_model), device=self.device, dtype=torch.float32)
        mask = torch.empty((B, M), device=self.device, dtype=torch.int32)
        m_out = ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mg_encode(self._h, _lib.cur_stream(), B, Lt, _lib.ptr(ids), _lib.ptr(box), _lib.ptr(px),
                                      _lib.ptr(am), _lib.ptr(out), _lib.ptr(mask), ctypes.addressof(m_out))
        _lib.check(rc, "mg_encode")
        assert m_out.value == M
        return out, mask

    def generate(self, input_ids, bbox, pixel_values, attention_mask=None, num_beams=1, max_length=512,
                 return_logits=False, trim=True):
        """Device-resident inputs -> LongTensor (B, T) of token ids (column 0 = decoder start id)."""
        ids, box, px, am, B, Lt = self._prep(input_ids, bbox, pixel_values, attention_mask, self.device)
        out = torch.empty((B, max_length), device=self.device, dtype=torch.int64)
        lens = torch.empty((B,), device=self.device, dtype=torch.int32)
        logits = None
        if return_logits:
            logits = torch.zeros((B, max_length - 1, self.cfg.vocab_size), device=self.device, dtype=torch.float32)
        steps = ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mg_generate(self._h, _lib.cur_stream(), B, Lt, _lib.ptr(ids), _lib.ptr(box), _lib.ptr(px),
                                        _lib.ptr(am), int(num_beams), int(max_length), _lib.ptr(out), _lib.ptr(lens),
                                        _lib.ptr(logits), ctypes.addressof(steps))
        _lib.check(rc, "mg_generate")
        self.last_steps = steps.value
        if trim:
            out = out[:, : _stop_column(out, lens, max_length)]
        if return_logits:
            return out, logits[:, : out.shape[1] - 1]
        return out

    def generate_host(self, input_ids, bbox, pixel_values, attention_mask=None, num_beams=1, max_length=512,
                      trim=True):
        """HOST tensors in (pinned recommended), HOST LongTensor out; H2D/D2H happen inside the C call."""
        ids, box, px, am, B, Lt = self._prep(input_ids, bbox, pixel_values, attention_mask, "cpu")
        out = torch.empty((B, max_length), dtype=torch.int64).pin_memory()
        lens = torch.empty((B,), dtype=torch.int32)
        steps = ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mg_generate_host(self._h, _lib.cur_stream(), B, Lt, _lib.ptr(ids), _lib.ptr(box),
                                             _lib.ptr(px), _lib.ptr(am), int(num_beams), int(max_length),
                                             _lib.ptr(out), _lib.ptr(lens), ctypes.addressof(steps))
        _lib.check(rc, "mg_generate_host")
        self.last_steps = steps.value
        if trim:
            out = out[:, : _stop_column(out, lens, max_length)]
        return out

    def prefetch_host(self, input_ids, bbox, pixel_values, attention_mask=None, max_length=512):
        """start the H2D copy of the NEXT batch (pinned HOST tensors, the very objects later passed to generate_host)
        on the library's copy stream; returns immediately so the copy overlaps the decode of the batch in flight"""
        ids, box, px, am, B, Lt = self._prep(input_ids, bbox, pixel_values, attention_mask, "cpu")
        for t, o in ((ids, input_ids), (box, bbox), (px, pixel_values)):
            if t.data_ptr() != o.data_ptr():
                raise _lib.MgError("prefetch_host needs contiguous int64 / float32 host tensors (no conversion copy)")
        with torch.cuda.device(self.device):
            rc = _lib.lib().mg_prefetch_host(self._h, B, Lt, _lib.ptr(ids), _lib.ptr(box), _lib.ptr(px), _lib.ptr(am),
                                             int(max_length))
        _lib.check(rc, "mg_prefetch_host")

    # ------------------------------------------------------------------ encoder run-ahead (a stream of batches)
    def encode_ahead(self, input_ids, bbox, pixel_values, attention_mask=None) -> bool:
        """queue the encoder of the NEXT batch on its own small SM partition and return at once; the generate /
        generate_dist call that later gets the very same device tensors (int64 / float32, contiguous, unchanged) takes
        the finished encoder memory.  False if SM partitioning is unavailable (nothing queued)."""
        ids, box, px, am, B, Lt = self._prep(input_ids, bbox, pixel_values, attention_mask, self.device)
        self._ahead_keep = (getattr(self, "_ahead_keep", ()) + ((ids, box, px, am),))[-2:]  # the library reads them later
        armed = ctypes.c_int32(0)
        L = _lib.lib()
        L.mg_encode_ahead.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5
        with torch.cuda.device(self.device):
            rc = L.mg_encode_ahead(self._h, _lib.cur_stream(), B, Lt, _lib.ptr(ids), _lib.ptr(box), _lib.ptr(px),
                                   _lib.ptr(am), ctypes.addressof(armed))
        _lib.check(rc, "mg_encode_ahead")
        return bool(armed.value)

    def encode_ahead_host(self, input_ids, bbox, pixel_values, attention_mask=None, max_length=512) -> bool:
        """prefetch_host + run-ahead encoder on the staged copy; consumed by the generate_host call that passes the
        same pinned HOST tensors"""
        ids, box, px, am, B, Lt = self._prep(input_ids, bbox, pixel_values, attention_mask, "cpu")
        for t, o in ((ids, input_ids), (box, bbox), (px, pixel_values)):
            if t.data_ptr() != o.data_ptr():
                raise _lib.MgError("encode_ahead_host needs contiguous int64 / float32 host tensors (no conversion copy)")
        armed = ctypes.c_int32(0)
        L = _lib.lib()
        L.mg_encode_ahead_host.argtypes = ([ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4 +
                                           [ctypes.c_int, ctypes.c_void_p])
        with torch.cuda.device(self.device):
            rc = L.mg_encode_ahead_host(self._h, B, Lt, _lib.ptr(ids), _lib.ptr(box), _lib.ptr(px), _lib.ptr(am),
                                        int(max_length), ctypes.addressof(armed))
        _lib.check(rc, "mg_encode_ahead_host")
        return bool(armed.value)

    def ahead_reset(self):
        """end of a stream of batches: wait for a run-ahead encoder in flight, drop batches nobody asked for"""
        L = _lib.lib()
        L.mg_ahead_reset.argtypes = [ctypes.c_void_p]
        with torch.cuda.device(self.device):
            _lib.check(L.mg_ahead_reset(self._h), "mg_ahead_reset")

    def last_ahead(self):
        """(encoder ms on its partition for the batch the last generate call took from a slot, SMs encoder, SMs decoder)"""
        ms, a, b = ctypes.c_float(0), ctypes.c_int32(0), ctypes.c_int32(0)
        L = _lib.lib()
        L.mg_last_ahead.argtypes = [ctypes.c_void_p] * 4
        _lib.check(L.mg_last_ahead(self._h, ctypes.addressof(ms), ctypes.addressof(a), ctypes.addressof(b)), "mg_last_ahead")
        return {"encoder_ms": ms.value, "sms_encoder": a.value, "sms_decoder": b.value}

    # ------------------------------------------------------------------ multi-GPU (one process per GPU)
    def comm_init_from_torch(self, group=None):
        """join an NCCL communicator owned by the library; the 128-byte id travels over torch.distributed"""
        import torch.distributed as dist

        L = _lib.lib()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _lib.check(L.mg_nccl_unique_id(buf), "mg_nccl_unique_id")
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        L.mg_comm_init.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
        with torch.cuda.device(self.device):
            _lib.check(L.mg_comm_init(self._h, world, rank, box[0]), "mg_comm_init")
        self.world, self.rank = world, rank

    def dist_mode(self) -> int:
        """0 single GPU, 1 ncclAllGather of the token ids per decode step, 2 NVLink peer stores fused into the selection kernel"""
        L = _lib.lib()
        L.mg_dist_mode.argtypes = [ctypes.c_void_p]
        return int(L.mg_dist_mode(self._h))

    def generate_dist(self, input_ids, bbox, pixel_values, attention_mask=None, max_length=512, num_beams=1):
        """this rank's shard in, ids of the WHOLE batch (world*B_local, max_length) out, on every rank"""
        ids, box, px, am, B, Lt = self._prep(input_ids, bbox, pixel_values, attention_mask, self.device)
        out = torch.empty((self.world * B, max_length), device=self.device, dtype=torch.int64)
        steps = ctypes.c_int32(0)
        L = _lib.lib()
        L.mg_generate_dist.argtypes = ([ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] +
                                       [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p])
        with torch.cuda.device(self.device):
            rc = L.mg_generate_dist(self._h, _lib.cur_stream(), B, Lt, _lib.ptr(ids), _lib.ptr(box), _lib.ptr(px),
                                    _lib.ptr(am), int(num_beams), int(max_length), _lib.ptr(out), ctypes.addressof(steps))
        _lib.check(rc, "mg_generate_dist")
        self.last_steps = steps.value
        return out

    def generate_sharded(self, input_ids, bbox, pixel_values, attention_mask=None, max_length=512, num_beams=1):
        """the WHOLE batch in (same tensors on every rank), ids of the whole batch (n, max_length) out on every rank:
        contiguous shards (parallel.shard_range), padded to equal size because the library's exchange needs equal
        shards, decoded through mg_generate_dist, padding rows dropped"""
        from . import parallel

        n = input_ids.shape[0]
        lo, hi = parallel.shard_range(n, self.world, self.rank)
        local = {"input_ids": input_ids[lo:hi], "bbox": bbox[lo:hi], "pixel_values": pixel_values[lo:hi],
                 "attention_mask": None if attention_mask is None else attention_mask[lo:hi]}
        local = parallel.pad_shard(local, parallel.shard_rows(n, self.world))
        out = self.generate_dist(**local, max_length=max_length, num_beams=num_beams)
        return parallel.unpad_gathered(out, n, self.world)

    def forward_logits(self, input_ids, bbox, pixel_values, decoder_input_ids, attention_mask=None):
        """teacher-forced logits (B, T, vocab) for decoder_input_ids (B, T) — `model(**batch).logits`"""
        ids, box, px, am, B, Lt = self._prep(input_ids, bbox, pixel_values, attention_mask, self.device)
        dec = decoder_input_ids.to(device=self.device, dtype=torch.int64).contiguous()
        T = dec.shape[1]
        logits = torch.empty((B, T, self.cfg.vocab_size), device=self.device, dtype=torch.float32)
        L = _lib.lib()
        L.mg_forward_logits.argtypes = ([ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] +
                                        [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p])
        with torch.cuda.device(self.device):
            rc = L.mg_forward_logits(self._h, _lib.cur_stream(), B, Lt, _lib.ptr(ids), _lib.ptr(box), _lib.ptr(px),
                                     _lib.ptr(am), _lib.ptr(dec), T, _lib.ptr(logits))
        _lib.check(rc, "mg_forward_logits")
        return logits

    def profile_cross_attn(self, reps: int = 3):
        L = _lib.lib()
        L.mg_profile_cross_attn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3
        ms, by, n = ctypes.c_float(0), ctypes.c_int64(0), ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.check(L.mg_profile_cross_attn(self._h, _lib.cur_stream(), reps, ctypes.addressof(ms),
                                               ctypes.addressof(by), ctypes.addressof(n)), "mg_profile_cross_attn")
        return {"ms_per_launch": ms.value, "bytes_per_launch": by.value, "launches": n.value}

    def last_decode_loop(self):
        ms, n, f = ctypes.c_float(0), ctypes.c_int32(0), ctypes.c_int32(0)
        _lib.check(_lib.lib().mg_last_decode_loop(self._h, ctypes.addressof(ms), ctypes.addressof(n),
                                                  ctypes.addressof(f)), "mg_last_decode_loop")
        p50, p99 = ctypes.c_float(0), ctypes.c_float(0)
        _lib.check(_lib.lib().mg_last_decode_latency(self._h, ctypes.addressof(p50), ctypes.addressof(p99)),
                   "mg_last_decode_latency")
        return {"loop_ms": ms.value, "steps": n.value, "fused": bool(f.value), "step_p50_ms": p50.value,
                "step_p99_ms": p99.value}

    def last_memory_len(self):
        """(encoder memory length, positions the decoder holds K/V for after dropping masked ones)"""
        a, b = ctypes.c_int32(0), ctypes.c_int32(0)
        L = _lib.lib()
        L.mg_last_memory_len.argtypes = [ctypes.c_void_p] * 3
        _lib.check(L.mg_last_memory_len(self._h, ctypes.addressof(a), ctypes.addressof(b)), "mg_last_memory_len")
        return a.value, b.value

    def launch_count(self) -> int:
        k = ctypes.c_int64(0)
        L = _lib.lib()
        L.mg_launch_count.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        _lib.check(L.mg_launch_count(self._h, ctypes.addressof(k)), "mg_launch_count")
        return k.value

    def last_stats(self):
        e, d, k = ctypes.c_float(0), ctypes.c_float(0), ctypes.c_int64(0)
        _lib.check(_lib.lib().mg_last_stats(self._h, ctypes.addressof(e), ctypes.addressof(d), ctypes.addressof(k)),
                   "mg_last_stats")
        return {"encode_ms": e.value, "decode_ms": d.value, "kernels": k.value}


def _stop_column(out: torch.Tensor, lens: torch.Tensor, max_length: int) -> int:
    """GenerationMixin stops once every row has emitted EOS: the returned width is the longest row."""
    return int(min(max_length, int(lens.max().item())))
