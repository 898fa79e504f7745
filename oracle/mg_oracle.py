"""CPU oracle for the MarkushGrapher-2 forward+generate hot path.  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
module, and only as the checker / the timed CPU baseline. The product path (markushgrapher_b200/) never does.

PARITY UNPINNED.  The reference's model arithmetic lives in two un-vendored, un-pinned forks that are absent
from /root/reference (lucas-morin/transformers `models/markushgrapher`, lucas-morin/MolScribe; cloned at
install time, reference setup.sh:37-44) and the reference ships no tests, fixtures or golden vectors
(SURVEY.md §4, §8c). This oracle therefore *composes the stock transformers==5.5.0 modules of the same
lineage* — it does not re-derive their math — and the golden vectors under tests/golden/ are generated from
it by oracle/make_golden.py:

  * VTL encoder / decoder : transformers.models.udop.modeling_udop.UdopForConditionalGeneration
                            (UdopStack :1025-1256, UdopBlock :692-781, UdopAttention :431-622,
                            combine_image_text_embeddings :133-214, UdopCellEmbeddings :784-807,
                            RelativePositionBias* :816-993) — UDOP-large is the checkpoint MarkushGrapher
                            was initialised from (reference README.md:298).
  * OCSR encoder          : transformers.models.swin.modeling_swin.SwinModel with the MolScribe/timm
                            `swin_base_patch4_window12_384` geometry (reference requirements.txt:25,
                            setup.sh:79-84; sub-module name reference utils_model_loading.py:23).
  * projector             : Linear -> GELU -> Linear (2.10 M params = 831 M - UDOP-large - Swin-B,
                            reference README.md:217; name reference begin.py:151).
  * fusion (me-lf-stack-1): encoder memory = cat[e1, e2], mask = cat[1, mask] (reference README.md:210-215).
  * generate()            : greedy loop below restates transformers/generation/utils.py::_sample
                            (:2658-2842: fp32 logits of the last position, first-max argmax, finished rows
                            emit pad, stop when every row has emitted EOS or at max_length including the start
                            token); `hf_generate` calls the stock GenerationMixin itself (greedy and beam) and
                            tests/test_oracle.py checks the two agree.
Call-site contract followed: reference markushgrapher/utils/ocsr/utils_evaluation.py:269-285
(generate(input_ids, bbox, pixel_values, labels, num_beams, max_length)) and
reference markushgrapher/core/trainers/curriculumTrainer.py:647-657 (model(**batch).logits).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from transformers import SwinConfig, SwinModel, UdopConfig, UdopForConditionalGeneration
from transformers.models.udop.modeling_udop import BaseModelOutputWithAttentionMask


@dataclasses.dataclass
class MGConfig:
    """Dimensions of the path. `full()` = MarkushGrapher-2 (UDOP-large + Swin-B + MLP)."""

    vocab_size: int = 33201
    d_model: int = 1024
    d_kv: int = 64
    d_ff: int = 4096
    num_layers: int = 24
    num_decoder_layers: int = 24
    num_heads: int = 16
    rel_buckets: int = 32
    rel_max_distance: int = 128
    max_2d: int = 1024
    image_size: int = 512
    patch_size: int = 16
    ln_eps: float = 1e-6
    # OCSR (Swin) branch
    swin_image: int = 384
    swin_patch: int = 4
    swin_embed: int = 128
    swin_depths: tuple = (2, 2, 18, 2)
    swin_heads: tuple = (4, 8, 16, 32)
    swin_window: int = 12
    swin_ln_eps: float = 1e-5
    proj_hidden: int = 1024

    @staticmethod
    def full() -> "MGConfig":
        return MGConfig()

    @staticmethod
    def small() -> "MGConfig":
        """mid-size: every kernel path exercised (3 Swin stages, shifted windows, 4 heads), seconds on CPU"""
        return MGConfig(vocab_size=2051, d_model=256, d_ff=512, num_layers=3, num_decoder_layers=3, num_heads=4,
                        image_size=128, swin_image=192, swin_embed=32, swin_depths=(2, 2, 2),
                        swin_heads=(1, 2, 4), proj_hidden=256)

    @staticmethod
    def tiny() -> "MGConfig":
        return MGConfig(vocab_size=515, d_model=128, d_ff=256, num_layers=2, num_decoder_layers=2, num_heads=2,
                        image_size=64, swin_image=96, swin_embed=32, swin_depths=(2, 2), swin_heads=(1, 2),
                        proj_hidden=128)

    @property
    def n_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def swin_dim(self) -> int:
        return self.swin_embed * 2 ** (len(self.swin_depths) - 1)

    @property
    def swin_tokens(self) -> int:
        g = self.swin_image // self.swin_patch // 2 ** (len(self.swin_depths) - 1)
        return g * g

    def udop_config(self) -> UdopConfig:
        return UdopConfig(vocab_size=self.vocab_size, d_model=self.d_model, d_kv=self.d_kv, d_ff=self.d_ff,
                          num_layers=self.num_layers, num_decoder_layers=self.num_decoder_layers,
                          num_heads=self.num_heads, relative_attention_num_buckets=self.rel_buckets,
                          relative_attention_max_distance=self.rel_max_distance, dropout_rate=0.0,
                          layer_norm_epsilon=self.ln_eps, feed_forward_proj="relu",
                          max_2d_position_embeddings=self.max_2d, image_size=self.image_size,
                          patch_size=self.patch_size, num_channels=3, decoder_start_token_id=0)

    def swin_config(self) -> SwinConfig:
        return SwinConfig(image_size=self.swin_image, patch_size=self.swin_patch, num_channels=3,
                          embed_dim=self.swin_embed, depths=list(self.swin_depths), num_heads=list(self.swin_heads),
                          window_size=self.swin_window, mlp_ratio=4.0, qkv_bias=True, hidden_dropout_prob=0.0,
                          attention_probs_dropout_prob=0.0, drop_path_rate=0.0, hidden_act="gelu",
                          use_absolute_embeddings=False, layer_norm_eps=self.swin_ln_eps)


class MGOracle(nn.Module):
    """fp32 CPU restatement. Sub-module names follow the reference contract:
    .encoder.molscribe_encoder, .encoder.molscribe_projector, .decoder, .lm_head."""

    def __init__(self, cfg: MGConfig):
        super().__init__()
        self.cfg = cfg
        self.udop = UdopForConditionalGeneration(cfg.udop_config())
        self.udop.generation_config.decoder_start_token_id = 0
        self.udop.generation_config.eos_token_id = 1
        self.udop.generation_config.pad_token_id = 0
        # separate (untied) LM head: the reference saves `model.lm_head` as its own module
        # (reference utils_model_loading.py:41); the d_model^-0.5 logit scale of the tied UDOP head is kept.
        self.udop.lm_head.weight = nn.Parameter(self.udop.lm_head.weight.detach().clone())
        self.encoder = self.udop.encoder
        self.decoder = self.udop.decoder
        self.lm_head = self.udop.lm_head
        self.encoder.molscribe_encoder = SwinModel(cfg.swin_config(), add_pooling_layer=False)
        self.encoder.molscribe_projector = nn.Sequential(
            nn.Linear(cfg.swin_dim, cfg.proj_hidden), nn.GELU(), nn.Linear(cfg.proj_hidden, cfg.d_model))
        self.eval()

    # ------------------------------------------------------------------ weights
    @torch.no_grad()
    def init_nondegenerate(self, seed: int = 0) -> "MGOracle":
        """Seeded random init with O(1) activations everywhere, non-trivial norm weights / bias tables and an
        untied LM head, so greedy decodes are diverse (stock init emits 0,0,0,... forever: SURVEY.md §7 hard
        part 3). Same bytes are handed to the CUDA library by the tests."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        cfg = self.cfg

        def rn(shape, std):
            return torch.randn(shape, generator=g) * std

        seen = set()
        for name, p in self.named_parameters():
            if id(p) in seen:
                continue
            seen.add(id(p))
            shp = tuple(p.shape)
            leaf = name.split(".")[-2] if name.count(".") else name
            if name.endswith("shared.weight") or "embed_tokens" in name:
                p.copy_(rn(shp, 1.0))
            elif name.endswith("lm_head.weight"):
                p.copy_(rn(shp, 1.0))
            elif "relative_attention_bias" in name:
                p.copy_(rn(shp, 0.5))
            elif "relative_position_bias_table" in name:
                p.copy_(rn(shp, 0.5))
            elif "position_embeddings" in name:  # cell 2d x / y tables
                p.copy_(rn(shp, 0.3))
            elif name.endswith("layer_norm.weight") or name.endswith("final_layer_norm.weight"):
                p.copy_(1.0 + rn(shp, 0.1))
            elif "layernorm" in name or ".norm." in name or name.endswith("norm.weight") or name.endswith("norm.bias"):
                p.copy_((1.0 if name.endswith("weight") else 0.0) + rn(shp, 0.1))
            elif p.dim() >= 2:
                fan_in = p[0].numel()
                std = fan_in ** -0.5
                if leaf == "q" and "Attention" in name:  # T5 attention has no 1/sqrt(d) scale
                    std = (fan_in * cfg.d_kv) ** -0.5 * 1.5
                if leaf in ("wo", "o") or leaf == "dense" and ".output." in name:
                    std *= 0.5  # keep the residual stream from blowing up over 24 layers
                p.copy_(rn(shp, std))
            else:  # biases
                p.copy_(rn(shp, 0.1))
        return self

    def export_state(self) -> dict:
        """name -> fp32 tensor, de-duplicated HF names (what mg_load_weight is fed)."""
        out = {}
        for k, v in self.state_dict().items():
            if k.startswith("udop."):
                k = k[len("udop."):]
            elif k.startswith("encoder.") or k.startswith("decoder.") or k.startswith("lm_head."):
                continue  # aliases of udop.*
            if "relative_position_index" in k:
                continue
            out[k] = v.detach().to(torch.float32).contiguous()
        return out

    # ------------------------------------------------------------------ forward pieces
    @torch.no_grad()
    def ocsr_pixels(self, pixel_values: torch.Tensor) -> torch.Tensor:
        """VTL pixels (B,3,512,512) -> Swin input (B,3,384,384): bilinear, align_corners=False, no antialias.
        (inferred: MolScribe's encoder is fixed at 384^2 and window 12 does not divide 512/4)"""
        s = self.cfg.swin_image
        if pixel_values.shape[-1] == s and pixel_values.shape[-2] == s:
            return pixel_values
        return F.interpolate(pixel_values, size=(s, s), mode="bilinear", align_corners=False, antialias=False)

    @torch.no_grad()
    def encode(self, input_ids, bbox, pixel_values, attention_mask=None, return_parts=False):
        """-> (memory (B, M, d), mask (B, M)),  M = swin_tokens + Lt + n_patches"""
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        enc = self.encoder(input_ids=input_ids, bbox=bbox.to(torch.float32), pixel_values=pixel_values,
                           attention_mask=attention_mask, return_dict=True)
        e2, m2 = enc.last_hidden_state, enc.attention_mask
        sw = self.encoder.molscribe_encoder(pixel_values=self.ocsr_pixels(pixel_values)).last_hidden_state
        e1 = self.encoder.molscribe_projector(sw)
        mem = torch.cat([e1, e2], dim=1)
        mask = torch.cat([torch.ones(e1.shape[:2], dtype=m2.dtype), m2], dim=1)
        if return_parts:
            return mem, mask, {"swin": sw, "e1": e1, "e2": e2, "mask2": m2}
        return mem, mask

    @torch.no_grad()
    def logits_teacher_forced(self, memory, mask, decoder_input_ids):
        out = self.decoder(input_ids=decoder_input_ids, encoder_hidden_states=memory, encoder_attention_mask=mask,
                           use_cache=False, return_dict=True)
        h = out.last_hidden_state * (self.cfg.d_model ** -0.5)
        return self.lm_head(h)

    @torch.no_grad()
    def forward_logits(self, input_ids, bbox, pixel_values, labels, attention_mask=None):
        """model(**batch).logits of the reference (curriculumTrainer.py:655): decoder inputs = shift_right(labels)"""
        dec_in = self.udop._shift_right(labels)
        mem, mask = self.encode(input_ids, bbox, pixel_values, attention_mask)
        return self.logits_teacher_forced(mem, mask, dec_in)

    @torch.no_grad()
    def generate_greedy(self, input_ids, bbox, pixel_values, attention_mask=None, max_length=512,
                        return_logits=False, memory=None, mask=None):
        """greedy restatement of GenerationMixin._sample with a KV cache. Returns LongTensor (B, T<=max_length),
        column 0 = decoder start id 0."""
        if memory is None:
            memory, mask = self.encode(input_ids, bbox, pixel_values, attention_mask)
        B = memory.shape[0]
        ids = torch.zeros((B, 1), dtype=torch.long)
        unfinished = torch.ones(B, dtype=torch.long)
        past = None
        all_logits = []
        cur = ids
        while ids.shape[1] < max_length:
            out = self.decoder(input_ids=cur, encoder_hidden_states=memory, encoder_attention_mask=mask,
                               past_key_values=past, use_cache=True, return_dict=True)
            past = out.past_key_values
            h = out.last_hidden_state[:, -1, :] * (self.cfg.d_model ** -0.5)
            logits = self.lm_head(h).float()
            if return_logits:
                all_logits.append(logits)
            nxt = torch.argmax(logits, dim=-1)
            nxt = nxt * unfinished + 0 * (1 - unfinished)  # finished rows emit pad (id 0)
            ids = torch.cat([ids, nxt[:, None]], dim=1)
            unfinished = unfinished & (nxt != 1).long()
            cur = nxt[:, None]
            if unfinished.max() == 0:
                break
        if return_logits:
            return ids, torch.stack(all_logits, dim=1)
        return ids

    @torch.no_grad()
    def hf_generate(self, input_ids, bbox, pixel_values, attention_mask=None, max_length=512, num_beams=1,
                    memory=None, mask=None):
        """the stock GenerationMixin loop itself (greedy or beam) on this oracle's encoder memory"""
        if memory is None:
            memory, mask = self.encode(input_ids, bbox, pixel_values, attention_mask)
        enc = BaseModelOutputWithAttentionMask(last_hidden_state=memory, attention_mask=mask)
        return self.udop.generate(encoder_outputs=enc, max_length=max_length, num_beams=num_beams, do_sample=False,
                                  early_stopping=False, length_penalty=1.0)


# ---------------------------------------------------------------------------------------- synthetic inputs
def make_inputs(cfg: MGConfig, batch: int, text_len: int, seed: int = 1234, ragged: bool = False,
                sep_box: float = 1.0):
    """Seeded synthetic batch of the shape SURVEY.md §8d specifies.
    pixel_values: white canvas with black line segments and filled rectangles, normalised to [-1, 1];
    input_ids: 14-id prompt prefix + sep(1) + OCR ids + sep(1); bbox rows in [0,1] (prefix rows 0, sep rows sep_box).
    ragged=True: per-row text lengths in [text_len/4, text_len], padded with id 0 / mask 0 / zero boxes."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    H = cfg.image_size
    px = torch.ones(batch, 3, H, H)
    for b in range(batch):
        for _ in range(40):
            x0, y0 = [int(v) for v in torch.randint(0, H - 2, (2,), generator=g)]
            ln = int(torch.randint(4, max(5, H // 4), (1,), generator=g))
            if torch.rand(1, generator=g).item() < 0.5:
                px[b, :, y0:y0 + 2, x0:min(H, x0 + ln)] = 0.0
            else:
                px[b, :, y0:min(H, y0 + ln), x0:x0 + 2] = 0.0
        for _ in range(12):
            x0, y0 = [int(v) for v in torch.randint(0, H - 8, (2,), generator=g)]
            w = int(torch.randint(3, max(4, H // 12), (1,), generator=g))
            h = int(torch.randint(2, max(3, H // 32), (1,), generator=g))
            px[b, :, y0:y0 + h, x0:x0 + w] = torch.rand(1, generator=g).item() * 0.5
    px = (px - 0.5) / 0.5
    n_prefix = min(14, max(1, text_len // 4))
    ids = torch.zeros(batch, text_len, dtype=torch.long)
    box = torch.zeros(batch, text_len, 4)
    mask = torch.zeros(batch, text_len, dtype=torch.long)
    for b in range(batch):
        L = text_len
        if ragged:
            L = int(torch.randint(max(n_prefix + 3, text_len // 4), text_len + 1, (1,), generator=g))
        hi_id = min(32000, cfg.vocab_size)
        ids[b, :n_prefix] = torch.randint(3, hi_id, (n_prefix,), generator=g)
        ids[b, n_prefix] = 1
        box[b, n_prefix] = sep_box
        n_ocr = L - n_prefix - 2
        ids[b, n_prefix + 1:n_prefix + 1 + n_ocr] = torch.randint(3, hi_id, (n_ocr,), generator=g)
        x0 = torch.rand(n_ocr, generator=g) * 0.88 + 0.02
        y0 = torch.rand(n_ocr, generator=g) * 0.88 + 0.02
        w = torch.rand(n_ocr, generator=g) * 0.07 + 0.01
        h = torch.rand(n_ocr, generator=g) * 0.02 + 0.01
        box[b, n_prefix + 1:n_prefix + 1 + n_ocr] = torch.stack(
            [x0, y0, (x0 + w).clamp(max=1.0), (y0 + h).clamp(max=1.0)], dim=-1)
        ids[b, L - 1] = 1
        box[b, L - 1] = sep_box
        mask[b, :L] = 1
    return {"input_ids": ids, "bbox": box, "pixel_values": px, "attention_mask": mask}


def build(cfg: MGConfig, seed: int = 0) -> MGOracle:
    torch.manual_seed(seed)
    return MGOracle(cfg).init_nondegenerate(seed)
