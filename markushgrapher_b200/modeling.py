"""Host-side mirror of the fork's `transformers.models.markushgrapher` model class, as exercised by
markushgrapher.core. All arithmetic happens in libmg_b200.so through MGEngine; this file only owns
the parameters (so `.state_dict()`, `.to()`, `safe_load`, sub-module names work) and the call signatures.

Contract sources: reference markushgrapher/core/common/begin.py:115-172 (construction, init_molscribe_weights,
safe_load, sub-module names), markushgrapher/utils/ocsr/utils_evaluation.py:269-285 (generate),
markushgrapher/core/trainers/curriculumTrainer.py:647-657 (forward -> .logits),
markushgrapher/utils/model/utils_model_loading.py:20-41 (encoder.molscribe_encoder / .molscribe_projector /
decoder / lm_head).
"""
from __future__ import annotations

import os
import warnings
from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn as nn

from .configuration import MarkushgrapherConfig, expected_weights
from .engine import MGEngine


class _Node(nn.Module):
    """a bare container module; parameters are attached by dotted state_dict name"""


def _attach(root: nn.Module, dotted: str, param: nn.Parameter) -> None:
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    mod.register_parameter(parts[-1], param)


def swin_timm_to_hf(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """timm-0.4.12 `SwinTransformer` names (MolScribe's encoder checkpoint) -> the stock HF Swin names this repo
    uses; fused qkv is split into query/key/value."""
    out = {}
    for k, v in sd.items():
        for pre in ("module.", "encoder.", "transformer."):
            if k.startswith(pre):
                k = k[len(pre):]
        if k.startswith("patch_embed.proj."):
            out["embeddings.patch_embeddings.projection." + k.split(".")[-1]] = v
        elif k.startswith("patch_embed.norm."):
            out["embeddings.norm." + k.split(".")[-1]] = v
        elif k.startswith("norm."):
            out["layernorm." + k.split(".")[-1]] = v
        elif k.startswith("layers."):
            p = k.split(".")
            s = p[1]
            if p[2] == "downsample":
                out[f"encoder.layers.{s}.downsample." + ".".join(p[3:])] = v
                continue
            b, rest = p[3], ".".join(p[4:])
            base = f"encoder.layers.{s}.blocks.{b}."
            if rest.startswith("attn.qkv."):
                q, kk, vv = v.chunk(3, dim=0)
                leaf = rest.split(".")[-1]
                out[base + "attention.self.query." + leaf] = q
                out[base + "attention.self.key." + leaf] = kk
                out[base + "attention.self.value." + leaf] = vv
            elif rest == "attn.relative_position_bias_table":
                out[base + "attention.self.relative_position_bias_table"] = v
            elif rest.startswith("attn.proj."):
                out[base + "attention.output.dense." + rest.split(".")[-1]] = v
            elif rest.startswith("norm1."):
                out[base + "layernorm_before." + rest.split(".")[-1]] = v
            elif rest.startswith("norm2."):
                out[base + "layernorm_after." + rest.split(".")[-1]] = v
            elif rest.startswith("mlp.fc1."):
                out[base + "intermediate.dense." + rest.split(".")[-1]] = v
            elif rest.startswith("mlp.fc2."):
                out[base + "output.dense." + rest.split(".")[-1]] = v
    return out


class MarkushgrapherForConditionalGeneration(nn.Module):
    config_class = MarkushgrapherConfig

    def __init__(self, config: MarkushgrapherConfig):
        super().__init__()
        self.config = config
        for name, shp in expected_weights(config).items():
            t = torch.empty(shp, dtype=torch.float32)
            if name.endswith("norm.weight") or name.endswith("layernorm.weight"):
                t.fill_(1.0)
            elif t.dim() >= 2:
                nn.init.normal_(t, std=0.02)
            else:
                t.zero_()
            _attach(self, name, nn.Parameter(t, requires_grad=False))
        self._engine: Optional[MGEngine] = None
        self._engine_key = None
        self.precision = 0
        self.eval()

    # ------------------------------------------------------------------ HF-style plumbing
    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    # NB: no `.module` attribute here. The reference tests `hasattr(model, "module")` (utils_evaluation.py:269) to detect
    # a DDP wrapper and then decodes greedily; on the bare model it must take the predict.yaml branch
    # (`beam_search: True` -> generate(num_beams=5), :278-285). A DDP wrapper supplies its own `.module`.

    @classmethod
    def from_pretrained(cls, path: str, config: Optional[MarkushgrapherConfig] = None, **kw):
        config = config or MarkushgrapherConfig.from_pretrained(path)
        model = cls(config)
        sd = None
        for fn in ("model.safetensors", "pytorch_model.bin", "model.pt"):
            f = os.path.join(path, fn)
            if os.path.exists(f):
                if fn.endswith(".safetensors"):
                    from safetensors.torch import load_file

                    sd = load_file(f)
                else:
                    sd = torch.load(f, map_location="cpu")
                break
        if sd is None:
            raise FileNotFoundError(f"no weights (model.safetensors / pytorch_model.bin) under {path}")
        missing = model.safe_load(model, sd)
        if missing:  # (a tied head absent from the file is aliased to shared.weight inside safe_load)
            raise KeyError(f"{len(missing)} parameters of the path are missing from the checkpoint under {path} "
                           f"(or have another shape), e.g. {missing[:5]}")
        return model

    def save_pretrained(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        self.config.save_pretrained(path)
        torch.save({k: v.detach().cpu() for k, v in self.state_dict().items()}, os.path.join(path, "pytorch_model.bin"))

    def safe_load(self, module: nn.Module, state_dict: Dict[str, torch.Tensor]):
        """load what matches by name and shape (tolerates `module.` / `udop.` prefixes); returns missing names"""
        own = module.state_dict()
        clean = {}
        for k, v in state_dict.items():
            for pre in ("module.", "udop.", "model."):
                if k.startswith(pre):
                    k = k[len(pre):]
            clean[k] = v
        aliases = {"encoder.embed_tokens.weight": "shared.weight", "decoder.embed_tokens.weight": "shared.weight",
                   "patch_embed.proj.weight": "encoder.embed_patches.proj.weight",
                   "patch_embed.proj.bias": "encoder.embed_patches.proj.bias"}
        for a, b in aliases.items():
            if a in clean and b not in clean:
                clean[b] = clean[a]
        # tied head: HF checkpoints saved with tie_word_embeddings=True carry no lm_head.weight (safetensors drops the
        # duplicate), the head IS the embedding matrix (TF/models/udop/modeling_udop.py:1587-1590)
        if "lm_head.weight" in own and "lm_head.weight" not in clean and "shared.weight" in clean:
            if getattr(self.config, "tie_word_embeddings", True):
                clean["lm_head.weight"] = clean["shared.weight"]
        missing = []
        with torch.no_grad():
            for k, p in own.items():
                if k in clean and tuple(clean[k].shape) == tuple(p.shape):
                    p.copy_(clean[k].to(p.dtype))
                else:
                    missing.append(k)
        self._engine_key = None  # weights changed: rebuild the device copy lazily
        return missing

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self._engine_key = None
        return r

    def init_molscribe_weights(self, path: Optional[str] = None) -> bool:
        """reference begin.py:137-138: load the pretrained MolScribe Swin encoder if its checkpoint is on disk"""
        path = path or os.path.join("external", "MolScribe", "ckpts", "swin_base_char_aux_1m680k.pth")
        if not os.path.exists(path):
            warnings.warn(f"MolScribe checkpoint {path} not found; OCSR encoder keeps its current weights")
            return False
        ck = torch.load(path, map_location="cpu")
        sd = ck.get("encoder", ck) if isinstance(ck, dict) else ck
        self.safe_load(self.encoder.molscribe_encoder, swin_timm_to_hf(sd))
        return True

    # ------------------------------------------------------------------ the hot path
    def _get_engine(self) -> MGEngine:
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("markushgrapher_b200 runs on CUDA (sm_100a) only: call model.to('cuda') first; "
                               "there is no CPU fallback")
        key = (str(dev), self.precision)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            state = {k: v for k, v in self.state_dict().items()}
            self._engine = MGEngine(self.config, state, precision=self.precision, device=dev)
            self._engine_key = key
        return self._engine

    @property
    def engine(self) -> MGEngine:
        """the device engine behind this model (built on first use): `encode_ahead` / `generate_host` / ... for callers that
        want more than the reference's `generate` contract"""
        return self._get_engine()

    @torch.no_grad()
    def generate(self, input_ids=None, bbox=None, pixel_values=None, attention_mask=None, labels=None,
                 num_beams: int = 1, max_length: int = 512, **kwargs) -> torch.Tensor:
        """-> LongTensor (B, T<=max_length); `labels` and unknown generation kwargs are tolerated and ignored,
        like the reference call site relies on (utils_evaluation.py:279-285)"""
        eng = self._get_engine()
        if pixel_values.dim() == 3:
            pixel_values = pixel_values[None]
        return eng.generate(input_ids, bbox, pixel_values, attention_mask, num_beams=num_beams,
                            max_length=max_length).to(input_ids.device)

    @torch.no_grad()
    def forward(self, input_ids=None, bbox=None, pixel_values=None, attention_mask=None, labels=None,
                decoder_input_ids=None, decoder_attention_mask=None, image=None, **kwargs):
        """teacher-forced pass: returns an object with `.logits` (B, T, vocab) and `.loss` when labels are given
        (decoder inputs = shift_right(labels), start id 0, -100 -> pad; TF/models/udop/modeling_udop.py:309-329)"""
        if decoder_input_ids is None:
            if labels is None:
                raise ValueError("either decoder_input_ids or labels is required")
            decoder_input_ids = labels.new_zeros(labels.shape)
            decoder_input_ids[:, 1:] = labels[:, :-1]
            decoder_input_ids[:, 0] = self.config.decoder_start_token_id
            decoder_input_ids.masked_fill_(decoder_input_ids == -100, self.config.pad_token_id)
        eng = self._get_engine()
        logits = eng.forward_logits(input_ids, bbox, pixel_values, decoder_input_ids, attention_mask)
        loss = None
        if labels is not None:
            loss = nn.functional.cross_entropy(logits.view(-1, logits.shape[-1]), labels.to(logits.device).view(-1),
                                               ignore_index=-100)
        return SimpleNamespace(logits=logits, loss=loss)
