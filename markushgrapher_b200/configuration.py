"""MarkushgrapherConfig — host-side mirror of the fork's config class (reference begin.py:115-121 mutates
`image_size`, `architecture_variant`, `output_attentions` on it) plus the dims of the B200 path."""
from __future__ import annotations

import json
import os
from typing import Dict, Tuple


class MarkushgrapherConfig:
    model_type = "markushgrapher"

    def __init__(self, vocab_size=33201, d_model=1024, d_kv=64, d_ff=4096, num_layers=24, num_decoder_layers=None,
                 num_heads=16, relative_attention_num_buckets=32, relative_attention_max_distance=128,
                 max_2d_position_embeddings=1024, image_size=512, patch_size=16, layer_norm_epsilon=1e-6,
                 swin_image=384, swin_patch=4, swin_embed=128, swin_depths=(2, 2, 18, 2), swin_heads=(4, 8, 16, 32),
                 swin_window=12, swin_ln_eps=1e-5, proj_hidden=1024, architecture_variant="me-lf-stack-1",
                 decoder_start_token_id=0, eos_token_id=1, pad_token_id=0, tie_word_embeddings=True,
                 output_attentions=False, **kwargs):
        self.vocab_size = vocab_size
        self.d_model = d_model
        self.d_kv = d_kv
        self.d_ff = d_ff
        self.num_layers = num_layers
        self.num_decoder_layers = num_decoder_layers if num_decoder_layers is not None else num_layers
        self.num_heads = num_heads
        self.relative_attention_num_buckets = relative_attention_num_buckets
        self.relative_attention_max_distance = relative_attention_max_distance
        self.max_2d_position_embeddings = max_2d_position_embeddings
        self.image_size = image_size
        self.patch_size = patch_size
        self.layer_norm_epsilon = layer_norm_epsilon
        self.swin_image = swin_image
        self.swin_patch = swin_patch
        self.swin_embed = swin_embed
        self.swin_depths = tuple(swin_depths)
        self.swin_heads = tuple(swin_heads)
        self.swin_window = swin_window
        self.swin_ln_eps = swin_ln_eps
        self.proj_hidden = proj_hidden
        self.architecture_variant = architecture_variant
        self.decoder_start_token_id = decoder_start_token_id
        self.eos_token_id = eos_token_id
        self.pad_token_id = pad_token_id
        self.tie_word_embeddings = tie_word_embeddings
        self.output_attentions = output_attentions
        self.extra = dict(kwargs)

    # ---- aliases used by the engine / C config
    rel_buckets = property(lambda s: s.relative_attention_num_buckets)
    rel_max_distance = property(lambda s: s.relative_attention_max_distance)
    max_2d = property(lambda s: s.max_2d_position_embeddings)
    ln_eps = property(lambda s: s.layer_norm_epsilon)
    hidden_size = property(lambda s: s.d_model)

    @property
    def logit_scale(self) -> float:
        # UDOP scales the decoder output by d_model^-0.5 when the head is (nominally) tied
        # (transformers/models/udop/modeling_udop.py:1587-1588)
        return self.d_model ** -0.5 if self.tie_word_embeddings else 1.0

    @property
    def swin_dim(self) -> int:
        return self.swin_embed * 2 ** (len(self.swin_depths) - 1)

    @property
    def swin_tokens(self) -> int:
        g = self.swin_image // self.swin_patch // 2 ** (len(self.swin_depths) - 1)
        return g * g

    @property
    def n_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    # ---- (de)serialisation in the HF style
    def to_dict(self) -> dict:
        d = {k: v for k, v in self.__dict__.items() if k != "extra"}
        d["swin_depths"] = list(self.swin_depths)
        d["swin_heads"] = list(self.swin_heads)
        d["model_type"] = self.model_type
        return d

    def save_pretrained(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(self.to_dict(), f, indent=1)

    @classmethod
    def from_pretrained(cls, path: str, **kwargs) -> "MarkushgrapherConfig":
        fn = os.path.join(path, "config.json") if os.path.isdir(path) else path
        with open(fn) as f:
            d = json.load(f)
        d.pop("model_type", None)
        d.update(kwargs)
        return cls(**d)

    @classmethod
    def from_dims(cls, c) -> "MarkushgrapherConfig":
        """from any object with the short dim names (e.g. the oracle's MGConfig)"""
        return cls(vocab_size=c.vocab_size, d_model=c.d_model, d_kv=c.d_kv, d_ff=c.d_ff, num_layers=c.num_layers,
                   num_decoder_layers=c.num_decoder_layers, num_heads=c.num_heads,
                   relative_attention_num_buckets=c.rel_buckets, relative_attention_max_distance=c.rel_max_distance,
                   max_2d_position_embeddings=c.max_2d, image_size=c.image_size, patch_size=c.patch_size,
                   layer_norm_epsilon=c.ln_eps, swin_image=c.swin_image, swin_patch=c.swin_patch,
                   swin_embed=c.swin_embed, swin_depths=c.swin_depths, swin_heads=c.swin_heads,
                   swin_window=c.swin_window, swin_ln_eps=c.swin_ln_eps, proj_hidden=c.proj_hidden)


def expected_weights(cfg) -> Dict[str, Tuple[int, ...]]:
    """name -> shape of every parameter the path consumes (HF naming of stock UDOP / Swin modules)."""
    d, dff, H, V = cfg.d_model, cfg.d_ff, cfg.num_heads, cfg.vocab_size
    nb = cfg.rel_buckets
    w: Dict[str, Tuple[int, ...]] = {}
    w["shared.weight"] = (V, d)
    w["lm_head.weight"] = (V, d)
    w["encoder.embed_patches.proj.weight"] = (d, 3, cfg.patch_size, cfg.patch_size)
    w["encoder.embed_patches.proj.bias"] = (d,)
    w["encoder.cell_2d_embedding.x_position_embeddings.weight"] = (cfg.max_2d, d)
    w["encoder.cell_2d_embedding.y_position_embeddings.weight"] = (cfg.max_2d, d)
    for i in range(3):
        w[f"encoder.relative_bias.biases.{i}.relative_attention_bias.weight"] = (nb, H)
    for i in range(cfg.num_layers):
        p = f"encoder.block.{i}.layer."
        for n in "qkvo":
            w[p + f"0.SelfAttention.{n}.weight"] = (d, d)
        w[p + "0.layer_norm.weight"] = (d,)
        w[p + "1.DenseReluDense.wi.weight"] = (dff, d)
        w[p + "1.DenseReluDense.wo.weight"] = (d, dff)
        w[p + "1.layer_norm.weight"] = (d,)
    w["encoder.final_layer_norm.weight"] = (d,)
    for i in range(cfg.num_decoder_layers):
        p = f"decoder.block.{i}.layer."
        for n in "qkvo":
            w[p + f"0.SelfAttention.{n}.weight"] = (d, d)
            w[p + f"1.EncDecAttention.{n}.weight"] = (d, d)
        w[p + "0.layer_norm.weight"] = (d,)
        w[p + "1.layer_norm.weight"] = (d,)
        w[p + "2.DenseReluDense.wi.weight"] = (dff, d)
        w[p + "2.DenseReluDense.wo.weight"] = (d, dff)
        w[p + "2.layer_norm.weight"] = (d,)
    w["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"] = (nb, H)
    w["decoder.final_layer_norm.weight"] = (d,)
    sp = "encoder.molscribe_encoder."
    C = cfg.swin_embed
    w[sp + "embeddings.patch_embeddings.projection.weight"] = (C, 3, cfg.swin_patch, cfg.swin_patch)
    w[sp + "embeddings.patch_embeddings.projection.bias"] = (C,)
    w[sp + "embeddings.norm.weight"] = (C,)
    w[sp + "embeddings.norm.bias"] = (C,)
    ntab = (2 * cfg.swin_window - 1) ** 2
    ns = len(cfg.swin_depths)
    for s in range(ns):
        for b in range(cfg.swin_depths[s]):
            p = sp + f"encoder.layers.{s}.blocks.{b}."
            for ln in ("layernorm_before", "layernorm_after"):
                w[p + ln + ".weight"] = (C,)
                w[p + ln + ".bias"] = (C,)
            for n in ("query", "key", "value"):
                w[p + f"attention.self.{n}.weight"] = (C, C)
                w[p + f"attention.self.{n}.bias"] = (C,)
            w[p + "attention.self.relative_position_bias_table"] = (ntab, cfg.swin_heads[s])
            w[p + "attention.output.dense.weight"] = (C, C)
            w[p + "attention.output.dense.bias"] = (C,)
            w[p + "intermediate.dense.weight"] = (4 * C, C)
            w[p + "intermediate.dense.bias"] = (4 * C,)
            w[p + "output.dense.weight"] = (C, 4 * C)
            w[p + "output.dense.bias"] = (C,)
        if s + 1 < ns:
            p = sp + f"encoder.layers.{s}.downsample."
            w[p + "reduction.weight"] = (2 * C, 4 * C)
            w[p + "norm.weight"] = (4 * C,)
            w[p + "norm.bias"] = (4 * C,)
            C *= 2
    w[sp + "layernorm.weight"] = (C,)
    w[sp + "layernorm.bias"] = (C,)
    w["encoder.molscribe_projector.0.weight"] = (cfg.proj_hidden, C)
    w["encoder.molscribe_projector.0.bias"] = (cfg.proj_hidden,)
    w["encoder.molscribe_projector.2.weight"] = (d, cfg.proj_hidden)
    w["encoder.molscribe_projector.2.bias"] = (d,)
    return w


def random_state(cfg, seed: int = 0, device="cuda"):
    """Seeded random-init weights of the right shapes with O(1) activations (bench / smoke use; there are no
    checkpoints in this environment). Same recipe family as the oracle's init but generated on `device`."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for name, shp in expected_weights(cfg).items():
        def rn(std):
            return torch.randn(shp, generator=g, device=device, dtype=torch.float32) * std
        leaf = name.split(".")[-2] if name.count(".") else name
        if name in ("shared.weight", "lm_head.weight"):
            t = rn(1.0)
        elif "relative_attention_bias" in name or "relative_position_bias_table" in name:
            t = rn(0.5)
        elif "position_embeddings" in name:
            t = rn(0.3)
        elif "norm" in name:
            t = (1.0 if name.endswith("weight") else 0.0) + rn(0.1)
        elif len(shp) >= 2:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            std = fan_in ** -0.5
            if leaf == "q" and "Attention" in name:
                std = (fan_in * cfg.d_kv) ** -0.5 * 1.5
            if leaf in ("wo", "o") or (leaf == "dense" and ".output." in name):
                std *= 0.5
            t = rn(std)
        else:
            t = rn(0.1)
        out[name] = t
    return out
