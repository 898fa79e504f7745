"""CPU: the drop-in surface the reference imports (begin.py:7-13) exists and behaves (no compute without a GPU)."""
import numpy as np
import pytest
import torch

from markushgrapher_b200.configuration import MarkushgrapherConfig, expected_weights


def small_cfg():
    return MarkushgrapherConfig(vocab_size=515, d_model=128, d_ff=256, num_layers=2, num_heads=2, image_size=64,
                                swin_image=96, swin_embed=32, swin_depths=(2, 2), swin_heads=(1, 2), proj_hidden=128)


def test_reference_import_line_resolves():
    import markushgrapher_b200.hf_shim  # noqa: F401
    from transformers.models.markushgrapher import (  # the exact names of reference begin.py:7-13
        MarkushgrapherConfig as C, MarkushgrapherForConditionalGeneration as M, MarkushgrapherImageProcessor as I,
        MarkushgrapherProcessor as P, MarkushgrapherTokenizer as T)

    assert all(x is not None for x in (C, M, I, P, T))


def test_model_object_contract(tmp_path):
    from markushgrapher_b200.modeling import MarkushgrapherForConditionalGeneration as M

    cfg = small_cfg()
    cfg.image_size = 64
    cfg.architecture_variant = "me-lf-stack-1"
    cfg.output_attentions = True            # reference begin.py:119-121 mutates these
    m = M(cfg)
    sd = m.state_dict()
    assert set(sd) == set(expected_weights(cfg))
    # sub-modules the reference names (utils_model_loading.py:23-41, begin.py:151,166)
    assert len(m.encoder.molscribe_encoder.state_dict()) > 0
    assert len(m.encoder.molscribe_projector.state_dict()) == 4
    assert len(m.decoder.state_dict()) > 0 and "weight" in m.lm_head.state_dict()
    # safe_load on a sub-module, save / from_pretrained round trip
    proj = {k: torch.randn_like(v) for k, v in m.encoder.molscribe_projector.state_dict().items()}
    assert m.safe_load(m.encoder.molscribe_projector, proj) == []
    m.save_pretrained(str(tmp_path))
    m2 = M.from_pretrained(str(tmp_path), config=MarkushgrapherConfig.from_pretrained(str(tmp_path)))
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k]), k
    assert m.init_molscribe_weights() is False     # checkpoint not on disk here -> warns, keeps weights
    assert m.device.type == "cpu"
    # the reference detects a DDP wrapper with hasattr(model, "module") (utils_evaluation.py:269) and then decodes
    # greedily; the bare model must NOT look wrapped or predict.yaml's beam search is silently downgraded
    assert not hasattr(m, "module")
    with pytest.raises(RuntimeError):              # no CPU fallback
        m.generate(input_ids=torch.zeros(1, 4, dtype=torch.long), bbox=torch.zeros(1, 4, 4),
                   pixel_values=torch.zeros(1, 3, 64, 64))


def test_tied_head_checkpoint_and_missing_weights(tmp_path):
    """a checkpoint saved with tie_word_embeddings=True has no lm_head.weight: the head aliases shared.weight;
    anything else missing is an error, not a warning"""
    from markushgrapher_b200.modeling import MarkushgrapherForConditionalGeneration as M

    cfg = small_cfg()
    m = M(cfg)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    del sd["lm_head.weight"]
    m.config.save_pretrained(str(tmp_path))
    torch.save(sd, str(tmp_path / "pytorch_model.bin"))
    m2 = M.from_pretrained(str(tmp_path))
    assert torch.equal(m2.state_dict()["lm_head.weight"], sd["shared.weight"])
    untied = small_cfg()
    untied.tie_word_embeddings = False
    with pytest.raises(KeyError):
        M.from_pretrained(str(tmp_path), config=untied)
    del sd["decoder.final_layer_norm.weight"]
    torch.save(sd, str(tmp_path / "pytorch_model.bin"))
    with pytest.raises(KeyError):
        M.from_pretrained(str(tmp_path))


def test_swin_timm_name_conversion():
    from markushgrapher_b200.modeling import swin_timm_to_hf

    C = 8
    sd = {"patch_embed.proj.weight": torch.zeros(C, 3, 4, 4), "patch_embed.norm.weight": torch.zeros(C),
          "layers.0.blocks.1.attn.qkv.weight": torch.arange(3 * C * C).float().view(3 * C, C),
          "layers.0.blocks.1.attn.qkv.bias": torch.zeros(3 * C), "layers.0.blocks.1.attn.proj.weight": torch.zeros(C, C),
          "layers.0.blocks.1.norm1.weight": torch.zeros(C), "layers.0.blocks.1.mlp.fc1.weight": torch.zeros(4 * C, C),
          "layers.0.blocks.1.mlp.fc2.bias": torch.zeros(C), "layers.0.blocks.1.attn.relative_position_bias_table": torch.zeros(529, 1),
          "layers.0.downsample.reduction.weight": torch.zeros(2 * C, 4 * C), "norm.bias": torch.zeros(8 * C),
          "layers.0.blocks.1.attn.relative_position_index": torch.zeros(4)}
    hf = swin_timm_to_hf(sd)
    assert hf["encoder.layers.0.blocks.1.attention.self.key.weight"].equal(sd["layers.0.blocks.1.attn.qkv.weight"][C:2 * C])
    for k in ("embeddings.patch_embeddings.projection.weight", "embeddings.norm.weight", "layernorm.bias",
              "encoder.layers.0.blocks.1.attention.output.dense.weight", "encoder.layers.0.blocks.1.layernorm_before.weight",
              "encoder.layers.0.blocks.1.intermediate.dense.weight", "encoder.layers.0.blocks.1.output.dense.bias",
              "encoder.layers.0.downsample.reduction.weight", "encoder.layers.0.blocks.1.attention.self.relative_position_bias_table"):
        assert k in hf, k


def test_image_processor_contract():
    from PIL import Image
    from markushgrapher_b200.processing import MarkushgrapherImageProcessor

    ip = MarkushgrapherImageProcessor(apply_ocr=False, size={"height": 512, "width": 512})
    img = Image.fromarray((np.random.RandomState(0).rand(300, 200, 3) * 255).astype(np.uint8))
    px = ip(img)["pixel_values"]
    assert px.shape == (1, 3, 512, 512) and px.dtype == torch.float32
    assert px.min() >= -1.0 and px.max() <= 1.0
