"""GPU: the non-bench BASELINE.json configs as parity / property cases at the full model size.
configs[2]: batch-256 USPTO-shape (ragged OCR text 32..256 tokens), VTL encoder on.
configs[4]: IP5-M-shape (text up to 512 tokens), beam=4, <=768 decoder tokens."""
import numpy as np
import pytest
import torch

from markushgrapher_b200.configuration import MarkushgrapherConfig, random_state
from markushgrapher_b200.engine import MGEngine
from oracle import mg_oracle as O

pytestmark = pytest.mark.gpu


def host_expected_mask(cfg, input_ids, bbox, attention_mask):
    """numpy restatement of which encoder-memory positions are valid (integer work: must match bit-exactly):
    [swin tokens: 1] + [text mask] + [surviving patches: 1 ... then 0 padding] (modeling_udop.py:149-214)"""
    B, Lt = input_ids.shape
    n = cfg.image_size // cfg.patch_size
    out = np.zeros((B, cfg.swin_tokens + Lt + n * n), dtype=np.int32)
    bb = bbox.numpy().astype(np.float32)
    for b in range(B):
        fx = np.floor((bb[b, :, 0] + bb[b, :, 2]) / np.float32(2.0) * np.float32(n)).astype(np.int64).clip(0, n - 1)
        fy = np.floor((bb[b, :, 1] + bb[b, :, 3]) / np.float32(2.0) * np.float32(n)).astype(np.int64).clip(0, n - 1)
        removed = np.unique(fx + fy * n)
        out[b, : cfg.swin_tokens] = 1
        out[b, cfg.swin_tokens: cfg.swin_tokens + Lt] = attention_mask[b].numpy()
        out[b, cfg.swin_tokens + Lt: cfg.swin_tokens + Lt + (n * n - len(removed))] = 1
    return out


@pytest.fixture(scope="module")
def full_engine():
    cfg = MarkushgrapherConfig()
    dev = torch.device("cuda", 0)
    eng = MGEngine(cfg, random_state(cfg, 0, dev), device=dev)
    yield cfg, eng
    eng.close()


def test_config3_batch256_uspto_shape_encoder(full_engine):
    cfg, eng = full_engine
    ocfg = O.MGConfig.full()
    inp = O.make_inputs(ocfg, 256, 256, seed=1237, ragged=True)  # per-image text length in [64, 256]
    mem, mask = eng.encode(**inp)
    assert mem.shape == (256, 144 + 256 + 1024, 1024)
    assert torch.isfinite(mem).all()
    assert np.array_equal(mask.cpu().numpy(), host_expected_mask(cfg, inp["input_ids"], inp["bbox"], inp["attention_mask"]))
    # images are independent: the same image encoded alone gives the same hidden states (batch/chunk invariance)
    for i in (0, 77, 255):
        one = {k: v[i:i + 1] for k, v in inp.items()}
        m1, k1 = eng.encode(**one)
        valid = k1[0].bool()
        err = ((mem[i][valid] - m1[0][valid]).double().norm() / m1[0][valid].double().norm()).item()
        assert err < 1e-4, (i, err)


def test_config5_ip5m_shape_beam4_768(full_engine):
    cfg, eng = full_engine
    ocfg = O.MGConfig.full()
    inp = O.make_inputs(ocfg, 3, 512, seed=1239, ragged=True)  # text up to the 512-token cap
    ids = eng.generate(**inp, num_beams=4, max_length=768, trim=False)
    assert ids.shape == (3, 768) and (ids[:, 0] == 0).all()
    assert (ids >= 0).all() and (ids < cfg.vocab_size).all()
    for row in ids.cpu():
        pos = (row == 1).nonzero()
        if len(pos):
            assert (row[pos[0, 0] + 1:] == 1).all()   # stock beam search fills with `pad or eos` = eos for pad id 0


def test_full_size_beam4_matches_stock_beam_search():
    """beam=4 at the full dims against GenerationMixin._beam_search on the oracle (short decode, one image)"""
    ocfg = O.MGConfig.full()
    oracle = O.build(ocfg, seed=0)
    eng = MGEngine(ocfg, oracle.export_state())
    inp = O.make_inputs(ocfg, 1, 64, seed=1240)
    ref = oracle.hf_generate(**inp, max_length=10, num_beams=4)
    ids = eng.generate(**inp, max_length=10, num_beams=4)
    assert ids.shape == ref.shape
    assert torch.equal(ids.cpu(), ref)
    eng.close()
