"""GPU: CUDA path vs the committed golden fixtures (tests/golden/*.npz, written by oracle/make_golden.py).
The oracle is only used here to rebuild the seeded weights/pixels; outputs are compared with the stored bytes."""
import os

import numpy as np
import pytest
import torch

from markushgrapher_b200.engine import MGEngine
from oracle import make_golden
from oracle import mg_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_cuda_matches_golden(name):
    cfg_name, B, Lt, ragged, seed, max_len = make_golden.CASES[name]
    cfg = getattr(O.MGConfig, cfg_name)()
    g = np.load(os.path.join(GOLD, name + ".npz"))
    oracle = O.build(cfg, seed=0)  # weights only
    inp = O.make_inputs(cfg, B, Lt, seed=seed, ragged=ragged)
    assert np.array_equal(inp["input_ids"].numpy(), g["input_ids"])
    assert abs(inp["pixel_values"].double().sum().item() - g["pixel_checksum"][0]) < 1e-6
    eng = MGEngine(cfg, oracle.export_state())
    mem, mask = eng.encode(**inp)
    assert np.array_equal(mask.cpu().numpy(), g["memory_mask"])
    ref = torch.from_numpy(g["memory"]).double()
    err = ((mem.cpu().double() - ref).norm() / ref.norm()).item()
    assert err < 1e-3, err
    ids, logits = eng.generate(**inp, max_length=max_len, return_logits=True)
    assert np.array_equal(ids.cpu().numpy(), g["greedy_ids"])
    np.testing.assert_allclose(logits[:, 0].cpu().numpy(), g["step0_logits"], rtol=0, atol=5e-4)
    np.testing.assert_allclose(logits[:, -1].cpu().numpy(), g["last_logits"], rtol=0, atol=5e-4)
    beam = eng.generate(**inp, max_length=max_len, num_beams=4)
    assert np.array_equal(beam.cpu().numpy(), g["beam4_ids"])
    tf = eng.forward_logits(inp["input_ids"], inp["bbox"], inp["pixel_values"],
                            torch.from_numpy(g["greedy_ids"][:, :-1]), inp["attention_mask"])
    np.testing.assert_allclose(tf[:, -1].cpu().numpy(), g["tf_last_logits"], rtol=0, atol=5e-4)
    eng.close()
