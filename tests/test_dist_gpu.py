"""GPU: the NCCL path of the decode loop (per-step all-gather of token ids). With one visible GPU this runs a
single-rank communicator; under torchrun (WORLD_SIZE>1, one GPU per rank) it checks the cross-rank gather."""
import ctypes
import os

import pytest
import torch

from markushgrapher_b200 import _lib
from markushgrapher_b200.engine import MGEngine
from oracle import mg_oracle as O

pytestmark = pytest.mark.gpu


def test_single_rank_nccl_generate_matches_plain_generate():
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        pytest.skip("single-process test")
    cfg = O.MGConfig.tiny()
    oracle = O.build(cfg, seed=0)
    eng = MGEngine(cfg, oracle.export_state())
    L = _lib.lib()
    buf = ctypes.create_string_buffer(128)
    _lib.check(L.mg_nccl_unique_id(buf), "mg_nccl_unique_id")
    L.mg_comm_init.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
    _lib.check(L.mg_comm_init(eng._h, 1, 0, buf.raw), "mg_comm_init")
    eng.world, eng.rank = 1, 0
    inp = O.make_inputs(cfg, 3, 12, seed=8)
    ref = oracle.generate_greedy(**inp, max_length=18)
    plain = eng.generate(**inp, max_length=18, trim=False)
    dist_ids = eng.generate_dist(**inp, max_length=18)
    assert eng.dist_mode() == 2          # single rank: the peer-store exchange path (its own buffer is the only peer)
    assert torch.equal(dist_ids, plain)
    assert torch.equal(dist_ids.cpu()[:, : ref.shape[1]], ref)
    # beam search through the sharded entry (single rank): same ids as the plain call
    beam_plain = eng.generate(**inp, max_length=18, num_beams=4, trim=False)
    beam_dist = eng.generate_dist(**inp, max_length=18, num_beams=4)
    assert torch.equal(beam_dist, beam_plain)
    eng.close()
