"""GPU: the two decode paths (fused persistent decode-step kernel / per-operation kernel chain) and the kv24 storage
format of the cross K/V.  Properties that do not need the oracle are checked at the full model size and batch 32,
where every CTA of the fused kernel carries four concurrent self-attention items and four cross-attention items."""
import ctypes
import os

import numpy as np
import pytest
import torch

from markushgrapher_b200 import _lib
from markushgrapher_b200.configuration import MarkushgrapherConfig, random_state
from markushgrapher_b200.engine import MGEngine
from oracle import mg_oracle as O

pytestmark = pytest.mark.gpu


class _env:
    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def kv24_reference(x: np.ndarray) -> np.ndarray:
    """numpy statement of the format: fp32 rounded to nearest-even at bit 8 of the word, low 8 bits dropped"""
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7F + ((u >> 8) & 1)) >> 8 << 8
    return (u & 0xFFFFFFFF).astype(np.uint32).view(np.float32)


def test_kv24_roundtrip_bit_exact_and_idempotent():
    B, H, Mp = 2, 3, 40
    g = torch.Generator().manual_seed(5)
    kt = torch.randn(B, H, 64 * Mp, generator=g) * torch.logspace(-6, 6, 64 * Mp)
    v = torch.randn(B, H, 64 * Mp, generator=g)
    v[0, 0, :8] = torch.tensor([0.0, -0.0, 1.0, -1.0, 2.0 ** -126, 3.0e38, 1.0 + 2.0 ** -16, 1.0 + 2.0 ** -15])
    out = torch.empty(2, B, H, 64 * Mp, device="cuda")
    ktd, vd = kt.cuda(), v.cuda()
    _lib.lib().mg_op_kv24_roundtrip.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    rc = _lib.lib().mg_op_kv24_roundtrip(_lib.cur_stream(), B, H, Mp, _lib.ptr(ktd), _lib.ptr(vd), _lib.ptr(out))
    _lib.check(rc, "mg_op_kv24_roundtrip")
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.array_equal(got[0].view(np.uint32), kv24_reference(kt.numpy()).view(np.uint32))   # bit exact
    assert np.array_equal(got[1].view(np.uint32), kv24_reference(v.numpy()).view(np.uint32))
    rel = np.abs(got[0] - kt.numpy()) / np.abs(kt.numpy())
    assert rel.max() <= 2.0 ** -16 + 1e-12            # 15 stored mantissa bits: half an ulp = 2^-16 relative
    out2 = torch.empty_like(out)
    a, b = out[0].contiguous(), out[1].contiguous()
    _lib.check(_lib.lib().mg_op_kv24_roundtrip(_lib.cur_stream(), B, H, Mp, _lib.ptr(a), _lib.ptr(b), _lib.ptr(out2)),
               "mg_op_kv24_roundtrip")
    assert torch.equal(out2, out)                                                                # idempotent


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_kernel_chain_path_token_identical_to_oracle(name):
    """MG_DECODE=chain: the per-operation kernel chain (what batches > 32 and beam search always use)"""
    cfg = getattr(O.MGConfig, name)()
    oracle = O.build(cfg, seed=0)
    inp = O.make_inputs(cfg, 4, 16, seed=15)
    ids_ref, lg_ref = oracle.generate_greedy(**inp, max_length=40, return_logits=True)
    for kv24 in ("1", "0"):
        with _env(MG_DECODE="chain", MG_KV24=kv24):
            eng = MGEngine(cfg, oracle.export_state())
            ids, lg = eng.generate(**inp, max_length=40, return_logits=True)
            eng.close()
        err = ((lg.cpu().double() - lg_ref.double()).norm() / lg_ref.double().norm()).item()
        assert err < 1e-3, (kv24, err)
        assert torch.equal(ids.cpu(), ids_ref), kv24


@pytest.mark.parametrize("path", ["mega", "chain"])
def test_masked_memory_positions_are_dropped_without_changing_results(path):
    """ragged text (up to 3/4 of a row is padding) + patches absorbed by OCR tokens: the decoder-side memory keeps the
    valid positions only (MG_COMPACT=0 keeps everything); ids identical, logits equal to fp32 reordering noise, both
    equal to the oracle, for greedy and beam search"""
    cfg = O.MGConfig.small()
    oracle = O.build(cfg, seed=0)
    inp = O.make_inputs(cfg, 5, 40, seed=61, ragged=True)
    ids_ref, lg_ref = oracle.generate_greedy(**inp, max_length=24, return_logits=True)
    beam_ref = oracle.hf_generate(**inp, max_length=20, num_beams=3)
    out = {}
    for compact in ("1", "0"):
        with _env(MG_DECODE=path, MG_COMPACT=compact):
            eng = MGEngine(cfg, oracle.export_state())
            ids, lg = eng.generate(**inp, max_length=24, return_logits=True)
            m_enc, m_dec = eng.last_memory_len()
            beam = eng.generate(**inp, max_length=20, num_beams=3)
            eng.close()
        out[compact] = (ids.cpu(), lg.cpu(), beam.cpu(), m_enc, m_dec)
    assert out["1"][3] == out["0"][3] and out["0"][4] >= out["0"][3]      # MG_COMPACT=0: every (padded) position kept
    assert out["1"][4] < out["1"][3] - 8, out["1"][3:]                     # positions really dropped
    assert torch.equal(out["1"][0], ids_ref) and torch.equal(out["0"][0], ids_ref)
    assert torch.equal(out["1"][2], beam_ref) and torch.equal(out["0"][2], beam_ref)
    d = (out["1"][1].double() - out["0"][1].double()).norm() / out["0"][1].double().norm()
    assert d < 1e-5, d
    assert ((out["1"][1].double() - lg_ref.double()).norm() / lg_ref.double().norm()) < 1e-3


def test_full_size_batch32_paths_agree():
    """batch 32, full dims, 200 tokens: the fused kernel (148 CTAs, 4 concurrent items each), the kernel chain with
    kv24 and the kernel chain with fp32 cross K/V must emit the same token ids; the first two also agree on the
    logits to fp32 reordering noise."""
    import bench

    cfg = MarkushgrapherConfig()
    dev = torch.device("cuda", 0)
    inp = {k: v.to(dev) for k, v in bench.synth_inputs(512, 32, 64, 1234, cfg.vocab_size).items()}
    res = {}
    for tag, env in (("fused", dict(MG_DECODE="mega", MG_KV24="1")), ("chain", dict(MG_DECODE="chain", MG_KV24="1")),
                     ("chain_fp32", dict(MG_DECODE="chain", MG_KV24="0"))):
        with _env(**env):
            eng = MGEngine(cfg, random_state(cfg, 0, dev), device=dev)
            ids = eng.generate(**inp, max_length=200, trim=False)
            ids_s, lg = eng.generate(**{k: v[:4] for k, v in inp.items()}, max_length=6, return_logits=True)
            launches = eng.last_stats()["kernels"]
            assert eng.last_decode_loop()["fused"] == (tag == "fused")
            eng.close()
        res[tag] = (ids.cpu(), lg.cpu(), launches)
        torch.cuda.empty_cache()
    assert res["fused"][0].shape == (32, 200)
    assert torch.equal(res["fused"][0], res["chain"][0])
    assert torch.equal(res["chain"][0], res["chain_fp32"][0])
    d = (res["fused"][1].double() - res["chain"][1].double()).norm() / res["chain"][1].double().norm()
    assert d < 1e-5, d
    d32 = (res["chain"][1].double() - res["chain_fp32"][1].double()).norm() / res["chain_fp32"][1].double().norm()
    assert d32 < 1e-4, d32
    assert res["fused"][2] < res["chain"][2]   # the fused path really ran: two launches per token
