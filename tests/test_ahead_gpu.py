"""Encoder run-ahead (mg_encode_ahead / mg_encode_ahead_host): the encoder of batch i+1 on its own SM partition while
batch i decodes on the rest -- token ids must be those of plain generate calls AND of the CPU oracle."""
import pytest
import torch

from markushgrapher_b200.engine import MGEngine
from oracle import mg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair():
    cfg = O.MGConfig.small()
    torch.set_num_threads(8)
    oracle = O.build(cfg, seed=0)
    eng = MGEngine(cfg, oracle.export_state())
    yield cfg, oracle, eng
    eng.close()


def test_run_ahead_device_inputs(pair):
    cfg, oracle, eng = pair
    dev = eng.device
    sets = [{k: v.to(dev) for k, v in O.make_inputs(cfg, 4, 16, seed=31 + i).items()} for i in range(3)]
    ref = [oracle.generate_greedy(**{k: v.cpu() for k, v in s.items()}, max_length=24) for s in sets]
    plain = [eng.generate(**s, max_length=24, trim=False).cpu() for s in sets]
    armed = eng.encode_ahead(**sets[0])
    if not armed:
        pytest.skip("SM partitioning (CUDA green contexts) unavailable on this device / driver")
    out = []
    for i in range(3):
        if i + 1 < 3:
            assert eng.encode_ahead(**sets[i + 1])   # batch i+1 is encoded while batch i decodes
        out.append(eng.generate(**sets[i], max_length=24, trim=False).cpu())
        info = eng.last_ahead()
        assert info["encoder_ms"] > 0 and info["sms_encoder"] >= 8 and info["sms_decoder"] >= 64
    eng.ahead_reset()
    for i in range(3):
        assert torch.equal(out[i], plain[i])
        n = ref[i].shape[1]
        assert torch.equal(out[i][:, :n], ref[i])
    # a batch encoded ahead and never asked for is dropped; later calls encode themselves on the whole chip
    assert eng.encode_ahead(**sets[2])
    eng.ahead_reset()
    again = eng.generate(**sets[1], max_length=24, trim=False).cpu()
    assert torch.equal(again, plain[1])
    assert eng.last_ahead()["encoder_ms"] == 0.0


def test_run_ahead_host_inputs_and_beams(pair):
    cfg, oracle, eng = pair
    host = [{k: v.pin_memory() for k, v in O.make_inputs(cfg, 3, 12, seed=41 + i).items()} for i in range(2)]
    plain = [eng.generate_host(**h, max_length=20, num_beams=nb, trim=False) for h, nb in zip(host, (1, 3))]
    if not eng.encode_ahead_host(**host[0], max_length=20):
        pytest.skip("SM partitioning (CUDA green contexts) unavailable on this device / driver")
    assert eng.encode_ahead_host(**host[1], max_length=20)
    a = eng.generate_host(**host[0], max_length=20, trim=False)               # greedy, fused step on the large partition
    b = eng.generate_host(**host[1], max_length=20, num_beams=3, trim=False)  # beam search takes a run-ahead batch too
    eng.ahead_reset()
    assert torch.equal(a, plain[0]) and torch.equal(b, plain[1])
    # refilled host buffers: the staged copy and its encoding are not reused
    host[0]["input_ids"].copy_(host[1]["input_ids"])
    host[0]["bbox"].copy_(host[1]["bbox"])
    host[0]["pixel_values"].copy_(host[1]["pixel_values"])
    c = eng.generate_host(**host[0], max_length=20, trim=False)
    d = eng.generate_host(**host[1], max_length=20, trim=False)
    assert torch.equal(c, d)
