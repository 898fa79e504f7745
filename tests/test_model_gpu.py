"""GPU parity of the full hot path (through the C ABI) against the CPU oracle on seeded inputs."""
import pytest
import torch

from markushgrapher_b200.engine import MGEngine
from oracle import mg_oracle as O

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(scope="module", params=["tiny", "small"])
def pair(request):
    cfg = getattr(O.MGConfig, request.param)()
    torch.set_num_threads(8)
    oracle = O.build(cfg, seed=0)
    eng = MGEngine(cfg, oracle.export_state())
    yield cfg, oracle, eng
    eng.close()


@pytest.mark.parametrize("B,Lt,ragged", [(2, 12, False), (3, 21, True)])
def test_encoder_hidden_states(pair, B, Lt, ragged):
    cfg, oracle, eng = pair
    inp = O.make_inputs(cfg, B, Lt, seed=7 + B, ragged=ragged)
    mem_ref, mask_ref = oracle.encode(**inp)
    mem, mask = eng.encode(**inp)
    assert torch.equal(mask.cpu().long(), mask_ref.long())          # integer work: bit exact
    n_sw = eng.swin_tokens
    e_swin = rel_err(mem[:, :n_sw], mem_ref[:, :n_sw])
    valid = mask_ref.bool()
    e_vtl = rel_err(mem.cpu()[valid], mem_ref[valid])
    e_all = rel_err(mem, mem_ref)
    print(f"encoder rel err: swin {e_swin:.2e} valid {e_vtl:.2e} all {e_all:.2e}")
    assert e_swin < 1e-3 and e_vtl < 1e-3 and e_all < 1e-3            # north-star tolerance: 1e-3 fp32 relative
    assert e_all < 2e-4                                              # what the split-bf16 path actually delivers


@pytest.mark.parametrize("B,Lt,max_length", [(2, 12, 24), (4, 16, 40)])
def test_greedy_token_identical(pair, B, Lt, max_length):
    cfg, oracle, eng = pair
    inp = O.make_inputs(cfg, B, Lt, seed=11 + B)
    ids_ref, logits_ref = oracle.generate_greedy(**inp, max_length=max_length, return_logits=True)
    ids, logits = eng.generate(**inp, max_length=max_length, return_logits=True)
    top2 = logits_ref.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1]).min().item()
    n = min(logits.shape[1], logits_ref.shape[1])
    err = rel_err(logits[:, :n], logits_ref[:, :n])
    print(f"logit rel err {err:.2e}, min top-2 margin {margin:.2e}, distinct tokens {ids_ref.unique().numel()}")
    assert ids_ref.unique().numel() > max_length // 3, "degenerate oracle decode"
    assert err < 1e-3
    assert ids.shape == ids_ref.shape
    assert torch.equal(ids.cpu(), ids_ref), (ids.cpu(), ids_ref)   # token ids: bit exact


def test_generate_host_path(pair):
    cfg, oracle, eng = pair
    inp = O.make_inputs(cfg, 2, 12, seed=3)
    ids_ref = oracle.generate_greedy(**inp, max_length=16)
    ids = eng.generate_host(**{k: v.pin_memory() for k, v in inp.items()}, max_length=16)
    assert not ids.is_cuda
    assert torch.equal(ids, ids_ref)


def test_prefetch_host_overlapped_staging(pair):
    """mg_prefetch_host: the next batch's inputs travel on the copy stream while another batch decodes; batches are
    matched by host pointer, unmatched calls stage inside the call, a prefetched batch is never overwritten"""
    cfg, oracle, eng = pair
    a = {k: v.pin_memory() for k, v in O.make_inputs(cfg, 2, 12, seed=3).items()}
    b = {k: v.pin_memory() for k, v in O.make_inputs(cfg, 3, 12, seed=4).items()}
    ref_a, ref_b = oracle.generate_greedy(**a, max_length=16), oracle.generate_greedy(**b, max_length=16)
    eng.prefetch_host(**a, max_length=16)
    eng.prefetch_host(**b, max_length=16)           # both slots hold a pending batch now
    assert torch.equal(eng.generate_host(**a, max_length=16), ref_a)
    eng.prefetch_host(**a, max_length=16)           # recycles a's slot, b stays staged
    assert torch.equal(eng.generate_host(**b, max_length=16), ref_b)
    assert torch.equal(eng.generate_host(**a, max_length=16), ref_a)
    assert torch.equal(eng.generate_host(**b, max_length=16), ref_b)   # nothing staged: copies inside the call


def test_drop_in_model_class_generate_and_logits():
    """the reference-facing object (MarkushgrapherForConditionalGeneration) end to end: generate() as called at
    utils_evaluation.py:279-285 and model(**batch).logits as called at curriculumTrainer.py:655"""
    from markushgrapher_b200.configuration import MarkushgrapherConfig
    from markushgrapher_b200.modeling import MarkushgrapherForConditionalGeneration

    cfg = O.MGConfig.tiny()
    oracle = O.build(cfg, seed=0)
    model = MarkushgrapherForConditionalGeneration(MarkushgrapherConfig.from_dims(cfg))
    assert model.safe_load(model, oracle.export_state()) == []
    model = model.to("cuda")
    inp = O.make_inputs(cfg, 2, 12, seed=31)
    enc = {k: v.to("cuda") for k, v in inp.items()}
    labels = torch.randint(3, cfg.vocab_size, (2, 9))
    ids = model.generate(**enc, labels=labels.to("cuda"), num_beams=1, max_length=14)
    ids_ref = oracle.generate_greedy(**inp, max_length=14)
    assert torch.equal(ids.cpu(), ids_ref)
    labels[1, 6:] = -100
    out = model(**enc, labels=labels.to("cuda"))
    ref = oracle.forward_logits(inp["input_ids"], inp["bbox"], inp["pixel_values"], labels, inp["attention_mask"])
    assert out.logits.shape == ref.shape
    assert rel_err(out.logits, ref) < 1e-4
    ref_loss = torch.nn.functional.cross_entropy(ref.view(-1, ref.shape[-1]), labels.view(-1), ignore_index=-100)
    assert abs(out.loss.item() - ref_loss.item()) < 1e-3


@pytest.mark.parametrize("B,Lt,max_length,nb", [(2, 12, 20, 4), (3, 16, 28, 5), (1, 12, 16, 2)])
def test_beam_search_token_identical(pair, B, Lt, max_length, nb):
    """num_beams > 1 (the reference's predict.yaml default is beam_search: True -> 5 beams) against the stock
    GenerationMixin._beam_search run on the oracle"""
    cfg, oracle, eng = pair
    inp = O.make_inputs(cfg, B, Lt, seed=40 + B + nb, ragged=(B == 3))
    ref = oracle.hf_generate(**inp, max_length=max_length, num_beams=nb)
    ids = eng.generate(**inp, max_length=max_length, num_beams=nb)
    assert ids.shape == ref.shape, (ids.shape, ref.shape)
    assert torch.equal(ids.cpu(), ref), (ids.cpu(), ref)


def test_beam_search_with_early_eos(pair):
    """make EOS likely so hypotheses finish at different lengths: exercises the finished-beam merge, the length
    normalisation and the early-stop heuristic"""
    cfg, oracle, eng0 = pair
    import copy
    o2 = copy.deepcopy(oracle)
    with torch.no_grad():
        o2.lm_head.weight[1] = o2.lm_head.weight[1] * 0.0 + o2.lm_head.weight[7] * 1.5 + o2.lm_head.weight[11] * 1.5
    eng = MGEngine(cfg, o2.export_state())
    inp = O.make_inputs(cfg, 4, 14, seed=77)
    for nb in (3, 5):
        ref = o2.hf_generate(**inp, max_length=40, num_beams=nb)
        ids = eng.generate(**inp, max_length=40, num_beams=nb)
        assert (ref == 1).any(), "test did not produce any EOS"
        assert ids.shape == ref.shape, (ids.shape, ref.shape)
        assert torch.equal(ids.cpu(), ref), (ids.cpu(), ref)
    eng.close()
