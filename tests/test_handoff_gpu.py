"""GPU: the ChemicalOCR hand-off end to end and the predict.yaml decode settings.
VLM string -> cells_from_ocr_string -> MarkushgrapherProcessor.from_cells -> MarkushgrapherForConditionalGeneration
.generate on the device, against the oracle fed the same tensors."""
import os

import numpy as np
import pytest
import torch

from markushgrapher_b200.configuration import MarkushgrapherConfig
from markushgrapher_b200.modeling import MarkushgrapherForConditionalGeneration
from markushgrapher_b200.processing import (MarkushgrapherImageProcessor, MarkushgrapherProcessor, MarkushgrapherTokenizer,
                                            cells_from_ocr_string)
from oracle import mg_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    import sentencepiece as spm

    path = str(tmp_path_factory.mktemp("tok"))
    rng = np.random.RandomState(0)
    words = ["Question", "Answering", "What", "markush", "structure", "is", "in", "the", "image", "alkyl", "aryl", "R1",
             "R2", "halogen", "C1-C6", "methyl", "ethyl", "group", "hydrogen", "wherein", "selected", "from"]
    with open(os.path.join(path, "corpus.txt"), "w") as f:
        for _ in range(400):
            f.write(" ".join(rng.choice(words, size=8)) + ".\n")
    spm.SentencePieceTrainer.Train(input=os.path.join(path, "corpus.txt"), model_prefix=os.path.join(path, "spiece"),
                                   vocab_size=120, model_type="unigram", pad_id=0, eos_id=1, unk_id=2, bos_id=-1,
                                   minloglevel=2, hard_vocab_limit=False)
    tok = MarkushgrapherTokenizer.from_pretrained(path)
    vocab = tok.sp.GetPieceSize() + 5 * 100 + 501 + 200
    tok.vocab_size = vocab
    ocfg = O.MGConfig(vocab_size=vocab, d_model=128, d_ff=256, num_layers=2, num_decoder_layers=2, num_heads=2,
                      image_size=64, swin_image=96, swin_embed=32, swin_depths=(2, 2), swin_heads=(1, 2), proj_hidden=128)
    oracle = O.build(ocfg, seed=0)
    model = MarkushgrapherForConditionalGeneration(MarkushgrapherConfig.from_dims(ocfg))
    assert model.safe_load(model, oracle.export_state()) == []
    proc = MarkushgrapherProcessor(MarkushgrapherImageProcessor(apply_ocr=False, size={"height": 64, "width": 64}), tok)
    return oracle, model.to("cuda"), proc


def page():
    from PIL import Image, ImageDraw

    img = Image.new("RGB", (300, 200), "white")
    dr = ImageDraw.Draw(img)
    for i in range(5):
        dr.line([(30 + 40 * i, 80), (50 + 40 * i, 120 if i % 2 else 40)], fill="black", width=2)
    return img


@pytest.mark.parametrize("fmt", ["legacy", "new"])
def test_ocr_string_to_generate(setup, fmt):
    oracle, model, proc = setup
    if fmt == "legacy":
        s = ("<ocr><loc_0><loc_0><loc_500><loc_500>\n<loc_40><loc_50><loc_160><loc_70>R1 alkyl\n"
             "<loc_200><loc_300><loc_420><loc_330>halogen methyl group\n<loc_10><loc_10><loc_20><loc_20>   </ocr> bye")
    else:
        s = "<ocr>0>0>500>500>40>50>160>70>R1 alkyl\n200>300>420>330>halogen methyl group\n7>8>R2</ocr>"
    cells = cells_from_ocr_string(s)
    assert [c["text"] for c in cells] == ["R1 alkyl", "halogen methyl group"]
    enc = proc.from_cells(page(), cells)
    assert enc["input_ids"].shape[0] == 1 and enc["bbox"].shape[1] == enc["input_ids"].shape[1]
    ref = oracle.generate_greedy(enc["input_ids"], enc["bbox"], enc["pixel_values"], enc["attention_mask"], max_length=14)
    ids = model.generate(**{k: v.cuda() for k, v in enc.items()}, num_beams=1, max_length=14)
    assert torch.equal(ids.cpu(), ref)


def test_predict_yaml_sequence_beam5(setup):
    """reference config/predict.yaml: beam_search True -> utils_evaluation.py:278-281 calls generate(num_beams=5,
    max_length=512); the bare model has no `.module`, so that is the branch taken.  Batch of one, as the reference's
    per-sample loop (utils_evaluation.py:140-176) builds it, `labels` passed and ignored."""
    oracle, model, proc = setup
    assert not hasattr(model, "module")
    enc = proc.from_cells(page(), cells_from_ocr_string("<ocr>0>0>500>500>40>50>160>70>R1 alkyl</ocr>"))
    ref = oracle.hf_generate(enc["input_ids"], enc["bbox"], enc["pixel_values"], enc["attention_mask"], max_length=40,
                             num_beams=5)
    ids = model.generate(**{k: v.cuda() for k, v in enc.items()}, labels=torch.tensor([[5, 6, 1]]).cuda(), num_beams=5,
                         max_length=40)
    assert ids.shape == ref.shape and torch.equal(ids.cpu(), ref)
