"""CPU: BASELINE.json configs[0] plumbing -- the reference's OWN entry code drives the drop-in.

The reference's real `markushgrapher/core/common/begin.py` is imported from /root/reference (tests/ref_import.py stubs
only absent third-party packages), its `parse_hf_arguments` (begin.py:32-58) parses the reference's `config/predict.yaml`
UNCHANGED, and its `load_markushgrapher` (begin.py:85-179) builds tokenizer / processor / model from a tiny saved
checkpoint through `transformers.models.markushgrapher` = markushgrapher_b200.hf_shim.  Then the call site of
utils_evaluation.py:269-285 is replayed on the returned objects.  Without a GPU the last step must fail loudly (there is
no CPU fallback); the same sequence with compute is tests/test_model_gpu.py::test_predict_yaml_sequence_beam5.
"""
import os
import sys

import numpy as np
import pytest
import torch

import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference is only present in the build container")


def tiny_checkpoint(path):
    import sentencepiece as spm

    from markushgrapher_b200.configuration import MarkushgrapherConfig
    from markushgrapher_b200.modeling import MarkushgrapherForConditionalGeneration

    corpus = os.path.join(path, "corpus.txt")
    rng = np.random.RandomState(0)
    words = ["Question", "Answering", "What", "markush", "structure", "is", "in", "the", "image", "alkyl", "aryl", "R1",
             "R2", "halogen", "C1-C6", "methyl", "ethyl", "group", "hydrogen", "wherein", "selected", "from"]
    with open(corpus, "w") as f:
        for _ in range(400):
            f.write(" ".join(rng.choice(words, size=8)) + ".\n")
    spm.SentencePieceTrainer.Train(input=corpus, model_prefix=os.path.join(path, "spiece"), vocab_size=120,
                                   model_type="unigram", pad_id=0, eos_id=1, unk_id=2, bos_id=-1,
                                   minloglevel=2, hard_vocab_limit=False)
    sp = spm.SentencePieceProcessor()
    sp.Load(os.path.join(path, "spiece.model"))
    vocab = sp.GetPieceSize() + 5 * 100 + 501 + 200  # + the UDOP special-token blocks (processing.MarkushgrapherTokenizer)
    cfg = MarkushgrapherConfig(vocab_size=vocab, d_model=128, d_ff=256, num_layers=2, num_heads=2, image_size=64,
                               swin_image=96, swin_embed=32, swin_depths=(2, 2), swin_heads=(1, 2), proj_hidden=128)
    MarkushgrapherForConditionalGeneration(cfg).save_pretrained(path)
    return cfg


def test_reference_begin_loads_the_drop_in_from_predict_yaml(tmp_path, monkeypatch):
    ref_import.install()
    from markushgrapher.core.common import begin  # the reference's own module

    assert begin.__file__.startswith(ref_import.REF)
    import markushgrapher_b200.modeling as mb
    assert begin.MarkushgrapherForConditionalGeneration is mb.MarkushgrapherForConditionalGeneration

    # 1) the reference's predict.yaml, byte for byte
    monkeypatch.setattr(sys, "argv", ["markushgrapher.eval", os.path.join(ref_import.REF, "config", "predict.yaml")])
    model_args, data_args, training_args = begin.parse_hf_arguments()
    assert model_args.model_name_or_path == "./models/markushgrapher-2"
    assert model_args.tokenizer_path == model_args.model_name_or_path            # "auto" expansion, begin.py:52-56
    assert model_args.architecture_variant == "me-lf-stack-1" and model_args.beam_search is True
    assert data_args.image_size == 512 and data_args.max_seq_length == 512 and data_args.max_seq_length_decoder == 512
    assert training_args.do_predict and not training_args.do_train

    # 2) a per-run copy pointing at a tiny checkpoint, as scripts/inference/inference.sh:186-243 writes one
    ck = tmp_path / "ckpt"
    ck.mkdir()
    tiny_checkpoint(str(ck))
    text = open(os.path.join(ref_import.REF, "config", "predict.yaml")).read()
    text = text.replace("./models/markushgrapher-2", str(ck)).replace("output_dir: auto", f"output_dir: {tmp_path / 'out'}")
    run_yaml = tmp_path / "predict.yaml"
    run_yaml.write_text(text)
    monkeypatch.setattr(sys, "argv", ["markushgrapher.eval", str(run_yaml)])
    model_args, data_args, training_args = begin.parse_hf_arguments()
    monkeypatch.chdir(tmp_path)  # init_molscribe_weights looks for external/MolScribe/ckpts relative to the cwd
    device = torch.device("cpu")  # world_size 1, no GPU in this container (begin.get_device would say the same)
    assert begin.get_device().type == ("cuda" if torch.cuda.is_available() else "cpu")
    tokenizer, processors, model = begin.load_markushgrapher(
        model_args, data_args, training_args, device, use_pretrained_molscribe=data_args.use_pretrained_molscribe)  # eval.py:49-56
    assert isinstance(model, mb.MarkushgrapherForConditionalGeneration)
    assert model.config.image_size == 512 and model.config.architecture_variant == "me-lf-stack-1"
    assert model.config.output_attentions is True                                   # begin.py:119-121
    assert set(processors) == {"no_ocr"}

    # 3) the call site (utils_evaluation.py:140-176, 269-285): one 512x512 synthetic molecule PNG, batch of one
    from PIL import Image, ImageDraw

    img = Image.new("RGB", (512, 512), "white")
    dr = ImageDraw.Draw(img)
    for i in range(6):
        dr.line([(60 + 60 * i, 200), (90 + 60 * i, 250 if i % 2 else 150)], fill="black", width=2)
    png = tmp_path / "mol.png"
    img.save(png)
    enc = processors["no_ocr"](images=Image.open(png).convert("RGB"),
                               text=["Question Answering. What markush structure is in the image?"],
                               text_pair=[["R1", "alkyl", "group"]],
                               boxes=[[[0.1, 0.1, 0.2, 0.15], [0.3, 0.1, 0.4, 0.15], [0.5, 0.1, 0.6, 0.15]]],
                               return_tensors="pt", padding=False, truncation=False)
    encoding = {"input_ids": enc["input_ids"].long().to(device), "bbox": enc["bbox"].float().to(device),
                "pixel_values": enc["pixel_values"].to(device), "labels": torch.tensor([[5, 6, 1]])}
    assert encoding["pixel_values"].shape == (1, 3, 512, 512) and encoding["bbox"].shape[-1] == 4
    assert not hasattr(model, "module")        # -> the predict.yaml branch: beam_search True -> num_beams=5
    num_beams = 5 if model_args.beam_search else 1
    if torch.cuda.is_available():
        ids = model.to("cuda").generate(**{k: v.to("cuda") for k, v in encoding.items()}, num_beams=num_beams, max_length=16)
        assert ids.shape[0] == 1 and ids[0, 0] == 0
    else:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model.generate(**encoding, num_beams=num_beams, max_length=512)
