"""CPU: the OCR-cell packing mirror (processing.prepare_cells_to_text) against golden vectors produced by the
REFERENCE's own function (tests/golden/cells_reference.json, oracle/make_cells_golden.py): words equal, token counts
equal, boxes bit-identical (doubles / truncated ints)."""
import json
import os

import pytest

from markushgrapher_b200 import processing

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cells_reference.json")


class ReplayTokenizer:
    def __init__(self, pieces):
        self.pieces = pieces

    def tokenize(self, text):
        return list(self.pieces[text])


@pytest.fixture(scope="module")
def gold():
    with open(GOLD, encoding="utf-8") as f:
        return json.load(f)


def test_prepare_cells_to_text_matches_reference(gold):
    tok = ReplayTokenizer(gold["pieces"])
    n_words = 0
    for c in gold["cases"]:
        words, boxes, n_tok = processing.prepare_cells_to_text(c["cells"], tok, c["w"], c["h"], c["normalize_bbox"],
                                                               c["max_sequence_length"])
        assert words == c["words"]
        assert n_tok == c["token_idx"]
        assert [list(b) for b in boxes] == c["bboxes"]          # exact: same doubles, same truncated ints
        n_words += len(words)
    assert n_words > 3000


def test_collate_cells_normalises_by_image_size(gold):
    from PIL import Image

    tok = ReplayTokenizer(gold["pieces"])
    c = next(c for c in gold["cases"] if c["normalize_bbox"] and len(c["words"]) > 5)
    im = Image.new("RGB", (1, 1))
    _, instruction, words, boxes = processing.collate_cells(im, c["cells"], tok, "What markush structure is in the image?")
    assert instruction == "Question Answering. What markush structure is in the image?"
    assert words == c["words"] and boxes == c["bboxes"]
