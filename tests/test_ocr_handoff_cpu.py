"""CPU: the ChemicalOCR hand-off -- VLM string -> OCR cells -> model inputs -- pinned against golden
vectors produced by the REFERENCE's own clean_ocr_text / parse_ocr_string (oracle/make_ocr_golden.py,
reference markushgrapher/ocr/chemical_ocr.py:165-223)."""
import json
import os

from markushgrapher_b200.processing import cells_from_ocr_string, clean_ocr_text, parse_ocr_string

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ocr_reference.json")


def test_parse_ocr_string_matches_reference():
    cases = json.load(open(GOLD, encoding="utf-8"))["cases"]
    assert len(cases) >= 60 and sum(len(c["words"]) for c in cases) > 300
    n_legacy = n_new = 0
    for c in cases:
        assert clean_ocr_text(c["raw"]) == c["cleaned"], c["raw"][:80]
        words, boxes = parse_ocr_string(c["cleaned"])
        assert words == c["words"], c["raw"][:80]
        assert boxes == c["boxes"], c["raw"][:80]            # float-exact: the same integer / 500 divisions
        cells = cells_from_ocr_string(c["raw"])
        assert [x["text"] for x in cells] == c["words"] and [x["bbox"] for x in cells] == c["boxes"]
        n_legacy += "<loc_" in c["cleaned"]
        n_new += "<loc_" not in c["cleaned"]
    assert n_legacy > 10 and n_new > 10


def test_empty_and_malformed_strings():
    assert parse_ocr_string("") == ([], [])
    assert parse_ocr_string("<ocr></ocr>") == ([], [])
    assert parse_ocr_string("<ocr>no boxes at all</ocr>") == ([], [])
    assert cells_from_ocr_string("chatter <ocr>0>0>500>500>10>20>30>40>R1</ocr> more") == [
        {"bbox": [10 / 500, 20 / 500, 30 / 500, 40 / 500], "text": "R1"}]
