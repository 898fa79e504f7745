"""Golden vectors for the OCR-cell packing step: outputs of the REFERENCE's own
`prepare_cells_to_text` (reference markushgrapher/core/common/data_preprocessing.py:59-104, which splits every OCR
cell box into per-sub-word boxes, :24-48).

TEST INFRASTRUCTURE; runs only in the build container (imports /root/reference).  The reference module imports
`markushgrapher.core.common.utils`, whose own imports (matplotlib, a private torch symbol) are absent here; the two
helpers it needs (`check_max_values`, `normalize_bbox_format`, utils.py:212-222) are compiled straight from the
reference source into a stand-in module at generation time -- nothing is copied into this repository.  The
sentencepiece model is absent as well: a deterministic sub-word splitter with the same piece conventions (U+2581
word-start marker, a lone marker piece) stands in; the committed golden file carries its pieces per text so the
tests replay them without the reference.

usage: python oracle/make_cells_golden.py
"""
import ast
import json
import os
import random
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


class StubSP:
    """sentencepiece-like: words split on spaces, each word cut into pieces of 1-4 chars, first piece carries '▁';
    with some probability a lone '▁' piece precedes a word (as sentencepiece does before digits / symbols)"""

    def __init__(self):
        self.memo = {}

    def tokenize(self, text):
        if text not in self.memo:
            rnd = random.Random(hash(text) & 0xFFFF)
            out = []
            for w in text.split():
                if rnd.random() < 0.15:
                    out.append("▁")
                    first = False
                else:
                    first = True
                i = 0
                while i < len(w):
                    n = rnd.randint(1, 4)
                    out.append(("▁" if first else "") + w[i:i + n])
                    first = False
                    i += n
            self.memo[text] = out
        return list(self.memo[text])


def load_reference():
    src = open(os.path.join(REF, "markushgrapher/core/common/utils.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("check_max_values", "normalize_bbox_format")]
    mod = types.ModuleType("markushgrapher.core.common.utils")
    exec(compile(ast.Module(body=keep, type_ignores=[]), "reference utils.py (two helpers)", "exec"), mod.__dict__)
    sys.path.insert(0, REF)
    import markushgrapher.core.common  # noqa: F401  (package __init__)
    sys.modules["markushgrapher.core.common.utils"] = mod
    from markushgrapher.core.common import data_preprocessing

    return data_preprocessing


def main():
    try:
        dp = load_reference()
    except Exception as e:  # the package __init__ may import the absent forks: load the module file directly
        import importlib.util

        src = open(os.path.join(REF, "markushgrapher/core/common/utils.py")).read()
        tree = ast.parse(src)
        keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("check_max_values", "normalize_bbox_format")]
        for name in ("markushgrapher", "markushgrapher.core", "markushgrapher.core.common"):
            sys.modules.setdefault(name, types.ModuleType(name))
        mod = types.ModuleType("markushgrapher.core.common.utils")
        exec(compile(ast.Module(body=keep, type_ignores=[]), "reference utils.py (two helpers)", "exec"), mod.__dict__)
        sys.modules["markushgrapher.core.common.utils"] = mod
        spec = importlib.util.spec_from_file_location("markushgrapher.core.common.data_preprocessing",
                                                      os.path.join(REF, "markushgrapher/core/common/data_preprocessing.py"))
        dp = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(dp)
        print("note: loaded data_preprocessing.py directly (", type(e).__name__, ")")
    tok = StubSP()
    rnd = random.Random(7)
    vocab = ["R1", "R2", "alkyl", "C1-C6", "is", "selected", "from", "H,", "halogen", "OMe", "(I)", "wherein", "n=1-3",
             "X", "=", "O", "or", "S", "phenyl", "12", "3.5", "µm", " ", "  "]
    cases = []
    for ci in range(24):
        n_cells = rnd.choice([0, 1, 3, 12, 40, 120])
        cells = []
        for _ in range(n_cells):
            text = " ".join(rnd.choice(vocab) for _ in range(rnd.randint(1, 5))) if rnd.random() > 0.05 else "   "
            x0, y0 = rnd.random() * 0.9, rnd.random() * 0.9
            x1, y1 = min(x0 + rnd.random() * 0.3, 1.0 if rnd.random() < 0.9 else 1.02), min(y0 + rnd.random() * 0.05, 1.0)
            cells.append({"text": text, "bbox": [x0, y0, x1, y1]})
        for normalize_bbox in (True, False):
            w, h = (1, 1) if normalize_bbox else (512, 512)
            max_len = rnd.choice([512, 64])
            words, boxes, n_tok = dp.prepare_cells_to_text(cells, tok, w, h, normalize_bbox, max_len)
            cases.append({"cells": cells, "w": w, "h": h, "normalize_bbox": normalize_bbox, "max_sequence_length": max_len,
                          "words": words, "bboxes": [list(b) for b in boxes], "token_idx": n_tok})
    out = {"pieces": tok.memo, "cases": cases}
    path = os.path.join(ROOT, "tests", "golden", "cells_reference.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)
    print("wrote", path, len(cases), "cases,", sum(len(c["words"]) for c in cases), "words")


if __name__ == "__main__":
    main()
