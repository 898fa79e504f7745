"""Golden vectors for the ChemicalOCR hand-off: outputs of the REFERENCE's own `clean_ocr_text` and
`parse_ocr_string` (reference markushgrapher/ocr/chemical_ocr.py:165-223) on synthetic VLM strings in both formats.

TEST INFRASTRUCTURE; runs only in the build container.  chemical_ocr.py imports vllm / datasets at module level, so the
two pure functions are compiled straight from the reference source file at generation time (ast) -- nothing is copied
into this repository.

usage: python oracle/make_ocr_golden.py   -> tests/golden/ocr_reference.json
"""
import ast
import json
import os
import random
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def load_reference():
    src = open(os.path.join(REF, "markushgrapher/ocr/chemical_ocr.py")).read()
    keep = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("parse_ocr_string", "clean_ocr_text")]
    assert len(keep) == 2
    ns = {"re": re}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "reference chemical_ocr.py (two functions)", "exec"), ns)
    return ns["clean_ocr_text"], ns["parse_ocr_string"]


def main():
    clean, parse = load_reference()
    rnd = random.Random(11)
    vocab = ["R1", "R2 =", "alkyl", "C1-C6 alkyl", "is selected from", "H, halogen", "OMe", "(I)", "wherein n=1-3", "X = O or S",
             "phenyl", "12", "3.5 µm", "a>b", "<sub>2</sub>", "N>", "5>6", "  padded  "]
    cases = []
    for ci in range(60):
        n = rnd.choice([0, 1, 2, 5, 17, 60])
        legacy = ci % 2 == 0
        lines = []
        for k in range(n):
            x1, y1 = rnd.randint(0, 480), rnd.randint(0, 480)
            x2, y2 = min(500, x1 + rnd.randint(1, 120)), min(500, y1 + rnd.randint(1, 30))
            text = rnd.choice(vocab) if rnd.random() > 0.08 else rnd.choice(["", "   "])
            kind = rnd.random()
            if legacy:
                if kind < 0.06:
                    lines.append(f"<loc_{x1}><loc_{y1}>{text}")                       # too few numbers
                elif kind < 0.12:
                    lines.append(f"<loc_3><loc_4><loc_{x1}><loc_{y1}><loc_{x2}><loc_{y2}>{text}")   # extra numbers: last four win
                else:
                    lines.append(f"<loc_{x1}><loc_{y1}><loc_{x2}><loc_{y2}>{text}")
            else:
                if kind < 0.06:
                    lines.append(f"{x1}>{y1}>{text}")
                elif kind < 0.12:
                    lines.append(f"  {x1}>{y1}>{x2}>{y2}>{text}  ")
                else:
                    lines.append(f"{x1}>{y1}>{x2}>{y2}>{text}")
        if legacy:
            body = ("<loc_0><loc_0><loc_500><loc_500>" if rnd.random() < 0.8 else "") + ("\n" if rnd.random() < 0.7 else "") + "\n".join(lines)
        else:
            first = ("0>0>500>500>" if rnd.random() < 0.8 else "")
            body = first + "\n".join(lines)
        wrap = rnd.random()
        if wrap < 0.5:
            raw = f"<ocr>{body}</ocr>"
        elif wrap < 0.75:
            raw = f"Assistant: here it is <ocr>{body}</ocr><end_of_utterance> trailing"
        elif wrap < 0.9:
            raw = f"<ocr>{body}"
        else:
            raw = body
        if rnd.random() < 0.25:
            raw += rnd.choice(["\n", " \n", "\n\n"])
        cleaned = clean(raw)
        words, boxes = parse(cleaned)
        cases.append({"raw": raw, "cleaned": cleaned, "words": words, "boxes": boxes})
    path = os.path.join(ROOT, "tests", "golden", "ocr_reference.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump({"cases": cases}, f, ensure_ascii=False)
    print("wrote", path, len(cases), "cases,", sum(len(c["words"]) for c in cases), "words")


if __name__ == "__main__":
    main()
