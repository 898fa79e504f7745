"""GPU: the batched detokeniser kernels against golden strings from the reference's own function."""
import json
import os

import pytest
import torch

from markushgrapher_b200.detok import BatchedDetokenizer

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "detok_reference.json")


def test_batched_detokeniser_byte_exact_vs_reference_golden():
    with open(GOLD, encoding="utf-8") as f:
        gold = json.load(f)
    for case in gold["cases"]:
        dt = BatchedDetokenizer(gold["pieces"], case["vocabulary"], case["vocabulary_inverse"], case["encode_index"])
        rows = case["ids"]
        T = max(1, max(len(r) for r in rows))
        ids = torch.zeros(len(rows), T, dtype=torch.long)
        lens = torch.tensor([len(r) for r in rows], dtype=torch.int32)
        for i, r in enumerate(rows):
            ids[i, : len(r)] = torch.tensor(r, dtype=torch.long)
        got = dt.decode(ids.cuda(), lens.cuda())                  # ragged rows, empty rows included
        assert got == case["expected"]
        full = [r for r in rows if len(r) == T]                   # rows that fill the matrix: lens omitted
        if full:
            got2 = dt.decode(torch.tensor(full, dtype=torch.long).cuda())
            assert got2 == [e for r, e in zip(rows, case["expected"]) if len(r) == T]
        dt.close()
