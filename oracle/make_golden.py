"""Generates tests/golden/*.npz from the CPU oracle (oracle/mg_oracle.py).  TEST INFRASTRUCTURE.

The reference ships no golden vectors for this path (parity unpinned), so these fixtures pin
the oracle itself: tests/test_oracle_cpu.py regenerates them on every CPU run, tests/test_golden_gpu.py checks
the CUDA path against the committed bytes without running the oracle.

    python oracle/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mg_oracle as O  # noqa: E402

CASES = {
    # name: (config, batch, text_len, ragged, input seed, max_length)
    "tiny_b2": ("tiny", 2, 12, False, 101, 20),
    "tiny_ragged_b3": ("tiny", 3, 21, True, 102, 16),
    "small_b2": ("small", 2, 16, False, 103, 24),
}


def run_case(name):
    cfg_name, B, Lt, ragged, seed, max_len = CASES[name]
    cfg = getattr(O.MGConfig, cfg_name)()
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    model = O.build(cfg, seed=0)
    inp = O.make_inputs(cfg, B, Lt, seed=seed, ragged=ragged)
    mem, mask = model.encode(**inp)
    ids, logits = model.generate_greedy(**inp, max_length=max_len, return_logits=True)
    beam = model.hf_generate(**inp, max_length=max_len, num_beams=4)
    labels = ids[:, 1:].clone()
    tf_logits = model.forward_logits(inp["input_ids"], inp["bbox"], inp["pixel_values"], labels,
                                     inp["attention_mask"])
    top2 = logits.topk(2, dim=-1).values
    return {
        "input_ids": inp["input_ids"].numpy(), "bbox": inp["bbox"].numpy(),
        "attention_mask": inp["attention_mask"].numpy(),
        "pixel_checksum": np.array([inp["pixel_values"].double().sum().item()]),
        "memory": mem.numpy().astype(np.float32), "memory_mask": mask.numpy().astype(np.int32),
        "greedy_ids": ids.numpy(), "step0_logits": logits[:, 0].numpy().astype(np.float32),
        "last_logits": logits[:, -1].numpy().astype(np.float32),
        "tf_last_logits": tf_logits[:, -1].numpy().astype(np.float32),
        "beam4_ids": beam.numpy(), "min_top2_margin": np.array([(top2[..., 0] - top2[..., 1]).min().item()]),
    }


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in CASES:
        d = run_case(name)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, {k: v.shape for k, v in d.items()}, "margin", d["min_top2_margin"])


if __name__ == "__main__":
    main()
