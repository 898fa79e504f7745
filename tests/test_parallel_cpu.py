"""CPU / gloo / world_size 2: the image-sharding + token-id all-gather host logic of the N>1 path."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from markushgrapher_b200.parallel import gather_token_ids, generate_sharded, shard_range


def test_shard_range_covers_everything():
    for n in (1, 7, 32, 1000, 1024):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = shard_range(n, world, r)
                got += list(range(lo, hi))
            assert got == list(range(n))


def test_pad_and_unpad_equal_shards():
    """the library's sharded generate needs equal shards: pad with the last image, drop the padding rows afterwards"""
    from markushgrapher_b200.parallel import pad_shard, shard_rows, unpad_gathered

    for n, world in ((7, 2), (9, 4), (8, 8), (5, 3)):
        rows = shard_rows(n, world)
        full = torch.arange(n)[:, None] * 10 + torch.arange(3)[None, :]            # "ids" of image i: 10 i + t
        gathered = []
        for r in range(world):
            lo, hi = shard_range(n, world, r)
            local = pad_shard({"input_ids": full[lo:hi], "attention_mask": None}, rows)
            assert local["input_ids"].shape[0] == rows and local["attention_mask"] is None
            assert torch.equal(local["input_ids"][hi - lo:], full[hi - 1: hi].expand(rows - (hi - lo), 3))
            gathered.append(local["input_ids"])
        assert torch.equal(unpad_gathered(torch.cat(gathered), n, world), full)
    with pytest.raises(ValueError):
        pad_shard({"input_ids": torch.zeros(0, 3)}, 1)


def _fake_generate(input_ids, max_length, **kw):
    # deterministic stand-in for the engine: "decodes" row i to i, i+1, ... and stops rank-dependently
    b = input_ids.shape[0]
    t = 4 + int(input_ids[0, 0]) % 3
    return input_ids[:, :1] + torch.arange(t)[None, :].expand(b, t)


def _worker(rank, world, port, n_total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = {"input_ids": torch.arange(n_total)[:, None].repeat(1, 3)}
        out = generate_sharded(_fake_generate, full, max_length=8)
        assert out.shape[0] == n_total
        assert torch.equal(out[:, 0], torch.arange(n_total))        # global image order restored
        lo, hi = shard_range(n_total, world, rank)
        mine = _fake_generate(full["input_ids"][lo:hi], 8)
        assert torch.equal(out[lo:hi, : mine.shape[1]], mine)
        assert (out[lo:hi, mine.shape[1]:] == 0).all()              # padded to the longest rank
        ids = gather_token_ids(torch.full((1, 2), rank, dtype=torch.int64), world)
        assert ids[:, 0].tolist() == list(range(world))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 8])
def test_gloo_world2_gather(n_total):
    port = 29500 + os.getpid() % 1000 + n_total
    mp.spawn(_worker, args=(2, port, n_total), nprocs=2, join=True)
