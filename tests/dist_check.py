"""torchrun --nproc-per-node N tests/dist_check.py : every rank decodes its shard, the per-step NCCL all-gather
must give every rank the ids of the whole batch == the oracle's ids for the whole batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from markushgrapher_b200.engine import MGEngine
from markushgrapher_b200.parallel import shard_range
from oracle import mg_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = O.MGConfig.tiny()
oracle = O.build(cfg, seed=0)
eng = MGEngine(cfg, oracle.export_state(), device=torch.device("cuda", local))
eng.comm_init_from_torch()
n = 2 * world
inp = O.make_inputs(cfg, n, 12, seed=8)
lo, hi = shard_range(n, world, rank)
ids = eng.generate_dist(**{k: v[lo:hi] for k, v in inp.items()}, max_length=18).cpu()
ref = oracle.generate_greedy(**inp, max_length=18)
ok = torch.equal(ids[:, : ref.shape[1]], ref)
print(f"rank {rank}/{world}: all-gathered ids {tuple(ids.shape)} match oracle for the whole batch: {ok}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
