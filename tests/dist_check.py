"""torchrun --nproc-per-node N tests/dist_check.py : every rank decodes its shard, the per-step NCCL all-gather
must give every rank the ids of the whole batch == the oracle's ids for the whole batch -- greedy, beam search
(beams of an image stay on its GPU; stock GenerationMixin._beam_search on the oracle is the reference), and the
shard-shape handshake (unequal shards must fail on every rank with a message, not hang)."""
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
print(f"rank {rank}/{world}: exchange mode {eng.dist_mode()} (2 = NVLink peer stores, 1 = ncclAllGather per step): "
      f"ids {tuple(ids.shape)} match oracle for the whole batch: {ok}")
# the other exchange path (the environment switch is read when the communicator is created)
os.environ["MG_DIST"] = "nccl"
eng_n = MGEngine(cfg, oracle.export_state(), device=torch.device("cuda", local))
eng_n.comm_init_from_torch()
del os.environ["MG_DIST"]
ids_n = eng_n.generate_dist(**{k: v[lo:hi] for k, v in inp.items()}, max_length=18).cpu()
okn = eng_n.dist_mode() == 1 and torch.equal(ids_n, ids)
print(f"rank {rank}/{world}: exchange mode {eng_n.dist_mode()}: same ids as the peer-store path: {okn}")
ok = ok and okn
# early EOS: ranks' rows finish on different steps, all ranks must stop on the same step and pad alike
import copy
o3 = copy.deepcopy(oracle)
with torch.no_grad():
    o3.lm_head.weight[1] = sum(o3.lm_head.weight[r] for r in (7, 11, 13)) * 1.2
eng3 = MGEngine(cfg, o3.export_state(), device=torch.device("cuda", local))
eng3.comm_init_from_torch()
inp3 = O.make_inputs(cfg, 3 * world, 14, seed=21)
lo3, hi3 = shard_range(3 * world, world, rank)
ids3 = eng3.generate_dist(**{k: v[lo3:hi3] for k, v in inp3.items()}, max_length=128).cpu()
ref3 = o3.generate_greedy(**inp3, max_length=128)
ok3 = torch.equal(ids3[:, : ref3.shape[1]], ref3) and bool((ids3[:, ref3.shape[1]:] == 0).all()) and eng3.last_steps < 127
print(f"rank {rank}/{world}: early-EOS decode (oracle width {ref3.shape[1]}, steps run {eng3.last_steps}) matches: {ok3}")
ok = ok and ok3
# beam search, 4 beams: early-EOS made likely so that ranks finish their searches on different steps
import copy
o2 = copy.deepcopy(oracle)
with torch.no_grad():
    o2.lm_head.weight[1] = o2.lm_head.weight[1] * 0.0 + o2.lm_head.weight[7] * 1.5 + o2.lm_head.weight[11] * 1.5
eng2 = MGEngine(cfg, o2.export_state(), device=torch.device("cuda", local))
eng2.comm_init_from_torch()
inp2 = O.make_inputs(cfg, n, 14, seed=77)
bids = eng2.generate_dist(**{k: v[lo:hi] for k, v in inp2.items()}, max_length=40, num_beams=4).cpu()
bref = o2.hf_generate(**inp2, max_length=40, num_beams=4)
okb = torch.equal(bids[:, : bref.shape[1]], bref) and bool((bids[:, bref.shape[1]:] == 1).all())
print(f"rank {rank}/{world}: beam-4 ids {tuple(bids.shape)} match stock beam search for the whole batch: {okb}")
ok = ok and okb
# a batch that does not divide by the ranks through the engine's shard / pad / unpad helper
n_odd = 2 * world + 1
inp_odd = O.make_inputs(cfg, n_odd, 12, seed=9)
ids_odd = eng.generate_sharded(**inp_odd, max_length=18).cpu()
ref_odd = oracle.generate_greedy(**inp_odd, max_length=18)
ok_odd = ids_odd.shape[0] == n_odd and torch.equal(ids_odd[:, : ref_odd.shape[1]], ref_odd)
print(f"rank {rank}/{world}: {n_odd} images over {world} ranks (padded shards): match oracle: {ok_odd}")
ok = ok and ok_odd
# unequal shards: every rank gets an error naming the offender
if world > 1:
    from markushgrapher_b200._lib import MgError
    nb = 2 if rank == 0 else 1
    try:
        eng.generate_dist(**{k: v[:nb] for k, v in inp.items()}, max_length=18)
        ok = False
        print(f"rank {rank}: unequal shards were NOT rejected")
    except MgError as e:
        print(f"rank {rank}: unequal shards rejected: {str(e)[:120]}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
