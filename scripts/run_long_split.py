"""One 60 s recording split by chunks across the ranks of a torchrun launch (SURVEY 8e optional row): checks the
result against the unsplit forward on every rank and times both (CUDA events; development probe)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from clearconverse_b200 import SepformerSeparation, sharding, synth, weights
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
sep = SepformerSeparation(weights.random_init_state_dicts(0), device=dev, precision="bf16")
mix = synth.synth_mixture(480000, 3)[0].to(dev)
want = sep.separate_batch(mix[None])[0]
group = dist.group.WORLD if world > 1 else None
got = sharding.separate_long(sep, mix, rank, world, group=group, parts=world)
same = bool(torch.equal(got, want))
def timed(fn, n=10):
    for _ in range(3): fn()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
t_split = timed(lambda: sharding.separate_long(sep, mix, rank, world, group=group, parts=world))
t_one = timed(lambda: sep.separate_batch(mix[None]))
flag = torch.tensor([1 if same else 0], device=dev)
if world > 1: dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "bit_identical_on_every_rank": bool(flag.item()), "ms_split": round(t_split, 3), "ms_one_gpu": round(t_one, 3),
                      "recording_s": 60}))
if world > 1: dist.destroy_process_group()
