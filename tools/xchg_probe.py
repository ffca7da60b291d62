"""Exchange / combine cost of the row-sharded fold step alone (torchrun, N ranks): the all_gather of 13 x 128 B per rank on the
step's side stream and on the default stream, the two device combines, and both together.  Measured on 2 B200s:
all_gather 0.035 / 0.017 ms, combines 0.33 ms, both 0.26 ms per step — the rest of a sharded step's time over its device
work is host-side launch latency and rank skew, not communication."""
import os, sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "tools")
import torch, torch.distributed as dist
import fold_step as F
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = F.ShardedGpuFoldStep(19, rank, world, device=local)
gathered = torch.empty(world * g.n_commits * 128, dtype=torch.uint8, device=dev)
part = g.step(); torch.cuda.synchronize()
def timeit(fn, n=20):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def ag():
    with torch.cuda.stream(g.stream):
        dist.all_gather_into_tensor(gathered, part)
def ag_default():
    dist.all_gather_into_tensor(gathered, part)
def comb():
    g.combine(gathered, world)
def both():
    ag(); comb()
for _ in range(3): both()
res = {"all_gather_side_stream_ms": timeit(ag), "all_gather_default_stream_ms": timeit(ag_default), "combine_ms": timeit(comb), "both_ms": timeit(both)}
if rank == 0: print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()
