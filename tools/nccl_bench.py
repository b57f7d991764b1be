"""NCCL collectives on the training path's flat buffer sizes: all-reduce vs reduce-scatter + all-gather (in place /
out of place, fp32 / bf16). Launch with torch.distributed.run."""
import os
import torch
import torch.distributed as dist

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl')
n = 654_000_000 // (1024 * world) * 1024 * world
buf = torch.randn(n, device='cuda')
shard = n // world
own = buf[rank * shard:(rank + 1) * shard]
sep = torch.empty(shard, device='cuda')
half = torch.empty(n, device='cuda', dtype=torch.bfloat16)
hown = half[rank * shard:(rank + 1) * shard]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def t(fn, name, reps=5):
    fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device='cuda')
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print('%-44s %.3f ms' % (name, float(ms)), flush=True)


t(lambda: dist.all_reduce(buf, op=dist.ReduceOp.AVG), 'all_reduce fp32 %.2f GB' % (n * 4 / 1e9))
t(lambda: dist.reduce_scatter_tensor(own, buf, op=dist.ReduceOp.AVG), 'reduce_scatter in place')
t(lambda: dist.reduce_scatter_tensor(sep, buf, op=dist.ReduceOp.AVG), 'reduce_scatter out of place')
t(lambda: dist.all_gather_into_tensor(buf, own), 'all_gather fp32 in place')
t(lambda: dist.all_gather_into_tensor(buf, sep), 'all_gather fp32 out of place')
t(lambda: dist.all_gather_into_tensor(half, hown), 'all_gather bf16 in place')
s = torch.zeros(1, device='cuda', dtype=torch.float64)
t(lambda: dist.all_reduce(s), 'all_reduce 8 bytes', reps=20)
dist.destroy_process_group()
