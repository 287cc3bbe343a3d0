import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
def attempt(tag, rows, feat, keep):
    try:
        t = sm.empty(rows, feat, dtype=torch.bfloat16, device=dev)
        hdl = sm.rendezvous(t, dist.group.WORLD)
        print(f'rank {rank} {tag}: rows {rows} ok buffer_size {hdl.buffer_size} offset {getattr(hdl, "offset", None)}', flush=True)
        keep.append((t, hdl))
    except Exception as e:
        print(f'rank {rank} {tag}: rows {rows} FAILED {e!r}'[:300], flush=True)
keep = []
for i in range(3): attempt(f'small{i}', 78136 * world, 64, keep)
keep.clear(); torch.cuda.empty_cache(); dist.barrier()
big = torch.empty((rank + 1) * (1 << 28), dtype=torch.uint8, device=dev)      # rank-dependent ordinary allocations in between
del big
for i in range(3): attempt(f'big{i}', 1283712 * world, 64, keep)
dist.barrier()
dist.destroy_process_group()
