"""Probe: torch symmetric memory on this box - rendezvous, peer buffers, barrier, copy-engine peer pulls vs NCCL all-gather.
run: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/symm_probe.py"""
import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
per, F = 1_250_000, 64
try:
    buf = sm.empty(world * per, F, dtype=torch.bfloat16, device=dev)
    hdl = sm.rendezvous(buf, dist.group.WORLD)
    print(rank, 'rendezvous ok', type(hdl).__name__, [m for m in dir(hdl) if not m.startswith('_')][:30], flush=True)
    buf[rank * per:(rank + 1) * per].fill_(rank + 1)
    hdl.barrier(channel=0)
    peers = [hdl.get_buffer(q, (world * per, F), torch.bfloat16) for q in range(world)]
    copy = torch.cuda.Stream()
    def pull():
        ev = torch.cuda.Event(); ev.record()
        copy.wait_event(ev)
        with torch.cuda.stream(copy):
            for i in range(1, world):
                q = (rank + i) % world
                buf[q * per:(q + 1) * per].copy_(peers[q][q * per:(q + 1) * per], non_blocking=True)
        torch.cuda.current_stream().wait_stream(copy)
    for _ in range(3):
        hdl.barrier(channel=0); pull(); hdl.barrier(channel=0)
    torch.cuda.synchronize()
    ok = all(float(buf[q * per].float().mean()) == q + 1 for q in range(world))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        hdl.barrier(channel=0); pull(); hdl.barrier(channel=0)
    b.record(); torch.cuda.synchronize()
    t_pull = a.elapsed_time(b) / 10
    full = torch.empty(world * per, F, dtype=torch.bfloat16, device=dev)
    loc = full[rank * per:(rank + 1) * per]
    for _ in range(3): dist.all_gather_into_tensor(full, loc)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): dist.all_gather_into_tensor(full, loc)
    b.record(); torch.cuda.synchronize()
    t_nccl = a.elapsed_time(b) / 10
    recv = (world - 1) * per * F * 2 / 1e9
    # pull next to a memory-bound kernel
    big = torch.empty(1 << 28, dtype=torch.float32, device=dev)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): big.mul_(1.0001)
    b.record(); torch.cuda.synchronize(); t_k = a.elapsed_time(b) / 10
    a.record()
    for _ in range(10):
        hdl.barrier(channel=0)
        ev = torch.cuda.Event(); ev.record(); copy.wait_event(ev)
        with torch.cuda.stream(copy):
            for i in range(1, world):
                q = (rank + i) % world
                buf[q * per:(q + 1) * per].copy_(peers[q][q * per:(q + 1) * per], non_blocking=True)
        big.mul_(1.0001)
        torch.cuda.current_stream().wait_stream(copy)
        hdl.barrier(channel=0)
    b.record(); torch.cuda.synchronize(); t_both = a.elapsed_time(b) / 10
    print(f'rank {rank}: data ok {ok}; pull+2 barriers {t_pull:.3f} ms ({recv / t_pull * 1e3:.0f} GB/s in); nccl all_gather {t_nccl:.3f} ms '
          f'({recv / t_nccl * 1e3:.0f} GB/s in); 2 GiB RMW kernel alone {t_k:.3f} ms, with pull underneath {t_both:.3f} ms', flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, 'SYMM FAILED', repr(e), flush=True)
dist.destroy_process_group()
