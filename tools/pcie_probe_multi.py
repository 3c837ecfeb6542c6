"""Pinned-memory PCIe bandwidth with EVERY rank copying at once (launch with torchrun): pure
copies, no kernels, no Python in the loop -- the ceiling the end-to-end pipeline runs under when
N GPUs of one box move pages at the same time.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/pcie_probe_multi.py

Every rank: 768 MB pinned in + 896 MB pinned out (the sizes of one 256-page step), H2D alone,
D2H alone, both at once; 5 repetitions between barriers; rank 0 prints per-rank and aggregate
GB/s, plus the time of a host memset over the pinned buffer with all ranks at once (host memory
bandwidth available to one rank while the others do the same).
"""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))

n_in, n_out = 768 * 1024 * 1024, 896 * 1024 * 1024
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
h_out.fill_(2)
d_a = torch.empty(n_in, dtype=torch.uint8, device='cuda')
d_b = torch.empty(n_out, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(h2d, d2h, reps=5):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / reps
    barrier()
    return t


def gather(value):
    v = torch.tensor([value], dtype=torch.float64, device='cuda')
    if world > 1:
        out = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(out, v)
        return [float(o[0]) for o in out]
    return [value]


for name, a, b in (('H2D', True, False), ('D2H', False, True), ('both', True, True)):
    run(a, b, 1)
    t = run(a, b)
    times = gather(t)
    if rank == 0:
        moved = (n_in if a else 0) + (n_out if b else 0)
        per_rank = [moved / x / 1e9 for x in times]
        print(f'{world} ranks {name:5s}: per rank {min(per_rank):5.1f} .. {max(per_rank):5.1f} GB/s '
              f'(sum of directions), aggregate {sum(per_rank):6.1f} GB/s, slowest step {max(times) * 1e3:.1f} ms')

# host memory bandwidth: every rank rewrites its pinned input buffer at the same time
barrier()
t0 = time.perf_counter()
for _ in range(3):
    h_in.fill_(3)
t = (time.perf_counter() - t0) / 3
times = gather(t)
if rank == 0:
    print(f'{world} ranks host fill of 768 MB: {min(times) * 1e3:.0f} .. {max(times) * 1e3:.0f} ms per rank '
          f'({n_in / max(times) / 1e9:.1f} GB/s for the slowest), cpus {os.cpu_count()}')
if world > 1:
    dist.destroy_process_group()
