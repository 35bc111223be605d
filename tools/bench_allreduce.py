"""NCCL all-reduce wire time for the gradient volume of cfg2 (591 M fp32 values), idle GPUs: torchrun --nproc-per-node N."""
import os, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 591_399_238
for chunks in (1, 24, 96):
    bufs = [torch.ones(n // chunks, device=dev) for _ in range(chunks)]
    for _ in range(2):
        for b in bufs: dist.all_reduce(b)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        for b in bufs: dist.all_reduce(b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    if dist.get_rank() == 0:
        W = dist.get_world_size()
        print(f"world {W}: {chunks:3d} x {n // chunks * 4 / 1e6:7.1f} MB  {ms:6.2f} ms  algbw {n * 4 / ms / 1e6:6.1f} GB/s  busbw {n * 4 / ms / 1e6 * 2 * (W - 1) / W:6.1f} GB/s", flush=True)
    del bufs
dist.destroy_process_group()
