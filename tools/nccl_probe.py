"""all-reduce time of one flat fp32 gradient buffer (base model: 163 M floats = 651 MiB) under torchrun; prints ms and
bus bandwidth.  Run under different NCCL_* environment settings to see what the fabric gives:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/nccl_probe.py [tag]"""
import os
import sys

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
for n in (170_655_744, 16 * 2 ** 20):
    buf = torch.randn(n, device="cuda")
    for _ in range(3):
        dist.all_reduce(buf, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    s.record()
    for _ in range(5):
        dist.all_reduce(buf, op=dist.ReduceOp.AVG)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    if dist.get_rank() == 0:
        gb = n * 4 / 1e9
        print(f"[{tag}] world={world} {gb:.3f} GB: {ms:.3f} ms  algbw {gb / ms * 1e3:.1f} GB/s  busbw {gb / ms * 1e3 * 2 * (world - 1) / world:.1f} GB/s", flush=True)
if dist.get_rank() == 0 and world >= 2:
    a = torch.randn(2 ** 28, device="cuda:0")
    b = torch.empty(2 ** 28, device="cuda:1")
    for _ in range(2):
        b.copy_(a)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        b.copy_(a)
    e.record()
    torch.cuda.synchronize()
    print(f"[{tag}] peer copy 1 GiB: {2 ** 30 / (s.elapsed_time(e) / 5) / 1e6:.1f} GB/s", flush=True)
dist.barrier()
dist.destroy_process_group()
