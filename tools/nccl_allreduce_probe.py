"""How fast is the fp32 mean all-reduce of dW_lm_head ([151936, 3584] = 2.18 GB) on this box, as NCCL is configured by the
environment? torchrun --nproc-per-node N tools/nccl_allreduce_probe.py [tag]"""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
x = torch.ones(151936, 3584, dtype=torch.float32, device=dev)
for _ in range(3):
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
torch.cuda.synchronize()
dist.barrier()
times = []
for _ in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
t = torch.tensor([sorted(times)[len(times) // 2]], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    gb = x.numel() * 4 / 1e9
    print(f"{sys.argv[1] if len(sys.argv) > 1 else 'default':28s} all-reduce {gb:.2f} GB fp32 on {world} GPUs: {ms:7.3f} ms  algbw {gb / ms * 1e3:6.1f} GB/s  "
          f"busbw {gb / ms * 1e3 * 2 * (world - 1) / world:6.1f} GB/s", flush=True)
dist.destroy_process_group()
