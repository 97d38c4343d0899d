"""Probe (GPU box, torchrun): does torch symmetric memory rendezvous work here, and how fast are peer pulls?"""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 64 << 20
t = symm.empty(n, dtype=torch.uint8, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok", hdl.rank, hdl.world_size, [hex(p) for p in hdl.buffer_ptrs][:2], t.data_ptr() % 256, flush=True)
t.fill_(rank + 1)
hdl.barrier(0)
peer = (rank + 1) % world
src = hdl.get_buffer(peer, (n,), torch.uint8, 0)
dst = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(3):
    dst.copy_(src)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    dst.copy_(src)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(rank, "peer copy_ (CE) 64 MiB:", ms, "ms", n / ms / 1e6, "GB/s", "value", int(dst[0]), flush=True)
# SM-based pull: elementwise kernel reading peer memory
a = src.view(torch.float32)
out = torch.empty_like(a)
for _ in range(3):
    torch.add(a, 0.0, out=out)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    torch.add(a, 0.0, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(rank, "peer SM read 64 MiB:", ms, "ms", n / ms / 1e6, "GB/s", flush=True)
# barrier latency
hdl.barrier(0)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    hdl.barrier(0)
e1.record()
torch.cuda.synchronize()
print(rank, "barrier:", e0.elapsed_time(e1) / 20 * 1e3, "us", flush=True)
dist.barrier()
dist.destroy_process_group()
