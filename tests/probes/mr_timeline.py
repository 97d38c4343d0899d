"""Probe (GPU box, torchrun): per-stage timeline of the sharded step on every rank, for a few configurations.

    torchrun --nproc-per-node N tests/probes/mr_timeline.py [B] [D]
"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from synergy_clip_b200 import ops  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
os.environ.setdefault("NCCL_MAX_CTAS", "16")
dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 768
bl = B // world
g = torch.Generator(device=dev).manual_seed(1234 + rank)
embs = [torch.randn(bl, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(3)]
t3 = torch.full((3,), 2.6592, device=dev)
g3 = torch.ones(3, device=dev)


def run(name, **kw):
    cfg = ops.TriContrastiveConfig(process_group=dist.group.WORLD, math="f16", **kw)
    fn = lambda: ops.forward_backward_raw(*embs, t3, g3, cfg)  # noqa: E731
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    host_ms = (time.perf_counter() - t0) / 20 * 1e3
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # one traced step: marks on the compute stream
    marks = []

    def trace(nm):
        ev = torch.cuda.Event(enable_timing=True); ev.record(); marks.append((nm, ev))

    dist.barrier(); torch.cuda.synchronize()
    ops._TRACE = trace
    fn()
    ops._TRACE = None
    torch.cuda.synchronize()
    tl = " ".join(f"{n}={marks[0][1].elapsed_time(ev):.3f}" for n, ev in marks[1:])
    extra = getattr(ops, "_LAST_COMM_EVENTS", None)
    comm = ""
    if extra:
        comm = " | comm: " + " ".join(f"{n}={marks[0][1].elapsed_time(ev):.3f}" for n, ev in extra)
    if rank in (0, world - 1):
        print(f"[{name}] rank {rank}: step(max over ranks)={ms.item():.3f} ms host-enqueue={host_ms:.3f} ms | {tl}{comm}", flush=True)
    dist.barrier()


variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["default", "no-overlap", "nccl"]
for v in variants:
    if v == "default":
        run("p2p " + v)
    elif v == "no-overlap":
        run("p2p no-overlap", overlap=False)
    elif v == "nccl":
        run("nccl", transport="nccl")
    elif v == "ce":
        run("p2p copy-engine push", push="ce")
    elif v == "sm":
        run("p2p kernel push", push="sm")
    elif v.startswith("comm_sms="):
        run("p2p " + v, comm_sms=int(v.split("=")[1]))
dist.destroy_process_group()
