"""Probe (1 GPU): the reference's own loss tail (oracle/reference_tail.py restates model.py:247-272 verbatim) run by
PyTorch eager on the same B200 -- "the existing kernel to beat on the same box"  -- next to the fused op."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_tail  # noqa: E402
from synergy_clip_b200 import ops  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for (b, d) in [(8192, 512), (16384, 768)]:
    g = torch.Generator(device="cuda").manual_seed(1234)
    base = [torch.randn(b, d, device="cuda", generator=g) for _ in range(3)]
    for dt in (torch.float32, torch.bfloat16):
        leaves = [x.to(dt).requires_grad_(True) for x in base]
        scales = [torch.tensor(2.6592, device="cuda", requires_grad=True) for _ in range(3)]

        def eager():
            for p in (*leaves, *scales):
                p.grad = None
            it, ta, ai = reference_tail.tail_losses(*leaves, *scales)
            (it + ta + ai).backward()

        try:
            ms = timed(eager, 5)
            print(f"eager reference tail B={b} D={d} {str(dt).split('.')[-1]}: {ms:.2f} ms/step "
                  f"({b / ms * 1e3:.3e} samples/s, peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB)", flush=True)
        except torch.OutOfMemoryError:
            print(f"eager reference tail B={b} D={d} {dt}: out of memory", flush=True)
        torch.cuda.reset_peak_memory_stats()
    ten = [x.bfloat16() for x in base]
    t3 = torch.full((3,), 2.6592, device="cuda")
    g3 = torch.ones(3, device="cuda")
    ms = timed(lambda: ops.forward_backward_raw(*ten, t3, g3, ops.TriContrastiveConfig(math="f16")), 10)
    print(f"fused op              B={b} D={d} bfloat16: {ms:.2f} ms/step ({b / ms * 1e3:.3e} samples/s)", flush=True)
