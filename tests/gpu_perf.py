"""Performance probes (run on the GPU box; not collected by pytest).

    python tests/gpu_perf.py step B D [iters]       # fwd+bwd timing with stage breakdown
    python tests/gpu_perf.py gemm M N K [iters]     # the tile GEMM for the four operand-major combinations
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

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


def step(b, d, iters):
    ten = [torch.randn(b, d, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    t3 = torch.full((3,), 2.6592, device="cuda")
    g3 = torch.ones(3, device="cuda")
    cfg = ops.TriContrastiveConfig(math="f16", stash={"1": True, "0": False}.get(os.environ.get("SCLIP_STASH", "auto"), "auto"))
    ms = timed(lambda: ops.forward_backward_raw(*ten, t3, g3, cfg), iters)
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        ops.forward_backward_raw(*ten, t3, g3, cfg)
    host_ms = (time.perf_counter() - t0) / iters * 1e3  # enqueue time only (no sync inside the loop)
    torch.cuda.synchronize()
    print(f"HOST enqueue {host_ms:.3f} ms/step", flush=True)
    marks = []

    def trace(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((name, ev))

    ops._TRACE = trace
    ops.forward_backward_raw(*ten, t3, g3, cfg)
    ops._TRACE = None
    torch.cuda.synchronize()
    stages = {n2: a.elapsed_time(b2) for (n1, a), (n2, b2) in zip(marks[:-1], marks[1:]) if n2 not in ("begin", "backward_begin")}
    tf = 18.0 * b * b * d / (ms * 1e-3) / 1e12
    print(f"STEP B={b} D={d}: {ms:.3f} ms  {tf:.1f} TFLOP/s algorithmic ({100 * tf / 1608.8:.1f}%)  " +
          "  ".join(f"{k}={v:.3f}" for k, v in stages.items()), flush=True)


def gemm(m, n, k, iters):
    a = (torch.randn(m, k, device="cuda") * 0.1).half()
    b = (torch.randn(n, k, device="cuda") * 0.1).half()
    for a_mn in (False, True):
        for b_mn in (False, True):
            aa = a.t().contiguous() if a_mn else a
            bb = b.t().contiguous() if b_mn else b
            ms = timed(lambda: ops.gemm_f16(aa, bb, a_mn=a_mn, b_mn=b_mn), iters)
            print(f"GEMM m={m} n={n} k={k} a_mn={int(a_mn)} b_mn={int(b_mn)}: {ms:.3f} ms  "
                  f"{2.0 * m * n * k / (ms * 1e-3) / 1e12:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    kind = sys.argv[1]
    nums = [int(x) for x in sys.argv[2:]]
    if kind == "step":
        step(nums[0], nums[1], nums[2] if len(nums) > 2 else 10)
    else:
        gemm(nums[0], nums[1], nums[2], nums[3] if len(nums) > 3 else 10)
