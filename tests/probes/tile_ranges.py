"""Probe (1 GPU): device time of forward tile launches over column-tile sub-ranges, with and without SM reservation."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from synergy_clip_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
D = int(sys.argv[2]) if len(sys.argv) > 2 else 768
dev = torch.device("cuda", 0)
embs = [torch.randn(B, D, device=dev).to(torch.bfloat16) for _ in range(3)]
t3 = torch.full((3,), 2.6592, device=dev)
cfg = ops.TriContrastiveConfig(math="f16")
pb, _, _ = ops._make_problem(embs[0], cfg)
ws = ops._POOL.acquire(pb, dev)
be = ops._BACKEND
be.prologue(ws, *embs)
be.forward_diag(ws, t3)
nt = ws.lay.col_tiles


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


CASES = [(0, nt, 0, True), (0, nt, 0, False)] if os.environ.get("SHORT") == "1" else None
for (lo, hi, sms, stash) in CASES or [(0, nt, 0, True), (0, nt // 2, 0, True), (nt // 2, nt, 0, True), (0, nt // 2, 128, True),
                             (0, nt // 8, 0, True), (0, nt // 8, 128, True), (nt // 8, nt // 4, 0, True),
                             (0, nt // 2, 0, False), (0, nt // 2, 128, False)]:
    prev = be.set_max_sms(sms)
    us = timed(lambda: be.forward_tiles_cols(ws, t3, 7, lo, hi, stash))
    be.set_max_sms(prev)
    tiles = 3 * (ws.lay.row_tiles // 2) * (hi - lo)
    print(f"B={B} D={D} cols[{lo},{hi}) max_sms={sms or 148} stash={int(stash)}: {us:.1f} us  {tiles} cluster tiles  "
          f"{us / max(tiles / ((sms or 148) // 2), 1):.2f} us per round", flush=True)
