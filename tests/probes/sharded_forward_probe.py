"""Probe (ONE GPU): the forward tile launch of one rank of an 8-rank problem (rows_local 4096 of 32768, dim 768) in
isolation -- no peers, no pushes, the "landed" flags preset -- for the launch variants the sharded path can choose from.
Separates what the kernel costs on this shape from what the concurrent exchange costs."""
import ctypes
import os
import sys
from ctypes import byref

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from synergy_clip_b200 import _lib  # noqa: E402

lib = _lib.load()
world, rank, bl, d = 8, 3, 4096, 768
pb = _lib.Problem(bl, bl * world, rank * bl, d, _lib.SCLIP_BF16, _lib.MATH_F16, world, 0)
lay = _lib.plan(pb)
ws = torch.empty(int(lay.total_bytes) + 256, dtype=torch.uint8, device="cuda")
ws = ws[(-ws.data_ptr()) % 256:]
ws[int(lay.sync):int(lay.sync) + 256].zero_()
p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
g = torch.Generator(device="cuda").manual_seed(1)
# fill the whole xhat / diag_all with plausible data: normalised random rows
xh = torch.randn(3, bl * world, d, device="cuda", generator=g)
xh = (xh / xh.norm(dim=-1, keepdim=True)).half()
ws[int(lay.xhat):int(lay.xhat) + xh.numel() * 2].view(torch.float16).copy_(xh.view(-1))
t3 = torch.full((3,), 2.6592, device="cuda")
dg = (xh[[0, 1, 2]].float() * xh[[1, 2, 0]].float()).sum(-1) * float(torch.exp(t3[0]))
ws[int(lay.diag_all):int(lay.diag_all) + dg.numel() * 4].view(torch.float32).copy_(dg.view(-1))
ws[int(lay.sync):int(lay.sync) + 64].view(torch.int32).fill_(1 << 20)  # every shard has "landed"
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(name, fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"FWDPROBE {name}: {ms:.3f} ms  ({6.0 * bl * bl * world * d / (ms * 1e-3) / 1e12:.0f} TFLOP/s)", flush=True)


ct = lay.col_tiles
for stash in (1, 0):
    for sms in (0, 128):
        timed(f"stash={stash} plain order max_sms={sms}",
              lambda: _lib.check(lib.sclip_forward_tiles_cols(byref(pb), p(ws), p(t3), 7, 0, ct, stash, sms, 0, st), "fwd"))
        timed(f"stash={stash} wave order (WAIT_PEERS) max_sms={sms}",
              lambda: _lib.check(lib.sclip_forward_tiles_cols(byref(pb), p(ws), p(t3), 7, 0, 0, stash | 4, sms, 1, st), "fwd"))
