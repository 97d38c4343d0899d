"""Small stash-mode and recompute-mode steps for compute-sanitizer (GPU box):
    compute-sanitizer --tool memcheck python tests/probes/sanitize_case.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from synergy_clip_b200 import ops  # noqa: E402

for (b, d, dt, math) in [(300, 768, torch.bfloat16, "f16"), (520, 512, torch.bfloat16, "f16"), (200, 256, torch.float32, "f16x3")]:
    ten = [torch.randn(b, d, device="cuda").to(dt) for _ in range(3)]
    t3 = torch.tensor([2.6592, 2.9, 4.7], device="cuda")  # the last pair takes the s >= 44 kernel
    g3 = torch.tensor([1.0, 0.5, 0.25], device="cuda")
    out = ops.forward_backward_raw(*ten, t3, g3, ops.TriContrastiveConfig(math=math, grads_fp32=True))
    torch.cuda.synchronize()
    print(b, d, math, [round(float(x), 4) for x in out[0]], float(out[1].abs().sum()), flush=True)
a = torch.randn(37, 768, device="cuda")
bb = torch.randn(10, 768, device="cuda")
print(ops.cosine_logits(a, bb, torch.tensor(2.0, device="cuda")).sum().item())
