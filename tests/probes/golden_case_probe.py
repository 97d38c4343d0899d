"""Probe (GPU box): every error component of one golden case, stash and recompute backward.

    python tests/probes/golden_case_probe.py b300x512_ln100
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_util  # noqa: E402
from synergy_clip_b200 import ops  # noqa: E402

for name in sys.argv[1:]:
    meta, embs, data = golden_util.load_case(name)
    dtype = torch.bfloat16 if meta["bf16_inputs"] else torch.float32
    ten = [torch.from_numpy(np.ascontiguousarray(e)).cuda().to(dtype) for e in embs]
    t3 = torch.tensor(meta["t3"], dtype=torch.float32, device="cuda")
    g3 = torch.tensor(meta["g3"], dtype=torch.float32, device="cuda")
    print(name, "t3", meta["t3"], "g3", meta["g3"], "golden dscale", data["dscale"], "loss", data["loss"])
    for st in (True, False):
        cfg = ops.TriContrastiveConfig(math="f16", grads_fp32=True, stash=st)
        loss3, dimg, dtxt, daud, dt3 = ops.forward_backward_raw(*ten, t3, g3, cfg)
        torch.cuda.synchronize()
        res = {"loss": loss3.double().cpu().numpy(), "dscale": dt3.double().cpu().numpy(), "dimg": dimg.double().cpu().numpy(),
               "dtxt": dtxt.double().cpu().numpy(), "daud": daud.double().cpu().numpy()}
        errs = golden_util.golden_errors(meta, data, res)
        print("  stash" if st else "  recompute", "dscale got", res["dscale"])
        print("   ", {k: float(f"{v:.3g}") for k, v in errs.items()})
