"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference ``model.py``.

Only usable in the build container (``/root/reference`` does not exist on the
GPU box).  ``model.py:22-23`` imports ``pytorch_msssim`` and ``piqa`` which are
absent here and irrelevant to the contrastive tail, so two empty stub modules
are pre-seeded in ``sys.modules``; nothing else is patched.  Used by
``tests/golden/make_golden.py`` to produce the committed golden vectors and by
``tests/test_oracle.py`` (skipped when the tree is absent).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SCLIP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


def load_reference_model():
    """Return the reference's ``model`` module (cached under ``_sclip_reference_model``)."""
    cached = sys.modules.get("_sclip_reference_model")
    if cached is not None:
        return cached
    if not available():
        raise FileNotFoundError(f"reference tree not present at {REFERENCE_ROOT}")
    for name in ("pytorch_msssim", "piqa"):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            for attr in ("ssim", "ms_ssim", "SSIM", "MS_SSIM"):
                setattr(stub, attr, None)
            sys.modules[name] = stub
    import importlib.util

    spec = importlib.util.spec_from_file_location("_sclip_reference_model", os.path.join(REFERENCE_ROOT, "model.py"))
    module = importlib.util.module_from_spec(spec)
    sys.modules["_sclip_reference_model"] = module
    spec.loader.exec_module(module)
    return module


def reference_tail(img, txt, aud, t3, g3=(1.0, 1.0, 1.0), dtype=None):
    """Drive reference ``clip_loss`` through the statement sequence of model.py:247-272."""
    import torch

    ref = load_reference_model()
    dtype = dtype or torch.float64
    leaves = [torch.as_tensor(e).detach().to(dtype).clone().requires_grad_(True) for e in (img, txt, aud)]
    scales = [torch.tensor(float(t), dtype=dtype, requires_grad=True) for t in t3]
    i, t, a = leaves
    i_n = i / i.norm(p=2, dim=-1, keepdim=True)
    t_n = t / t.norm(p=2, dim=-1, keepdim=True)
    a_n = a / a.norm(p=2, dim=-1, keepdim=True)
    l_it = torch.matmul(i_n, t_n.t()) * scales[0].exp()
    l_ta = torch.matmul(t_n, a_n.t()) * scales[1].exp()
    l_ai = torch.matmul(a_n, i_n.t()) * scales[2].exp()
    losses = (ref.clip_loss(l_it), ref.clip_loss(l_ta), ref.clip_loss(l_ai))
    sum(float(g) * l for g, l in zip(g3, losses)).backward()
    return {
        "loss": torch.stack([l.detach() for l in losses]).numpy(),
        "dscale": torch.stack([s.grad for s in scales]).numpy(),
        "dimg": leaves[0].grad.numpy(),
        "dtxt": leaves[1].grad.numpy(),
        "daud": leaves[2].grad.numpy(),
    }
