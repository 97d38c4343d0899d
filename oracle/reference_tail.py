"""TEST INFRASTRUCTURE ONLY -- PyTorch-CPU port of the reference loss tail.

Same operator sequence the reference executes (``/root/reference/model.py``):
``norm``/``div`` (:248-250), ``exp`` of the three log-temperatures and scaled
``matmul`` (:254-265), ``F.cross_entropy`` on the logits and their transpose
with ``arange`` labels (:52-58, :269-271), gradients by autograd.  It exists so
that (a) the numpy closed form has a second, autograd-derived opinion on any
machine (the reference tree itself is not present on the GPU box) and (b)
``bench.py`` can time "the reference's PyTorch CPU loss path" on the GPU box's
host cores (``cpu_baseline.kind == "port"``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _symmetric_ce(similarity: torch.Tensor) -> torch.Tensor:
    # model.py:52-58 -- mean CE over rows plus mean CE over columns, halved
    labels = torch.arange(similarity.shape[0], device=similarity.device)
    rows = F.cross_entropy(similarity, labels)
    cols = F.cross_entropy(similarity.t(), labels)
    return (rows + cols) / 2.0


def tail_losses(img, txt, aud, t_it, t_ta, t_ai):
    """model.py:247-272 on raw (B, D) projection outputs -> (IT, TA, AI) losses."""
    img = img / img.norm(p=2, dim=-1, keepdim=True)
    txt = txt / txt.norm(p=2, dim=-1, keepdim=True)
    aud = aud / aud.norm(p=2, dim=-1, keepdim=True)
    sim_it = torch.matmul(img, txt.t()) * t_it.exp()
    sim_ta = torch.matmul(txt, aud.t()) * t_ta.exp()
    sim_ai = torch.matmul(aud, img.t()) * t_ai.exp()
    return _symmetric_ce(sim_it), _symmetric_ce(sim_ta), _symmetric_ce(sim_ai)


def tail_forward_backward(img, txt, aud, t3, g3=(1.0, 1.0, 1.0), dtype=torch.float64):
    """Run the port with autograd; returns the same dict layout as closed_form.tri_contrastive."""
    leaves = [torch.as_tensor(e).detach().to(dtype).clone().requires_grad_(True) for e in (img, txt, aud)]
    scales = [torch.tensor(float(t), dtype=dtype, requires_grad=True) for t in t3]
    losses = tail_losses(*leaves, *scales)
    total = sum(float(g) * l for g, l in zip(g3, losses))
    total.backward()
    return {
        "loss": torch.stack([l.detach() for l in losses]).numpy(),
        "dscale": torch.stack([s.grad for s in scales]).numpy(),
        "dimg": leaves[0].grad.numpy(),
        "dtxt": leaves[1].grad.numpy(),
        "daud": leaves[2].grad.numpy(),
    }


def timed_step(img, txt, aud, scales):
    """One fwd+bwd of the tail exactly as main_pretraining.py:166-173 drives it (unit weights)."""
    for p in (img, txt, aud, *scales):
        p.grad = None
    it, ta, ai = tail_losses(img, txt, aud, *scales)
    (it + ta + ai).backward()
    return float(it.detach()) + float(ta.detach()) + float(ai.detach())
