"""GPU: the projection heads fused with the tail (`synergy_clip_b200.projection`) against the reference's
expression -- `nn.Linear(bias=False)` x 3 (model.py:76-78, 234-245) followed by the loss tail (model.py:247-272) --
evaluated in fp64: losses, gradients of the pooler outputs, of the projection weights and of the log-temperatures."""
import numpy as np
import pytest
import torch

from oracle import closed_form
from tests import golden_util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,hidden,dim,dtype_name", [
    (300, (96, 128, 64), 256, "float32"),
    (520, (768, 768, 768), 768, "bfloat16"),   # Base heads: hidden 768 -> projection 768 (config.py:13,39,89,113)
])
def test_projected_tail_matches_linear_plus_oracle(b, hidden, dim, dtype_name):
    from synergy_clip_b200.projection import projected_tri_contrastive

    dtype = getattr(torch, dtype_name)
    g = torch.Generator().manual_seed(3)
    base = torch.randn(b, 32, generator=g)  # a shared latent so that positives exist
    pools, weights = [], []
    for h in hidden:
        mix = torch.randn(32, h, generator=g)
        pools.append((torch.tanh(base @ mix * 0.3 + 0.3 * torch.randn(b, h, generator=g))).to(dtype))
        weights.append((torch.randn(dim, h, generator=g) / h ** 0.5).to(dtype))
    t3, w3 = (2.6592, 2.9, 2.2), (0.3, 0.7, 1.1)
    # reference expression in fp64 on the same (rounded) inputs
    p64 = [p.double().numpy() for p in pools]
    w64 = [w.double().numpy() for w in weights]
    embs = [p @ w.T for p, w in zip(p64, w64)]
    want = closed_form.tri_contrastive(*embs, t3, w3)
    demb = [want["dimg"], want["dtxt"], want["daud"]]
    want_dw = [d.T @ p for d, p in zip(demb, p64)]
    want_dp = [d @ w for d, w in zip(demb, w64)]

    pl = [p.cuda().requires_grad_(True) for p in pools]
    wl = [w.cuda().requires_grad_(True) for w in weights]
    ts = [torch.tensor(t, device="cuda", requires_grad=True) for t in t3]
    losses = projected_tri_contrastive(*pl, *wl, *ts)
    sum(w * l for w, l in zip(w3, losses)).backward()
    got_loss = np.array([l.item() for l in losses])
    # fp16 operands in the heads: 1e-3 (the bf16 contract); gradients read back from bf16 leaves carry bf16 rounding
    tol, gtol = 1e-3, (1e-3 if dtype == torch.float32 else 4e-3)
    assert np.max(np.abs(got_loss - want["loss"]) / want["loss"]) < tol
    for leaf, ref in zip(pl, want_dp):
        assert leaf.grad.dtype == dtype
        assert golden_util.rel(leaf.grad.double().cpu().numpy(), ref) < gtol
    for leaf, ref in zip(wl, want_dw):
        assert golden_util.rel(leaf.grad.double().cpu().numpy(), ref) < gtol
    got_dt = np.array([t.grad.item() for t in ts])
    assert np.max(np.abs(got_dt - want["dscale"])) / np.max(np.abs(want["dscale"])) < tol
