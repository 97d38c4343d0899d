"""TEST INFRASTRUCTURE ONLY -- fp64 numpy oracle for the tri-modal contrastive tail.

Independent restatement (no autograd) of what the reference computes in
``Tri_CLIP.forward`` (``/root/reference/model.py:247-272``) through
``clip_loss`` / ``contrastive_loss`` (``model.py:52-58``), together with the
gradients PyTorch's autograd derives for it.  The closed form is SURVEY.md
section 8(a); it is pinned against the reference's own code by
``tests/golden/make_golden.py`` (vectors in ``tests/golden/*.npz``).

Pair / role table (``model.py:255,260,265``):
    IT: rows = image, cols = text
    TA: rows = text,  cols = audio
    AI: rows = audio, cols = image
"""
from __future__ import annotations

import numpy as np

PAIRS = (("IT", 0, 1), ("TA", 1, 2), ("AI", 2, 0))  # (name, row modality, col modality); 0=img 1=txt 2=aud


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what a bf16 I/O tensor holds)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32).reshape(x.shape)


def _logsumexp(a: np.ndarray, axis: int) -> np.ndarray:
    m = a.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def l2_normalise(x: np.ndarray):
    """model.py:248-250 -- x / x.norm(p=2, dim=-1, keepdim=True); no epsilon."""
    n = np.sqrt((x * x).sum(axis=-1, keepdims=True))
    return x / n, n


def pair_forward_backward(xh: np.ndarray, yh: np.ndarray, t: float, g: float):
    """One modality pair on already-normalised fp64 rows.

    Returns loss, d(loss*g)/dxh, d(loss*g)/dyh, d(loss*g)/dt.
    model.py:254-255 (s = t.exp(); logits = xh @ yh.T * s), model.py:55-58
    (clip_loss = (CE(L) + CE(L.T)) / 2 with labels arange(B), mean reduction).
    """
    b = xh.shape[0]
    s = np.exp(t)
    logits = s * (xh @ yh.T)
    lse_r = _logsumexp(logits, axis=1)
    lse_c = _logsumexp(logits, axis=0)
    diag = np.diagonal(logits)
    loss = 0.5 * ((lse_r - diag).mean() + (lse_c - diag).mean())
    p_r = np.exp(logits - lse_r[:, None])
    p_c = np.exp(logits - lse_c[None, :])
    gmat = (p_r + p_c) / (2.0 * b)
    gmat[np.arange(b), np.arange(b)] -= 1.0 / b
    gmat *= g
    dxh = s * (gmat @ yh)
    dyh = s * (gmat.T @ xh)
    dt = float((gmat * logits).sum())
    return float(loss), dxh, dyh, dt


def tri_contrastive(img, txt, aud, t3, g3=(1.0, 1.0, 1.0), want_grads=True):
    """Full tail in fp64.

    img/txt/aud: (B, D) arrays (any float dtype; promoted to fp64 exactly).
    t3: the three learnable log-temperatures (logit_scale_for_IT/TA/AI, model.py:80-82).
    g3: upstream gradients of the three returned losses (alpha/beta/gamma over
        accumulation_steps in main_pretraining.py:166,172).

    Returns dict(loss=(3,), dscale=(3,), dimg, dtxt, daud) -- gradients of
    sum_p g3[p] * loss_p with respect to the *un-normalised* embeddings and t3.
    """
    embs = [np.asarray(e, dtype=np.float64) for e in (img, txt, aud)]
    normed = [l2_normalise(e) for e in embs]
    hats = [n[0] for n in normed]
    norms = [n[1] for n in normed]
    losses = np.zeros(3)
    dscale = np.zeros(3)
    dhat = [np.zeros_like(h) for h in hats]
    for p, (_, r, c) in enumerate(PAIRS):
        loss, dxh, dyh, dt = pair_forward_backward(hats[r], hats[c], float(t3[p]), float(g3[p]))
        losses[p] = loss
        dscale[p] = dt
        if want_grads:
            dhat[r] += dxh
            dhat[c] += dyh
    out = {"loss": losses, "dscale": dscale}
    if want_grads:
        grads = []
        for h, n, d in zip(hats, norms, dhat):
            # backward of x / ||x||: (d - xh * <xh, d>) / ||x||
            grads.append((d - h * (h * d).sum(axis=-1, keepdims=True)) / n)
        out["dimg"], out["dtxt"], out["daud"] = grads
    return out


def synthetic_embeddings(b: int, d: int, seed: int, planted: float = 0.0):
    """Deterministic test inputs (numpy PCG64 stream; no torch dependency).

    planted > 0 mixes a shared per-sample direction into the three modalities so the
    positive-pair cosine is about ``planted`` ("trained-like" distribution).
    """
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((b, d)).astype(np.float32)
    out = []
    for _ in range(3):
        e = rng.standard_normal((b, d)).astype(np.float32)
        if planted > 0.0:
            w = np.float32(np.sqrt(planted / (1.0 - planted)))
            e = e + w * base
        out.append(e)
    return out
