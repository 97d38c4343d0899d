"""TEST INFRASTRUCTURE ONLY -- the fp64 oracle evaluated block by block, for batches whose B x B matrices do not fit.

Same mathematics as ``oracle/closed_form.py`` (the closed form of SURVEY.md section 8(a) for
``/root/reference/model.py:247-272`` + ``model.py:52-58``), written with torch so that it can run in fp64 on the GPU
that hosts the test: the three similarity matrices are produced ``block`` rows at a time and never held whole, the
column log-sum-exps are merged online, and gradients are produced only for a sample of rows (a gradient row of a
modality needs one row of G of its row-role pair and one column of G of its column-role pair, plus the complete
row / column statistics).  ``tests/test_oracle.py`` pins it against ``closed_form`` -- which is itself pinned against
the outputs of the unmodified reference -- on batches small enough for both.
"""
from __future__ import annotations

import torch

PAIRS = ((0, 1), (1, 2), (2, 0))  # (row modality, column modality) of IT, TA, AI (model.py:255,260,265)


def tri_contrastive_rows(img, txt, aud, t3, g3, rows, block: int = 2048, device=None):
    """Losses (3,), dlogit_scale (3,) and the gradient rows ``rows`` of the three un-normalised embeddings, all fp64.

    img / txt / aud: (B, D) tensors or arrays (any float dtype; promoted exactly).  t3: the three log-temperatures;
    g3: upstream gradients of the three losses.  Returns a dict of CPU fp64 tensors
    ``loss, dscale, dimg_rows, dtxt_rows, daud_rows`` (the last three of shape (len(rows), D)).
    """
    dev = torch.device(device) if device is not None else (img.device if isinstance(img, torch.Tensor) else "cpu")
    x = [torch.as_tensor(e).to(device=dev, dtype=torch.float64) for e in (img, txt, aud)]
    b, d = x[0].shape
    rows = torch.as_tensor(rows, dtype=torch.long, device=dev)
    norms = [e.norm(dim=-1, keepdim=True) for e in x]
    hats = [e / n for e, n in zip(x, norms)]  # model.py:248-250, no epsilon
    loss = torch.zeros(3, dtype=torch.float64)
    dscale = torch.zeros(3, dtype=torch.float64)
    dhat_rows = [torch.zeros((rows.numel(), d), dtype=torch.float64, device=dev) for _ in range(3)]
    for p, (r, c) in enumerate(PAIRS):
        s = float(torch.exp(torch.tensor(float(t3[p]), dtype=torch.float64)))
        g = float(g3[p])
        xr, yc = hats[r], hats[c]
        lse_row = torch.empty(b, dtype=torch.float64, device=dev)
        diag = torch.empty(b, dtype=torch.float64, device=dev)
        cmax = torch.full((b,), -float("inf"), dtype=torch.float64, device=dev)
        csum = torch.zeros(b, dtype=torch.float64, device=dev)
        for lo in range(0, b, block):  # pass 1: row statistics, online column statistics, positive pairs
            hi = min(lo + block, b)
            logits = s * (xr[lo:hi] @ yc.T)
            lse_row[lo:hi] = torch.logsumexp(logits, dim=1)
            diag[lo:hi] = logits[torch.arange(hi - lo, device=dev), torch.arange(lo, hi, device=dev)]
            m = torch.maximum(cmax, logits.max(dim=0).values)
            csum = csum * torch.exp(cmax - m) + torch.exp(logits - m[None, :]).sum(dim=0)
            cmax = m
        lse_col = cmax + torch.log(csum)
        loss[p] = (0.5 * ((lse_row - diag).mean() + (lse_col - diag).mean())).cpu()
        dt = torch.zeros((), dtype=torch.float64, device=dev)
        for lo in range(0, b, block):  # pass 2: G = g ((P_row + P_col) / 2B - I / B) block by block
            hi = min(lo + block, b)
            logits = s * (xr[lo:hi] @ yc.T)
            gmat = (torch.exp(logits - lse_row[lo:hi, None]) + torch.exp(logits - lse_col[None, :])) * (g / (2.0 * b))
            gmat[torch.arange(hi - lo, device=dev), torch.arange(lo, hi, device=dev)] -= g / b
            dt += (gmat * logits).sum()
            inblk = (rows >= lo) & (rows < hi)
            if bool(inblk.any()):  # row role: rows of G of the sampled rows that live in this block
                dhat_rows[r][inblk] += s * (gmat[rows[inblk] - lo] @ yc)
            dhat_rows[c] += s * (gmat[:, rows].T @ xr[lo:hi])  # column role: columns of G at the sampled indices
        dscale[p] = dt.cpu()
    out = {"loss": loss, "dscale": dscale}
    for m, key in enumerate(("dimg_rows", "dtxt_rows", "daud_rows")):
        h = hats[m][rows]
        dd = dhat_rows[m]
        out[key] = ((dd - h * (h * dd).sum(dim=-1, keepdim=True)) / norms[m][rows]).cpu()
    return out
