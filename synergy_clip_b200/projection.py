"""Projection heads + contrastive tail as one autograd node.

The reference applies three bias-free ``nn.Linear(hidden -> projection_dim)`` heads to the encoders' pooler outputs
(``model.py:76-78``, applied ``:234,237,245``) and hands the results to the loss tail (``:247-272``).  Here the three
projection GEMMs, the tail, and the backward GEMMs of the heads (``dW = dEmb^T . pooled``, ``dpooled = dEmb . W``) all
run on the library's tcgen05 tile kernels (``sclip_gemm_f16``: fp16 operands, fp32 accumulation and output):

* the heads run at tensor-core rate from fp32 / bf16 pooler outputs without an autocast region around the model, and
  the embeddings reach the tail in fp32 (the tail rounds the *normalised* rows to fp16 operands itself), so nothing is
  rounded to bf16 on the way;
* neither the (B, D) embeddings nor their gradients are kept as autograd-visible tensors between the heads and the
  tail: one node saves the pooler outputs and the weights.

What is NOT fused: the L2 normalisation stays the prologue of the tail (``prologue3_kernel``) instead of becoming the
epilogue of the projection GEMM.  A row's norm needs the whole row in one tile, and a 768-wide fp32 accumulator row
does not fit the 512 TMEM columns of an SM -- the two n tiles of a row live in different CTA pairs.

The gradient of the embeddings is scaled by a power of two before it is rounded to fp16 (its entries are of the order
1e-6 .. 1e-3, below the normal range of fp16); the scale is computed on the device, there is no host synchronisation.
PyTorch supplies the casts and that scaling (plumbing); every contraction is a library call.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .ops import (_DEFAULT, TriContrastiveConfig, _POOL, _backward_impl, _check_inputs, _forward_impl, _make_problem,
                  _on_device, _use_p2p, gemm_f16)

__all__ = ["projected_tri_contrastive"]


def _fp16_scale(t: torch.Tensor) -> torch.Tensor:
    """Power of two s such that max |t| * s lands near 2^13 (comfortably inside fp16's normal range)."""
    amax = t.abs().amax().clamp_min(1e-30)
    return torch.exp2(torch.floor(torch.log2(8192.0 / amax)))


class _ProjectedTriContrastive(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pool_img, pool_txt, pool_aud, w_img, w_txt, w_aud, t3, cfg):
        pools = [p.contiguous() for p in (pool_img, pool_txt, pool_aud)]
        weights = [w.contiguous() for w in (w_img, w_txt, w_aud)]
        with _on_device(pools[0]):
            pool16 = [p.detach().to(torch.float16) for p in pools]
            w16 = [w.detach().to(torch.float16) for w in weights]
            embs = [gemm_f16(p, w) for p, w in zip(pool16, w16)]  # (B, D) fp32 = pooled . W^T   (model.py:234,237,245)
            _check_inputs(*embs)
            pb, _, _ = _make_problem(embs[0], cfg)
            ws = _POOL.acquire_sharded(pb, embs[0].device, cfg.process_group, _use_p2p(cfg, embs[0]))
            try:
                loss3 = _forward_impl(ws, *embs, t3, cfg, keep=True)
            except Exception:
                _POOL.release(ws)
                raise
        ctx.ws = ws
        ctx.cfg = cfg
        ctx.embs = embs
        ctx.pool16 = pool16
        ctx.w16 = w16
        ctx.dtypes = ([p.dtype for p in pools], [w.dtype for w in weights])
        ctx.save_for_backward(t3)
        return loss3

    @staticmethod
    def backward(ctx, g3):
        (t3,) = ctx.saved_tensors
        ws, ctx.ws = ctx.ws, None
        if ws is None:
            raise _lib.SclipError("the projected contrastive objective was already back-propagated once")
        embs = ctx.embs
        with _on_device(embs[0]):
            try:
                demb = _backward_impl(ws, *embs, t3, g3.to(torch.float32).contiguous(), ctx.cfg)
            finally:
                _POOL.release(ws)
            *d3, dt3 = demb
            dpools, dweights = [], []
            for d, p16, w16, pdt, wdt in zip(d3, ctx.pool16, ctx.w16, *ctx.dtypes):
                d = d.float()
                scale = _fp16_scale(d)
                d16 = (d * scale).to(torch.float16)
                # dW (D, H) = dEmb^T . pooled: both operands stored [k = sample][.], i.e. MN-major
                dweights.append((gemm_f16(d16, p16, a_mn=True, b_mn=True) / scale).to(wdt))
                # dpooled (B, H) = dEmb . W: W stored [k = D][H] is an MN-major B operand
                dpools.append((gemm_f16(d16, w16, a_mn=False, b_mn=True) / scale).to(pdt))
        return (*dpools, *dweights, dt3, None)


def projected_tri_contrastive(pool_img: torch.Tensor, pool_txt: torch.Tensor, pool_aud: torch.Tensor,
                              w_img: torch.Tensor, w_txt: torch.Tensor, w_aud: torch.Tensor, t_IT: torch.Tensor,
                              t_TA: torch.Tensor, t_AI: torch.Tensor, config: Optional[TriContrastiveConfig] = None):
    """``(IT_loss, TA_loss, AI_loss)`` of ``model.py:234-272`` from the three pooler outputs (B, H_m), the three
    projection weights (D, H_m) (``nn.Linear.weight`` layout, no bias: ``model.py:76-78``) and the three
    ``logit_scale_for_*`` parameters.  Differentiable with respect to all nine."""
    cfg = config or _DEFAULT
    if cfg.math not in ("auto", "f16"):
        raise ValueError("the fused projection heads run with fp16 tensor-core operands (math='f16'); keep nn.Linear in "
                         "front of fused_tri_contrastive for the fp32 parity mode")
    if cfg.math == "auto":
        cfg = TriContrastiveConfig(process_group=cfg.process_group, math="f16", grad_scale=cfg.grad_scale,
                                   grads_fp32=True, overlap=cfg.overlap, comm_sms=cfg.comm_sms, stash=cfg.stash,
                                   transport=cfg.transport, check_status=cfg.check_status)
    for p, w in ((pool_img, w_img), (pool_txt, w_txt), (pool_aud, w_aud)):
        if not p.is_cuda or not w.is_cuda:
            raise _lib.SclipError("projected_tri_contrastive only runs on a CUDA (sm_100a) device and has no CPU fallback")
        if p.dim() != 2 or w.dim() != 2 or p.shape[1] != w.shape[1] or p.shape[1] % 8 or w.shape[0] % 8:
            raise ValueError("pooler outputs (B, H) and projection weights (D, H) with H and D multiples of 8 are expected")
    t3 = torch.stack([t_IT.reshape(()), t_TA.reshape(()), t_AI.reshape(())]).to(device=pool_img.device,
                                                                               dtype=torch.float32)
    loss3 = _ProjectedTriContrastive.apply(pool_img, pool_txt, pool_aud, w_img, w_txt, w_aud, t3, cfg)
    return loss3[0], loss3[1], loss3[2]
