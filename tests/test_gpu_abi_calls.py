"""GPU: the two-call C ABI (`sclip_forward` / `sclip_backward`) bound with ctypes exactly as INTEGRATION.md section 3
shows it to a maintainer -- no `synergy_clip_b200.ops` in between."""
import ctypes
from ctypes import byref, c_int, c_void_p

import numpy as np
import pytest
import torch

from oracle import closed_form
from tests import golden_util

pytestmark = pytest.mark.gpu


def _bind():
    from synergy_clip_b200 import _lib

    return _lib, _lib.load()


@pytest.mark.parametrize("b,d,dtype_name,math,tol", [
    (700, 768, "bfloat16", 0, 1e-3),    # dim >= 512: the forward stashes, the backward converts
    (700, 512, "bfloat16", 0, 1e-3),
    (700, 256, "bfloat16", 0, 1e-3),    # dim < 512: the backward recomputes
    (260, 256, "float32", 1, 1e-5),     # fp32 parity mode never stashes
])
def test_two_call_abi_matches_oracle(b, d, dtype_name, math, tol):
    _lib, lib = _bind()
    dtype = getattr(torch, dtype_name)
    embs = closed_form.synthetic_embeddings(b, d, 11, 0.15)
    if dtype == torch.bfloat16:
        embs = [closed_form.round_to_bf16(e) for e in embs]
    t3v, g3v = (2.6592, 2.4, 3.0), (1.0, 0.5, 0.25)
    want = closed_form.tri_contrastive(*embs, t3v, g3v)
    img, txt, aud = [torch.from_numpy(e).cuda().to(dtype) for e in embs]
    pb = _lib.Problem(rows_local=b, rows_global=b, row_offset=0, dim=d, dtype=0 if dtype == torch.float32 else 1,
                      math=math, world=1, parity=0)
    lay = _lib.Layout()
    assert lib.sclip_plan(byref(pb), byref(lay)) == 0
    ws = torch.empty(int(lay.total_bytes) + 256, dtype=torch.uint8, device="cuda")
    ws = ws[(-ws.data_ptr()) % 256:]
    ws[int(lay.sync):int(lay.sync) + 256].zero_()  # the owner zeroes the `sync` area once
    st = c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: c_void_p(t.data_ptr())  # noqa: E731
    t3 = torch.tensor(t3v, dtype=torch.float32, device="cuda")
    g3 = torch.tensor(g3v, dtype=torch.float32, device="cuda")
    loss3, dt3 = torch.empty(3, device="cuda"), torch.empty(3, device="cuda")
    grads = [torch.empty(b, d, dtype=torch.float32, device="cuda") for _ in range(3)]
    launches0 = lib.sclip_kernel_launches()
    rc = lib.sclip_forward(byref(pb), p(ws), p(img), p(txt), p(aud), p(t3), c_int(1), p(loss3), st)
    assert rc == 0, lib.sclip_last_error()
    rc = lib.sclip_backward(byref(pb), p(ws), p(img), p(txt), p(aud), p(t3), p(g3), *[p(g) for g in grads], c_int(1), p(dt3), st)
    assert rc == 0, lib.sclip_last_error()
    torch.cuda.synchronize()
    assert lib.sclip_kernel_launches() - launches0 <= 11
    assert np.max(np.abs(loss3.double().cpu().numpy() - want["loss"]) / want["loss"]) < tol
    for g, key in zip(grads, ("dimg", "dtxt", "daud")):
        assert golden_util.rel(g.double().cpu().numpy(), want[key]) < tol, key
    assert np.max(np.abs(dt3.double().cpu().numpy() - want["dscale"])) / np.max(np.abs(want["dscale"])) < tol
    # a forward serves ONE backward (a stash is converted in place): the second call is refused, nothing is launched
    before = lib.sclip_kernel_launches()
    rc = lib.sclip_backward(byref(pb), p(ws), p(img), p(txt), p(aud), p(t3), p(g3), *[p(g) for g in grads], c_int(1), p(dt3), st)
    assert rc == -1 and b"one backward" in lib.sclip_last_error()
    assert lib.sclip_kernel_launches() == before
    # evaluation: keep_for_backward = 0 leaves nothing for a backward either
    assert lib.sclip_forward(byref(pb), p(ws), p(img), p(txt), p(aud), p(t3), c_int(0), p(loss3), st) == 0
    assert lib.sclip_backward(byref(pb), p(ws), p(img), p(txt), p(aud), p(t3), p(g3), *[p(g) for g in grads], c_int(1), p(dt3), st) == -1
    status = (ctypes.c_int32 * 4)()
    assert lib.sclip_read_status(byref(pb), p(ws), status, st) == 0 and status[0] == 0


@pytest.mark.parametrize("b,d,planted,t,swap,flag,tol", [
    (300, 768, 0.15, 2.6592, False, 0, 1e-3),  # the default temperature: the stash is converted
    (300, 768, 0.1, 3.6889, True, 1, 1e-3),    # s = 40, one text row replaced by another sample's image: that negative
                                               # pair beats its positive pairs by ~36 nats -> the stash saturates (bit 0)
    # s = 43, positive-pair cosine 0.35: losses 7e-4 (bit 1).  Everything that is left of the gradient is proportional
    # to exp(L_ij - L_ii) of far negatives, and fp16 OPERANDS put ~s * 2^-11 of absolute error on every logit: the
    # recompute route -- the most accurate one this math mode has -- lands at 1.01e-3 here (measured), so this case
    # documents the limit of SCLIP_MATH_F16 rather than the 1e-3 bar (SCLIP_MATH_F16X3 is the route below such losses)
    (300, 768, 0.35, 3.7612, False, 2, 2e-3),
    (300, 768, 0.15, 4.6052, False, 4, 1e-3),  # s = 100 (bit 2): the stash route left dlogit_scale 2.1e-3 off here
    (300, 512, 0.0, 4.6052, False, 4, 1e-3),   # s = 100, untrained: the stash route was 1.5 % off on one pair
])
def test_stash_fallback_to_recompute(b, d, planted, t, swap, flag, tol):
    """The forward's fp16 stash E~ = exp(L_ij - (L_ii + L_jj)/2)/16 cannot carry every regime to 1e-3: it saturates when
    a negative pair beats both positive pairs by more than ln(16 * 65504), it has no bits left for what remains of the
    gradient when the softmax is extremely peaked, and at large scales the rounding of a few dominant elements shows in
    dlogit_scale.  The library notices each case on the device (status word 1) and `sclip_backward_scale` then
    recomputes G' from the similarities: the results must meet the oracle either way."""
    _lib, lib = _bind()
    embs = [closed_form.round_to_bf16(e) for e in closed_form.synthetic_embeddings(b, d, 5, planted)]
    if swap:
        embs[1][7] = embs[0][3]
    t3v, g3v = (t, t, t), (1.0, 0.5, 0.25)
    want = closed_form.tri_contrastive(*embs, t3v, g3v)
    img, txt, aud = [torch.from_numpy(e).cuda().bfloat16() for e in embs]
    pb = _lib.Problem(rows_local=b, rows_global=b, row_offset=0, dim=d, dtype=1, math=0, world=1, parity=0)
    lay = _lib.Layout()
    assert lib.sclip_plan(byref(pb), byref(lay)) == 0
    ws = torch.empty(int(lay.total_bytes) + 256, dtype=torch.uint8, device="cuda")
    ws = ws[(-ws.data_ptr()) % 256:]
    ws[int(lay.sync):int(lay.sync) + 256].zero_()
    st = c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda x: c_void_p(x.data_ptr())  # noqa: E731
    t3 = torch.tensor(t3v, dtype=torch.float32, device="cuda")
    g3 = torch.tensor(g3v, dtype=torch.float32, device="cuda")
    loss3, dt3 = torch.empty(3, device="cuda"), torch.empty(3, device="cuda")
    grads = [torch.empty(b, d, dtype=torch.float32, device="cuda") for _ in range(3)]
    assert lib.sclip_forward(byref(pb), p(ws), p(img), p(txt), p(aud), p(t3), c_int(1), p(loss3), st) == 0
    rc = lib.sclip_backward(byref(pb), p(ws), p(img), p(txt), p(aud), p(t3), p(g3), *[p(g) for g in grads], c_int(1), p(dt3), st)
    assert rc == 0, lib.sclip_last_error()
    status = (ctypes.c_int32 * 4)()
    assert lib.sclip_read_status(byref(pb), p(ws), status, st) == 0
    assert status[0] == 0 and status[1] == flag, list(status)
    errs = {"loss": np.max(np.abs(loss3.double().cpu().numpy() - want["loss"]) / want["loss"]),
            "dscale": np.max(np.abs(dt3.double().cpu().numpy() - want["dscale"])) / np.max(np.abs(want["dscale"]))}
    for g, key in zip(grads, ("dimg", "dtxt", "daud")):
        errs[key] = golden_util.rel(g.double().cpu().numpy(), want[key])
    assert max(errs.values()) < tol, errs


def test_status_word_reports_a_non_finite_forward():
    """`logit_scale` is unclamped in the reference (model.py:80-82): at exp(t) = e^100 every exponential overflows.  The
    losses are non-finite (as the reference's are) and the device status word says why."""
    from synergy_clip_b200 import _lib, ops

    embs = [torch.from_numpy(e).cuda() for e in closed_form.synthetic_embeddings(128, 64, 3)]
    t3 = torch.full((3,), 100.0, device="cuda")
    g3 = torch.ones(3, device="cuda")
    with pytest.raises(_lib.SclipError, match="not finite"):
        ops.forward_backward_raw(*embs, t3, g3, ops.TriContrastiveConfig(math="f16", check_status=True))
    loss3 = ops.forward_backward_raw(*embs, t3, g3, ops.TriContrastiveConfig(math="f16"))[0]
    assert not torch.isfinite(loss3).all()
