"""Stage-by-stage diagnostics of the CUDA path against the numpy oracle (run on the GPU box).

    python tests/gpu_diag.py [--quick]

Prints one line per check; exits non-zero when a check fails.  Not collected by pytest.
"""
import ctypes
import math
import os
import sys
import time
from ctypes import byref

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import closed_form  # noqa: E402
from synergy_clip_b200 import _lib, ops  # noqa: E402

FAIL = []


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-300))


def report(name, err, tol):
    ok = bool(err < tol) and math.isfinite(err)
    print(f"{'PASS' if ok else 'FAIL'} {name}: err={err:.3e} tol={tol:.1e}", flush=True)
    if not ok:
        FAIL.append(name)


def check_gemm():
    torch.manual_seed(0)
    dev = "cuda"
    for (m, n, k) in [(128, 256, 64), (128, 256, 256), (200, 512, 320), (384, 768, 1000)]:
        a = (torch.randn(m, k, device=dev) * 0.5).half()
        b = (torch.randn(n, k, device=dev) * 0.5).half()
        ref = a.double() @ b.double().t()
        for a_mn in (False, True):
            for b_mn in (False, True):
                aa = a.t().contiguous() if a_mn else a
                bb = b.t().contiguous() if b_mn else b
                try:
                    c = ops.gemm_f16(aa, bb, a_mn=a_mn, b_mn=b_mn, alpha=0.5)
                    torch.cuda.synchronize()
                    err = rel(c.cpu().numpy(), 0.5 * ref.cpu().numpy())
                except Exception as e:  # noqa: BLE001
                    print("EXC", repr(e))
                    err = float("inf")
                report(f"gemm m{m} n{n} k{k} a_mn={int(a_mn)} b_mn={int(b_mn)}", err, 2e-3)


def run_case(b, d, dtype, t3, g3, seed, planted, math_mode, tol_loss, tol_grad, label):
    dev = "cuda"
    embs = closed_form.synthetic_embeddings(b, d, seed, planted)
    if dtype == torch.bfloat16:
        embs = [closed_form.round_to_bf16(e) for e in embs]
    want = closed_form.tri_contrastive(*embs, t3, g3)
    ten = [torch.from_numpy(e).to(dev).to(dtype) for e in embs]
    t3d = torch.tensor(t3, dtype=torch.float32, device=dev)
    g3d = torch.tensor(g3, dtype=torch.float32, device=dev)
    cfg = ops.TriContrastiveConfig(math=math_mode, grads_fp32=True)
    loss3, dimg, dtxt, daud, dt3 = ops.forward_backward_raw(*ten, t3d, g3d, cfg)
    torch.cuda.synchronize()
    report(f"{label} loss", float(np.max(np.abs(loss3.cpu().numpy() - want["loss"]) / np.abs(want["loss"]))), tol_loss)
    report(f"{label} dscale", float(np.max(np.abs(dt3.cpu().numpy() - want["dscale"])) / np.max(np.abs(want["dscale"]))), tol_grad)
    for nm, g in (("dimg", dimg), ("dtxt", dtxt), ("daud", daud)):
        report(f"{label} {nm}", rel(g.float().cpu().numpy(), want[nm]), tol_grad)


def check_prologue_and_stats(b=300, d=512):
    """Forward internals on one ragged case: xhat, lse_row, lse_col, diag."""
    dev = "cuda"
    embs = closed_form.synthetic_embeddings(b, d, 5, 0.2)
    t3 = (2.6592, 2.0, 3.0)
    ten = [torch.from_numpy(e).to(dev) for e in embs]
    t3d = torch.tensor(t3, dtype=torch.float32, device=dev)
    cfg = ops.TriContrastiveConfig(math="f16")
    pb, _, _ = ops._make_problem(ten[0], cfg)
    ws = ops._Workspace(pb, ten[0].device)
    loss3 = ops._forward_impl(ws, *ten, t3d, cfg)
    torch.cuda.synchronize()
    lay = ws.lay
    xhat = ws.view(lay.xhat, (3, b, d), torch.float16).float().cpu().numpy()
    hats = [closed_form.l2_normalise(e.astype(np.float64))[0] for e in embs]
    for m in range(3):
        report(f"prologue xhat[{m}]", rel(xhat[m], hats[m]), 1e-3)
    lse_row = ws.view(lay.lse_row, (3, b), torch.float32).cpu().numpy()
    lse_col = ws.view(lay.lse_col, (3, b), torch.float32).cpu().numpy()
    diag = ws.view(lay.diag, (3, b), torch.float32).cpu().numpy()
    for p, (_, r, c) in enumerate(closed_form.PAIRS):
        logits = math.exp(t3[p]) * hats[r] @ hats[c].T
        report(f"stats lse_row[{p}]", float(np.max(np.abs(lse_row[p] - closed_form._logsumexp(logits, 1)))), 5e-3)
        report(f"stats lse_col[{p}]", float(np.max(np.abs(lse_col[p] - closed_form._logsumexp(logits, 0)))), 5e-3)
        report(f"stats diag[{p}]", float(np.max(np.abs(diag[p] - np.diagonal(logits)))), 5e-3)
    print("loss3", loss3.cpu().numpy(), "status", ws.view(lay.status, (4,), torch.int32).cpu().numpy())


def main():
    quick = "--quick" in sys.argv
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    _lib.load()
    check_gemm()
    check_prologue_and_stats()
    s0 = 2.6592
    ln100 = math.log(100.0)
    run_case(256, 512, torch.float32, (s0, s0, s0), (1.0, 1.0, 1.0), 11, 0.0, "f16", 1e-3, 1e-3, "256x512 fp32/f16")
    run_case(35, 768, torch.float32, (s0, s0, s0), (0.25, 0.5, 0.125), 13, 0.0, "f16", 1e-3, 1e-3, "35x768 fp32/f16")
    run_case(1000, 1024, torch.bfloat16, (ln100,) * 3, (1.0, 0.5, 0.25), 20, 0.08, "f16", 1e-3, 1e-3, "1000x1024 bf16 ln100")
    run_case(2048, 512, torch.bfloat16, (s0, s0, s0), (0.25, 0.5, 0.125), 17, 0.0, "f16", 1e-3, 1e-3, "2048x512 bf16")
    run_case(256, 512, torch.float32, (s0, 2.7, 2.55), (0.25, 0.5, 0.125), 12, 0.0, "f16x3", 1e-5, 1e-5, "256x512 fp32/f16x3")
    run_case(300, 512, torch.float32, (ln100,) * 3, (1.0, 1.0, 1.0), 16, 0.1, "f16x3", 1e-5, 1e-5, "300x512 fp32/f16x3 ln100")
    if not quick:
        run_case(4096, 768, torch.bfloat16, (s0, s0, s0), (1.0, 1.0, 1.0), 19, 0.0, "f16", 1e-3, 1e-3, "4096x768 bf16")
    # timing of the headline shapes (CUDA events, after warm-up)
    for (b, d) in [(8192, 512)] + ([] if quick else [(32768, 768)]):
        ten = [torch.randn(b, d, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
        t3d = torch.full((3,), s0, device="cuda")
        g3d = torch.ones(3, device="cuda")
        cfg = ops.TriContrastiveConfig(math="f16")
        for _ in range(3):
            ops.forward_backward_raw(*ten, t3d, g3d, cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            out = ops.forward_backward_raw(*ten, t3d, g3d, cfg)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        tf = 18.0 * b * b * d / (ms * 1e-3) / 1e12
        print(f"TIMING B={b} D={d}: {ms:.3f} ms/step  {b / (ms * 1e-3):.3e} samples/s  {tf:.1f} TFLOP/s algorithmic "
              f"({100 * tf / 1608.8:.1f}% of measured bf16 peak) loss={out[0].cpu().numpy()}", flush=True)
    print("FAILED:", FAIL if FAIL else "none")
    return 1 if FAIL else 0


if __name__ == "__main__":
    sys.exit(main())
