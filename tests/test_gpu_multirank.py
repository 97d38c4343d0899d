"""GPU, world_size >= 2: the row-sharded CUDA path against the oracle on the concatenated global batch,
with both transports -- peer-memory pulls over symmetric memory ("p2p") and torch.distributed NCCL collectives
("nccl").  Skipped on single-GPU boxes."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import closed_form

pytestmark = pytest.mark.gpu

T3 = (2.6592, 2.9, 2.2)
W3 = (0.3, 0.7, 1.1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rows_local, dim, dtype_name, math_mode, grad_scale, transport, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from synergy_clip_b200 import ops

        dtype = getattr(torch, dtype_name)
        b = rows_local * world
        embs = closed_form.synthetic_embeddings(b, dim, 321, 0.2)
        if dtype == torch.bfloat16:
            embs = [closed_form.round_to_bf16(e) for e in embs]
        sl = slice(rank * rows_local, (rank + 1) * rows_local)
        ten = [torch.from_numpy(e[sl].copy()).cuda().to(dtype) for e in embs]
        t3 = torch.tensor(T3, dtype=torch.float32, device="cuda")
        g3 = torch.tensor(W3, dtype=torch.float32, device="cuda")
        cfg = ops.TriContrastiveConfig(process_group=dist.group.WORLD, math=math_mode, grad_scale=grad_scale,
                                       grads_fp32=True, transport=transport)
        for _ in range(3):  # repeated steps reuse the workspace: the cross-step ordering of the exchanges is exercised
            loss3, dimg, dtxt, daud, dt3 = ops.forward_backward_raw(*ten, t3, g3, cfg)
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=loss3.double().cpu().numpy(),
                 dscale=dt3.double().cpu().numpy(), dimg=dimg.double().cpu().numpy(),
                 dtxt=dtxt.double().cpu().numpy(), daud=daud.double().cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
@pytest.mark.parametrize("rows_local,dim,dtype_name,math_mode,grad_scale,tol", [
    (320, 256, "float32", "f16x3", "ddp", 1e-5),
    (1000, 512, "bfloat16", "f16", "ddp", 1e-3),
    (96, 64, "float32", "f16", "sum", 1e-3),
    (512, 768, "bfloat16", "f16", "ddp", 1e-3),   # whole 256-column tiles per rank: the pipelined (wave) schedule
])
def test_ranks_match_global_batch_oracle(tmp_path, rows_local, dim, dtype_name, math_mode, grad_scale, tol, transport):
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    mp.spawn(_worker, args=(world, _free_port(), rows_local, dim, dtype_name, math_mode, grad_scale, transport,
                            str(tmp_path)),
             nprocs=world, join=True)
    embs = closed_form.synthetic_embeddings(rows_local * world, dim, 321, 0.2)
    if dtype_name == "bfloat16":
        embs = [closed_form.round_to_bf16(e) for e in embs]
    want = closed_form.tri_contrastive(*embs, T3, W3)
    ranks = [dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(world)]
    for r in ranks:
        assert np.max(np.abs(r["loss"] - want["loss"]) / want["loss"]) < tol
    mult = world if grad_scale == "ddp" else 1
    for key in ("dimg", "dtxt", "daud"):
        got = np.concatenate([r[key] for r in ranks], axis=0) / mult
        err = np.sqrt(((got - want[key]) ** 2).sum() / (want[key] ** 2).sum())
        assert err < tol, (key, err)
    dscale = np.mean([r["dscale"] for r in ranks], axis=0) if grad_scale == "ddp" else ranks[0]["dscale"]
    assert np.max(np.abs(dscale - want["dscale"])) / np.max(np.abs(want["dscale"])) < tol


def _worker_no_set_device(rank, world, rows, dim, out_dir):
    """What the reference driver does: mp.spawn, `.to(rank)`, NO torch.cuda.set_device (main_pretraining.py:64,137-138,
    285-293) -- on every rank > 0 the current device stays 0 while the embeddings live on cuda:rank."""
    from synergy_clip_b200 import fused_tri_contrastive

    assert torch.cuda.current_device() == 0
    dev = torch.device("cuda", rank)
    embs = closed_form.synthetic_embeddings(rows, dim, 500 + rank, 0.2)
    leaves = [torch.from_numpy(e).to(dev).requires_grad_(True) for e in embs]
    ts = [torch.tensor(t, device=dev, requires_grad=True) for t in T3]
    losses = fused_tri_contrastive(*leaves, *ts)  # local batch, like the reference (no process group)
    sum(w * l for w, l in zip(W3, losses)).backward()
    with torch.no_grad():  # evaluation path and the scorers on the same device
        again = fused_tri_contrastive(*leaves, *ts)
    assert torch.cuda.current_device() == 0  # the op must not leak a device switch either
    np.savez(os.path.join(out_dir, f"nd{rank}.npz"), loss=np.array([l.item() for l in losses]),
             loss_eval=np.array([l.item() for l in again]), dimg=leaves[0].grad.double().cpu().numpy(),
             dscale=np.array([t.grad.item() for t in ts]))


def test_ranks_that_never_call_set_device(tmp_path):
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    rows, dim = 300, 256
    mp.spawn(_worker_no_set_device, args=(world, rows, dim, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        embs = closed_form.synthetic_embeddings(rows, dim, 500 + r, 0.2)
        want = closed_form.tri_contrastive(*embs, T3, W3)
        got = dict(np.load(tmp_path / f"nd{r}.npz"))
        assert np.max(np.abs(got["loss"] - want["loss"]) / want["loss"]) < 1e-5
        assert np.max(np.abs(got["loss_eval"] - want["loss"]) / want["loss"]) < 1e-5
        assert np.sqrt(((got["dimg"] - want["dimg"]) ** 2).sum() / (want["dimg"] ** 2).sum()) < 1e-5
        assert np.max(np.abs(got["dscale"] - want["dscale"])) / np.max(np.abs(want["dscale"])) < 1e-5
