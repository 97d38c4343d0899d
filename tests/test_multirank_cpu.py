"""CPU, world_size 2 over gloo: the sharded choreography of synergy_clip_b200.ops (all-gather of normalised
shards, column-statistics merge, reduce-scatter of column-role gradients, DDP gradient scaling) reproduces
the reference loss tail evaluated on the concatenated global batch (SURVEY 8e parity definition).  The CUDA
stage calls are replaced by the fp64 test double in tests/emulated_backend.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import closed_form

WORLD = 2
T3 = (2.6592, 2.9, 2.2)
W3 = (0.3, 0.7, 1.1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rows_local, dim, dtype_name, math_mode, grad_scale, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from synergy_clip_b200 import fused_tri_contrastive, ops
        from tests.emulated_backend import EmulatedBackend

        ops._BACKEND = EmulatedBackend()
        dtype = getattr(torch, dtype_name)
        b = rows_local * world
        embs = closed_form.synthetic_embeddings(b, dim, 321, 0.2)
        if dtype == torch.bfloat16:
            embs = [closed_form.round_to_bf16(e) for e in embs]
        sl = slice(rank * rows_local, (rank + 1) * rows_local)
        leaves = [torch.from_numpy(e[sl].copy()).to(dtype).requires_grad_(True) for e in embs]
        ts = [torch.tensor(t, requires_grad=True) for t in T3]
        cfg = ops.TriContrastiveConfig(process_group=dist.group.WORLD, math=math_mode, grad_scale=grad_scale,
                                       grads_fp32=True)
        losses = fused_tri_contrastive(*leaves, *ts, config=cfg)
        sum(w * l for w, l in zip(W3, losses)).backward()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"),
                 loss=np.array([l.item() for l in losses]), dscale=np.array([t.grad.item() for t in ts]),
                 dimg=leaves[0].grad.double().numpy(), dtxt=leaves[1].grad.double().numpy(),
                 daud=leaves[2].grad.double().numpy())
    finally:
        dist.destroy_process_group()


# gtol: gradients read back through autograd land in `.grad` of a bf16 leaf, i.e. rounded to bf16 (2^-9 relative
# quantisation, Frobenius ~1.7e-3); the fp32 emission itself is checked against 1e-3 in the GPU parity tests
@pytest.mark.parametrize("dtype_name,math_mode,grad_scale,tol,gtol", [
    ("float32", "f16x3", "ddp", 2e-6, 2e-6),
    ("float32", "f16", "sum", 1e-3, 1e-3),
    ("bfloat16", "f16", "ddp", 1e-3, 4e-3),
])
def test_two_ranks_match_global_batch_oracle(tmp_path, dtype_name, math_mode, grad_scale, tol, gtol):
    rows_local, dim = 96, 64
    mp.spawn(_worker, args=(WORLD, _free_port(), rows_local, dim, dtype_name, math_mode, grad_scale, str(tmp_path)),
             nprocs=WORLD, join=True)
    embs = closed_form.synthetic_embeddings(rows_local * WORLD, dim, 321, 0.2)
    if dtype_name == "bfloat16":
        embs = [closed_form.round_to_bf16(e) for e in embs]
    want = closed_form.tri_contrastive(*embs, T3, W3)
    ranks = [dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(WORLD)]
    for r in ranks:  # every rank reports the global-batch losses
        assert np.max(np.abs(r["loss"] - want["loss"]) / want["loss"]) < tol
    mult = WORLD if grad_scale == "ddp" else 1
    for key in ("dimg", "dtxt", "daud"):
        got = np.concatenate([r[key] for r in ranks], axis=0) / mult
        err = np.sqrt(((got - want[key]) ** 2).sum() / (want[key] ** 2).sum())
        assert err < gtol, (key, err)
    if grad_scale == "ddp":  # DDP averages the parameter gradient over ranks
        dscale = np.mean([r["dscale"] for r in ranks], axis=0)
    else:                    # "sum": every rank already holds the all-reduced total
        dscale = ranks[0]["dscale"]
        assert np.allclose(ranks[0]["dscale"], ranks[1]["dscale"], rtol=1e-6)
    assert np.max(np.abs(dscale - want["dscale"])) / np.max(np.abs(want["dscale"])) < tol


def test_single_rank_emulation_matches_oracle():
    """The same test double with world_size 1 (no process group): guards the double itself."""
    from synergy_clip_b200 import fused_tri_contrastive, ops
    from tests.emulated_backend import EmulatedBackend

    saved = ops._BACKEND
    ops._BACKEND = EmulatedBackend()
    try:
        embs = closed_form.synthetic_embeddings(70, 48, 5, 0.1)
        want = closed_form.tri_contrastive(*embs, T3, W3)
        leaves = [torch.from_numpy(e).requires_grad_(True) for e in embs]
        ts = [torch.tensor(t, requires_grad=True) for t in T3]
        losses = fused_tri_contrastive(*leaves, *ts, config=ops.TriContrastiveConfig(math="f16x3"))
        sum(w * l for w, l in zip(W3, losses)).backward()
        assert np.max(np.abs(np.array([l.item() for l in losses]) - want["loss"]) / want["loss"]) < 2e-6
        for leaf, key in zip(leaves, ("dimg", "dtxt", "daud")):
            g = leaf.grad.double().numpy()
            assert np.sqrt(((g - want[key]) ** 2).sum() / (want[key] ** 2).sum()) < 2e-6, key
    finally:
        ops._BACKEND = saved


# ---- under the caller's DDP wrapper (main_pretraining.py:138): hooks fire on the projection heads and the three scales
class _Heads(torch.nn.Module):
    """The part of Tri_CLIP that sits on the path: three bias-free projection heads (model.py:76-78) and the three
    log-temperatures (model.py:80-82), fed with pooled features instead of encoders."""

    def __init__(self, hidden, dim, cfg):
        super().__init__()
        self.vision_projection = torch.nn.Linear(hidden, dim, bias=False)
        self.text_projection = torch.nn.Linear(hidden, dim, bias=False)
        self.audio_projection = torch.nn.Linear(hidden, dim, bias=False)
        self.logit_scale_for_IT = torch.nn.Parameter(torch.tensor(2.6592))
        self.logit_scale_for_TA = torch.nn.Parameter(torch.tensor(2.6592))
        self.logit_scale_for_AI = torch.nn.Parameter(torch.tensor(2.6592))
        self.cfg = cfg

    def forward(self, pv, pt, pa):
        from synergy_clip_b200 import fused_tri_contrastive

        return fused_tri_contrastive(self.vision_projection(pv), self.text_projection(pt), self.audio_projection(pa),
                                     self.logit_scale_for_IT, self.logit_scale_for_TA, self.logit_scale_for_AI,
                                     config=self.cfg)


def _ddp_worker(rank, world, port, rows_local, hidden, dim, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from synergy_clip_b200 import ops
        from tests.emulated_backend import EmulatedBackend

        ops._BACKEND = EmulatedBackend()
        torch.manual_seed(11)  # same weights on every rank
        cfg = ops.TriContrastiveConfig(process_group=dist.group.WORLD, math="f16x3", grad_scale="ddp")
        model = torch.nn.parallel.DistributedDataParallel(_Heads(hidden, dim, cfg))
        g = torch.Generator().manual_seed(5)
        pooled = [torch.randn(rows_local * world, hidden, generator=g) for _ in range(3)]
        sl = slice(rank * rows_local, (rank + 1) * rows_local)
        it, ta, ai = model(*[p[sl] for p in pooled])
        (it * W3[0] + ta * W3[1] + ai * W3[2]).backward()  # DDP's hooks average the parameter gradients over ranks
        torch.save({k: p.grad.clone() for k, p in model.module.named_parameters()},
                   os.path.join(out_dir, f"ddp{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_ddp_wrapper_averages_to_the_global_batch_gradient(tmp_path):
    """With grad_scale="ddp" every rank hands W x its partial derivative to autograd, so DDP's mean over ranks is the
    exact gradient of the global-batch losses with respect to the shared parameters."""
    from oracle import reference_tail

    rows_local, hidden, dim = 48, 24, 32
    mp.spawn(_ddp_worker, args=(WORLD, _free_port(), rows_local, hidden, dim, str(tmp_path)), nprocs=WORLD, join=True)
    grads = [torch.load(tmp_path / f"ddp{r}.pt") for r in range(WORLD)]
    for k in grads[0]:  # after the all-reduce every rank holds the same averaged gradient
        assert torch.allclose(grads[0][k], grads[1][k], rtol=1e-6, atol=1e-9), k
    # single-process truth: the reference tail (autograd, fp64) on the whole batch through the same heads
    torch.manual_seed(11)
    ref = _Heads(hidden, dim, None).double()
    g = torch.Generator().manual_seed(5)
    pooled = [torch.randn(rows_local * WORLD, hidden, generator=g).double() for _ in range(3)]
    losses = reference_tail.tail_losses(ref.vision_projection(pooled[0]), ref.text_projection(pooled[1]),
                                        ref.audio_projection(pooled[2]), ref.logit_scale_for_IT,
                                        ref.logit_scale_for_TA, ref.logit_scale_for_AI)
    sum(w * l for w, l in zip(W3, losses)).backward()
    for k, p in ref.named_parameters():
        want, got = p.grad, grads[0][k].double()
        err = (got - want).norm() / want.norm()
        assert err < 1e-5, (k, float(err))


def _wave_major_order(world, rank, tiles_per_rank, npairs, row_tiles):
    """Python restatement of decode_forward (csrc/sclip_tc.cu) for SCLIP_FWD_WAIT_PEERS: linear tile id -> (wave, pair,
    row tile, column tile), grouped rasterisation inside a wave."""
    col_tiles = world * tiles_per_rank
    per_wave = npairs * row_tiles * tiles_per_rank
    out = []
    for t in range(world * per_wave):
        wave, r = divmod(t, per_wave)
        pair, r = divmod(r, row_tiles * tiles_per_rank)
        group, r = divmod(r, 8 * tiles_per_rank)
        first = group * 8
        gm = min(row_tiles - first, 8)
        ti, tj = first + r % gm, r // gm
        tj = (rank * tiles_per_rank + ((world - wave) % world) * tiles_per_rank + tj) % col_tiles
        out.append((wave, pair, ti, tj))
    return out


@pytest.mark.parametrize("world,row_tiles", [(2, 8), (3, 5), (4, 16), (8, 8), (16, 3)])
def test_wave_major_tile_order_matches_the_pull_order(world, row_tiles):
    """Host-side model of the single-launch peer-memory forward: every (pair, row tile, column tile) is taken exactly
    once, wave 0 is this rank's own columns and wave w only touches the columns of rank - w.  Every rank pushes its
    shard to rank + 1 first, rank + 2 second, ... (`sclip_push_shards`), so the shard of rank - w is the w-th to land
    here: the kernel never needs a shard earlier than the exchange delivers it."""
    tiles_per_rank, npairs = 3, 3
    for rank in range(world):
        order = _wave_major_order(world, rank, tiles_per_rank, npairs, row_tiles)
        assert len(set((p, ti, tj) for _, p, ti, tj in order)) == len(order) == npairs * row_tiles * world * tiles_per_rank
        waves = [w for w, _, _, _ in order]
        assert waves == sorted(waves)
        for wave, _, _, tj in order:
            assert tj // tiles_per_rank == (rank - wave) % world


def test_config_validates_its_knobs(monkeypatch):
    """TriContrastiveConfig: unknown values raise; `push="auto"` follows SCLIP_PUSH and defaults to the copy engines."""
    from synergy_clip_b200 import ops

    for kw in (dict(math="fp8"), dict(grad_scale="mean"), dict(stash="yes"), dict(transport="mpi"), dict(push="dma")):
        with pytest.raises(ValueError):
            ops.TriContrastiveConfig(**kw)
    monkeypatch.delenv("SCLIP_PUSH", raising=False)
    assert ops.TriContrastiveConfig().push == "ce"
    monkeypatch.setenv("SCLIP_PUSH", "sm")
    assert ops.TriContrastiveConfig().push == "sm"
    assert ops.TriContrastiveConfig(push="ce").push == "ce"
    monkeypatch.setenv("SCLIP_PUSH", "nvlink")
    with pytest.raises(ValueError):
        ops.TriContrastiveConfig()
