"""GPU (ONE device): the row-sharded peer-memory path, rank by rank, through the C ABI.

Every "rank" is a workspace of its own on the same device and `peer_ws` is the table of those workspaces -- exactly
what the symmetric-memory mapping gives a real multi-GPU run, except that the peers' bytes come through local loads
instead of NVLink loads.  That exercises `sclip_push_shards` (including the per-rank "landed" flags and the two-copy exchange buffers), the single-launch
forward that waits on those flags (`SCLIP_FWD_WAIT_PEERS`), `sclip_forward_loss_peers`, the role-split gradient GEMMs
and `sclip_pull_reduce_cols` with the same kernels, launch arguments and workspace layout as `ops._forward_p2p` /
`ops._backward_impl`; the result is compared with the oracle on the concatenated batch (SURVEY 8e parity definition).
The ranks run one after the other on one stream, producers before consumers (a kernel that waited for a rank launched
behind it would never finish), so no two kernels wait on each other.  The multi-process, multi-GPU version of the same
comparison is tests/test_gpu_multirank.py."""
import ctypes
from ctypes import byref

import numpy as np
import pytest
import torch

from oracle import closed_form
from tests import golden_util

pytestmark = pytest.mark.gpu

T3 = (2.6592, 2.9, 2.2)
W3 = (0.3, 0.7, 1.1)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class _Rank:
    def __init__(self, lib_mod, rank, world, rows_local, dim, dtype, math):
        self.pb = lib_mod.Problem(rows_local, rows_local * world, rank * rows_local, dim,
                                  lib_mod.SCLIP_F32 if dtype == torch.float32 else lib_mod.SCLIP_BF16, math, world, 0)
        self.lay = lib_mod.plan(self.pb)
        raw = torch.empty(int(self.lay.total_bytes) + 256, dtype=torch.uint8, device="cuda")
        skew = (-raw.data_ptr()) % 256
        self.blob = raw[skew:skew + int(self.lay.total_bytes)]
        self.blob[int(self.lay.sync):int(self.lay.sync) + 256].zero_()  # the owner zeroes the sync area once
        self.ptr = ctypes.c_void_p(self.blob.data_ptr())

    def view(self, offset, shape, dtype):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.blob[int(offset):int(offset) + n].view(dtype).view(*shape)


def _run_sharded(world, rows_local, dim, dtype, math_name, stash, single_launch, grad_mult, steps=2, convert=True,
                 push_blocks=8, t3_values=T3, status_out=None):
    from synergy_clip_b200 import _lib

    lib = _lib.load()
    math = _lib.MATH_F16X3 if math_name == "f16x3" else _lib.MATH_F16
    b = rows_local * world
    embs = closed_form.synthetic_embeddings(b, dim, 321, 0.2)
    if dtype == torch.bfloat16:
        embs = [closed_form.round_to_bf16(e) for e in embs]
    ranks = [_Rank(_lib, r, world, rows_local, dim, dtype, math) for r in range(world)]
    table = (ctypes.c_void_p * world)(*[r.blob.data_ptr() for r in ranks])
    shards = [[torch.from_numpy(e[r * rows_local:(r + 1) * rows_local].copy()).cuda().to(dtype) for e in embs]
              for r in range(world)]
    t3 = torch.tensor(t3_values, dtype=torch.float32, device="cuda")
    g3 = torch.tensor(W3, dtype=torch.float32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    col_tiles = ranks[0].lay.col_tiles
    out = None
    for step in range(1, steps + 1):  # two steps on the same workspaces: the epoch flags must not be stale
        loss = [torch.empty(3, device="cuda") for _ in range(world)]
        for r, rk in enumerate(ranks):
            rk.pb.parity = step & 1  # the two copies of the exchange buffers alternate
            _lib.check(lib.sclip_prologue(byref(rk.pb), rk.ptr, *[_p(x) for x in shards[r]], _p(t3), 1 if stash else 0, st),
                       "prologue")
        for rk in ranks:  # (a real run: on the side stream, concurrently with the tiles)
            _lib.check(lib.sclip_push_shards(byref(rk.pb), rk.ptr, table, push_blocks, 1024, step, st), "push_shards")
        for rk in ranks:
            if not single_launch:
                _lib.check(lib.sclip_wait_shards(byref(rk.pb), rk.ptr, step, st), "wait_shards")
            flags = (1 if stash else 0) | (4 if single_launch else 0)
            _lib.check(lib.sclip_forward_tiles_cols(byref(rk.pb), rk.ptr, _p(t3), 7, 0, 0 if single_launch else col_tiles,
                                                    flags, 100, step, st), "forward_tiles_cols")
            _lib.check(lib.sclip_forward_reduce(byref(rk.pb), rk.ptr, st), "forward_reduce")
        for r, rk in enumerate(ranks):  # (a real run: barrier first)
            _lib.check(lib.sclip_forward_loss_peers(byref(rk.pb), rk.ptr, table, _p(loss[r]), st), "forward_loss_peers")
        for rk in ranks:
            conv = 1 if (stash and convert and lib.sclip_gemm_converts_stash(byref(rk.pb))) else 0
            if stash and conv:
                _lib.check(lib.sclip_backward_factors(byref(rk.pb), rk.ptr, _p(t3), _p(g3), st), "backward_factors")
            elif stash:
                _lib.check(lib.sclip_backward_scale(byref(rk.pb), rk.ptr, _p(t3), _p(g3), st), "backward_scale")
            else:
                _lib.check(lib.sclip_backward_tiles(byref(rk.pb), rk.ptr, _p(t3), _p(g3), st), "backward_tiles")
            _lib.check(lib.sclip_backward_gemms_role(byref(rk.pb), rk.ptr, _p(t3), _p(g3), 1, conv, 0, st), "gemms column role")
        grads = []
        for r, rk in enumerate(ranks):  # (a real run: barrier first, pull-reduce on the side stream under the row role)
            _lib.check(lib.sclip_pull_reduce_cols(byref(rk.pb), rk.ptr, table, 16, 512, st), "pull_reduce_cols")
            conv = 1 if (stash and convert and lib.sclip_gemm_converts_stash(byref(rk.pb))) else 0
            _lib.check(lib.sclip_backward_gemms_role(byref(rk.pb), rk.ptr, _p(t3), _p(g3), 2, conv, 128, st), "gemms row role")
            d3 = [torch.empty((rows_local, dim), dtype=torch.float32, device="cuda") for _ in range(3)]
            dt = torch.empty(3, device="cuda")
            col = rk.view(rk.lay.col_contrib, (3, rows_local, dim), torch.float32)
            _lib.check(lib.sclip_backward_finish(byref(rk.pb), rk.ptr, *[_p(x) for x in shards[r]], _p(t3), _p(g3), _p(col),
                                                 ctypes.c_float(grad_mult), *[_p(x) for x in d3], 1, 1 if stash else 0,
                                                 _p(dt), st), "backward_finish")
            grads.append((d3, dt))
        torch.cuda.synchronize()
        status = (ctypes.c_int32 * 4)()
        _lib.check(lib.sclip_read_status(byref(ranks[0].pb), ranks[0].ptr, status, st), "read_status")
        assert status[0] == 0
        if status_out is not None:
            for rk in ranks:
                _lib.check(lib.sclip_read_status(byref(rk.pb), rk.ptr, status, st), "read_status")
                status_out.append(status[1])
        out = (loss, grads)
    return embs, out


@pytest.mark.parametrize("world,rows_local,dim,dtype_name,math_name,stash,single_launch,tol", [
    (2, 512, 768, "bfloat16", "f16", True, True, 1e-3),     # the bench configuration in small: stash + one-launch forward
                                                            # + conversion inside the role-split GEMMs
    (4, 256, 768, "bfloat16", "f16", True, True, 1e-3),
    (4, 256, 512, "bfloat16", "f16", False, True, 1e-3),    # recompute backward behind the one-launch forward
    (2, 320, 256, "float32", "f16x3", False, False, 1e-5),  # ragged shards (no 256 multiple): plain column order, fp32 parity
    (3, 200, 520, "bfloat16", "f16", True, False, 1e-3),    # odd world size, odd dim
])
def test_emulated_ranks_match_global_batch_oracle(world, rows_local, dim, dtype_name, math_name, stash, single_launch, tol):
    dtype = getattr(torch, dtype_name)
    mult = float(world)  # grad_scale="ddp": the caller's DDP wrapper averages over ranks (main_pretraining.py:138)
    embs, (loss, grads) = _run_sharded(world, rows_local, dim, dtype, math_name, stash, single_launch, mult)
    want = closed_form.tri_contrastive(*embs, T3, W3)
    for l in loss:  # every rank holds the complete losses, bit-identical
        assert np.max(np.abs(l.double().cpu().numpy() - want["loss"]) / want["loss"]) < tol
        assert torch.equal(l, loss[0])
    for m, key in enumerate(("dimg", "dtxt", "daud")):
        got = np.concatenate([g[0][m].double().cpu().numpy() for g in grads], axis=0) / mult
        assert golden_util.rel(got, want[key]) < tol, key
    dscale = np.mean([g[1].double().cpu().numpy() for g in grads], axis=0)  # DDP's mean over ranks
    assert np.max(np.abs(dscale - want["dscale"])) / np.max(np.abs(want["dscale"])) < tol


@pytest.mark.parametrize("world,rows_local,dim,dtype_name,math_name,stash", [
    (4, 256, 768, "bfloat16", "f16", True),
    (2, 256, 512, "float32", "f16x3", False),   # six operand segments (hi + lo)
])
def test_copy_engine_push_is_bit_identical_to_the_kernel_push(world, rows_local, dim, dtype_name, math_name, stash):
    """sclip_push_shards with max_blocks == 0 moves the shards with strided peer copies and publishes the flags from a
    one-thread kernel: same bytes in the same places, so everything downstream must be identical."""
    dtype = getattr(torch, dtype_name)
    a = _run_sharded(world, rows_local, dim, dtype, math_name, stash, True, 1.0, steps=2, push_blocks=8)[1]
    b = _run_sharded(world, rows_local, dim, dtype, math_name, stash, True, 1.0, steps=2, push_blocks=0)[1]
    for la, lb in zip(a[0], b[0]):
        assert torch.equal(la, lb)
    for (da, ta), (db, tb) in zip(a[1], b[1]):
        assert torch.equal(ta, tb)
        for x, y in zip(da, db):
            assert torch.equal(x, y)


def test_sharded_stash_falls_back_to_recompute_at_a_large_scale():
    """Every rank stashes in the forward, finds exp(t) >= 44 in its backward (status word 1, bit 2) and recomputes its
    strip of G' instead of converting the stash: the sharded result must still be the global-batch oracle's."""
    t3 = (float(np.log(100.0)), 2.9, float(np.log(60.0)))
    flags = []
    embs, (loss, grads) = _run_sharded(2, 256, 768, torch.bfloat16, "f16", True, True, 2.0, steps=2, convert=False,
                                       t3_values=t3, status_out=flags)
    assert flags and all(f & 4 for f in flags), flags
    want = closed_form.tri_contrastive(*embs, t3, W3)
    for l in loss:
        assert np.max(np.abs(l.double().cpu().numpy() - want["loss"]) / want["loss"]) < 1e-3
    for m, key in enumerate(("dimg", "dtxt", "daud")):
        got = np.concatenate([g[0][m].double().cpu().numpy() for g in grads], axis=0) / 2.0
        assert golden_util.rel(got, want[key]) < 1e-3, key
    dscale = np.mean([g[1].double().cpu().numpy() for g in grads], axis=0)
    assert np.max(np.abs(dscale - want["dscale"])) / np.max(np.abs(want["dscale"])) < 1e-3


def test_conversion_in_the_gemm_is_bit_identical_to_the_hbm_pass():
    """dim 768: the GEMM kernel that converts the stash in its A-operand path (tensor memory) must give exactly what the
    in-place HBM pass + the plain GEMM kernel give -- same fp32 expression, same rounding to fp16, same MMA order --
    for the row role and the column role separately (the sharded path launches them one by one)."""
    a = _run_sharded(2, 512, 768, torch.bfloat16, "f16", True, True, 2.0, steps=1, convert=True)[1]
    b = _run_sharded(2, 512, 768, torch.bfloat16, "f16", True, True, 2.0, steps=1, convert=False)[1]
    for (ga, dta), (gb, dtb) in zip(a[1], b[1]):
        for x, y in zip(ga, gb):
            assert torch.equal(x, y)
        assert torch.equal(dta, dtb)
