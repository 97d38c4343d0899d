"""Host side of the fused tri-modal contrastive objective.

``fused_tri_contrastive(img, txt, aud, t_IT, t_TA, t_AI)`` replaces lines 247-272 of the reference
``Tri_CLIP.forward`` (``/root/reference/model.py``) together with ``clip_loss`` / ``contrastive_loss``
(``model.py:52-58``): it returns the three 0-dim losses ``(IT_loss, TA_loss, AI_loss)`` and is
differentiable with respect to the three (un-normalised) embeddings and the three log-temperatures.

PyTorch is plumbing here: device memory (one workspace blob per problem shape), the current CUDA
stream, autograd wiring and -- when the batch is sharded over ranks -- the NCCL collectives of
``torch.distributed``.  All arithmetic happens in ``libsclip.so`` (``include/sclip.h``); there is
no CPU or eager fallback, and a missing library or a failed call raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import byref
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import MATH_F16, MATH_F16X3, Problem, SCLIP_BF16, SCLIP_F32

__all__ = ["fused_tri_contrastive", "TriContrastiveConfig", "cosine_logits", "gemm_f16", "workspace_bytes"]


class TriContrastiveConfig:
    """Knobs that do not exist in the reference (defaults reproduce it: local batch, no collectives).

    process_group : shard the *global* batch over the ranks of this group (row strips); ``None`` keeps the
                    reference behaviour (each rank's loss uses its local batch only, model.py:248-272).
    math          : "f16" (fp16 tensor-core operands, fp32 accumulate), "f16x3" (split operands, ~fp32
                    accuracy at 3x the tensor-core work) or "auto" (f16x3 for fp32 inputs, f16 for bf16).
    grad_scale    : "ddp"  -> gradients are multiplied by world_size because the caller's DDP wrapper averages
                              parameter gradients over ranks (main_pretraining.py:138);
                    "sum"  -> exact partial derivatives of the global-batch losses.
    grads_fp32    : emit fp32 embedding gradients even for bf16 inputs (used by the parity tests).
    overlap       : world_size > 1 only.  Run the collectives on a side stream under the tile kernels: the similarity
                    tiles start with this rank's own columns and then take the modalities in the order their
                    all-gathers complete; the column-role gradient GEMMs run first so that their reduce-scatter
                    overlaps the row-role GEMMs.  `comm_sms` SMs are left to the communication kernels meanwhile.
    """

    def __init__(self, process_group=None, math: str = "auto", grad_scale: str = "ddp", grads_fp32: bool = False,
                 overlap: bool = True, comm_sms: int = 20, stash="auto", transport: str = "auto",
                 check_status: bool = False, push: str = "auto"):
        if math not in ("auto", "f16", "f16x3"):
            raise ValueError(f"math={math!r}")
        if grad_scale not in ("ddp", "sum"):
            raise ValueError(f"grad_scale={grad_scale!r}")
        self.process_group = process_group
        self.math = math
        self.grad_scale = grad_scale
        self.grads_fp32 = grads_fp32
        self.overlap = overlap
        self.comm_sms = comm_sms
        # fp16-operand mode only: store the scaled exponentials of every tile in the forward ("stash") and convert them
        # in place in the backward (4 bytes of HBM traffic per logit) instead of recomputing the similarity matrices
        # (2 D flop per logit).  "auto": stash when D >= 512 -- measured on B200 (D = 512: 0.740 against 0.754 ms at 8192 rows,
        # 11.3-11.5 against 12.2 ms at 32768; below that the recompute is the cheaper side and stays the choice).
        if stash not in ("auto", True, False):
            raise ValueError(f"stash={stash!r}")
        self.stash = stash
        # world_size > 1: how the shards move between ranks.
        #   "p2p"  -- the workspace lives in symmetric memory (torch.distributed._symmetric_memory: every rank's blob
        #             mapped into every process over NVLink / NVSwitch) and the exchanges are kernels of the library
        #             that store to / load from the peers' workspaces (push all-gather of the operand shards under
        #             one flag-gated tile launch, pull reduce of the column-role gradients under the row-role GEMMs);
        #   "nccl" -- torch.distributed collectives (all-gather / reduce-scatter) on a side stream;
        #   "auto" -- "p2p" on CUDA when symmetric memory is available, else "nccl".
        if transport not in ("auto", "p2p", "nccl"):
            raise ValueError(f"transport={transport!r}")
        self.transport = transport
        # transport "p2p": what moves the operand shards into the peers' workspaces.
        #   "sm" -- a kernel of the library on `comm_sms` SMs (the similarity tiles get the rest);
        #   "ce" -- the copy engines (strided peer copies + one-thread flag kernels): the tiles keep every SM;
        #   "auto" -- the SCLIP_PUSH environment variable, else "ce" (8 GPUs, 32768 x 768, bench.py, three alternating
        #             pairs on two boxes: 2.25 / 2.32 / 2.35 ms against 2.39 / 2.46 / 2.42 ms with "sm").
        if push not in ("auto", "sm", "ce"):
            raise ValueError(f"push={push!r}")
        self.push = push if push != "auto" else os.environ.get("SCLIP_PUSH", "ce")
        if self.push not in ("sm", "ce"):
            raise ValueError(f"SCLIP_PUSH={self.push!r}")
        # Read the device status word after every forward (one host synchronisation per call) and raise if a row or
        # column log-sum-exp was not finite.  Off by default: the losses are non-finite too in that case, which the
        # caller's own `.item()` (main_pretraining.py:169-170) shows without an extra sync.
        self.check_status = check_status


_DEFAULT = TriContrastiveConfig()

# Optional stage tracer (bench.py / profiling only): a callable(stage_name) invoked after every stage launch.
_TRACE = None


def _mark(name: str) -> None:
    if _TRACE is not None:
        _TRACE(name)


_COMM_STREAMS = {}
_LAST_COMM_EVENTS = None  # probe only: (name, event) pairs of the side stream during a traced step


def _all_gather(pieces, group, coalesce):
    """pieces: [(output, this rank's input)]; one NCCL group launch when `coalesce`."""
    import torch.distributed as dist

    if coalesce and len(pieces) > 1:
        with dist._coalescing_manager(group=group):
            for out, inp in pieces:
                dist.all_gather_into_tensor(out, inp, group=group)
    else:
        for out, inp in pieces:
            dist.all_gather_into_tensor(out, inp, group=group)


def _reduce_scatter(pieces, group, coalesce):
    import torch.distributed as dist

    if coalesce and len(pieces) > 1:
        with dist._coalescing_manager(group=group):
            for out, inp in pieces:
                dist.reduce_scatter_tensor(out, inp, group=group)
    else:
        for out, inp in pieces:
            dist.reduce_scatter_tensor(out, inp, group=group)


def _comm_stream(device: torch.device) -> "torch.cuda.Stream":
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _COMM_STREAMS:
        _COMM_STREAMS[key] = torch.cuda.Stream(device=device, priority=-1)
    return _COMM_STREAMS[key]


def _sm_count(device: torch.device) -> int:
    return torch.cuda.get_device_properties(device).multi_processor_count


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _on_device:
    """Make the tensors' device the current CUDA device for the duration of a call.  The reference driver spawns its ranks
    with mp.spawn, moves the model with `.to(rank)` and never calls torch.cuda.set_device (main_pretraining.py:64,137-138),
    so on rank r > 0 the current device is still 0 while the embeddings live on cuda:r.  Streams, events, tensor-map
    encodes and kernel launches of the library all go to the current device: select the tensors' one first."""

    def __init__(self, t: torch.Tensor):
        self._ctx = torch.cuda.device(t.device) if t.is_cuda else None

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            return self._ctx.__exit__(*exc)
        return False


class _Workspace:
    """One workspace blob + typed views of the sub-buffers the host touches."""

    def __init__(self, pb: Problem, device: torch.device):
        self.pb = pb
        self.lay = _lib.plan(pb)
        raw = torch.empty(int(self.lay.total_bytes) + 256, dtype=torch.uint8, device=device)
        skew = (-raw.data_ptr()) % 256  # the ABI wants a 256-byte aligned blob
        self.blob = raw[skew:skew + int(self.lay.total_bytes)]
        self._init_sync()

    def _init_sync(self):
        """The flags / counters of the `sync` area live across calls: the owner zeroes them once (include/sclip.h)."""
        self.epoch = 0  # grows by one per sharded forward: the "shard landed" flags are compared against it
        self.blob[int(self.lay.sync):int(self.lay.sync) + 256].zero_()

    def view(self, offset: int, shape, dtype: torch.dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.blob[int(offset):int(offset) + nbytes].view(dtype).view(*shape)

    @property
    def ptr(self):
        return ctypes.c_void_p(self.blob.data_ptr())


class _SymmWorkspace(_Workspace):
    """Workspace blob in symmetric memory: `peer_ptrs[r]` is rank r's blob in this process' address space.
    Construction is a collective over the process group (every rank creates its workspaces in the same order)."""

    def __init__(self, pb: Problem, device: torch.device, group):
        import torch.distributed._symmetric_memory as symm

        self.pb = pb
        self.lay = _lib.plan(pb)
        nbytes = int(self.lay.total_bytes)
        raw = symm.empty(nbytes + 256, dtype=torch.uint8, device=device)
        if raw.data_ptr() % 256 != 0:
            raise _lib.SclipError("symmetric-memory allocation is not 256-byte aligned")
        self.hdl = symm.rendezvous(raw, group)
        self._raw = raw
        self.blob = raw[:nbytes]
        self._init_sync()
        ptrs = list(self.hdl.buffer_ptrs)
        self.peer_ptrs = (ctypes.c_void_p * len(ptrs))(*ptrs)
        self.group = group


_P2P_FAILED = False
_MAX_PEERS = 16  # SCLIP_MAX_PEERS (include/sclip.h)
_ROWS_CHECKED = set()


def _try_symm_workspace(pb: Problem, device, group):
    """Symmetric-memory workspace, or None when the platform refuses it (no peer mapping between these GPUs, handle
    exchange not permitted, ...).  The decision is taken by all ranks together -- a rank-local fallback would leave the
    others waiting in a barrier -- and is sticky: after one failure the NCCL transport is used for good."""
    global _P2P_FAILED
    import torch.distributed as dist

    if _P2P_FAILED:
        return None
    ws, err = None, None
    try:
        ws = _SymmWorkspace(pb, device, group)
    except Exception as e:  # noqa: BLE001
        err = e
    ok = torch.tensor([0 if ws is None else 1], device=device, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 1:
        return ws
    _P2P_FAILED = True
    import warnings

    warnings.warn(f"symmetric-memory workspace unavailable ({err!r}); the sharded path uses NCCL collectives instead")
    return None


def _symm_available() -> bool:
    try:
        import torch.distributed._symmetric_memory as symm  # noqa: F401

        return hasattr(symm, "empty") and hasattr(symm, "rendezvous")
    except Exception:  # noqa: BLE001
        return False


def _use_p2p(cfg, img) -> bool:
    if cfg.process_group is None or not img.is_cuda or _BACKEND.allows_cpu:
        return False
    if cfg.transport == "nccl":
        return False
    import torch.distributed as dist

    if dist.get_world_size(cfg.process_group) > _MAX_PEERS:  # SCLIP_MAX_PEERS: the pull kernels take at most 16 ranks
        if cfg.transport == "p2p":
            raise _lib.SclipError(f"transport='p2p' supports at most {_MAX_PEERS} ranks")
        return False
    if cfg.transport == "p2p":
        return True
    return _symm_available()


class _Pool:
    """Workspaces are recycled per problem shape; one is held from forward until its backward has run."""

    def __init__(self):
        self._free = {}
        self._lock = threading.Lock()

    @staticmethod
    def _key(pb: Problem, device, symm_group=None):
        return (device.index, pb.rows_local, pb.rows_global, pb.row_offset, pb.dim, pb.dtype, pb.math, pb.world,
                None if symm_group is None else id(symm_group))

    def acquire(self, pb: Problem, device, symm_group=None) -> _Workspace:
        with self._lock:
            lst = self._free.get(self._key(pb, device, symm_group))
            if lst:
                return lst.pop()
        if symm_group is not None:
            ws = _try_symm_workspace(pb, device, symm_group)
            if ws is not None:
                return ws
        return _Workspace(pb, device)

    def release(self, ws: _Workspace):
        with self._lock:
            self._free.setdefault(self._key(ws.pb, ws.blob.device, getattr(ws, "group", None)), []).append(ws)

    def acquire_sharded(self, pb: Problem, device, group, want_p2p: bool) -> _Workspace:
        """Sharded problems: a symmetric-memory workspace when the peer-memory transport is wanted and available."""
        if want_p2p and not _P2P_FAILED:
            return self.acquire(pb, device, group)
        return self.acquire(pb, device, None)

    def clear(self):
        with self._lock:
            self._free.clear()


_POOL = _Pool()


class _Lease:
    """A workspace held from a forward until its backward has run (released explicitly there; `__del__` only covers
    graphs that are dropped without a backward)."""

    def __init__(self, ws: _Workspace):
        self.ws = ws

    def release(self):
        ws, self.ws = self.ws, None
        if ws is not None:
            _POOL.release(ws)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def _make_problem(img: torch.Tensor, cfg: TriContrastiveConfig) -> Tuple[Problem, int, int]:
    import torch.distributed as dist

    rows, dim = img.shape
    world, rank = 1, 0
    if cfg.process_group is not None:
        world = dist.get_world_size(cfg.process_group)
        rank = dist.get_rank(cfg.process_group)
        # the row-strip layout assumes equal shards (rows_global = world * rows_local): checked once per shape
        key = (id(cfg.process_group), rows)
        if world > 1 and key not in _ROWS_CHECKED:
            counts = torch.tensor([rows, -rows], device=img.device, dtype=torch.int64)
            dist.all_reduce(counts, op=dist.ReduceOp.MAX, group=cfg.process_group)
            lo_hi = counts.tolist()
            if lo_hi[0] != rows or -lo_hi[1] != rows:
                raise ValueError(f"the sharded contrastive objective needs the same number of rows on every rank "
                                 f"(this rank: {rows}, range over ranks: {-lo_hi[1]}..{lo_hi[0]}); use drop_last=True")
            _ROWS_CHECKED.add(key)
    if img.dtype == torch.float32:
        dtype = SCLIP_F32
    elif img.dtype == torch.bfloat16:
        dtype = SCLIP_BF16
    else:
        raise TypeError(f"embeddings must be float32 or bfloat16, got {img.dtype}")
    math = cfg.math
    if math == "auto":
        math = "f16x3" if dtype == SCLIP_F32 else "f16"
    pb = Problem(rows_local=rows, rows_global=rows * world, row_offset=rows * rank, dim=dim, dtype=dtype,
                 math=MATH_F16X3 if math == "f16x3" else MATH_F16, world=world, parity=0)
    return pb, world, rank


def workspace_bytes(rows_local: int, dim: int, dtype=torch.bfloat16, world: int = 1, math: str = "auto") -> int:
    """Size of the workspace the library needs for a problem (pure host arithmetic, no GPU needed)."""
    d = SCLIP_F32 if dtype == torch.float32 else SCLIP_BF16
    if math == "auto":
        math = "f16x3" if d == SCLIP_F32 else "f16"
    pb = Problem(rows_local, rows_local * world, 0, dim, d, MATH_F16X3 if math == "f16x3" else MATH_F16, world, 0)
    return int(_lib.plan(pb).total_bytes)


def _check_inputs(img, txt, aud):
    for name, e in (("image", img), ("text", txt), ("audio", aud)):
        if not e.is_cuda and not _BACKEND.allows_cpu:
            raise _lib.SclipError(
                f"{name} embeddings are on {e.device}: the fused contrastive objective only runs on a CUDA (sm_100a) "
                "device and has no CPU fallback")
        if e.dim() != 2 or e.shape != img.shape or e.dtype != img.dtype or e.device != img.device:
            raise ValueError("image / text / audio embeddings must share one (B, D) shape, dtype and device")


class _CudaBackend:
    """The product path: every stage is one call into libsclip.so on the current CUDA stream."""

    allows_cpu = False

    def __init__(self):
        self._lib = None

    @property
    def lib(self):
        if self._lib is None:
            self._lib = _lib.load()
        return self._lib

    def prologue(self, ws, img, txt, aud, t3=None, diag=False):
        """normalise + cast; diag: also the positive-pair logits of this rank's rows (what a stash forward needs)"""
        _lib.check(self.lib.sclip_prologue(byref(ws.pb), ws.ptr, _ptr(img), _ptr(txt), _ptr(aud), _ptr(t3),
                                           1 if diag else 0, _stream()), "sclip_prologue")

    def forward_tiles_cols(self, ws, t3, pair_mask, tile_lo, tile_hi, stash=False, wrap=False, max_sms=0,
                           wait_epoch=None):
        flags = (1 if stash else 0) | (2 if wrap else 0) | (4 if wait_epoch is not None else 0)
        _lib.check(self.lib.sclip_forward_tiles_cols(byref(ws.pb), ws.ptr, _ptr(t3), int(pair_mask), int(tile_lo),
                                                     int(tile_hi), flags, int(max_sms), int(wait_epoch or 0),
                                                     _stream()),
                   "sclip_forward_tiles_cols")

    def backward_scale(self, ws, t3, g3):
        _lib.check(self.lib.sclip_backward_scale(byref(ws.pb), ws.ptr, _ptr(t3), _ptr(g3), _stream()),
                   "sclip_backward_scale")

    def backward_gemms_role(self, ws, t3, g3, role, max_sms=0, convert=False):
        _lib.check(self.lib.sclip_backward_gemms_role(byref(ws.pb), ws.ptr, _ptr(t3), _ptr(g3), int(role),
                                                      1 if convert else 0, int(max_sms), _stream()),
                   "sclip_backward_gemms_role")

    def gemm_converts_stash(self, ws):
        return bool(self.lib.sclip_gemm_converts_stash(byref(ws.pb)))

    def backward_factors(self, ws, t3, g3):
        _lib.check(self.lib.sclip_backward_factors(byref(ws.pb), ws.ptr, _ptr(t3), _ptr(g3), _stream()),
                   "sclip_backward_factors")

    def push_shards(self, ws, max_blocks, block_threads=1024, epoch=0):
        _lib.check(self.lib.sclip_push_shards(byref(ws.pb), ws.ptr, ws.peer_ptrs, int(max_blocks), int(block_threads),
                                              int(epoch), _stream()), "sclip_push_shards")

    def wait_shards(self, ws, epoch):
        _lib.check(self.lib.sclip_wait_shards(byref(ws.pb), ws.ptr, int(epoch), _stream()), "sclip_wait_shards")

    def forward_loss_peers(self, ws, loss3):
        _lib.check(self.lib.sclip_forward_loss_peers(byref(ws.pb), ws.ptr, ws.peer_ptrs, _ptr(loss3), _stream()),
                   "sclip_forward_loss_peers")

    def pull_reduce_cols(self, ws, max_blocks, block_threads=512):
        _lib.check(self.lib.sclip_pull_reduce_cols(byref(ws.pb), ws.ptr, ws.peer_ptrs, int(max_blocks),
                                                   int(block_threads), _stream()), "sclip_pull_reduce_cols")

    def forward_reduce(self, ws):
        _lib.check(self.lib.sclip_forward_reduce(byref(ws.pb), ws.ptr, _stream()), "sclip_forward_reduce")

    def forward_loss(self, ws, col_lse_all, loss3):
        _lib.check(self.lib.sclip_forward_loss(byref(ws.pb), ws.ptr, _ptr(col_lse_all), _ptr(loss3), _stream()),
                   "sclip_forward_loss")

    def backward_tiles(self, ws, t3, g3):
        _lib.check(self.lib.sclip_backward_tiles(byref(ws.pb), ws.ptr, _ptr(t3), _ptr(g3), _stream()),
                   "sclip_backward_tiles")

    def backward_finish(self, ws, img, txt, aud, t3, g3, col, mult, dimg, dtxt, daud, out_f32, dt3, stashed=False):
        _lib.check(
            self.lib.sclip_backward_finish(byref(ws.pb), ws.ptr, _ptr(img), _ptr(txt), _ptr(aud), _ptr(t3), _ptr(g3),
                                           _ptr(col), ctypes.c_float(mult), _ptr(dimg), _ptr(dtxt), _ptr(daud),
                                           int(out_f32), 1 if stashed else 0, _ptr(dt3), _stream()),
            "sclip_backward_finish")

    def read_status(self, ws):
        words = (ctypes.c_int32 * 4)()
        _lib.check(self.lib.sclip_read_status(byref(ws.pb), ws.ptr, words, _stream()), "sclip_read_status")
        return list(words)


# The stage executor.  The package ships exactly one (CUDA); the world_size > 1 CPU tests substitute a test double
# from tests/ to exercise the collective choreography below under the gloo backend.
_BACKEND = _CudaBackend()


def _check_status(ws: _Workspace, cfg: TriContrastiveConfig) -> None:
    if cfg.check_status and not _BACKEND.allows_cpu:
        words = _BACKEND.read_status(ws)
        if words[0] != 0:
            raise _lib.SclipError("a row or column log-sum-exp of the contrastive forward is not finite (non-finite "
                                  "embeddings, or exp(logit_scale) beyond the fp32 range): the losses are not finite")


def _forward_impl(ws: _Workspace, img, txt, aud, t3, cfg: TriContrastiveConfig, keep: bool = False) -> torch.Tensor:
    loss3 = _forward_stages(ws, img, txt, aud, t3, cfg, keep)
    _check_status(ws, cfg)
    return loss3


def _forward_stages(ws: _Workspace, img, txt, aud, t3, cfg: TriContrastiveConfig, keep: bool = False) -> torch.Tensor:
    """keep: a backward will follow on this workspace.  With fp16 operands the forward then also stores the scaled
    exponentials of every tile (the "stash") so that the backward does not recompute the similarity matrices."""
    be = _BACKEND
    pb, lay = ws.pb, ws.lay
    stash = bool(keep) and pb.math == MATH_F16 and (pb.dim >= 512 if cfg.stash == "auto" else bool(cfg.stash))
    ws.stashed = stash
    _mark("begin")
    loss3 = torch.empty(3, dtype=torch.float32, device=img.device)
    if isinstance(ws, _SymmWorkspace):
        return _forward_p2p(ws, img, txt, aud, t3, cfg, stash, loss3)
    be.prologue(ws, img, txt, aud, t3, diag=stash)
    _mark("prologue")
    if pb.world == 1:
        be.forward_tiles_cols(ws, t3, 7, 0, lay.col_tiles, stash)
        _mark("forward_tiles")
        be.forward_reduce(ws)
        be.forward_loss(ws, None, loss3)
        _mark("forward_finish")
        return loss3
    import torch.distributed as dist

    pg = cfg.process_group
    bl, bg, d, off = pb.rows_local, pb.rows_global, pb.dim, pb.row_offset
    # all-gather of the normalised row shards (each modality is column-side in one pair) and, for the stash, of the
    # positive-pair logits: one coalesced NCCL launch
    pieces = []
    bufs = [ws.view(lay.xhat, (3, bg, d), torch.float16)]
    if pb.math == MATH_F16X3:
        bufs.append(ws.view(lay.xhat_lo, (3, bg, d), torch.float16))
    for buf in bufs:
        for m in range(3):
            pieces.append((buf[m].view(-1), buf[m, off:off + bl].reshape(-1)))
    if stash:
        dg = ws.view(lay.diag_all, (3, bg), torch.float32)
        for p in range(3):
            pieces.append((dg[p].view(torch.float16), dg[p, off:off + bl].view(torch.float16)))

    overlap = cfg.overlap and img.is_cuda and bl % 256 == 0
    if not overlap:
        _all_gather(pieces, pg, img.is_cuda)
        _mark("all_gather")
        be.forward_tiles_cols(ws, t3, 7, 0, lay.col_tiles, stash)
    else:
        # the gather runs on the side stream under the tiles whose columns are this rank's own rows
        cur = torch.cuda.current_stream()
        comm = _comm_stream(img.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        landed = torch.cuda.Event()
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            _all_gather(pieces, pg, True)
            landed.record(comm)
        lo, hi = off // 256, (off + bl) // 256
        be.forward_tiles_cols(ws, t3, 7, lo, hi, stash, max_sms=_sm_count(img.device) - cfg.comm_sms)
        cur.wait_event(landed)
        be.forward_tiles_cols(ws, t3, 7, 0, lo, stash)
        be.forward_tiles_cols(ws, t3, 7, hi, lay.col_tiles, stash)
    _mark("forward_tiles")
    be.forward_reduce(ws)
    # column statistics: every rank holds the log-sum-exp over its own rows; merge them over ranks
    col_local = ws.view(lay.lse_col_local, (3, bg), torch.float32)
    col_all = torch.empty((pb.world, 3, bg), dtype=torch.float32, device=img.device)
    dist.all_gather_into_tensor(col_all.view(-1), col_local.reshape(-1), group=pg)
    be.forward_loss(ws, col_all, loss3)
    dist.all_reduce(loss3, group=pg)  # every rank reports the global-batch losses
    _mark("forward_finish")
    return loss3


def _forward_p2p(ws: "_SymmWorkspace", img, txt, aud, t3, cfg: TriContrastiveConfig, stash: bool, loss3: torch.Tensor):
    """world > 1, workspace in symmetric memory (every rank's blob mapped into every process over NVLink / NVSwitch).

    * all-gather of the operand shards = PUSH: after the prologue every rank writes its normalised shard straight into
      the peers' workspaces (`sclip_push_shards`, side stream, rank + 1 first) and publishes a per-source "landed" flag
      in each destination as it completes.  No barrier comes first: the exchange buffers exist twice and the steps
      alternate between the copies (`problem.parity`), so a peer that is still finishing the previous step reads the
      other copy.
    * the similarity tiles are ONE persistent launch (`SCLIP_FWD_WAIT_PEERS`): this rank's own columns first, then the
      peers' in arrival order (rank - 1, rank - 2, ...: everybody pushes to rank + 1 first), the TMA producer acquiring a rank's flag before the first tile on its columns -- no
      host-side waves, no per-wave launch / ramp / tail.  The push kernel runs on the `comm_sms` SMs the tiles leave.
    * statistics: one signal-pad barrier behind `sclip_forward_reduce`, then `sclip_forward_loss_peers` reads every
      rank's column statistics and row terms from the peers and computes the complete losses.

    Why a shard may be overwritten: rank r pushes step n + 1 into copy (n + 1) % 2 of peer q.  That copy was last read
    by q in step n - 1; q's push of step n -- which r has consumed, or it could not be in step n + 1 -- was enqueued
    behind q's whole step n - 1.  The column statistics a peer reads are rewritten one `sclip_forward_reduce` later,
    which lies behind the next step's tiles, which wait for every rank's next push, which lies behind that rank's
    `sclip_forward_loss_peers` of this step."""
    be = _BACKEND
    pb, lay, hdl = ws.pb, ws.lay, ws.hdl
    bl = pb.rows_local
    dev = ws.blob.device
    cur = torch.cuda.current_stream()
    comm = _comm_stream(dev)
    ws.epoch += 1
    pb.parity = ws.epoch & 1
    be.prologue(ws, img, txt, aud, t3, diag=stash)
    _mark("prologue")
    # the in-kernel wait needs the push kernels of all ranks to make progress next to the tile kernels: own SMs
    ce = cfg.push == "ce"  # copy engines move the shards: no SMs set aside
    pipelined = cfg.overlap and bl % 256 == 0 and (ce or cfg.comm_sms > 0)
    blocks = 0 if ce else 2 * max(cfg.comm_sms, 4)  # 1024-thread blocks, two per SM left free by the tile kernel
    ready = torch.cuda.Event()
    ready.record(cur)
    trace = _TRACE is not None
    pushed = torch.cuda.Event(enable_timing=trace)
    with torch.cuda.stream(comm):
        comm.wait_event(ready)
        be.push_shards(ws, blocks, 1024, ws.epoch)
        pushed.record(comm)
    if trace:
        global _LAST_COMM_EVENTS
        _LAST_COMM_EVENTS = [("pushes", pushed)]
    if pipelined:
        be.forward_tiles_cols(ws, t3, 7, 0, 0, stash, max_sms=0 if ce else _sm_count(dev) - cfg.comm_sms,
                              wait_epoch=ws.epoch)
    else:
        # ragged shards, or no SMs set aside for the pushes: wait for every shard first, then the plain column order
        be.wait_shards(ws, ws.epoch)
        be.forward_tiles_cols(ws, t3, 7, 0, lay.col_tiles, stash)
    cur.wait_event(pushed)  # this rank's own pushes are out before anything may rewrite the shard
    _mark("forward_tiles")
    be.forward_reduce(ws)
    _mark("forward_reduce")
    hdl.barrier(1)  # every rank's column statistics and row terms are complete
    _mark("forward_barrier1")
    be.forward_loss_peers(ws, loss3)  # merged column statistics + the complete losses, read from the peers
    _mark("forward_finish")
    return loss3


def _backward_impl(ws: _Workspace, img, txt, aud, t3, g3, cfg: TriContrastiveConfig):
    be = _BACKEND
    pb, lay = ws.pb, ws.lay
    out_f32 = 1 if (cfg.grads_fp32 or img.dtype == torch.float32) else 0
    gdtype = torch.float32 if out_f32 else img.dtype
    dimg, dtxt, daud = (torch.empty(img.shape, dtype=gdtype, device=img.device) for _ in range(3))
    dt3 = torch.empty(3, dtype=torch.float32, device=img.device)
    stashed = bool(getattr(ws, "stashed", False))
    # stash path: an in-place HBM pass converts the stash to G' before the gradient GEMMs (default).
    # SCLIP_CONVERT_IN_GEMM=1 selects the experimental GEMM that converts in its A-operand path through tensor memory
    # (dim 768 only).  It is bit-identical but SLOWER on B200 at 32768 x 768: 12.7-13.3 ms against 2.6 + 8.0 ms for
    # pass + GEMM, over five variants -- opt-in for A/B measurements, not the product path.
    convert = (stashed and not be.allows_cpu and os.environ.get("SCLIP_CONVERT_IN_GEMM", "0") == "1"
               and be.gemm_converts_stash(ws))
    _mark("backward_begin")
    if stashed:
        ws.stashed = False  # a stash serves one backward (the in-place pass consumes it)
        if convert:
            be.backward_factors(ws, t3, g3)
        else:
            be.backward_scale(ws, t3, g3)
    else:
        be.backward_tiles(ws, t3, g3)
    _mark("backward_tiles")
    col = None
    mult = 1.0
    if pb.world == 1:
        be.backward_gemms_role(ws, t3, g3, 0, convert=convert)
        _mark("backward_gemms")
    else:
        import torch.distributed as dist

        bl, bg, d = pb.rows_local, pb.rows_global, pb.dim
        part = ws.view(lay.dxhat_col, (3, bg, d), torch.float32)
        col = ws.view(lay.col_contrib, (3, bl, d), torch.float32)

        def scatter():  # reduce-scatter of the column-role partial gradients: one coalesced NCCL launch
            _reduce_scatter([(col[m].view(-1), part[m].view(-1)) for m in range(3)], cfg.process_group, img.is_cuda)

        if isinstance(ws, _SymmWorkspace):
            # column role first; its partial sums are then pulled and summed straight from the peers' workspaces
            # (NVLink loads, side stream) while the row-role GEMMs run on the remaining SMs
            cur = torch.cuda.current_stream()
            comm = _comm_stream(img.device)
            be.backward_gemms_role(ws, t3, g3, 1, convert=convert)
            _mark("backward_gemms_col")
            done = torch.cuda.Event()
            done.record(cur)
            trace = _TRACE is not None
            reduced = torch.cuda.Event(enable_timing=trace)
            bar_done = torch.cuda.Event(enable_timing=True) if trace else None
            with torch.cuda.stream(comm):
                comm.wait_event(done)
                ws.hdl.barrier(0)  # every rank's column-role partial sums are complete
                if trace:
                    bar_done.record(comm)
                if cfg.comm_sms == 0:
                    be.pull_reduce_cols(ws, 2 * _sm_count(img.device), 256)  # beside the row-role GEMM CTAs
                else:
                    be.pull_reduce_cols(ws, 4 * max(cfg.comm_sms, 4), 512)
                reduced.record(comm)
            if trace:
                global _LAST_COMM_EVENTS
                _LAST_COMM_EVENTS = (_LAST_COMM_EVENTS or []) + [("bwd_barrier", bar_done), ("pull_reduce", reduced)]
            be.backward_gemms_role(ws, t3, g3, 2, max_sms=(_sm_count(img.device) - cfg.comm_sms) if cfg.overlap else 0,
                                   convert=convert)
            _mark("backward_gemms_row")
            cur.wait_event(reduced)
            _mark("backward_gemms")
        elif not (cfg.overlap and img.is_cuda):
            be.backward_gemms_role(ws, t3, g3, 0, convert=convert)
            _mark("backward_gemms")
            scatter()
            _mark("reduce_scatter")
        else:
            cur = torch.cuda.current_stream()
            comm = _comm_stream(img.device)
            be.backward_gemms_role(ws, t3, g3, 1, convert=convert)  # column role first ...
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(done)
                scatter()                            # ... its reduce-scatter runs under the row-role GEMMs
                reduced = torch.cuda.Event()
                reduced.record(comm)
            be.backward_gemms_role(ws, t3, g3, 2, max_sms=_sm_count(img.device) - cfg.comm_sms, convert=convert)
            cur.wait_event(reduced)
            _mark("backward_gemms")
        if cfg.grad_scale == "ddp":
            mult = float(pb.world)
    be.backward_finish(ws, img, txt, aud, t3, g3, col, mult, dimg, dtxt, daud, out_f32, dt3, stashed)
    _mark("backward_finish")
    if pb.world > 1 and cfg.grad_scale == "sum":
        import torch.distributed as dist

        dist.all_reduce(dt3, group=cfg.process_group)
    return dimg, dtxt, daud, dt3


class _TriContrastive(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, txt, aud, t3, cfg):
        img, txt, aud = img.contiguous(), txt.contiguous(), aud.contiguous()
        pb, _, _ = _make_problem(img, cfg)
        with _on_device(img):
            ws = _POOL.acquire_sharded(pb, img.device, cfg.process_group, _use_p2p(cfg, img))
            lease = _Lease(ws)
            loss3 = _forward_impl(ws, img, txt, aud, t3, cfg, keep=True)
        ctx.save_for_backward(img, txt, aud, t3)
        ctx.lease = lease
        ctx.cfg = cfg
        ctx.stashed = ws.stashed
        return loss3

    @staticmethod
    def backward(ctx, g3):
        img, txt, aud, t3 = ctx.saved_tensors
        ws = ctx.lease.ws
        if ws is None or (ctx.stashed and not ws.stashed):
            raise _lib.SclipError("the contrastive objective was already back-propagated once: its workspace (stashed "
                                  "tiles converted in place) cannot serve a second backward (retain_graph)")
        g3 = g3.to(torch.float32).contiguous()
        with _on_device(img):
            dimg, dtxt, daud, dt3 = _backward_impl(ws, img, txt, aud, t3, g3, ctx.cfg)
        # the workspace goes back to the pool here, on the stream that ran the backward (not at garbage-collection time)
        ctx.lease.release()
        if not ctx.cfg.grads_fp32:
            dimg, dtxt, daud = dimg.to(img.dtype), dtxt.to(img.dtype), daud.to(img.dtype)
        return dimg, dtxt, daud, dt3, None


def fused_tri_contrastive(img: torch.Tensor, txt: torch.Tensor, aud: torch.Tensor, t_IT: torch.Tensor,
                          t_TA: torch.Tensor, t_AI: torch.Tensor, config: Optional[TriContrastiveConfig] = None):
    """(IT_loss, TA_loss, AI_loss) of model.py:247-272 for (B, D) projection outputs and the three
    ``logit_scale_for_*`` parameters (0-dim fp32, model.py:80-82)."""
    cfg = config or _DEFAULT
    _check_inputs(img, txt, aud)
    t3 = torch.stack([t_IT.reshape(()), t_TA.reshape(()), t_AI.reshape(())]).to(device=img.device, dtype=torch.float32)
    needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (img, txt, aud, t_IT, t_TA, t_AI))
    if needs_grad:
        loss3 = _TriContrastive.apply(img, txt, aud, t3, cfg)
    else:  # eval loops run under torch.no_grad() (main_pretraining.py:192-210): nothing is kept for backward
        img, txt, aud = img.contiguous(), txt.contiguous(), aud.contiguous()
        pb, _, _ = _make_problem(img, cfg)
        with _on_device(img):
            ws = _POOL.acquire_sharded(pb, img.device, cfg.process_group, _use_p2p(cfg, img))
            try:
                loss3 = _forward_impl(ws, img.detach(), txt.detach(), aud.detach(), t3.detach(), cfg)
            finally:
                _POOL.release(ws)
    return loss3[0], loss3[1], loss3[2]


def forward_backward_raw(img, txt, aud, t3, g3, config: Optional[TriContrastiveConfig] = None):
    """One fwd+bwd without autograd bookkeeping (bench / tests): returns (loss3, dimg, dtxt, daud, dt3)."""
    cfg = config or _DEFAULT
    _check_inputs(img, txt, aud)
    pb, _, _ = _make_problem(img, cfg)
    with _on_device(img):
        ws = _POOL.acquire_sharded(pb, img.device, cfg.process_group, _use_p2p(cfg, img))
        try:
            loss3 = _forward_impl(ws, img, txt, aud, t3, cfg, keep=True)
            dimg, dtxt, daud, dt3 = _backward_impl(ws, img, txt, aud, t3, g3, cfg)
        finally:
            _POOL.release(ws)
    return loss3, dimg, dtxt, daud, dt3


def gemm_f16(a: torch.Tensor, b: torch.Tensor, a_mn: bool = False, b_mn: bool = False, alpha: float = 1.0):
    """C = alpha * A @ B^T-style contraction on the tcgen05 tile kernel (tests, zero-shot scorers).

    a_mn=False: ``a`` is (M, K); a_mn=True: ``a`` is (K, M).  b_mn=False: ``b`` is (N, K); b_mn=True: ``b`` is (K, N).
    """
    lib = _lib.load()
    if a.dtype != torch.float16 or b.dtype != torch.float16 or not a.is_cuda:
        raise TypeError("gemm_f16 takes CUDA float16 operands")
    a, b = a.contiguous(), b.contiguous()
    m, k = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    n, kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if k != kb:
        raise ValueError("inner dimensions differ")
    c = torch.empty((m, n), dtype=torch.float32, device=a.device)
    with _on_device(a):
        _lib.check(lib.sclip_gemm_f16(_ptr(a), a.stride(0), int(a_mn), _ptr(b), b.stride(0), int(b_mn), _ptr(c), n, m, n,
                                      k, ctypes.c_float(alpha), _stream()), "sclip_gemm_f16")
    return c


def cosine_logits(a: torch.Tensor, b: torch.Tensor, log_scale: torch.Tensor, math: str = "auto",
                  out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``exp(log_scale) * unit(a) @ unit(b).T`` as a materialised (M, N) fp32 matrix -- the arithmetic of the reference's
    zero-shot scorers (``get_img_txt_sim_score`` / ``get_aud_txt_sim_score``, model.py:126-203) and of its
    ``return_logits`` branch (model.py:275-277) on the library's normalise + tile kernels.  Forward only (the reference
    calls these under ``torch.no_grad()``: ZS_task.py:338,344); CUDA only, no CPU fallback.

    Returns a contiguous (M, N) matrix, fp32 by default (what the kernel accumulates in); ``out_dtype=a.dtype`` gives
    the reference's contract (logits in the dtype of the embeddings), which the ``Tri_CLIP`` mirror asks for."""
    lib = _lib.load()
    if not a.is_cuda or not b.is_cuda:
        raise _lib.SclipError("cosine_logits only runs on a CUDA (sm_100a) device and has no CPU fallback")
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1] or a.dtype != b.dtype or a.device != b.device:
        raise ValueError("cosine_logits takes (M, D) and (N, D) matrices of one dtype on one device")
    if a.dtype == torch.float32:
        dtype = SCLIP_F32
    elif a.dtype == torch.bfloat16:
        dtype = SCLIP_BF16
    else:
        raise TypeError(f"embeddings must be float32 or bfloat16, got {a.dtype}")
    if math == "auto":
        math = "f16x3" if dtype == SCLIP_F32 else "f16"
    mode = MATH_F16X3 if math == "f16x3" else MATH_F16
    a, b = a.detach().contiguous(), b.detach().contiguous()
    m, d = a.shape
    n = b.shape[0]
    need = ctypes.c_uint64()
    _lib.check(lib.sclip_cosine_logits_scratch(m, n, d, mode, byref(need)), "sclip_cosine_logits_scratch")
    with _on_device(a):
        # scratch comes from the caching allocator per call: it is tied to the calling stream like any torch tensor
        # (a process-wide cache keyed by size alone would be shared by concurrent calls on different streams)
        raw = torch.empty(int(need.value) + 256, dtype=torch.uint8, device=a.device)
        skew = (-raw.data_ptr()) % 256
        scratch = raw[skew:skew + int(need.value)]
        ldc = (n + 3) // 4 * 4
        out = torch.empty((m, ldc), dtype=torch.float32, device=a.device)
        t = log_scale.detach().reshape(1).to(device=a.device, dtype=torch.float32)
        _lib.check(lib.sclip_cosine_logits(_ptr(a), _ptr(b), _ptr(t), m, n, d, dtype, mode, _ptr(scratch), _ptr(out),
                                           ldc, _stream()), "sclip_cosine_logits")
        res = out if ldc == n else out[:, :n].contiguous()
        return res if out_dtype in (None, torch.float32) else res.to(out_dtype)
