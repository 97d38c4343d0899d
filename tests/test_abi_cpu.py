"""CPU: the C-ABI library builds/loads, exports every symbol include/sclip.h declares, and its host-only entry
points (sclip_plan, argument validation) behave.  No compute calls are made here."""
import ctypes
import os
import re

import pytest

from synergy_clip_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sclip.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sclip_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 13
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_lib.EXPORTS)
    assert lib.sclip_abi_version() == _lib.ABI_VERSION == 2


def test_plan_layout_is_consistent():
    pb = _lib.Problem(4096, 32768, 8192, 768, _lib.SCLIP_BF16, _lib.MATH_F16, 8, 0)
    lay = _lib.plan(pb)
    assert lay.row_tiles == 32 and lay.col_tiles == 128 and lay.ld_g == 32768
    offs = [getattr(lay, f) for f, _ in _lib.Layout._fields_[1:18]]
    assert all(o % 256 == 0 for o in offs)
    assert lay.total_bytes > 3 * 4096 * 32768 * 2  # the three G' strips dominate
    assert lay.dxhat_col != lay.dxhat_row
    single = _lib.plan(_lib.Problem(256, 256, 0, 512, _lib.SCLIP_F32, _lib.MATH_F16X3, 1, 0))
    assert single.dxhat_col == single.dxhat_row and single.xhat_lo != single.xhat


@pytest.mark.parametrize("bad", [
    dict(rows_local=0), dict(dim=100), dict(dtype=7), dict(math=5), dict(world=0),
    dict(rows_local=128, rows_global=64), dict(row_offset=200),
])
def test_plan_rejects_bad_problems(bad):
    kw = dict(rows_local=128, rows_global=128, row_offset=0, dim=512, dtype=0, math=0, world=1, parity=0)
    kw.update(bad)
    lay = _lib.Layout()
    rc = _lib.load().sclip_plan(ctypes.byref(_lib.Problem(**kw)), ctypes.byref(lay))
    assert rc == -1
    assert len(_lib.load().sclip_last_error()) > 0


def test_converting_gemm_predicate_is_host_arithmetic():
    """sclip_gemm_converts_stash: 1 only for fp16 operands and 384-column gradient tiles (dim 768); a convert flag on any
    other problem is an argument error, not a launch."""
    lib = _lib.load()
    yes = _lib.Problem(512, 512, 0, 768, _lib.SCLIP_BF16, _lib.MATH_F16, 1, 0)
    assert lib.sclip_gemm_converts_stash(ctypes.byref(yes)) == 1
    for pb in (_lib.Problem(512, 512, 0, 512, _lib.SCLIP_BF16, _lib.MATH_F16, 1, 0),
               _lib.Problem(512, 512, 0, 768, _lib.SCLIP_F32, _lib.MATH_F16X3, 1, 0)):
        assert lib.sclip_gemm_converts_stash(ctypes.byref(pb)) == 0
        assert lib.sclip_backward_gemms_role(ctypes.byref(pb), None, None, None, 0, 1, 0, None) == -1
    assert lib.sclip_gemm_converts_stash(None) == 0
    assert lib.sclip_backward_factors(ctypes.byref(yes), None, None, None, None) == -1


def test_null_arguments_are_errors_not_crashes():
    lib = _lib.load()
    assert lib.sclip_plan(None, None) == -1
    pb = _lib.Problem(128, 128, 0, 512, 0, 0, 1, 0)
    assert lib.sclip_forward_tiles(ctypes.byref(pb), None, None, None) == -1
    assert lib.sclip_gemm_f16(None, 0, 0, None, 0, 0, None, 0, 1, 4, 1, 1.0, None) == -1


def test_op_refuses_cpu_tensors():
    import torch

    from synergy_clip_b200 import fused_tri_contrastive

    x = torch.randn(8, 16)
    t = torch.tensor(2.6592)
    with pytest.raises(_lib.SclipError):
        fused_tri_contrastive(x, x, x, t, t, t)


def test_peer_memory_and_scorer_calls_validate_their_arguments():
    """No device work happens before the argument checks: these run on a machine without a GPU."""
    lib = _lib.load()
    need = ctypes.c_uint64()
    assert lib.sclip_cosine_logits_scratch(0, 10, 512, 0, ctypes.byref(need)) == -1
    assert lib.sclip_cosine_logits_scratch(128, 1000, 516, 0, ctypes.byref(need)) == -1      # dim not a multiple of 8
    assert lib.sclip_cosine_logits_scratch(128, 1000, 512, 1, ctypes.byref(need)) == 0
    assert need.value >= 2 * (128 + 1000) * 512 * 2                                            # hi + lo operand copies
    assert lib.sclip_cosine_logits(None, None, None, 8, 8, 64, 0, 0, None, None, 8, None) == -1
    fake_ws = ctypes.c_void_p(1 << 20)                                                         # 256-byte aligned, never touched
    single = _lib.Problem(128, 128, 0, 512, 1, 0, 1, 0)
    table = (ctypes.c_void_p * 2)(1 << 20, 2 << 20)
    assert lib.sclip_push_shards(ctypes.byref(single), fake_ws, table, 8, 256, 1, None) == -1        # world must be >= 2
    sharded = _lib.Problem(128, 256, 128, 512, 1, 0, 2, 0)                                     # rank 1 of 2
    assert lib.sclip_push_shards(ctypes.byref(sharded), fake_ws, table, 8, 256, 1, None) == -1       # peer_ws[rank] != ws
    assert b"own workspace" in lib.sclip_last_error()
    assert lib.sclip_pull_reduce_cols(ctypes.byref(sharded), fake_ws, None, 8, 256, None) == -1
    assert lib.sclip_forward_loss_peers(ctypes.byref(sharded), fake_ws, None, None, None) == -1
    assert lib.sclip_wait_shards(ctypes.byref(single), fake_ws, 1, None) == -1
    bad_parity = _lib.Problem(128, 128, 0, 512, 1, 0, 1, 1)                                    # parity needs world > 1
    assert lib.sclip_plan(ctypes.byref(bad_parity), ctypes.byref(_lib.Layout())) == -1
    # the single-launch forward (SCLIP_FWD_WAIT_PEERS = 4) needs equal shards that are multiples of 256 rows
    t3 = ctypes.c_void_p(4 << 20)
    assert lib.sclip_forward_tiles_cols(ctypes.byref(sharded), fake_ws, t3, 7, 0, 0, 4, 128, 1, None) == -1
    assert b"multiples of 256" in lib.sclip_last_error()
    assert lib.sclip_forward_tiles_cols(ctypes.byref(single), fake_ws, t3, 7, 0, 0, 4, 128, 1, None) == -1
    # stash prologue flag (SCLIP_PRO_DIAG = 1) without the scales / a second backward without a forward
    x = ctypes.c_void_p(8 << 20)
    assert lib.sclip_prologue(ctypes.byref(single), fake_ws, x, x, x, None, 1, None) == -1
    assert lib.sclip_backward(ctypes.byref(single), fake_ws, x, x, x, t3, t3, x, x, x, 0, t3, None) == -1
    assert b"preceding sclip_forward" in lib.sclip_last_error()
    assert lib.sclip_read_status(ctypes.byref(single), fake_ws, None, None) == -1
    assert lib.sclip_kernel_launches() == 0                                                    # nothing was launched
