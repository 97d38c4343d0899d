"""GPU: parity at the HEADLINE shape (BASELINE.json north star: global batch 32768, dim 768, bf16).

The dense oracle needs nine 8 GiB fp64 matrices here, so the check uses ``oracle/blockwise.py``: the same closed form
evaluated 2048 rows at a time in fp64 on the GPU (pinned against the reference-generated golden vectors by
tests/test_oracle.py), with gradients for a sample of rows of each modality.  This is the only test that exercises the
6 GiB G' region (> 2^31-byte offsets), TMA coordinates up to 32768 and the K = 65536 accumulation chains of the
gradient GEMMs -- the kernels `bench.py` times."""
import numpy as np
import pytest
import torch

from oracle import blockwise
from tests import golden_util

pytestmark = pytest.mark.gpu

TOL = 1e-3  # north star: 1e-3 relative in bf16


def _inputs(b, d, planted):
    g = torch.Generator(device="cuda").manual_seed(1234)
    base = torch.randn(b, d, device="cuda", generator=g)
    w = (planted / (1.0 - planted)) ** 0.5
    return [(torch.randn(b, d, device="cuda", generator=g) + w * base).to(torch.bfloat16) for _ in range(3)]


@pytest.mark.parametrize("b,d,planted,t3,g3", [
    (32768, 768, 0.0, (2.6592, 2.6592, 2.6592), (1.0, 1.0, 1.0)),            # the bench workload itself
    (32768, 768, 0.2, (2.6592, 3.2, 2.9), (0.25, 0.5, 0.125)),               # trained-like: planted positives, weights
])
def test_headline_shape_matches_blockwise_oracle(b, d, planted, t3, g3):
    from synergy_clip_b200 import ops

    free, _ = torch.cuda.mem_get_info()
    if free < 24 << 30:
        pytest.skip("needs 24 GiB of free device memory")
    embs = _inputs(b, d, planted)
    t3d = torch.tensor(t3, dtype=torch.float32, device="cuda")
    g3d = torch.tensor(g3, dtype=torch.float32, device="cuda")
    cfg = ops.TriContrastiveConfig(math="f16", grads_fp32=True, check_status=True)
    loss3, dimg, dtxt, daud, dt3 = ops.forward_backward_raw(*embs, t3d, g3d, cfg)
    torch.cuda.synchronize()
    ops._POOL.clear()
    rng = np.random.default_rng(7)
    # rows from every region of the strip, including the first / last tile and both sides of the 2^31-byte offset
    rows = np.unique(np.concatenate([[0, 127, 128, 16383, 16384, 16385, b - 129, b - 1], rng.choice(b, 24, replace=False)]))
    want = blockwise.tri_contrastive_rows(*embs, t3, g3, rows, block=2048, device="cuda")
    got_loss = loss3.double().cpu().numpy()
    assert np.max(np.abs(got_loss - want["loss"].numpy()) / want["loss"].numpy()) < TOL
    got_dt = dt3.double().cpu().numpy()
    assert np.max(np.abs(got_dt - want["dscale"].numpy())) / np.max(np.abs(want["dscale"].numpy())) < TOL
    for got, key in ((dimg, "dimg_rows"), (dtxt, "dtxt_rows"), (daud, "daud_rows")):
        sample = got[torch.as_tensor(rows, device="cuda")].double().cpu().numpy()
        ref = want[key].numpy()
        assert golden_util.rel(sample, ref) < TOL, (key, golden_util.rel(sample, ref))
        per_row = np.sqrt(((sample - ref) ** 2).sum(1) / (ref ** 2).sum(1))
        assert per_row.max() < 2 * TOL, (key, per_row.max())  # no single sampled row is off either
    # the emitted bf16 gradients are the rounding of the fp32 emission (1.7e-3 quantisation floor by itself)
    cfg16 = ops.TriContrastiveConfig(math="f16", grads_fp32=False)
    out16 = ops.forward_backward_raw(*embs, t3d, g3d, cfg16)
    assert out16[1].dtype == torch.bfloat16 and torch.equal(out16[1], dimg.bfloat16())
    assert torch.equal(out16[0], loss3)  # and the step is deterministic
