"""CPU: the oracle (numpy closed form + torch port) against the reference-generated golden vectors."""
import numpy as np
import pytest

from oracle import closed_form, ref_import, reference_tail
from tests import golden_util

SMALL = [n for n in golden_util.case_names() if not n.startswith("b4096")]


@pytest.mark.parametrize("name", golden_util.case_names())
def test_closed_form_matches_reference_golden(name):
    meta, embs, data = golden_util.load_case(name)
    res = closed_form.tri_contrastive(*embs, meta["t3"], meta["g3"])
    errs = golden_util.golden_errors(meta, data, res)
    # fp64 vs fp64 (gradient fixtures are stored as fp32 => 6e-8 quantisation)
    assert errs["loss"] < 1e-12 and errs["dscale"] < 1e-9, errs
    for k, v in errs.items():
        assert v < (1e-9 if k.endswith(("_proj", "_fro")) else 2e-7), (k, v)


@pytest.mark.parametrize("name", SMALL)
def test_torch_port_matches_reference_golden(name):
    import torch

    meta, embs, data = golden_util.load_case(name)
    res = reference_tail.tail_forward_backward(*embs, meta["t3"], meta["g3"], dtype=torch.float64)
    errs = golden_util.golden_errors(meta, data, res)
    assert max(errs.values()) < 2e-7, errs
    res32 = reference_tail.tail_forward_backward(*embs, meta["t3"], meta["g3"], dtype=torch.float32)
    assert np.allclose(res32["loss"], data["loss_fp32"], rtol=2e-6, atol=0)


def test_anchor_value():
    # SURVEY 8c sanity anchor: random N(0,1), B=256, D=512, s=e^2.6592 -> clip_loss ~ 5.69 (ln 256 = 5.545)
    meta, embs, data = golden_util.load_case("cfg1_256x512_fp32")
    assert 5.6 < data["loss"].mean() < 5.9


def test_zero_row_is_nan_like_reference():
    # model.py:248 has no epsilon: a zero embedding row yields NaN losses; the oracle keeps that behaviour
    embs = closed_form.synthetic_embeddings(8, 16, 3)
    embs[0][2] = 0.0
    with np.errstate(all="ignore"):
        res = closed_form.tri_contrastive(*embs, (2.6592,) * 3, want_grads=False)
    assert np.isnan(res["loss"][0]) and np.isnan(res["loss"][2]) and np.isfinite(res["loss"][1])


def test_bf16_rounding_helper_matches_torch():
    import torch

    x = np.random.default_rng(0).standard_normal(4096).astype(np.float32)
    assert np.array_equal(closed_form.round_to_bf16(x), torch.from_numpy(x).bfloat16().float().numpy())


@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_live_reference_agrees_with_oracle():
    import torch

    embs = closed_form.synthetic_embeddings(64, 128, 99)
    t3, g3 = (2.6592, 2.9, 2.2), (0.3, 0.7, 1.1)
    ref = ref_import.reference_tail(*embs, t3, g3, dtype=torch.float64)
    mine = closed_form.tri_contrastive(*embs, t3, g3)
    for k in ("loss", "dscale", "dimg", "dtxt", "daud"):
        assert golden_util.rel(mine[k], ref[k]) < 1e-12, k


@pytest.mark.parametrize("name,block", [("cfg1_256x512_weighted", 100), ("ragged_35x768", 16), ("b2048x512_planted", 512),
                                        ("b300x512_ln100", 128)])
def test_blockwise_oracle_matches_reference_golden(name, block):
    """oracle/blockwise.py (the full-size oracle of tests/test_gpu_fullsize.py: B x B never held whole, gradients for a
    sample of rows) against the vectors the unmodified reference produced, and against the dense closed form."""
    from oracle import blockwise

    meta, embs, data = golden_util.load_case(name)
    rows = data["rows"]
    got = blockwise.tri_contrastive_rows(*embs, meta["t3"], meta["g3"], rows, block=block)
    assert np.max(np.abs(got["loss"].numpy() - data["loss"]) / np.abs(data["loss"])) < 1e-12
    assert np.max(np.abs(got["dscale"].numpy() - data["dscale"])) / np.max(np.abs(data["dscale"])) < 1e-9
    dense = closed_form.tri_contrastive(*embs, meta["t3"], meta["g3"])
    for key in ("dimg", "dtxt", "daud"):
        assert golden_util.rel(got[key + "_rows"].numpy(), data[key + "_rows"]) < 2e-7, key   # fixtures are fp32
        assert golden_util.rel(got[key + "_rows"].numpy(), dense[key][rows]) < 1e-12, key
