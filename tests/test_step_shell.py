"""the training-step shell around the path (main_pretraining.py:157-215).  The loop body is restated in
`tests/step_shell.py`; its log was recorded once with the UNMODIFIED reference `Tri_CLIP` on the CPU
(`tests/golden/step_shell.pt`, written by `tests/golden/make_step_golden.py`).  Here the same shell drives this
package's `Tri_CLIP` -- weighted losses, `.item()` reads, gradient accumulation, AdamW steps, the left-over step, the
no_grad evaluation loop -- and must reproduce what the reference logged."""
import os

import numpy as np
import pytest
import torch

from tests import step_shell

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_shell.pt")


def _model(monkeypatch, device):
    from synergy_clip_b200.model import Tri_CLIP

    step_shell.patch_tiny_encoders(lambda obj, name, value: monkeypatch.setattr(obj, name, value))
    gold = torch.load(GOLDEN, map_location="cpu")
    model = Tri_CLIP(step_shell.tiny_config())
    model.load_state_dict(gold["init"], strict=True)
    return model.to(device), gold["log"]


def test_shell_log_fixture_is_consistent():
    gold = torch.load(GOLDEN, map_location="cpu")
    assert len(gold["log"]["train"]) == 5 and len(gold["log"]["valid"]) == 2
    assert all(np.isfinite(v) for row in gold["log"]["train"] + gold["log"]["valid"] for v in row)
    # the weights alpha, beta, gamma = 0.15, 1, 1 (main_pretraining.py:97-99) show in the logged IT share
    assert gold["log"]["train"][0][0] < 0.2 * gold["log"]["train"][0][1]


@pytest.mark.gpu
def test_training_shell_reproduces_the_reference_log(monkeypatch):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model, want = _model(monkeypatch, "cuda")
    got = step_shell.run_shell(model, step_shell.synthetic_batches(5, 12, seed=7),
                               step_shell.synthetic_batches(2, 12, seed=8), torch.device("cuda"))
    # first micro-batch: nothing but the forward differs (fp32 parity mode of the fused op): 1e-5
    assert np.allclose(got["train"][0], want["train"][0], rtol=1e-5, atol=0)
    # later micro-batches and the evaluation loop include three AdamW steps on the fused op's gradients; AdamW divides
    # by sqrt(v): rounding-level gradient differences (CPU vs GPU encoders, 1e-6) grow to ~1e-4 on the losses
    for a, b in zip(got["train"][1:] + got["valid"], want["train"][1:] + want["valid"]):
        assert np.allclose(a, b, rtol=2e-3, atol=0), (a, b)
    assert np.allclose(got["scales"], want["scales"], rtol=0, atol=2e-4)
    # the logit scales moved (their gradients reached the optimizer through the fused op)
    assert all(abs(s - 2.6592) > 1e-4 for s in got["scales"])
