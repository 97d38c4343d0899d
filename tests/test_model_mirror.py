"""The host-side mirror of the reference ``model.py`` interface: constructor / attributes / state-dict keys /
return modes on CPU (tiny randomly initialised encoders, no network), and the fused pre-training branch on GPU."""
import types

import numpy as np
import pytest
import torch

from oracle import closed_form


class _Sub:  # the reference's sub-configs are plain classes with class attributes (config.py:1-118)
    output_attentions = False
    output_hidden_states = False


def _tiny_config(is_pt=True, return_logits=False, return_lhs=False, dim=64):
    vis = type("V", (_Sub,), {"hidden_size": 32})
    txt = type("T", (_Sub,), {"hidden_size": 48})
    aud = type("A", (_Sub,), {"hidden_size": 40})
    return types.SimpleNamespace(vision_config=vis, text_config=txt, audio_config=aud, projection_dim=dim,
                                 logit_scale_init_value=2.6592, return_dict=False, is_PT=is_pt,
                                 return_logits=return_logits, return_lhs=return_lhs)


@pytest.fixture()
def tiny_encoders(monkeypatch):
    import transformers
    from transformers import ASTConfig, ASTModel, BertConfig, BertModel, CLIPVisionConfig, CLIPVisionModel

    def vis(path):
        return CLIPVisionModel(CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=1,
                                                num_attention_heads=2, image_size=32, patch_size=16))

    def txt(path):
        return BertModel(BertConfig(hidden_size=48, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2,
                                    vocab_size=100, max_position_embeddings=40))

    def aud(path):
        return ASTModel(ASTConfig(hidden_size=40, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2,
                                  max_length=64, num_mel_bins=32, patch_size=16, frequency_stride=10, time_stride=10))

    monkeypatch.setattr(transformers.CLIPVisionModel, "from_pretrained", staticmethod(vis))
    monkeypatch.setattr(transformers.AutoModel, "from_pretrained", staticmethod(txt))
    monkeypatch.setattr(transformers.ASTModel, "from_pretrained", staticmethod(aud))


def _batch(b, device="cpu"):
    g = torch.Generator().manual_seed(0)
    return dict(pixel_values=torch.randn(b, 3, 32, 32, generator=g).to(device),
                input_ids=torch.randint(0, 100, (b, 12), generator=g).to(device),
                att_mask=torch.ones(b, 12, dtype=torch.long).to(device),
                input_values=torch.randn(b, 64, 32, generator=g).to(device))


def test_constructor_attributes_and_state_dict_keys(tiny_encoders):
    from synergy_clip_b200.model import Tri_CLIP

    m = Tri_CLIP(_tiny_config(), vision_model_path="v", text_model_path="t", audio_model_path="a")
    keys = set(m.state_dict().keys())
    for k in ("logit_scale_for_IT", "logit_scale_for_TA", "logit_scale_for_AI", "vision_projection.weight",
              "text_projection.weight", "audio_projection.weight"):
        assert k in keys
    assert not any(k in keys for k in ("vision_projection.bias", "text_projection.bias", "audio_projection.bias"))
    assert {k.split(".")[0] for k in keys} == {"vision_model", "text_model", "audio_model", "vision_projection",
                                               "text_projection", "audio_projection", "logit_scale_for_IT",
                                               "logit_scale_for_TA", "logit_scale_for_AI"}
    assert m.logit_scale_for_IT.dim() == 0 and abs(m.logit_scale_for_IT.item() - 2.6592) < 1e-6
    for attr in ("vision_model", "text_model", "audio_model"):
        assert isinstance(getattr(m, attr), torch.nn.Module)


def test_non_pretraining_return_modes_on_cpu(tiny_encoders):
    from synergy_clip_b200.model import Tri_CLIP, clip_loss

    b = 6
    m = Tri_CLIP(_tiny_config(is_pt=False, return_logits=True)).eval()
    with torch.no_grad():
        logits, img, txt, aud = m(**_batch(b))
    assert [tuple(l.shape) for l in logits] == [(b, b)] * 3
    assert torch.allclose(img.norm(dim=-1), torch.ones(b), atol=1e-5)
    assert torch.allclose(logits[2], aud @ img.t() * m.logit_scale_for_AI.exp(), atol=1e-5)
    assert clip_loss(logits[0]).dim() == 0
    m.config.return_logits = False
    m.config.return_lhs = True
    with torch.no_grad():
        lhs = m(**_batch(b))
    assert lhs[0].shape[0] == b and lhs[0].dim() == 3
    m.config.return_lhs = False
    with torch.no_grad():
        embs = m(**_batch(b))
    assert all(e.shape == (b, 64) for e in embs)
    with torch.no_grad():
        s = m.get_img_txt_sim_score(**{k: v for k, v in _batch(b).items() if k != "input_values"})
    assert s.shape == (b, b)


def test_pretraining_branch_has_no_cpu_fallback(tiny_encoders):
    from synergy_clip_b200 import _lib
    from synergy_clip_b200.model import Tri_CLIP

    m = Tri_CLIP(_tiny_config(is_pt=True))
    with pytest.raises(_lib.SclipError):
        m(**_batch(4))


@pytest.mark.gpu
def test_pretraining_step_on_gpu_matches_oracle(tiny_encoders):
    """The loop body of main_pretraining.py:163-173 on a tiny model: three weighted losses, backward through the
    projection heads and the three logit scales; compared with the numpy oracle fed the same embeddings."""
    from synergy_clip_b200.model import Tri_CLIP

    b = 35  # the reference's per-GPU batch (main_pretraining.py:79)
    m = Tri_CLIP(_tiny_config(is_pt=True)).cuda().eval()  # eval: no dropout, so the embeddings can be recomputed
    batch = _batch(b, "cuda")
    out = m(**batch)
    alpha, beta, gamma = 0.15, 1.0, 1.0
    loss = out[0] * alpha + out[1] * beta + out[2] * gamma
    (loss / 4).backward()
    with torch.no_grad():
        m.config.is_PT = False
        m.config.return_logits = False
        m.config.return_lhs = False
        img = m.get_image_features(batch["pixel_values"])
        txt = m.get_text_features(batch["input_ids"], batch["att_mask"], None)
        aud = m.get_audio_features(batch["input_values"], None)
    want = closed_form.tri_contrastive(img.cpu().numpy(), txt.cpu().numpy(), aud.cpu().numpy(), (2.6592,) * 3,
                                       (alpha / 4, beta / 4, gamma / 4))
    got = np.array([o.item() for o in out])
    assert np.max(np.abs(got - want["loss"]) / want["loss"]) < 1e-5
    dt = np.array([m.logit_scale_for_IT.grad.item(), m.logit_scale_for_TA.grad.item(), m.logit_scale_for_AI.grad.item()])
    assert np.max(np.abs(dt - want["dscale"])) / np.max(np.abs(want["dscale"])) < 1e-5
    # d loss / d projection weight = dEmb^T @ pooled : check through the vision head
    assert m.vision_projection.weight.grad is not None and torch.isfinite(m.vision_projection.weight.grad).all()
    assert m.vision_projection.weight.grad.abs().sum().item() > 0
    # eval loop (main_pretraining.py:192-210): forward only, under no_grad
    m.config.is_PT = True
    m.eval()
    with torch.no_grad():
        out2 = m(**batch)
    assert all(not o.requires_grad for o in out2)


@pytest.mark.skipif(not __import__("os").path.isfile("/root/reference/model.py"),
                    reason="reference tree only exists in the build container")
def test_dropin_shim_reexports_reference_names():
    """`from model import *` through dropin/model.py yields every name the reference scripts use, with
    the three contrastive-path names replaced by this package's."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, types\n"
        "for n in ('pytorch_msssim', 'piqa'):\n"  # absent here and irrelevant to the contrastive path
        "    m = types.ModuleType(n); m.ssim = m.ms_ssim = m.SSIM = m.MS_SSIM = object; sys.modules[n] = m\n"
        "from model import *\n"
        "import synergy_clip_b200.model as ours\n"
        "assert Tri_CLIP is ours.Tri_CLIP and clip_loss is ours.clip_loss and contrastive_loss is ours.contrastive_loss\n"
        "for name in ('AutoTokenizer', 'AutoProcessor', 'TXT_AUD_2_IMG', 'IMG_AUD_2_TXT', 'IMG_TXT_2_AUD', 'torchvision', 'Image', 'ssim'):\n"
        "    assert name in globals(), name\n"
        "print('ok')\n")
    env = dict(os.environ, SCLIP_REFERENCE_DIR="/root/reference",
               PYTHONPATH=os.pathsep.join([os.path.join(root, "dropin"), root]))
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]


@pytest.mark.gpu
def test_zero_shot_scorers_on_gpu_match_the_reference_expression(tiny_encoders):
    """ZS_task.py:338,344 call the scorers under no_grad on the GPU: that path runs on the library and
    must agree with the reference's own three statements (model.py:160-167), which the same method executes when
    autograd is enabled."""
    from synergy_clip_b200.model import Tri_CLIP

    m = Tri_CLIP(_tiny_config(is_pt=False, return_logits=True)).cuda().eval()
    batch = _batch(9, "cuda")
    it_args = {k: v for k, v in batch.items() if k != "input_values"}
    ta_args = {k: v for k, v in batch.items() if k != "pixel_values"}
    with torch.no_grad():
        fused_it = m.get_img_txt_sim_score(**it_args)
        fused_ta = m.get_aud_txt_sim_score(**ta_args)
        fused_logits, _, _, _ = m(**batch)
    plain_it = m.get_img_txt_sim_score(**it_args).detach()
    plain_ta = m.get_aud_txt_sim_score(**ta_args).detach()
    plain_logits, _, _, _ = m(**batch)
    for got, want in ((fused_it, plain_it), (fused_ta, plain_ta), *zip(fused_logits, plain_logits)):
        want = want.detach()
        assert got.shape == want.shape
        assert ((got - want).norm() / want.norm()).item() < 1e-5
