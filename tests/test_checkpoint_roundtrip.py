"""the checkpoint wire format.  `save_model` of the reference (main_pretraining.py:30-59) writes
`CLIP_model_*.tar` = {'model_state_dict': model.state_dict()} plus one `projection_head.tar` per modality; the consumers
(main_MMR.py:91-92, ZS_task.py:275-276, FT_image_task.py:119-120) `load_state_dict` them into a reference `Tri_CLIP`.
A model built from this package's mirror class must round-trip through those files with the UNMODIFIED reference
class on the other side, in both directions, and the two classes must then compute the same embeddings / logits.

CPU only; the reference class is imported from /root/reference (build container), so the cross-class half is skipped
where that tree is absent.  The self round trip runs everywhere."""
import os

import pytest
import torch

from oracle import ref_import
from tests.test_model_mirror import _batch, _tiny_config, tiny_encoders  # noqa: F401  (fixture)


def _save_like_reference(model, root, model_sz="base", text_des="caption"):
    """The statements of main_pretraining.py:30-59 (save_pretrained of the HF encoders left out: it is HF's own
    format, untouched by this package)."""
    os.makedirs(root, exist_ok=True)
    ckpt = os.path.join(root, f"CLIP_model_{model_sz}_{text_des}.tar")
    torch.save({"model_state_dict": model.state_dict()}, ckpt)
    heads = {}
    for modal, head in (("image", model.vision_projection), ("text", model.text_projection),
                        ("audio", model.audio_projection)):
        d = os.path.join(root, f"CLIP_{modal}_model_{model_sz}", text_des)
        os.makedirs(d, exist_ok=True)
        heads[modal] = os.path.join(d, "projection_head.tar")
        torch.save({"model_state_dict": head.state_dict()}, heads[modal])
    return ckpt, heads


def test_self_roundtrip_is_bit_exact(tiny_encoders, tmp_path):  # noqa: F811
    from synergy_clip_b200.model import Tri_CLIP

    torch.manual_seed(0)
    a = Tri_CLIP(_tiny_config(is_pt=False))
    with torch.no_grad():
        a.logit_scale_for_TA.fill_(3.25)
    ckpt, heads = _save_like_reference(a, str(tmp_path))
    torch.manual_seed(1)
    b = Tri_CLIP(_tiny_config(is_pt=False))
    b.load_state_dict(torch.load(ckpt, map_location="cpu")["model_state_dict"])  # main_MMR.py:91-92
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    # FT_text_task.py:97-98 style: a bare nn.Linear receives the projection head
    lin = torch.nn.Linear(48, 64, bias=False)
    lin.load_state_dict(torch.load(heads["text"], map_location="cpu")["model_state_dict"])
    assert torch.equal(lin.weight, a.text_projection.weight)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_roundtrip_with_the_unmodified_reference_class(tiny_encoders, tmp_path):  # noqa: F811
    from synergy_clip_b200.model import Tri_CLIP

    ref = ref_import.load_reference_model()
    torch.manual_seed(0)
    ours = Tri_CLIP(_tiny_config(is_pt=False, return_logits=True)).eval()
    with torch.no_grad():
        ours.logit_scale_for_IT.fill_(2.9)
    ckpt, _ = _save_like_reference(ours, str(tmp_path / "a"))
    torch.manual_seed(1)
    theirs = ref.Tri_CLIP(_tiny_config(is_pt=False, return_logits=True)).eval()
    # same key set, same order, same shapes: strict load in both directions
    assert list(theirs.state_dict().keys()) == list(ours.state_dict().keys())
    theirs.load_state_dict(torch.load(ckpt, map_location="cpu")["model_state_dict"], strict=True)
    batch = _batch(5)
    with torch.no_grad():
        lo, io, to, ao = ours(**batch)
        lt, it, tt, at = theirs(**batch)
    for x, y in zip((*lo, io, to, ao), (*lt, it, tt, at)):
        assert torch.allclose(x, y, atol=1e-6, rtol=1e-6)
    # and back: a checkpoint written by the reference class loads into the mirror
    with torch.no_grad():
        theirs.logit_scale_for_AI.fill_(1.75)
    ckpt2, _ = _save_like_reference(theirs, str(tmp_path / "b"))
    ours.load_state_dict(torch.load(ckpt2, map_location="cpu")["model_state_dict"], strict=True)
    assert abs(ours.logit_scale_for_AI.item() - 1.75) < 1e-7
    so, st = ours.state_dict(), theirs.state_dict()
    assert all(torch.equal(so[k], st[k]) for k in so)
