"""Drop-in ``model`` module for the reference's entry scripts (``from model import *``,
main_pretraining.py:16, main_MMR.py:20).

Put this directory first on PYTHONPATH and point SCLIP_REFERENCE_DIR at the reference checkout:

    SCLIP_REFERENCE_DIR=/path/to/Synergy-CLIP PYTHONPATH=/path/to/repo/dropin:/path/to/repo python main_pretraining.py ...

Everything the scripts obtain from ``model`` that is *not* on the contrastive path (the MMR decoders
``TXT_AUD_2_IMG`` / ``IMG_AUD_2_TXT`` / ``IMG_TXT_2_AUD``, ``AutoTokenizer``, ``AutoProcessor``, ``torchvision``,
``Image``, ``ssim`` ...) is re-exported unchanged from the reference's own ``model.py``; ``Tri_CLIP``,
``clip_loss`` and ``contrastive_loss`` are replaced by the B200 implementations.  Without
SCLIP_REFERENCE_DIR only the contrastive-path names (plus the tokenizer / processor factories the
pre-training script needs) are provided.
"""
import importlib.util
import os
import sys

_ref_dir = os.environ.get("SCLIP_REFERENCE_DIR")
if _ref_dir:
    _spec = importlib.util.spec_from_file_location("_sclip_reference_model", os.path.join(_ref_dir, "model.py"))
    _ref = importlib.util.module_from_spec(_spec)
    sys.modules["_sclip_reference_model"] = _ref
    _spec.loader.exec_module(_ref)
    globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("_")})
else:
    from transformers import AutoProcessor, AutoTokenizer  # noqa: F401  (main_pretraining.py:117-118)

from synergy_clip_b200.model import Tri_CLIP, clip_loss, contrastive_loss  # noqa: E402,F401
