"""synergy_clip_b200 -- B200-native (sm_100a) implementation of Synergy-CLIP's tri-modal contrastive
objective (reference model.py:52-58, 247-272) behind the reference's own Python interface.

    from synergy_clip_b200 import fused_tri_contrastive      # the fused op (three losses, differentiable)
    from synergy_clip_b200.model import Tri_CLIP, clip_loss  # drop-in replacement of the reference model.py names
"""
from .ops import TriContrastiveConfig, fused_tri_contrastive, gemm_f16, workspace_bytes  # noqa: F401

__all__ = ["fused_tri_contrastive", "TriContrastiveConfig", "gemm_f16", "workspace_bytes"]
