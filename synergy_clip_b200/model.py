"""Host-side mirror of the reference's ``model.py`` interface for the contrastive path.

``Tri_CLIP`` keeps the reference's constructor, attribute names, state-dict keys, forward signature and
return modes (``/root/reference/model.py:60-281``) so that ``main_pretraining.py`` (calls at :132-135,
:163-166, :208-210) and the checkpoint consumers (``main_MMR.py:87-109``, ``ZS_task.py:271-276``) can use
it in place of the reference class.  The only behavioural difference is *where* the pre-training branch
(``config.is_PT``) is computed: lines 247-272 of the reference (normalise, three scaled similarity matmuls,
three ``clip_loss``) are one call into the sm_100a library (``synergy_clip_b200.ops``), which has no CPU
fallback.  The zero-shot scorers and the ``return_logits`` branch materialise their logits like the reference;
under ``torch.no_grad()`` on the GPU (how ZS_task.py and friends call them) the normalise + scaled matmul run on
the same library (``ops.cosine_logits``), otherwise the reference's own statements are executed.

Knobs the reference does not have are read from the environment so that the scripts run unchanged:
  SCLIP_GLOBAL_BATCH=1   use the *global* batch as negatives: row-shard over the default process group
                         (all-gather / reduce-scatter inside the op).  Default 0 = the reference's local batch.
  SCLIP_MATH=auto|f16|f16x3   tensor-core operand mode (auto: f16x3 for fp32 embeddings, f16 for bf16).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import TriContrastiveConfig, cosine_logits, fused_tri_contrastive

__all__ = ["Tri_CLIP", "clip_loss", "contrastive_loss"]


def contrastive_loss(logits: torch.Tensor) -> torch.Tensor:
    """Mean cross-entropy of each row against its own index (model.py:52-53)."""
    target = torch.arange(logits.shape[0], device=logits.device)
    return F.cross_entropy(logits, target)


def clip_loss(similarity: torch.Tensor) -> torch.Tensor:
    """Symmetric InfoNCE on a materialised similarity matrix (model.py:55-58).  Kept for callers that already
    hold logits; the pre-training path never materialises them (see ``fused_tri_contrastive``)."""
    return 0.5 * (contrastive_loss(similarity) + contrastive_loss(similarity.t()))


def _unit(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(p=2, dim=-1, keepdim=True)  # no epsilon, like model.py:248-250


def _scaled_cosine(a: torch.Tensor, b: torch.Tensor, log_scale: torch.Tensor) -> torch.Tensor:
    """``matmul(unit(a), unit(b).t()) * log_scale.exp()`` (model.py:160-167, 195-202, 252-265).  The evaluation callers
    (ZS_task.py:338,344 and the ``return_logits`` consumers ZS_image_task.py:1479, ZS_audio_task.py:195) run under
    ``torch.no_grad()`` on the GPU: those go through the library's normalise + tile kernels.  With autograd enabled, or
    on the CPU, the reference's own three statements are executed unchanged."""
    fused = (a.is_cuda and b.is_cuda and a.device == b.device and a.dtype == b.dtype and not torch.is_grad_enabled()
             and a.dtype in (torch.float32, torch.bfloat16) and a.dim() == 2 and b.dim() == 2
             and a.shape[1] == b.shape[1] and a.shape[1] % 8 == 0)
    if fused:
        return cosine_logits(a, b, log_scale, out_dtype=a.dtype)  # contiguous, dtype of the embeddings like the reference
    return torch.matmul(_unit(a), _unit(b).t()) * log_scale.exp()


def _op_config() -> TriContrastiveConfig:
    group = None
    if os.environ.get("SCLIP_GLOBAL_BATCH", "0") == "1":
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            group = dist.group.WORLD
    return TriContrastiveConfig(process_group=group, math=os.environ.get("SCLIP_MATH", "auto"), grad_scale="ddp")


class Tri_CLIP(nn.Module):
    """Three encoders + three bias-free projection heads + three learnable log-temperatures (model.py:60-82)."""

    def __init__(self, config, vision_model_path="openai/clip-vit-base-patch16",
                 text_model_path="openai/clip-vit-base-patch16",
                 audio_model_path="MIT/ast-finetuned-audioset-10-10-0.4593"):
        super().__init__()
        from transformers import ASTModel, AutoModel, CLIPVisionModel

        self.config = config
        self.vision_config = config.vision_config
        self.text_config = config.text_config
        self.audio_config = config.audio_config
        self.vision_model = CLIPVisionModel.from_pretrained(vision_model_path)
        self.text_model = AutoModel.from_pretrained(text_model_path)
        self.audio_model = ASTModel.from_pretrained(audio_model_path)
        dim = config.projection_dim
        self.vision_projection = nn.Linear(self.vision_config.hidden_size, dim, bias=False)
        self.text_projection = nn.Linear(self.text_config.hidden_size, dim, bias=False)
        self.audio_projection = nn.Linear(self.audio_config.hidden_size, dim, bias=False)
        init = float(config.logit_scale_init_value)
        self.logit_scale_for_IT = nn.Parameter(torch.tensor(init))
        self.logit_scale_for_TA = nn.Parameter(torch.tensor(init))
        self.logit_scale_for_AI = nn.Parameter(torch.tensor(init))

    # ---- encoders (the reference repeats these calls in every method; here they are written once) ----------
    def _vision(self, pixel_values):
        cfg = self.vision_config
        return self.vision_model(pixel_values=pixel_values, output_attentions=cfg.output_attentions,
                                 output_hidden_states=cfg.output_hidden_states, return_dict=self.config.return_dict)

    def _text(self, input_ids, att_mask, pos_ids):
        cfg = self.text_config
        return self.text_model(input_ids=input_ids, attention_mask=att_mask, position_ids=pos_ids,
                               output_attentions=cfg.output_attentions, output_hidden_states=cfg.output_hidden_states,
                               return_dict=self.config.return_dict)

    def _audio(self, input_values, head_mask):
        cfg = self.audio_config
        return self.audio_model(input_values=input_values, head_mask=head_mask,
                                output_attentions=cfg.output_attentions, output_hidden_states=cfg.output_hidden_states,
                                return_dict=self.config.return_dict)

    # ---- feature getters (model.py:84-124): projected pooler outputs, not normalised ------------------------
    def get_image_features(self, pixel_values):
        return self.vision_projection(self._vision(pixel_values)[1])

    def get_text_features(self, input_ids, att_mask, pos_ids):
        return self.text_projection(self._text(input_ids, att_mask, pos_ids)[1])

    def get_audio_features(self, input_values, head_mask):
        return self.audio_projection(self._audio(input_values, head_mask)[1])

    # ---- zero-shot scorers (model.py:126-203): materialised logits ------------------------------------------
    def get_img_txt_sim_score(self, pixel_values=None, input_ids=None, att_mask=None, pos_ids=None):
        img = self.get_image_features(pixel_values)
        txt = self.get_text_features(input_ids, att_mask, pos_ids)
        return _scaled_cosine(img, txt, self.logit_scale_for_IT)

    def get_aud_txt_sim_score(self, input_ids=None, att_mask=None, pos_ids=None, input_values=None, head_mask=None):
        txt = self.get_text_features(input_ids, att_mask, pos_ids)
        aud = self.get_audio_features(input_values, head_mask)
        return _scaled_cosine(txt, aud, self.logit_scale_for_TA)

    # ---- forward (model.py:205-281) -------------------------------------------------------------------------
    def forward(self, pixel_values=None, input_ids=None, att_mask=None, pos_ids=None, input_values=None,
                head_mask=None):
        vision_out = self._vision(pixel_values)
        text_out = self._text(input_ids, att_mask, pos_ids)
        audio_out = self._audio(input_values, head_mask)
        if (self.config.is_PT and os.environ.get("SCLIP_FUSED_PROJECTION", "0") == "1" and vision_out[1].is_cuda
                and os.environ.get("SCLIP_REFERENCE_TAIL", "0") != "1"):
            # model.py:234-272 as one node: the three projection GEMMs, the tail and the heads' backward GEMMs on the
            # library's tile kernels (fp16 tensor-core operands, see projection.py)
            from .projection import projected_tri_contrastive

            return projected_tri_contrastive(vision_out[1], text_out[1], audio_out[1], self.vision_projection.weight,
                                             self.text_projection.weight, self.audio_projection.weight,
                                             self.logit_scale_for_IT, self.logit_scale_for_TA, self.logit_scale_for_AI,
                                             config=_op_config())
        img = self.vision_projection(vision_out[1])
        txt = self.text_projection(text_out[1])
        aud = self.audio_projection(audio_out[1])

        if self.config.is_PT:
            if os.environ.get("SCLIP_REFERENCE_TAIL", "0") == "1":
                # A/B switch for measurements (bench.py --workload step): the reference's own statements, model.py:247-272
                return (clip_loss(_scaled_cosine(img, txt, self.logit_scale_for_IT)),
                        clip_loss(_scaled_cosine(txt, aud, self.logit_scale_for_TA)),
                        clip_loss(_scaled_cosine(aud, img, self.logit_scale_for_AI)))
            # model.py:247-272 as one fused op: (IT_loss, TA_loss, AI_loss), each a differentiable 0-dim tensor
            return fused_tri_contrastive(img, txt, aud, self.logit_scale_for_IT, self.logit_scale_for_TA,
                                         self.logit_scale_for_AI, config=_op_config())

        if self.config.return_logits:
            logits = (_scaled_cosine(img, txt, self.logit_scale_for_IT),
                      _scaled_cosine(txt, aud, self.logit_scale_for_TA),
                      _scaled_cosine(aud, img, self.logit_scale_for_AI))
            return logits, _unit(img), _unit(txt), _unit(aud)
        img, txt, aud = _unit(img), _unit(txt), _unit(aud)
        if self.config.return_lhs:
            return vision_out[0], text_out[0], audio_out[0]
        return img, txt, aud
