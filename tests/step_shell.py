"""TEST INFRASTRUCTURE -- the loop body of the reference's training driver.

`run_shell` restates `main_pretraining.py:157-187` (training: three weighted losses, four `.item()` reads per
micro-batch, `loss / accumulation_steps`, `.backward()`, optimizer step every `accumulation_steps` micro-batches, the
left-over step at :185-187) and `:200-215` (evaluation under `torch.no_grad()`), with the data loader replaced by seeded
synthetic batches and DDP left out (one process).  It is run once with the UNMODIFIED reference `Tri_CLIP`
(`tests/golden/make_step_golden.py`, CPU) to record the logged losses, and by `tests/test_step_shell.py` with this
package's `Tri_CLIP` on the GPU, which must reproduce that log.
"""
from __future__ import annotations

import types

import torch


class _Sub:  # the reference's sub-configs are plain classes with class attributes (config.py:1-118)
    output_attentions = False
    output_hidden_states = False


def tiny_config(dim=64):
    vis = type("V", (_Sub,), {"hidden_size": 32})
    txt = type("T", (_Sub,), {"hidden_size": 48})
    aud = type("A", (_Sub,), {"hidden_size": 40})
    return types.SimpleNamespace(vision_config=vis, text_config=txt, audio_config=aud, projection_dim=dim,
                                 logit_scale_init_value=2.6592, return_dict=False, is_PT=True,
                                 return_logits=False, return_lhs=False)


def patch_tiny_encoders(monkeypatch_setattr):
    """`from_pretrained` needs the network; the shell runs on config-initialised tiny encoders without dropout."""
    import transformers
    from transformers import ASTConfig, ASTModel, BertConfig, BertModel, CLIPVisionConfig, CLIPVisionModel

    def vis(path):
        return CLIPVisionModel(CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=1,
                                                num_attention_heads=2, image_size=32, patch_size=16,
                                                attention_dropout=0.0))

    def txt(path):
        return BertModel(BertConfig(hidden_size=48, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2,
                                    vocab_size=100, max_position_embeddings=40, hidden_dropout_prob=0.0,
                                    attention_probs_dropout_prob=0.0))

    def aud(path):
        return ASTModel(ASTConfig(hidden_size=40, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2,
                                  max_length=64, num_mel_bins=32, patch_size=16, frequency_stride=10, time_stride=10,
                                  hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))

    monkeypatch_setattr(transformers.CLIPVisionModel, "from_pretrained", staticmethod(vis))
    monkeypatch_setattr(transformers.AutoModel, "from_pretrained", staticmethod(txt))
    monkeypatch_setattr(transformers.ASTModel, "from_pretrained", staticmethod(aud))


def synthetic_batches(n, batch, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        out.append((torch.randn(batch, 3, 32, 32, generator=g), torch.randn(batch, 64, 32, generator=g),
                    (torch.randint(0, 100, (batch, 12), generator=g), torch.ones(batch, 12, dtype=torch.long))))
    return out


def run_shell(model, train_batches, valid_batches, device, alpha=0.15, beta=1.0, gamma=1.0, accumulation_steps=2,
              lr=1e-3):
    """Returns {"train": [(IT, TA, AI) per micro-batch], "valid": [...], "scales": final logit scales}."""
    opt = torch.optim.AdamW(model.parameters(), lr=lr)  # main_pretraining.py:139
    log = {"train": [], "valid": []}
    model.train()
    opt.zero_grad()
    batch_idx = -1
    for batch_idx, (images, audios, (input_ids, att_mask)) in enumerate(train_batches):
        images, audios = images.to(device), audios.to(device)
        input_ids, att_mask = input_ids.to(device), att_mask.to(device)
        output = model(pixel_values=images, input_ids=input_ids, att_mask=att_mask, input_values=audios)
        IT, TA, AI = (output[0] * alpha), (output[1] * beta), (output[2] * gamma)
        loss = IT + TA + AI
        log["train"].append((IT.item(), TA.item(), AI.item()))
        loss = loss / accumulation_steps
        loss.backward()
        if (batch_idx + 1) % accumulation_steps == 0:
            opt.step()
            opt.zero_grad()
    if batch_idx % accumulation_steps != 0:  # main_pretraining.py:185-187
        opt.step()
        opt.zero_grad()
    model.eval()
    with torch.no_grad():
        for images, audios, (input_ids, att_mask) in valid_batches:
            images, audios = images.to(device), audios.to(device)
            input_ids, att_mask = input_ids.to(device), att_mask.to(device)
            output = model(pixel_values=images, input_ids=input_ids, att_mask=att_mask, input_values=audios)
            log["valid"].append((output[0].item() * alpha, output[1].item() * beta, output[2].item() * gamma))
    log["scales"] = [model.logit_scale_for_IT.item(), model.logit_scale_for_TA.item(), model.logit_scale_for_AI.item()]
    return log
