"""Generate the committed golden vectors by running the UNMODIFIED reference code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case the reference's ``model.clip_loss`` (model.py:55-58) is driven
through the statement sequence of model.py:247-272 in fp64 ("truth") and fp32
("what the reference executes"); inputs come from
``oracle.closed_form.synthetic_embeddings`` (numpy PCG64, reproducible on the GPU
box) and are optionally rounded to bf16 first.  Small cases store the full
gradients, large ones 16 sampled rows per modality plus Frobenius norms and two
random projections, so the fixtures stay small.
"""
from __future__ import annotations

import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import closed_form, ref_import  # noqa: E402

LN100 = math.log(100.0)

# name, B, D, seed, bf16-rounded inputs, t3, g3, planted cosine
CASES = [
    ("cfg1_256x512_fp32", 256, 512, 11, False, (2.6592, 2.6592, 2.6592), (1.0, 1.0, 1.0), 0.0),
    ("cfg1_256x512_weighted", 256, 512, 12, False, (2.6592, 2.70, 2.55), (0.25, 0.5, 0.125), 0.0),
    ("ragged_35x768", 35, 768, 13, False, (2.6592, 2.6592, 2.6592), (0.25, 0.5, 0.125), 0.0),
    ("ragged_14x1024", 14, 1024, 14, False, (2.6592, 2.6592, 2.6592), (1.0, 1.0, 1.0), 0.0),
    ("ragged_32x768_bf16", 32, 768, 15, True, (2.6592, 2.6592, 2.6592), (1.0, 0.5, 0.25), 0.0),
    ("b300x512_ln100", 300, 512, 16, False, (LN100, LN100, LN100), (1.0, 1.0, 1.0), 0.1),
    ("b2048x512_bf16", 2048, 512, 17, True, (2.6592, 2.6592, 2.6592), (0.25, 0.5, 0.125), 0.0),
    ("b2048x512_planted", 2048, 512, 18, False, (2.6592, 3.2, 3.9), (1.0, 1.0, 1.0), 0.3),
    ("b4096x768_bf16", 4096, 768, 19, True, (2.6592, 2.6592, 2.6592), (1.0, 1.0, 1.0), 0.0),
    ("b1000x1024_bf16_ln100", 1000, 1024, 20, True, (LN100, LN100, LN100), (1.0, 0.5, 0.25), 0.08),
    # trained-like batch, sharply peaked softmax (the positive-pair entry of G' is a small difference): stash path
    ("b300x768_bf16_peaked", 300, 768, 21, True, (math.log(43.5),) * 3, (0.5, 1.0, 0.25), 0.25),
    # one pair per forward kernel: s = 14.3 and s = 43.5 take the folded-exponent epilogue, s = 100 the per-tile maximum
    ("b300x768_bf16_mixed", 300, 768, 22, True, (2.6592, math.log(43.5), LN100), (0.5, 1.0, 0.25), 0.12),
]

FULL_GRAD_MAX_B = 64
N_SAMPLED_ROWS = 16


def case_inputs(b, d, seed, bf16, planted):
    embs = closed_form.synthetic_embeddings(b, d, seed, planted)
    if bf16:
        embs = [closed_form.round_to_bf16(e) for e in embs]
    return embs


def summarise(name, res, b, d, seed):
    out = {"loss": res["loss"].astype(np.float64), "dscale": res["dscale"].astype(np.float64)}
    rng = np.random.default_rng(seed + 7919)
    rows = np.sort(rng.choice(b, size=min(N_SAMPLED_ROWS, b), replace=False))
    proj = rng.standard_normal((d, 4))
    out["rows"] = rows.astype(np.int64)
    for key in ("dimg", "dtxt", "daud"):
        g = res[key].astype(np.float64)
        if b <= FULL_GRAD_MAX_B:
            out[key] = g.astype(np.float32)
        out[key + "_rows"] = g[rows].astype(np.float32)
        out[key + "_fro"] = np.array(np.sqrt((g * g).sum()))
        out[key + "_proj"] = (g @ proj).astype(np.float64)
    return out


def main(only=None):
    """only: comma-separated case names to (re)generate; the other fixtures and their manifest entries are kept."""
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    manifest = []
    keep = set()
    if only:
        keep = set(only.split(","))
        with open(os.path.join(HERE, "manifest.json")) as f:
            manifest = [c for c in json.load(f)["cases"] if c["name"] not in keep]
    for name, b, d, seed, bf16, t3, g3, planted in CASES:
        if only and name not in keep:
            continue
        embs = case_inputs(b, d, seed, bf16, planted)
        ref64 = ref_import.reference_tail(*embs, t3, g3, dtype=torch.float64)
        ref32 = ref_import.reference_tail(*embs, t3, g3, dtype=torch.float32)
        payload = summarise(name, ref64, b, d, seed)
        payload["loss_fp32"] = ref32["loss"].astype(np.float64)
        payload["dscale_fp32"] = ref32["dscale"].astype(np.float64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **payload)
        manifest.append(
            {"name": name, "B": b, "D": d, "seed": seed, "bf16_inputs": bf16, "t3": list(t3), "g3": list(g3),
             "planted": planted, "loss_fp64": [float(x) for x in ref64["loss"]]}
        )
        print(name, ref64["loss"], ref64["dscale"], flush=True)
    order = {c[0]: i for i, c in enumerate(CASES)}
    manifest.sort(key=lambda c: order.get(c["name"], 1 << 30))
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "reference": "model.py:52-58,247-272 (unmodified, imported)",
                   "cases": manifest}, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
