"""Helpers shared by the CPU and GPU parity tests: load a golden case, rebuild its inputs,
compare a result dict (loss, dscale, dimg, dtxt, daud) with the stored reference outputs."""
from __future__ import annotations

import json
import os

import numpy as np

from oracle import closed_form

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def manifest():
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as f:
        return json.load(f)["cases"]


def case_names():
    return [c["name"] for c in manifest()]


def load_case(name):
    meta = next(c for c in manifest() if c["name"] == name)
    data = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    embs = closed_form.synthetic_embeddings(meta["B"], meta["D"], meta["seed"], meta["planted"])
    if meta["bf16_inputs"]:
        embs = [closed_form.round_to_bf16(e) for e in embs]
    return meta, embs, data


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-300))


def golden_errors(meta, data, res):
    """Relative errors of ``res`` against the golden payload (Frobenius-relative for gradients)."""
    errs = {
        "loss": float(np.max(np.abs(res["loss"] - data["loss"]) / np.abs(data["loss"]))),
        # the three dlogit_scale values are compared like a gradient tensor (relative to its largest entry): a
        # component can be close to zero by cancellation (expected logit vs diagonal logit), where even the
        # reference's own fp32 run is only 1.3e-5 from its fp64 run (b2048x512_planted, third pair)
        "dscale": float(np.max(np.abs(res["dscale"] - data["dscale"])) / np.max(np.abs(data["dscale"]))),
    }
    rng = np.random.default_rng(meta["seed"] + 7919)
    rows = np.sort(rng.choice(meta["B"], size=min(16, meta["B"]), replace=False))
    proj = rng.standard_normal((meta["D"], 4))
    assert np.array_equal(rows, data["rows"])
    for key in ("dimg", "dtxt", "daud"):
        if key not in res:
            continue
        g = np.asarray(res[key], dtype=np.float64)
        if key in data:
            errs[key] = rel(g, data[key])
        errs[key + "_rows"] = rel(g[rows], data[key + "_rows"])
        errs[key + "_fro"] = abs(np.sqrt((g * g).sum()) - float(data[key + "_fro"])) / float(data[key + "_fro"])
        errs[key + "_proj"] = rel(g @ proj, data[key + "_proj"])
    return errs
